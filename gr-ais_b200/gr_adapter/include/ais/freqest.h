// gr::ais::freqest, B200 build (reference include/ais/freqest.h:37-50): vectors of fftlen
// complex spectrum bins in, one float (Hz) out per vector.
#ifndef B200AIS_GR_AIS_FREQEST_H
#define B200AIS_GR_AIS_FREQEST_H

#include <ais/api.h>
#ifdef B200AIS_HAVE_GNURADIO
#include <gnuradio/sync_block.h>
#endif

namespace gr {
namespace ais {

class AIS_API freqest : virtual public gr::sync_block
{
public:
    typedef boost::shared_ptr<freqest> sptr;
    static sptr make(float sample_rate, int data_rate, int fftlen);
};

} // namespace ais
} // namespace gr

#endif
