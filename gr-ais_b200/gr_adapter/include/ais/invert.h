// gr::ais::invert, B200 build (reference include/ais/invert.h:37-50): char in, char out,
// out = (in ^ 1) & 1.
#ifndef B200AIS_GR_AIS_INVERT_H
#define B200AIS_GR_AIS_INVERT_H

#include <ais/api.h>
#ifdef B200AIS_HAVE_GNURADIO
#include <gnuradio/sync_block.h>
#endif

namespace gr {
namespace ais {

class AIS_API invert : virtual public gr::sync_block
{
public:
    typedef boost::shared_ptr<invert> sptr;
    static sptr make();
};

} // namespace ais
} // namespace gr

#endif
