// gr::ais::msk_timing_recovery_cc, B200 build.  Same public surface as the reference block
// (reference include/ais/msk_timing_recovery_cc.h:46-70): make(sps, gain, limit, osps) and the
// gain / limit / sps accessors; complex in, complex (+ float error, float mu) out at osps
// samples per symbol; restarts its loop on time_est tags.
#ifndef B200AIS_GR_AIS_MSK_TIMING_RECOVERY_CC_H
#define B200AIS_GR_AIS_MSK_TIMING_RECOVERY_CC_H

#include <ais/api.h>
#ifdef B200AIS_HAVE_GNURADIO
#include <gnuradio/block.h>
#endif

namespace gr {
namespace ais {

class AIS_API msk_timing_recovery_cc : virtual public gr::block
{
public:
    typedef boost::shared_ptr<msk_timing_recovery_cc> sptr;

    static sptr make(float sps, float gain, float limit, int osps);

    virtual void set_gain(float gain) = 0;
    virtual float get_gain(void) = 0;
    virtual void set_limit(float limit) = 0;
    virtual float get_limit(void) = 0;
    virtual void set_sps(float sps) = 0;
    virtual float get_sps(void) = 0;
};

} // namespace ais
} // namespace gr

#endif
