// gr::ais::corr_est_cc, B200 build.  Same public surface as the reference block
// (reference include/ais/corr_est_cc.h:85-107): make(symbols, sps, mark_delay, threshold),
// symbols(), set_symbols(); complex in, 1-2 complex out; stream tags corr_start, phase_est,
// time_est, corr_est.  The arithmetic runs in libb200ais.so (include/b200ais.h).
#ifndef B200AIS_GR_AIS_CORR_EST_CC_H
#define B200AIS_GR_AIS_CORR_EST_CC_H

#include <ais/api.h>
#ifdef B200AIS_HAVE_GNURADIO
#include <gnuradio/sync_block.h>
#endif
#include <vector>

namespace gr {
namespace ais {

class AIS_API corr_est_cc : virtual public sync_block
{
public:
    typedef boost::shared_ptr<corr_est_cc> sptr;

    // symbols: the sync word as it appears on the air; sps: samples per symbol;
    // mark_delay: items after corr_start at which phase_est/time_est/corr_est are placed;
    // threshold: fraction of the template's autocorrelation peak (squared), default 0.9.
    static sptr make(const std::vector<gr_complex> &symbols, float sps, unsigned int mark_delay,
                     float threshold = 0.9);

    virtual std::vector<gr_complex> symbols() const = 0;
    virtual void set_symbols(const std::vector<gr_complex> &symbols) = 0;
};

} // namespace ais
} // namespace gr

#endif
