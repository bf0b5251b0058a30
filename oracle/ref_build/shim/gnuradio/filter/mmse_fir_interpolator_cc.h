// Shim of <gnuradio/filter/mmse_fir_interpolator_cc.h> (msk_timing_recovery_cc_impl.cc:50,103,170):
// the 8-tap, 128-step MMSE interpolator [G] forwards to the oracle's restatement and throws
// what GNU Radio throws for an out-of-range mu.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <gnuradio/block.h>

#include "ais_oracle.h"

namespace gr {
namespace filter {
class mmse_fir_interpolator_cc
{
public:
    unsigned ntaps() const { return 8; }
    unsigned nsteps() const { return 128; }
    gr_complex interpolate(const gr_complex input[], float mu) const
    {
        float re, im;
        if (ao_mmse_interpolate(reinterpret_cast<const float *>(input), mu, &re, &im))
            throw std::runtime_error("mmse_fir_interpolator_cc: imu out of bounds.\n");
        return gr_complex(re, im);
    }
};
} // namespace filter
} // namespace gr
