// Shim of <gnuradio/math.h>: gr::fast_atan2f (corr_est_cc_impl.cc:247) and gr::branchless_clip
// (msk_timing_recovery_cc_impl.cc:180,182) are GNU Radio kernels [G]; they forward to the
// oracle's restatements.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <gnuradio/block.h>

#include "ais_oracle.h"

namespace gr {
static inline float fast_atan2f(float y, float x) { return ao_fast_atan2f(y, x); }
static inline float branchless_clip(float x, float clip) { return ao_branchless_clip(x, clip); }
} // namespace gr
