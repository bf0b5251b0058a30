// Shim of <volk/volk.h> for corr_est_cc_impl.cc:113-116,122-123,191: aligned allocation and the
// magnitude-squared kernel [G] (the oracle's restatement).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <complex>
#include <cstdlib>

#include "ais_oracle.h"

typedef std::complex<float> lv_32fc_t;
static inline size_t volk_get_alignment(void) { return 64; }
static inline void *volk_malloc(size_t size, size_t alignment)
{
    void *p = 0;
    if (posix_memalign(&p, alignment, size ? size : alignment))
        return 0;
    return p;
}
static inline void volk_free(void *p) { free(p); }
static inline void volk_32fc_magnitude_squared_32f(float *out, const lv_32fc_t *in, unsigned int n)
{
    ao_mag_squared(reinterpret_cast<const float *>(in), (int)n, out);
}
