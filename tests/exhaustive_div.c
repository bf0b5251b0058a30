/* Every float x in [0, pi_f]: the reciprocal-based quotient the CUDA fast path uses
 * (q0 = y*r, e = fma(-q0, pi, y), q = fma(e, r, q0), r = RN(1/pi_f), y = x * 2^31) equals the
 * IEEE division y / pi_f that gr::fxpt::float_to_fixed performs.  Prints the mismatch count. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void)
{
    const float PI = 3.14159265358979323846f;
    const float R = 1.0f / PI;
    uint32_t lim;
    memcpy(&lim, &PI, 4);
    long long bad = 0, total = 0;
#pragma omp parallel for reduction(+ : bad, total) schedule(static)
    for (uint32_t u = 0; u <= lim; u++) {
        float x;
        memcpy(&x, &u, 4);
        float y = x * 2147483648.0f;
        float q = y / PI;
        float q0 = y * R;
        float e = fmaf(-q0, PI, y);
        float q1 = fmaf(e, R, q0);
        total++;
        if (q1 != q)
            bad++;
    }
    printf("%lld %lld\n", total, bad);
    return bad != 0;
}
