/*
 * ais_oracle.c -- CPU restatement of the gr-ais IQ-demod hot path (see ais_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  Parity status: see the header (the [R] blocks are pinned to the
 * reference sources compiled under oracle/ref_build; the [G] kernels are restated).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -mfma -fopenmp (oracle/Makefile).
 * -mfma only makes fmaf() a single instruction; -ffp-contract=off guarantees no
 * multiply-add is fused unless written as fmaf().
 */
#include "ais_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ tables */

static const float k_mmse[129][8] = {
#include "tables/mmse_taps.inc"
};
static const float k_atan[257] = {
#include "tables/atan_table.inc"
};
static const float k_sine[1024][2] = {
#include "tables/sine_table.inc"
};

const float *ao_mmse_taps(void) { return &k_mmse[0][0]; }
const float *ao_atan_table(void) { return k_atan; }
const float *ao_sine_table(void) { return &k_sine[0][0]; }

/* ----------------------------------------------------------- scalar pieces */

/* gr::fast_atan2f [G]: octant fold + 256-step linear table. */
float ao_fast_atan2f(float y, float x)
{
    const float TAN_MAP_RES = 0.003921569f; /* 1/255 */
    float y_abs = fabsf(y), x_abs = fabsf(x), z, base_angle, angle;
    if (!((y_abs > 0.0f) || (x_abs > 0.0f)))
        return 0.0f;
    if (y_abs < x_abs)
        z = y_abs / x_abs;
    else
        z = x_abs / y_abs;
    if (z < TAN_MAP_RES) {
        base_angle = z;
    } else {
        float alpha = z * 256.0f - 0.5f;
        int index = (int)alpha;
        alpha -= (float)index;
        base_angle = k_atan[index];
        base_angle += (k_atan[index + 1] - k_atan[index]) * alpha;
    }
    if (x_abs > y_abs) {
        if (x >= 0.0f) {
            angle = (y >= 0.0f) ? base_angle : -base_angle;
        } else {
            angle = 3.14159265358979323846f;
            if (y >= 0.0f)
                angle -= base_angle;
            else
                angle = base_angle - angle;
        }
    } else {
        if (y >= 0.0f) {
            angle = 1.57079632679489661923f;
            if (x >= 0.0f)
                angle -= base_angle;
            else
                angle += base_angle;
        } else {
            angle = -1.57079632679489661923f;
            if (x >= 0.0f)
                angle += base_angle;
            else
                angle -= base_angle;
        }
    }
    return angle;
}

/* std::abs(gr_complex) -> hypotf.  glibc >= 2.35 evaluates hypotf as the
 * double-precision expression below (exact products, one rounded add, one
 * rounded sqrt, one narrowing); written out so the GPU can match it bit for bit.
 * tests/test_oracle_units.py checks it equals this libm's hypotf. */
float ao_hypotf(float re, float im)
{
    double a = (double)re, b = (double)im;
    return (float)sqrt(a * a + b * b);
}

/* gr::branchless_clip [G] */
float ao_branchless_clip(float x, float clip)
{
    float x1 = fabsf(x + clip);
    float x2 = fabsf(x - clip);
    x1 -= x2;
    return 0.5f * x1;
}

/* gr::fxpt::float_to_fixed [G] (3.8: fold into [-pi,pi] first). An out-of-range
 * float->int32 conversion takes the x86 "integer indefinite" value INT32_MIN. */
int32_t ao_float_to_fixed(float x)
{
    const float PI = 3.14159265358979323846f;
    const float TWO_PI = 2.0f * PI;
    const float TWO_TO_THE_31 = 2147483648.0f;
    int d = (int)floor((double)(x / TWO_PI) + 0.5);
    x -= (float)d * TWO_PI;
    float v = x * TWO_TO_THE_31 / PI;
    if (!(v > -2147483904.0f && v < 2147483648.0f))
        return INT32_MIN;
    return (int32_t)v;
}

/* gr::fxpt::sincos [G]: 10-bit segment index, slope applied to (ux >> 1). */
void ao_fxpt_sincos(int32_t angle, float *s, float *c)
{
    uint32_t ux = (uint32_t)angle;
    uint32_t idx = ux >> 22;
    *s = k_sine[idx][0] * (float)(ux >> 1) + k_sine[idx][1];
    ux = (uint32_t)angle + 0x40000000u;
    idx = ux >> 22;
    *c = k_sine[idx][0] * (float)(ux >> 1) + k_sine[idx][1];
}

/* feedforward_agc_cc envelope [G]: the 0.4 literal is a double. */
float ao_agc_envelope(float re, float im)
{
    float r_abs = fabsf(re), i_abs = fabsf(im);
    if (r_abs > i_abs)
        return (float)((double)r_abs + 0.4 * (double)i_abs);
    return (float)((double)i_abs + 0.4 * (double)r_abs);
}

/* canonical complex product a*b (VOLK multiply kernels, FMA form):
 * re = fma(ar, br, -(ai*bi)), im = fma(ar, bi, ai*br) */
static inline void cmul(float ar, float ai, float br, float bi, float *re, float *im)
{
    *re = fmaf(ar, br, -(ai * bi));
    *im = fmaf(ar, bi, ai * br);
}

/* --------------------------------------------------- A0 / G7: the template */

/* firdes::gaussian [G] */
void ao_firdes_gaussian(double gain, double spb, double bt, int ntaps, float *taps)
{
    double scale = 0;
    double dt = 1.0 / spb;
    double s = 1.0 / (sqrt(log(2.0)) / (2 * M_PI * bt));
    double t0 = -0.5 * ntaps;
    for (int i = 0; i < ntaps; i++) {
        t0++;
        double ts = s * dt * t0;
        taps[i] = (float)exp(-0.5 * ts * ts);
        scale += taps[i];
    }
    for (int i = 0; i < ntaps; i++)
        taps[i] = (float)(taps[i] / scale * gain);
}

/* gmsk_mod [G]: NRZ -> interp_fir(sps, gaussian (*) rect) -> FM(sens=(pi/2)/sps),
 * all filter and phase state starting at zero (the throw-away flowgraph that
 * modulate_vector_bc runs, python/ais_demod.py:36-38). */
int ao_gmsk_template_bits(const uint8_t *bits, int nbits, int sps, float bt, float *out_iq)
{
    int ng = 4 * sps, nt = ng + sps - 1;
    float *g = (float *)malloc(sizeof(float) * ng);
    double *conv = (double *)calloc(nt, sizeof(double));
    float *taps = (float *)malloc(sizeof(float) * nt);
    if (!g || !conv || !taps)
        return -1;
    ao_firdes_gaussian(1.0, (double)sps, (double)bt, ng, g);
    /* numpy.convolve(gaussian, ones(sps)) in float64, handed to the float FIR */
    for (int i = 0; i < ng; i++)
        for (int j = 0; j < sps; j++)
            conv[i + j] += (double)g[i];
    for (int i = 0; i < nt; i++)
        taps[i] = (float)conv[i];
    float sens = (float)((M_PI / 2) / sps);
    float phase = 0.0f;
    const float F_PI = (float)M_PI;
    for (int n = 0; n < nbits * sps; n++) {
        /* interpolating FIR: y[n] = sum_k taps[k] * u[n-k], u = zero-stuffed NRZ */
        float acc = 0.0f;
        for (int k = n % sps; k < nt; k += sps) {
            int sym = (n - k) / sps;
            if (n - k < 0 || sym >= nbits)
                continue;
            float v = bits[sym] ? 1.0f : -1.0f;
            acc += taps[k] * v;
        }
        phase = phase + sens * acc;
        phase = fmodf(phase + F_PI, 2.0f * F_PI) - F_PI;
        float oi, oq;
        ao_fxpt_sincos(ao_float_to_fixed(phase), &oq, &oi);
        out_iq[2 * n] = oi;
        out_iq[2 * n + 1] = oq;
    }
    free(g);
    free(conv);
    free(taps);
    return nbits * sps;
}

int ao_gmsk_template_packed(const uint8_t *bytes, int nbytes, int sps, float bt, float *out_iq)
{
    uint8_t *bits = (uint8_t *)malloc((size_t)nbytes * 8);
    if (!bits)
        return -1;
    for (int i = 0; i < nbytes; i++)
        for (int b = 0; b < 8; b++)
            bits[i * 8 + b] = (bytes[i] >> (7 - b)) & 1; /* GR_MSB_FIRST */
    int r = ao_gmsk_template_bits(bits, nbytes * 8, sps, bt, out_iq);
    free(bits);
    return r;
}

/* ------------------------------------------------- G1: square / FFT / shift */

/* blocks.multiply_cc(x, x) (python/gmsk_sync.py:22,30-31) */
void ao_square(const float *x, float *out, int n)
{
    for (int i = 0; i < n; i++)
        cmul(x[2 * i], x[2 * i + 1], x[2 * i], x[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
}

/* Forward DFT, radix-2 decimation in time.  Canonical graph: stage s combines
 * two size-m/2 DFTs with twiddle W[j*n/m], W[k] = (float)cos(2 pi k/n) - i (float)sin(2 pi k/n)
 * (k = 0 and k = n/4 exact), butterfly  t = W*b (cmul order), out = a +/- t. */
int ao_fft_forward(const float *in, float *out, int n)
{
    int lg = 0;
    while ((1 << lg) < n)
        lg++;
    if ((1 << lg) != n || n < 2)
        return -1;
    float *w = (float *)malloc(sizeof(float) * n); /* n/2 complex twiddles */
    if (!w)
        return -1;
    for (int k = 0; k < n / 2; k++) {
        double ang = 2.0 * M_PI * (double)k / (double)n;
        w[2 * k] = (float)cos(ang);
        w[2 * k + 1] = (float)(-sin(ang));
    }
    w[0] = 1.0f;
    w[1] = 0.0f;
    if (n >= 4) {
        w[2 * (n / 4)] = 0.0f;
        w[2 * (n / 4) + 1] = -1.0f;
    }
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < lg; b++)
            r |= ((i >> b) & 1) << (lg - 1 - b);
        out[2 * r] = in[2 * i];
        out[2 * r + 1] = in[2 * i + 1];
    }
    for (int m = 2; m <= n; m <<= 1) {
        int half = m / 2, step = n / m;
        for (int k = 0; k < n; k += m) {
            for (int j = 0; j < half; j++) {
                float wr = w[2 * (j * step)], wi = w[2 * (j * step) + 1];
                float *a = &out[2 * (k + j)], *b = &out[2 * (k + j + half)];
                float tr, ti;
                cmul(wr, wi, b[0], b[1], &tr, &ti);
                float ar = a[0], ai = a[1];
                a[0] = ar + tr;
                a[1] = ai + ti;
                b[0] = ar - tr;
                b[1] = ai - ti;
            }
        }
    }
    free(w);
    return 0;
}

/* fft_vcc(forward, shift=True) output order [G]: halves swapped, DC at n/2 */
void ao_fft_shift(const float *in, float *out, int n)
{
    int len = (n + 1) / 2;
    memcpy(out, in + 2 * len, sizeof(float) * 2 * (n - len));
    memcpy(out + 2 * (n - len), in, sizeof(float) * 2 * len);
}

/* ------------------------------------------------------------- A8: freqest */

/* lib/freqest_impl.cc:41-48 */
void ao_freqest_init(ao_freqest *f, float sample_rate, int data_rate, int fftlen)
{
    f->offset = (int)((float)fftlen * ((float)data_rate / sample_rate));
    f->binsize = sample_rate / (float)fftlen;
    f->fftlen = fftlen;
}

/* lib/freqest_impl.cc:57-88.  maxpos is a local of work(): 0 at the start of
 * every call, carried from vector to vector inside the call (only maxenergy is
 * reset per vector). */
int ao_freqest_work(const ao_freqest *f, const float *spec, int nvec, float *out, int *maxpos_out)
{
    unsigned fftlen = (unsigned)f->fftlen;
    float maxenergy = 0;
    unsigned maxpos = 0;
    for (int i = 0; i < nvec; i++) {
        const float *in = spec + (size_t)2 * i * fftlen;
        maxenergy = 0;
        for (unsigned j = 0; j < fftlen - (unsigned)f->offset; j++) {
            float cur = ao_hypotf(in[2 * j], in[2 * j + 1]) +
                        ao_hypotf(in[2 * (j + f->offset)], in[2 * (j + f->offset) + 1]);
            if (cur > maxenergy) {
                maxenergy = cur;
                maxpos = j + (unsigned)(f->offset / 2);
            }
        }
        out[i] = ((float)maxpos - (float)(fftlen / 2)) * f->binsize / 2.0f;
        if (maxpos_out)
            maxpos_out[i] = (int)maxpos;
    }
    return nvec;
}

/* ------------------------------- G2: repeat + frequency_modulator_fc + mix */

void ao_nco_mix(float *phase, float sensitivity, const float *freq, int rep, const float *x, int n,
                float *out)
{
    const float F_PI = (float)M_PI;
    float ph = *phase;
    for (int i = 0; i < n; i++) {
        ph = ph + sensitivity * freq[i / rep];
        ph = fmodf(ph + F_PI, 2.0f * F_PI) - F_PI;
        float oi, oq;
        ao_fxpt_sincos(ao_float_to_fixed(ph), &oq, &oi);
        cmul(x[2 * i], x[2 * i + 1], oi, oq, &out[2 * i], &out[2 * i + 1]);
    }
    *phase = ph;
}

/* ------------------------------------------------ G3: feedforward_agc_cc */

void ao_agc_work(const float *in, int n, int nsamples, float reference, float *out)
{
    /* GNU Radio recomputes envelope(in[i+j]) inside the window loop; the values are the same,
     * so they are taken once per item here (this only makes the CPU arm of the benchmark faster) */
    int total = n + nsamples - 1;
    float *env = (float *)malloc(sizeof(float) * (size_t)(total > 0 ? total : 1));
    for (int i = 0; i < total; i++)
        env[i] = ao_agc_envelope(in[2 * i], in[2 * i + 1]);
    for (int i = 0; i < n; i++) {
        float max_env = 1e-4f;
        for (int j = 0; j < nsamples; j++) {
            float e = env[i + j];
            max_env = e > max_env ? e : max_env; /* std::max(max_env, e) */
        }
        float gain = reference / max_env;
        out[2 * i] = gain * in[2 * i];
        out[2 * i + 1] = gain * in[2 * i + 1];
    }
    free(env);
}

/* ----------------------------------------------------- A1-A4: corr_est_cc */

static int fft_filter_fftsize(int ntaps)
{
    /* kernel::fft_filter_ccc::compute_sizes [G]: fftsize = 2 * 2^ceil(log2 ntaps) */
    int p = 1;
    while (p < ntaps)
        p <<= 1;
    return 2 * p;
}

static void make_twiddles(int n, float *w)
{
    for (int k = 0; k < n / 2; k++) {
        double ang = 2.0 * M_PI * (double)k / (double)n;
        w[2 * k] = (float)cos(ang);
        w[2 * k + 1] = (float)(-sin(ang));
    }
    w[0] = 1.0f;
    w[1] = 0.0f;
    if (n >= 4) {
        w[2 * (n / 4)] = 0.0f;
        w[2 * (n / 4) + 1] = -1.0f;
    }
}

/* Forward DFT, radix-2 decimation in frequency: natural-order input, bit-reversed output.
 * Stage m: a' = a + b, b' = W_m^j * (a - b) (cmul order), m = n, n/2, ..., 2. */
static void fft_dif(float *x, int n, const float *w)
{
    for (int m = n; m >= 2; m >>= 1) {
        int half = m / 2, step = n / m;
        for (int k = 0; k < n; k += m)
            for (int j = 0; j < half; j++) {
                float *a = &x[2 * (k + j)], *b = &x[2 * (k + j + half)];
                float ar = a[0], ai = a[1], br = b[0], bi = b[1];
                a[0] = ar + br;
                a[1] = ai + bi;
                cmul(w[2 * (j * step)], w[2 * (j * step) + 1], ar - br, ai - bi, &b[0], &b[1]);
            }
    }
}

/* Inverse DFT (unnormalised), radix-2 decimation in time with conjugated twiddles:
 * bit-reversed input, natural-order output.  Stage m: t = conj(W_m^j) * b, a' = a + t, b' = a - t. */
static void ifft_dit(float *x, int n, const float *w)
{
    for (int m = 2; m <= n; m <<= 1) {
        int half = m / 2, step = n / m;
        for (int k = 0; k < n; k += m)
            for (int j = 0; j < half; j++) {
                float *a = &x[2 * (k + j)], *b = &x[2 * (k + j + half)];
                float tr, ti;
                cmul(w[2 * (j * step)], -w[2 * (j * step) + 1], b[0], b[1], &tr, &ti);
                float ar = a[0], ai = a[1];
                a[0] = ar + tr;
                a[1] = ai + ti;
                b[0] = ar - tr;
                b[1] = ai - ti;
            }
    }
}

int ao_fft_dif_inplace(float *x, int n)
{
    if (n < 2 || (n & (n - 1)))
        return -1;
    float *w = (float *)malloc(sizeof(float) * n);
    make_twiddles(n, w);
    fft_dif(x, n, w);
    free(w);
    return 0;
}

int ao_ifft_dit_inplace(float *x, int n)
{
    if (n < 2 || (n & (n - 1)))
        return -1;
    float *w = (float *)malloc(sizeof(float) * n);
    make_twiddles(n, w);
    ifft_dit(x, n, w);
    free(w);
    return 0;
}

/* kernel::fft_filter_ccc [G].  set_taps: taps scaled by 1/fftsize, zero padded, transformed
 * once; the tail keeps its contents and is resized to ntaps-1 (std::vector::resize). */
void ao_fftfilt_init(ao_fftfilt *f) { memset(f, 0, sizeof(*f)); }

int ao_fftfilt_set_taps(ao_fftfilt *f, const float *taps, int ntaps)
{
    int old = f->ntaps;
    int F = fft_filter_fftsize(ntaps);
    f->ntaps = ntaps;
    f->fftsize = F;
    f->nsamples = F - ntaps + 1;
    free(f->H);
    free(f->tw);
    f->H = (float *)calloc((size_t)F * 2, sizeof(float));
    f->tw = (float *)malloc(sizeof(float) * F);
    make_twiddles(F, f->tw);
    float scale = 1.0f / (float)F;
    for (int k = 0; k < ntaps; k++) {
        f->H[2 * k] = taps[2 * k] * scale;
        f->H[2 * k + 1] = taps[2 * k + 1] * scale;
    }
    fft_dif(f->H, F, f->tw);
    float *nt = (float *)calloc((size_t)(ntaps > 1 ? ntaps - 1 : 1) * 2, sizeof(float));
    if (f->tail) {
        int keep = (old < ntaps ? old : ntaps) - 1;
        if (keep > 0)
            memcpy(nt, f->tail, sizeof(float) * 2 * (size_t)keep);
        free(f->tail);
    }
    f->tail = nt;
    return f->nsamples;
}

/* filter(): blocks of nsamples items, zero padded to fftsize, forward FFT, multiplied by the
 * transformed taps (volk multiply order), inverse FFT, the first ntaps-1 outputs get the
 * previous block's tail added, the last ntaps-1 become the new tail.  Canonical transforms:
 * the radix-2 DIF / DIT pair above. */
int ao_fftfilt_filter(ao_fftfilt *f, int nitems, const float *x, float *out)
{
    int L = f->ntaps, F = f->fftsize, ns = f->nsamples;
    float *buf = (float *)malloc(sizeof(float) * 2 * (size_t)F);
    for (int i0 = 0; i0 < nitems; i0 += ns) {
        memcpy(buf, x + 2 * (size_t)i0, sizeof(float) * 2 * (size_t)ns);
        memset(buf + 2 * (size_t)ns, 0, sizeof(float) * 2 * (size_t)(F - ns));
        fft_dif(buf, F, f->tw);
        for (int p = 0; p < F; p++)
            cmul(buf[2 * p], buf[2 * p + 1], f->H[2 * p], f->H[2 * p + 1], &buf[2 * p], &buf[2 * p + 1]);
        ifft_dit(buf, F, f->tw);
        for (int j = 0; j < L - 1; j++) {
            buf[2 * j] += f->tail[2 * j];
            buf[2 * j + 1] += f->tail[2 * j + 1];
        }
        memcpy(out + 2 * (size_t)i0, buf, sizeof(float) * 2 * (size_t)ns);
        memcpy(f->tail, buf + 2 * (size_t)ns, sizeof(float) * 2 * (size_t)(L - 1));
    }
    free(buf);
    return nitems;
}

void ao_fftfilt_free(ao_fftfilt *f)
{
    free(f->H);
    free(f->tail);
    free(f->tw);
    memset(f, 0, sizeof(*f));
}

/* volk_32fc_magnitude_squared_32f [G] (corr_est_cc_impl.cc:191) */
void ao_mag_squared(const float *in, int n, float *out)
{
    for (int i = 0; i < n; i++) {
        float re = in[2 * i], im = in[2 * i + 1];
        out[i] = re * re + im * im;
    }
}

/* lib/corr_est_cc_impl.cc:48-117 */
int ao_corr_est_init(ao_corr_est *c, const float *symbols, int L, float sps, unsigned mark_delay,
                     float threshold)
{
    memset(c, 0, sizeof(*c));
    c->taps = (float *)malloc(sizeof(float) * 2 * (size_t)L);
    if (!c->taps)
        return -1;
    c->L = L;
    c->sps = sps;
    for (int i = 0; i < L; i++) { /* conj then reverse (:59-63) */
        c->taps[2 * (L - 1 - i)] = symbols[2 * i];
        c->taps[2 * (L - 1 - i) + 1] = -symbols[2 * i + 1];
    }
    c->mark_delay = mark_delay >= (unsigned)L ? (unsigned)L - 1 : mark_delay; /* :65-66 */
    float corr = 0; /* :71-74: abs(z*conj(z)) = re*re + im*im */
    for (int i = 0; i < L; i++) {
        float re = c->taps[2 * i], im = c->taps[2 * i + 1];
        corr += re * re + im * im;
    }
    c->thresh = threshold * corr * corr;
    ao_fftfilt_init(&c->f);
    ao_fftfilt_set_taps(&c->f, c->taps, L); /* :77 (the kernel's ctor) and :84 */
    return 0;
}

/* lib/corr_est_cc_impl.cc:132-162: the taps are replaced verbatim (no conj, no
 * reverse) and d_thresh is NOT recomputed. */
void ao_corr_est_set_symbols(ao_corr_est *c, const float *symbols, int L)
{
    free(c->taps);
    c->taps = (float *)malloc(sizeof(float) * 2 * (size_t)L);
    memcpy(c->taps, symbols, sizeof(float) * 2 * (size_t)L);
    c->L = L;
    ao_fftfilt_set_taps(&c->f, c->taps, L);
    c->mark_delay = c->mark_delay >= (unsigned)L ? (unsigned)L - 1 : c->mark_delay;
}

void ao_corr_est_free(ao_corr_est *c)
{
    free(c->taps);
    ao_fftfilt_free(&c->f);
    c->taps = 0;
}

/* float64 direct-form truth of the same filter on a fresh block: like fft_filter_ccc, it is
 * fed &in[hist_len] only, so the items before it (the block's tagging history) do not enter
 * the correlation of the first call; the filter's own tail, zero at construction, does. */
void ao_corr_direct_f64(const ao_corr_est *c, int n, const float *in, double *corr_out)
{
    const float *x = in + 2 * (size_t)c->L;
    for (int i = 0; i < n; i++) {
        double re = 0, im = 0;
        for (int k = 0; k < c->L; k++) {
            if (i - k < 0)
                break;
            double tr = c->taps[2 * k], ti = c->taps[2 * k + 1];
            double xr = x[2 * (i - k)], xi = x[2 * (i - k) + 1];
            re += tr * xr - ti * xi;
            im += tr * xi + ti * xr;
        }
        corr_out[2 * i] = re;
        corr_out[2 * i + 1] = im;
    }
}

static void push_tag(ao_tag *tags, int max_tags, int *ntags, uint64_t off, int key, int port,
                     double v)
{
    if (*ntags < max_tags) {
        tags[*ntags].offset = off;
        tags[*ntags].key = key;
        tags[*ntags].port = port;
        tags[*ntags].value = v;
    }
    (*ntags)++;
}

/* lib/corr_est_cc_impl.cc:164-279.  The correlation filter (:188) is GNU Radio's
 * kernel::fft_filter_ccc [G] (ao_fftfilt_filter above). */
int ao_corr_est_work(ao_corr_est *c, int n, const float *in, uint64_t nitems_written,
                     float *out0, float *corr, float *mag, int two_ports, ao_tag *tags,
                     int max_tags, int *ntags)
{
    int L = c->L, ns = c->f.nsamples;
    *ntags = 0;
    if (n % ns)
        return -1; /* set_output_multiple(nsamples) */
    float *corr_buf = corr ? corr : (float *)malloc(sizeof(float) * 2 * (size_t)(n > 0 ? n : 1));
    float *mag_buf = mag ? mag : (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    if (out0)
        memcpy(out0, in, sizeof(float) * 2 * (size_t)n);           /* :184 */
    ao_fftfilt_filter(&c->f, n, in + 2 * (size_t)L, corr_buf);      /* &in[hist_len] (:188) */
    ao_mag_squared(corr_buf, n, mag_buf);                          /* :191 */
    int isps = (int)(c->sps + 0.5f);
    int i = 0;
    while (i < n) {
        if (mag_buf[i] <= c->thresh) {
            i++;
            continue;
        }
        while ((i < (n - 1)) && (mag_buf[i] < mag_buf[i + 1]))
            i++;
        push_tag(tags, max_tags, ntags, nitems_written + i, AO_TAG_CORR_START, 0,
                 (double)mag_buf[i]);
        double center = 0.0;
        if (i > 0 && i < (n - 1)) {
            double nom = 0, den = 0;
            for (int s = 0; s < 3; s++) {
                nom += (float)(s + 1) * mag_buf[i + s - 1]; /* int*float product is a float */
                den += mag_buf[i + s - 1];
            }
            center = nom / den - 2.0;
        }
        float phase = ao_fast_atan2f(corr_buf[2 * i + 1], corr_buf[2 * i]);
        int index = i + (int)c->mark_delay;
        push_tag(tags, max_tags, ntags, nitems_written + index, AO_TAG_PHASE_EST, 0, (double)phase);
        push_tag(tags, max_tags, ntags, nitems_written + index, AO_TAG_TIME_EST, 0, center);
        push_tag(tags, max_tags, ntags, nitems_written + index, AO_TAG_CORR_EST, 0,
                 (double)mag_buf[i]);
        if (two_ports) {
            push_tag(tags, max_tags, ntags, nitems_written + i, AO_TAG_PHASE_EST, 1, (double)phase);
            push_tag(tags, max_tags, ntags, nitems_written + i, AO_TAG_TIME_EST, 1, center);
            push_tag(tags, max_tags, ntags, nitems_written + i, AO_TAG_CORR_EST, 1,
                     (double)mag_buf[i]);
        }
        i += isps;
    }
    if (!corr)
        free(corr_buf);
    if (!mag)
        free(mag_buf);
    return n;
}

/* ------------------------------------------ A5-A7: msk_timing_recovery_cc */

/* lib/msk_timing_recovery_cc_impl.cc:45-96 */
int ao_msk_init(ao_msk *m, float sps, float gain, float limit, int osps)
{
    memset(m, 0, sizeof(*m));
    m->limit = limit;
    m->mu = 0.5f;
    m->div = 0;
    m->osps = osps;
    m->sps = (float)((double)sps / 2.0);
    m->omega = m->sps;
    m->gain = gain;
    if (gain <= 0)
        return -1; /* std::out_of_range("Gain must be positive") */
    m->gain_omega = (float)((double)(gain * gain) * 0.25);
    if (osps != 1 && osps != 2)
        return -2; /* std::out_of_range("osps must be 1 or 2") */
    return 0;
}

/* :98-105 */
int ao_msk_forecast(const ao_msk *m, int noutput_items)
{
    return (int)ceil(((double)((float)noutput_items * m->sps * 2.0f)) + 3.0 * (double)m->sps + 8.0);
}

/* mmse_fir_interpolator_cc::interpolate [G]: imu = rint(mu*128); 8-tap dot
 * product with the reversed table row.  Canonical summation: four two-term fmaf
 * partial sums p_j = in[j]*t[j] + in[j+4]*t[j+4], combined (p0+p1)+(p2+p3). */
int ao_mmse_interpolate(const float *s /* 8 complex */, float mu, float *vr, float *vi)
{
    int imu = (int)rintf(mu * 128.0f);
    if (imu < 0 || imu > 128)
        return -1;
    const float *row = k_mmse[imu];
    float pr[4], pi[4];
    for (int j = 0; j < 4; j++) {
        float t0 = row[7 - j], t1 = row[3 - j];
        pr[j] = fmaf(s[2 * (j + 4)], t1, s[2 * j] * t0);
        pi[j] = fmaf(s[2 * (j + 4) + 1], t1, s[2 * j + 1] * t0);
    }
    *vr = (pr[0] + pr[1]) + (pr[2] + pr[3]);
    *vi = (pi[0] + pi[1]) + (pi[2] + pi[3]);
    return 0;
}

/* :107-206 */
int ao_msk_general_work(ao_msk *m, int noutput_items, int ninput_items, const float *in,
                        uint64_t nitems_read, const ao_tag *tags, int ntags, float *out,
                        float *out_err, float *out_mu, int *consumed)
{
    int oidx = 0, iidx = 0;
    int ninp = (int)((double)ninput_items - 3.0 * (double)m->sps); /* :119 */
    if (ninp <= 0) {
        *consumed = 0;
        return 0;
    }
    /* get_tags_in_range(time_est) over [read, read+ninp), kept in offset order (:125-130) */
    int nt = 0;
    const ao_tag **tv = (const ao_tag **)malloc(sizeof(*tv) * (size_t)(ntags > 0 ? ntags : 1));
    for (int t = 0; t < ntags; t++)
        if (tags[t].key == AO_TAG_TIME_EST && tags[t].port == 0 && tags[t].offset >= nitems_read &&
            tags[t].offset < nitems_read + (uint64_t)ninp)
            tv[nt++] = &tags[t];
    for (int a = 1; a < nt; a++) { /* stable insertion sort by offset */
        const ao_tag *key = tv[a];
        int b = a - 1;
        while (b >= 0 && tv[b]->offset > key->offset) {
            tv[b + 1] = tv[b];
            b--;
        }
        tv[b + 1] = key;
    }
    int thead = 0;
    float err_out = 0;
    float s8[16];
    while (oidx < noutput_items && iidx < ninp) {
        if (thead < nt) {
            int offset = (int)(tv[thead]->offset - nitems_read);
            if ((offset >= iidx) && ((float)offset < ((float)iidx + m->sps))) {
                float center = (float)tv[thead]->value;
                if (center != center) {
                    thead++; /* NaN: drop the tag, skip the reset (:144-147) */
                } else {
                    m->mu = center;
                    iidx = offset;
                    if (m->mu < 0) {
                        m->mu = m->mu + 1.0f;
                        iidx--;
                    }
                    m->div = 0;
                    m->omega = m->sps;
                    m->dly2_re = m->dly1_re;
                    m->dly2_im = m->dly1_im;
                    thead++;
                }
            }
        }
        for (int k = 0; k < 8; k++) {
            int idx = iidx + k;
            if (idx >= 0) {
                s8[2 * k] = in[2 * idx];
                s8[2 * k + 1] = in[2 * idx + 1];
            } else { /* in[-1]: the item before the read pointer */
                s8[2 * k] = (idx == -1) ? m->prev_re : 0.0f;
                s8[2 * k + 1] = (idx == -1) ? m->prev_im : 0.0f;
            }
        }
        float vr, vi;
        if (ao_mmse_interpolate(s8, m->mu, &vr, &vi) != 0) {
            free(tv);
            *consumed = iidx;
            return -1; /* mmse interpolator would throw */
        }
        /* std::complex products, GCC order: (ac - bd, ad + bc), no contraction */
        float sq_re = vr * vr - vi * vi, sq_im = vr * vi + vi * vr;
        float d_re = m->dly2_re * m->dly2_re - m->dly2_im * m->dly2_im;
        float d_im = -(m->dly2_re * m->dly2_im + m->dly2_im * m->dly2_re);
        float nl_re = sq_re * d_re - sq_im * d_im;
        float nl_im = sq_re * d_im + sq_im * d_re;
        err_out = nl_re - m->diff1_re;
        if (m->div % 2) {
            err_out = ao_branchless_clip(err_out, 3.0f);
            m->omega = m->omega + m->gain_omega * err_out;
            m->omega = m->sps + ao_branchless_clip(m->omega - m->sps, m->limit);
            m->mu = m->mu + m->gain * err_out;
        }
        if (!(m->div % 2) || m->osps == 2) {
            out[2 * oidx] = vr;
            out[2 * oidx + 1] = vi;
            if (out_err)
                out_err[oidx] = err_out;
            if (out_mu)
                out_mu[oidx] = m->mu;
            oidx++;
        }
        m->div++;
        m->dly1_re = vr;
        m->dly1_im = vi;
        m->dly2_re = vr;
        m->dly2_im = vi;
        m->diff1_re = nl_re;
        m->diff1_im = nl_im;
        m->mu = m->mu + m->omega;
        float fl = floorf(m->mu);
        iidx += (int)fl;
        m->mu = m->mu - fl;
    }
    free(tv);
    if (iidx > 0) {
        m->prev_re = in[2 * (iidx - 1)];
        m->prev_im = in[2 * (iidx - 1) + 1];
    }
    *consumed = iidx;
    return oidx;
}

/* ------------------------------------------------- G4-G6, A9: the bit tail */

/* quadrature_demod_cf [G]: y = gain * fast_atan2f(Im, Re) of x[n]*conj(x[n-1]),
 * product in the VOLK multiply-conjugate FMA form */
void ao_quad_demod(float *prev, const float *in, int n, float gain, float *out)
{
    float pr = prev[0], pi = prev[1];
    for (int i = 0; i < n; i++) {
        float ar = in[2 * i], ai = in[2 * i + 1];
        float re = fmaf(ar, pr, ai * pi);
        float im = fmaf(ai, pr, -(ar * pi));
        out[i] = gain * ao_fast_atan2f(im, re);
        pr = ar;
        pi = ai;
    }
    prev[0] = pr;
    prev[1] = pi;
}

void ao_binary_slicer(const float *in, int n, uint8_t *out)
{
    for (int i = 0; i < n; i++)
        out[i] = in[i] >= 0 ? 1 : 0;
}

/* diff_decoder_bb [G]: (in[i] - in[i-1]) % modulus in unsigned arithmetic */
void ao_diff_decoder(uint8_t *prev, const uint8_t *in, int n, unsigned modulus, uint8_t *out)
{
    uint8_t p = *prev;
    for (int i = 0; i < n; i++) {
        out[i] = (uint8_t)(((unsigned)in[i] - (unsigned)p) % modulus);
        p = in[i];
    }
    *prev = p;
}

/* lib/invert_impl.cc:54-68 */
void ao_invert(const uint8_t *in, int n, uint8_t *out)
{
    for (int i = 0; i < n; i++)
        out[i] = (in[i] ^ 0x01) & 0x01;
}

/* ------------------------------------------- block provider: the oracle's own */

static void *ob_corr_new(const float *symbols, int L, float sps, unsigned mark_delay, float threshold)
{
    ao_corr_est *c = (ao_corr_est *)malloc(sizeof(*c));
    if (c && ao_corr_est_init(c, symbols, L, sps, mark_delay, threshold)) {
        free(c);
        c = 0;
    }
    return c;
}
static void ob_corr_delete(void *h)
{
    if (h) {
        ao_corr_est_free((ao_corr_est *)h);
        free(h);
    }
}
static int ob_corr_output_multiple(void *h) { return ((ao_corr_est *)h)->f.nsamples; }
static int ob_corr_set_symbols(void *h, const float *symbols, int L)
{
    ao_corr_est_set_symbols((ao_corr_est *)h, symbols, L);
    return 0;
}
static int ob_corr_work(void *h, int n, const float *in, uint64_t nitems_written, float *out0,
                        float *corr, float *mag, int two_ports, ao_tag *tags, int max_tags, int *ntags)
{
    return ao_corr_est_work((ao_corr_est *)h, n, in, nitems_written, out0, corr, mag, two_ports, tags,
                            max_tags, ntags);
}
static void *ob_msk_new(float sps, float gain, float limit, int osps, int *status)
{
    ao_msk *m = (ao_msk *)malloc(sizeof(*m));
    int rc = m ? ao_msk_init(m, sps, gain, limit, osps) : -1;
    if (status)
        *status = rc;
    if (rc) {
        free(m);
        m = 0;
    }
    return m;
}
static void ob_free(void *h) { free(h); }
static float ob_msk_get_sps(void *h) { return ((ao_msk *)h)->sps; }
static int ob_msk_general_work(void *h, int noutput_items, int ninput_items, const float *in,
                               uint64_t nitems_read, const ao_tag *tags, int ntags, float *out,
                               float *out_err, float *out_mu, int *consumed)
{
    return ao_msk_general_work((ao_msk *)h, noutput_items, ninput_items, in, nitems_read, tags, ntags,
                               out, out_err, out_mu, consumed);
}
static void *ob_freqest_new(float sample_rate, int data_rate, int fftlen)
{
    ao_freqest *f = (ao_freqest *)malloc(sizeof(*f));
    if (f)
        ao_freqest_init(f, sample_rate, data_rate, fftlen);
    return f;
}
static int ob_freqest_work(void *h, const float *spec, int nvec, float *out)
{
    return ao_freqest_work((const ao_freqest *)h, spec, nvec, out, 0);
}

const ao_blocks *ao_blocks_oracle(void)
{
    static const ao_blocks b = { "oracle",
                                 ob_corr_new,
                                 ob_corr_delete,
                                 ob_corr_output_multiple,
                                 ob_corr_set_symbols,
                                 ob_corr_work,
                                 ob_msk_new,
                                 ob_free,
                                 ob_msk_get_sps,
                                 ob_msk_general_work,
                                 ob_freqest_new,
                                 ob_free,
                                 ob_freqest_work,
                                 ao_invert };
    return &b;
}

/* ----------------------------------------------------------- the chain */

int ao_default_corr_chunk(int L)
{
    int ns = fft_filter_fftsize(L) - L + 1;
    return (24576 / ns) * ns; /* set_output_multiple(nsamples) under set_max_noutput_items(24576) */
}

/* square -> stream_to_vector -> fft_vcc(shift) -> freqest -> repeat -> FM -> mix over nvec
 * whole vectors (python/gmsk_sync.py:22-37); freqest sees them in ONE work() call */
static void freq_sync(const ao_blocks *blk, void *fe, const ao_chain_cfg *cfg, const float *x, int n1,
                      float *phase, float *mixed, float *fhat_out)
{
    int fftlen = cfg->fftlen, nvec = n1 / fftlen;
    float *sq = (float *)malloc(sizeof(float) * 2 * (size_t)fftlen);
    float *sp = (float *)malloc(sizeof(float) * 2 * (size_t)fftlen);
    float *spec = (float *)malloc(sizeof(float) * 2 * (size_t)fftlen * (size_t)(nvec + 1));
    float *fh = (float *)malloc(sizeof(float) * (size_t)(nvec + 1));
    for (int b = 0; b < nvec; b++) {
        ao_square(x + 2 * (size_t)b * fftlen, sq, fftlen);
        ao_fft_forward(sq, sp, fftlen);
        ao_fft_shift(sp, spec + 2 * (size_t)b * fftlen, fftlen);
    }
    blk->freqest_work(fe, spec, nvec, fh);
    if (fhat_out)
        memcpy(fhat_out, fh, sizeof(float) * (size_t)nvec);
    float sens = (float)(-1.0 / ((double)cfg->sample_rate / (2 * M_PI))); /* gmsk_sync.py:27 */
    ao_nco_mix(phase, sens, fh, fftlen, x, n1, mixed);
    free(sq);
    free(sp);
    free(spec);
    free(fh);
}

/* python/ais_demod.py:34-56 + python/gmsk_sync.py:22-37 over one record, every
 * block freshly constructed.  Canonical scheduling: freqest sees all vectors in
 * one work() call; corr_est runs in work chunks of cfg->corr_chunk items (multiple
 * of nsamples; the remainder shorter than nsamples is left unprocessed, as the
 * scheduler would); msk_timing_recovery sees corr_est's whole output in one
 * general_work() call with unbounded noutput_items. */
int ao_demod_chain_with(const ao_blocks *blk, const ao_chain_cfg *cfg, const float *symbols, int L,
                        const float *x, int n, ao_chain_out *o)
{
    if (!blk)
        blk = ao_blocks_oracle();
    int rc = 0;
    int fftlen = cfg->fftlen;
    int n1 = (cfg->stages & AO_STAGE_FREQSYNC) ? (n / fftlen) * fftlen : n;
    float *mixed = (float *)malloc(sizeof(float) * 2 * (size_t)(n1 + 1));
    int hist = (cfg->stages & AO_STAGE_AGC) ? cfg->agc_nsamples - 1 : 0;
    float *agc_in = (float *)calloc((size_t)(n1 + hist + 1) * 2, sizeof(float));
    float *X = (float *)calloc((size_t)(n1 + L + 1) * 2, sizeof(float)); /* L zeros of history first */
    float *agc = X + 2 * (size_t)L;
    if (!mixed || !agc_in || !X)
        return -1;

    if (cfg->stages & AO_STAGE_FREQSYNC) {
        void *fe = blk->freqest_new(cfg->sample_rate, cfg->data_rate, fftlen);
        float phase = 0.0f;
        freq_sync(blk, fe, cfg, x, n1, &phase, mixed, o->fhat);
        blk->freqest_delete(fe);
    } else {
        memcpy(mixed, x, sizeof(float) * 2 * (size_t)n1);
    }
    if (o->mixed)
        memcpy(o->mixed, mixed, sizeof(float) * 2 * (size_t)n1);

    if (cfg->stages & AO_STAGE_AGC) {
        memcpy(agc_in + 2 * (size_t)hist, mixed, sizeof(float) * 2 * (size_t)n1);
        ao_agc_work(agc_in, n1, cfg->agc_nsamples, cfg->agc_reference, agc);
    } else {
        memcpy(agc, mixed, sizeof(float) * 2 * (size_t)n1);
    }
    if (o->agc)
        memcpy(o->agc, agc, sizeof(float) * 2 * (size_t)n1);

    void *ce = blk->corr_new(symbols, L, cfg->sps, cfg->mark_delay, cfg->threshold);
    int ns = blk->corr_output_multiple(ce);
    int chunk = cfg->corr_chunk > 0 ? cfg->corr_chunk : ao_default_corr_chunk(L);
    chunk = (chunk / ns) * ns;
    if (chunk <= 0)
        chunk = ns;
    float *out0 = (float *)malloc(sizeof(float) * 2 * (size_t)(n1 + 1));
    int n2 = 0;
    o->ntags = 0;
    while (n1 - n2 >= ns) {
        int nn = n1 - n2;
        nn = nn > chunk ? chunk : (nn / ns) * ns;
        int nt = 0;
        blk->corr_work(ce, nn, X + 2 * (size_t)n2, (uint64_t)n2, out0 + 2 * (size_t)n2,
                       o->corr ? o->corr + 2 * (size_t)n2 : 0, o->mag ? o->mag + n2 : 0, 0,
                       o->tags + (o->ntags < o->max_tags ? o->ntags : o->max_tags),
                       o->max_tags - (o->ntags < o->max_tags ? o->ntags : o->max_tags), &nt);
        o->ntags += nt;
        n2 += nn;
    }
    blk->corr_delete(ce);
    if (o->ntags > o->max_tags)
        rc = -3; /* tag buffer too small */

    int mrc = 0;
    void *mk = blk->msk_new(cfg->sps, cfg->gain, cfg->limit, cfg->osps, &mrc);
    if (mrc)
        rc = mrc;
    int maxsym = o->max_bits;
    float *sym = (float *)malloc(sizeof(float) * 2 * (size_t)(maxsym + 1));
    float *err = (float *)malloc(sizeof(float) * (size_t)(maxsym + 1));
    float *mu = (float *)malloc(sizeof(float) * (size_t)(maxsym + 1));
    float *soft = (float *)malloc(sizeof(float) * (size_t)(maxsym + 1));
    uint8_t *b0 = (uint8_t *)malloc((size_t)maxsym + 1), *b1 = (uint8_t *)malloc((size_t)maxsym + 1);
    int consumed = 0, k = 0;
    if (!mrc) {
        int usable = o->ntags < o->max_tags ? o->ntags : o->max_tags;
        k = blk->msk_general_work(mk, maxsym, n2, out0, 0, o->tags, usable, sym, err, mu, &consumed);
        if (k < 0) {
            rc = -4;
            k = 0;
        }
        blk->msk_delete(mk);
    }
    float prev[2] = { 0, 0 };
    uint8_t dprev = 0;
    ao_quad_demod(prev, sym, k, (float)(M_PI / 2), soft);
    ao_binary_slicer(soft, k, b0);
    ao_diff_decoder(&dprev, b0, k, 2, b1);
    blk->invert_work(b1, k, o->bits);
    o->nbits = k;
    o->n1 = n1;
    o->n2 = n2;
    o->consumed = consumed;
    if (o->sym)
        memcpy(o->sym, sym, sizeof(float) * 2 * (size_t)k);
    if (o->err)
        memcpy(o->err, err, sizeof(float) * (size_t)k);
    if (o->mu)
        memcpy(o->mu, mu, sizeof(float) * (size_t)k);
    if (o->soft)
        memcpy(o->soft, soft, sizeof(float) * (size_t)k);
    free(mixed);
    free(agc_in);
    free(X);
    free(out0);
    free(sym);
    free(err);
    free(mu);
    free(soft);
    free(b0);
    free(b1);
    return rc;
}

int ao_demod_chain(const ao_chain_cfg *cfg, const float *symbols, int L, const float *x, int n,
                   ao_chain_out *o)
{
    return ao_demod_chain_with(0, cfg, symbols, L, x, n, o);
}

int ao_demod_chain_batch_with(const ao_blocks *blk, const ao_chain_cfg *cfg, const float *symbols,
                              int L, const float *x, int channels, int n, uint8_t *bits,
                              int max_bits, int *nbits, ao_tag *tags, int max_tags, int *ntags,
                              int nthreads)
{
    int status = 0;
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < channels; c++) {
        ao_chain_out o;
        memset(&o, 0, sizeof(o));
        o.bits = bits + (size_t)c * max_bits;
        o.max_bits = max_bits;
        o.tags = tags + (size_t)c * max_tags;
        o.max_tags = max_tags;
        int rc = ao_demod_chain_with(blk, cfg, symbols, L, x + 2 * (size_t)c * n, n, &o);
        nbits[c] = o.nbits;
        ntags[c] = o.ntags;
        if (rc) {
#pragma omp critical
            if (!status)
                status = rc;
        }
    }
    return status;
}

int ao_demod_chain_batch(const ao_chain_cfg *cfg, const float *symbols, int L, const float *x,
                         int channels, int n, uint8_t *bits, int max_bits, int *nbits,
                         ao_tag *tags, int max_tags, int *ntags, int nthreads)
{
    return ao_demod_chain_batch_with(0, cfg, symbols, L, x, channels, n, bits, max_bits, nbits, tags,
                                     max_tags, ntags, nthreads);
}

/* ----------------------------------------------------- the chain as a stream */

int ao_stream_init(ao_stream *s, const ao_blocks *blk, const ao_chain_cfg *cfg, const float *symbols,
                   int L)
{
    memset(s, 0, sizeof(*s));
    if (!blk)
        blk = ao_blocks_oracle();
    s->blk = blk;
    s->cfg = *cfg;
    s->L = L;
    s->ce = blk->corr_new(symbols, L, cfg->sps, cfg->mark_delay, cfg->threshold);
    if (!s->ce)
        return -1;
    s->ns = blk->corr_output_multiple(s->ce);
    int rc = 0;
    s->mk = blk->msk_new(cfg->sps, cfg->gain, cfg->limit, cfg->osps, &rc);
    if (rc) {
        blk->corr_delete(s->ce);
        s->ce = 0;
        return rc;
    }
    s->fe = blk->freqest_new(cfg->sample_rate, cfg->data_rate, cfg->fftlen);
    s->xcarry = (float *)calloc((size_t)cfg->fftlen * 2, sizeof(float));
    s->agc_hist = (float *)calloc((size_t)(cfg->agc_nsamples > 1 ? cfg->agc_nsamples - 1 : 1) * 2, sizeof(float));
    s->acarry = (float *)calloc((size_t)(L + s->ns) * 2, sizeof(float)); /* L zeros of history */
    s->ocarry = (float *)calloc(64 * 2, sizeof(float));
    s->captags = 64;
    s->tags = (ao_tag *)calloc((size_t)s->captags, sizeof(ao_tag));
    return 0;
}

ao_stream *ao_stream_new_with(const ao_blocks *blk, const ao_chain_cfg *cfg, const float *symbols, int L)
{
    ao_stream *s = (ao_stream *)malloc(sizeof(ao_stream));
    if (s && ao_stream_init(s, blk, cfg, symbols, L)) {
        free(s);
        s = 0;
    }
    return s;
}

ao_stream *ao_stream_new(const ao_chain_cfg *cfg, const float *symbols, int L)
{
    return ao_stream_new_with(0, cfg, symbols, L);
}

void ao_stream_delete(ao_stream *s)
{
    if (s) {
        ao_stream_free(s);
        free(s);
    }
}

void ao_stream_free(ao_stream *s)
{
    if (s->blk) {
        if (s->ce)
            s->blk->corr_delete(s->ce);
        if (s->mk)
            s->blk->msk_delete(s->mk);
        if (s->fe)
            s->blk->freqest_delete(s->fe);
    }
    free(s->xcarry);
    free(s->agc_hist);
    free(s->acarry);
    free(s->ocarry);
    free(s->tags);
    memset(s, 0, sizeof(*s));
}

/* corr_est_cc::set_symbols on the running stream (same length: the chain's buffers are sized by it) */
int ao_stream_set_symbols(ao_stream *s, const float *symbols, int L)
{
    if (L != s->L)
        return -1;
    s->blk->corr_set_symbols(s->ce, symbols, L);
    s->ns = s->blk->corr_output_multiple(s->ce);
    return 0;
}

int ao_stream_work(ao_stream *s, const float *x, int n, uint8_t *bits, int max_bits, int *nbits,
                   ao_tag *tags_out, int max_tags, int *ntags_out)
{
    const ao_chain_cfg *cfg = &s->cfg;
    const ao_blocks *blk = s->blk;
    const int L = s->L, fftlen = cfg->fftlen, W = cfg->agc_nsamples, ns = s->ns;
    int rc = 0;
    *nbits = 0;
    *ntags_out = 0;
    /* ---- freq sync on whole vectors; the rest of the input waits in xcarry ---- */
    int navail = s->nxcarry + n;
    float *xin = (float *)malloc(sizeof(float) * 2 * (size_t)(navail + 1));
    memcpy(xin, s->xcarry, sizeof(float) * 2 * (size_t)s->nxcarry);
    memcpy(xin + 2 * (size_t)s->nxcarry, x, sizeof(float) * 2 * (size_t)n);
    int n1 = (cfg->stages & AO_STAGE_FREQSYNC) ? (navail / fftlen) * fftlen : navail;
    float *mixed = (float *)malloc(sizeof(float) * 2 * (size_t)(n1 + 1));
    if (cfg->stages & AO_STAGE_FREQSYNC)
        freq_sync(blk, s->fe, cfg, xin, n1, &s->nco_phase, mixed, 0); /* one work(): maxpos starts at 0 */
    else
        memcpy(mixed, xin, sizeof(float) * 2 * (size_t)n1);
    s->nxcarry = navail - n1;
    memcpy(s->xcarry, xin + 2 * (size_t)n1, sizeof(float) * 2 * (size_t)s->nxcarry);
    free(xin);
    /* ---- AGC over every mixed item, history carried ---- */
    int hist = (cfg->stages & AO_STAGE_AGC) ? W - 1 : 0;
    int nac_old = s->nacarry;
    s->acarry = (float *)realloc(s->acarry, sizeof(float) * 2 * (size_t)(L + nac_old + n1 + 1));
    float *agc_out = s->acarry + 2 * (size_t)(L + nac_old);
    if (cfg->stages & AO_STAGE_AGC) {
        float *buf = (float *)malloc(sizeof(float) * 2 * (size_t)(hist + n1 + 1));
        memcpy(buf, s->agc_hist, sizeof(float) * 2 * (size_t)hist);
        memcpy(buf + 2 * (size_t)hist, mixed, sizeof(float) * 2 * (size_t)n1);
        ao_agc_work(buf, n1, W, cfg->agc_reference, agc_out);
        memcpy(s->agc_hist, buf + 2 * (size_t)n1, sizeof(float) * 2 * (size_t)hist);
        free(buf);
    } else {
        memcpy(agc_out, mixed, sizeof(float) * 2 * (size_t)n1);
    }
    free(mixed);
    /* ---- corr_est in work chunks over whole output multiples ---- */
    int avail = nac_old + n1;
    int chunk = cfg->corr_chunk > 0 ? cfg->corr_chunk : ao_default_corr_chunk(L);
    chunk = (chunk / ns) * ns;
    if (chunk <= 0)
        chunk = ns;
    s->ocarry = (float *)realloc(s->ocarry, sizeof(float) * 2 * (size_t)(s->nocarry + avail + 64));
    int done = 0, ntag_new = 0;
    while (avail - done >= ns) {
        int nn = avail - done;
        nn = nn > chunk ? chunk : (nn / ns) * ns;
        int nt = 0;
        if (s->ntags + 8 * (nn / 5 + 2) > s->captags) {
            s->captags = s->ntags + 8 * (nn / 5 + 2) + 64;
            s->tags = (ao_tag *)realloc(s->tags, sizeof(ao_tag) * (size_t)s->captags);
        }
        blk->corr_work(s->ce, nn, s->acarry + 2 * (size_t)done, s->written,
                       s->ocarry + 2 * (size_t)s->nocarry, 0, 0, 0, s->tags + s->ntags,
                       s->captags - s->ntags, &nt);
        for (int k = 0; k < nt; k++) { /* this call's tags for the caller */
            if (ntag_new < max_tags)
                tags_out[ntag_new] = s->tags[s->ntags + k];
            ntag_new++;
        }
        s->ntags += nt;
        s->nocarry += nn;
        s->written += (uint64_t)nn;
        done += nn;
    }
    *ntags_out = ntag_new;
    if (ntag_new > max_tags)
        rc = -3;
    /* keep L items of history + what corr_est has not taken */
    memmove(s->acarry, s->acarry + 2 * (size_t)done, sizeof(float) * 2 * (size_t)(L + avail - done));
    s->nacarry = avail - done;
    /* ---- msk over everything corr_est has produced and msk has not consumed ---- */
    int maxsym = max_bits;
    float *sym = (float *)malloc(sizeof(float) * 2 * (size_t)(maxsym + 1));
    float *soft = (float *)malloc(sizeof(float) * (size_t)(maxsym + 1));
    uint8_t *b0 = (uint8_t *)malloc((size_t)maxsym + 1), *b1 = (uint8_t *)malloc((size_t)maxsym + 1);
    int consumed = 0;
    int k = blk->msk_general_work(s->mk, maxsym, s->nocarry, s->ocarry, s->read, s->tags, s->ntags, sym,
                                  0, 0, &consumed);
    if (k < 0) {
        rc = -4;
        k = 0;
    } else if (k >= maxsym && consumed < (int)((double)s->nocarry - 3.0 * (double)blk->msk_get_sps(s->mk))) {
        rc = -7; /* output row too small */
    }
    memmove(s->ocarry, s->ocarry + 2 * (size_t)consumed, sizeof(float) * 2 * (size_t)(s->nocarry - consumed));
    s->nocarry -= consumed;
    s->read += (uint64_t)consumed;
    int keep = 0; /* drop the tags msk can no longer see */
    for (int t = 0; t < s->ntags; t++)
        if (s->tags[t].offset >= s->read)
            s->tags[keep++] = s->tags[t];
    s->ntags = keep;
    /* ---- bit tail with carried history ---- */
    ao_quad_demod(s->qprev, sym, k, (float)(M_PI / 2), soft);
    ao_binary_slicer(soft, k, b0);
    ao_diff_decoder(&s->dprev, b0, k, 2, b1);
    blk->invert_work(b1, k, bits);
    *nbits = k;
    free(sym);
    free(soft);
    free(b0);
    free(b1);
    return rc;
}
