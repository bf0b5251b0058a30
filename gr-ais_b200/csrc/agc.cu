// agc.cu -- NCO mix (square_and_fft_sync_cc back half) fused with feedforward_agc_cc.
//
// Replaces (paths relative to /root/reference):
//   python/gmsk_sync.py:27-28,33-37   frequency_modulator_fc -> multiply_cc(x, nco)
//   python/ais_demod.py:35            analog.feedforward_agc_cc(512, 2)
//
// One block produces TILE consecutive AGC outputs of one channel.  Output t needs the
// mixed samples y[t-W+1 .. t] (W = agc window), so the block re-mixes a halo of W-1
// samples in front of its tile from the phase checkpoints written by k_nco_phase; the
// mixed stream itself never goes to HBM.
#include <cstdlib>

#include "device_math.cuh"
#include "internal.h"

namespace b200ais {

namespace {

constexpr int kAgcThreads = 256;
constexpr int kAgcTile = 2048;

__global__ void __launch_bounds__(kAgcThreads)
k_mix_agc(const float2 *__restrict__ x, size_t x_stride, int channels, int n1, int fftlen,
          const float *__restrict__ fhat, int vstride, const float *__restrict__ ckpt, int seg,
          float sens,
          int stages, int W, float reference, const float2 *__restrict__ sine,
          float2 *__restrict__ out, size_t out_stride)
{
    extern __shared__ float2 ys[]; // [span] mixed samples, then env[span], then tmp[span]
    const int c = channel_index();
    if (c >= channels)
        return;
    const int t0 = blockIdx.x * kAgcTile;
    const int tile = min(kAgcTile, n1 - t0);
    const bool do_mix = stages & B200AIS_STAGE_FREQSYNC;
    const bool do_agc = stages & B200AIS_STAGE_AGC;
    const int halo = do_agc ? W - 1 : 0;
    const int lo = t0 - halo; // first mixed sample index needed (may be negative: zeros)
    const int span = tile + halo;
    float *env = reinterpret_cast<float *>(ys + (kAgcTile + halo));
    const float2 *xc = x + (size_t)c * x_stride;

    if (do_mix) {
        // segments are aligned to multiples of seg in absolute sample index
        const int s_first = (lo < 0 ? 0 : lo) / seg;
        const int s_last = (t0 + tile - 1) / seg;
        for (int sgi = s_first + threadIdx.x; sgi <= s_last; sgi += blockDim.x) {
            const int n0 = sgi * seg;
            float ph = ckpt[(size_t)sgi * channels + c];
            const float inc = sens * fhat[(size_t)c * vstride + n0 / fftlen];
            for (int i = 0; i < seg; i++) {
                ph = nco_step(ph, inc);
                const int n = n0 + i;
                if (n >= lo && n < t0 + tile) {
                    float sn, cs;
                    fxpt_sincos(float_to_fixed(ph), sine, &sn, &cs);
                    ys[n - lo] = cmul_fma(xc[n], make_float2(cs, sn));
                }
            }
        }
        for (int i = threadIdx.x; i < span; i += blockDim.x)
            if (lo + i < 0)
                ys[i] = make_float2(0.0f, 0.0f);
    } else {
        for (int i = threadIdx.x; i < span; i += blockDim.x) {
            const int n = lo + i;
            ys[i] = n >= 0 ? xc[n] : make_float2(0.0f, 0.0f);
        }
    }
    __syncthreads();
    float2 *oc = out + (size_t)c * out_stride;
    if (!do_agc) {
        for (int i = threadIdx.x; i < tile; i += blockDim.x)
            oc[t0 + i] = ys[i];
        return;
    }
    for (int i = threadIdx.x; i < span; i += blockDim.x)
        env[i] = agc_envelope(ys[i].x, ys[i].y);
    __syncthreads();
    if (W == 512) {
        // van Herk / Gil-Werman: per aligned block of W, prefix maxima P and suffix maxima S;
        // max over [i, i+W-1] = max(S[i], P[i+W-1]).  One warp per block, 16 elements per lane,
        // lane carries combined with shuffles.  max() is exact, so any evaluation order is.
        float *sfx = env + (kAgcTile + halo);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int nblocks = (span + 511) >> 9;
        for (int blk = warp; blk < nblocks; blk += kAgcThreads / 32) {
            const int base = blk * 512 + lane * 16;
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
                v[k] = (base + k < span) ? env[base + k] : 0.0f;
            float pre[16], suf[16];
            pre[0] = v[0];
#pragma unroll
            for (int k = 1; k < 16; k++)
                pre[k] = fmaxf(pre[k - 1], v[k]);
            suf[15] = v[15];
#pragma unroll
            for (int k = 14; k >= 0; k--)
                suf[k] = fmaxf(suf[k + 1], v[k]);
            // exclusive scans of the lane totals (envelopes are >= 0, so 0 is the identity)
            float up = pre[15], dn = suf[0];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float a = __shfl_up_sync(0xffffffffu, up, o);
                const float b = __shfl_down_sync(0xffffffffu, dn, o);
                if (lane >= o)
                    up = fmaxf(up, a);
                if (lane + o < 32)
                    dn = fmaxf(dn, b);
            }
            float cup = __shfl_up_sync(0xffffffffu, up, 1);
            float cdn = __shfl_down_sync(0xffffffffu, dn, 1);
            if (lane == 0)
                cup = 0.0f;
            if (lane == 31)
                cdn = 0.0f;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (base + k < span) {
                    env[base + k] = fmaxf(cup, pre[k]);
                    sfx[base + k] = fmaxf(cdn, suf[k]);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < tile; i += blockDim.x) {
            const float m = fmaxf(sfx[i], env[i + W - 1]);
            const float max_env = fmaxf(1e-4f, m);
            const float gain = reference / max_env;
            const float2 y = ys[i];
            oc[t0 + i] = make_float2(gain * y.x, gain * y.y);
        }
        return;
    }
    // sparse-table doubling: after the pass for width k, env[i] = max(env0[i .. i+k-1]) (clipped)
    int p = 1;
    constexpr int kMaxPer = (kAgcTile + 2047 + kAgcThreads - 1) / kAgcThreads;
    while (p * 2 <= W) {
        float v[kMaxPer];
#pragma unroll
        for (int k = 0; k < kMaxPer; k++) {
            const int i = threadIdx.x + k * kAgcThreads;
            if (i < span) {
                float a = env[i];
                float b = (i + p < span) ? env[i + p] : a;
                v[k] = fmaxf(a, b);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kMaxPer; k++) {
            const int i = threadIdx.x + k * kAgcThreads;
            if (i < span)
                env[i] = v[k];
        }
        __syncthreads();
        p *= 2;
    }
    // out[t] = y[t-W+1] * (reference / max(1e-4, max env over y[t-W+1 .. t]))
    for (int i = threadIdx.x; i < tile; i += blockDim.x) {
        // window starts at local index i (absolute t-W+1) and is W long: two width-p blocks
        float m = fmaxf(env[i], env[i + W - p]);
        float max_env = fmaxf(1e-4f, m);
        float gain = reference / max_env;
        float2 y = ys[i];
        oc[t0 + i] = make_float2(gain * y.x, gain * y.y);
    }
}

// ---- fast path for the reference's constants: feedforward_agc_cc(512, 2), 16-sample
// phase checkpoints.  One block = 4096 consecutive samples (512 of history + 3584 outputs);
// thread t owns samples [16t, 16t+16): it re-runs the NCO from its checkpoint, mixes, takes
// the envelopes and -- because a warp's 32 x 16 samples are exactly one aligned 512 block --
// finishes the van Herk prefix/suffix maxima with warp shuffles without leaving registers.
constexpr int kFSpan = 4096;
constexpr int kFHalo = 512;
constexpr int kFOut = kFSpan - kFHalo;
constexpr int kFPad = kFSpan + kFSpan / 16; // 1-in-16 padding: a thread's 16 items hit 16 banks

// kHist: stream mode, AGC history in / out.  (The paired sine table stays in global memory / L1:
// a copy in shared memory costs the third CTA per SM and measured 6 % slower.)
template <bool kHist>
__global__ void __launch_bounds__(256, 3)
k_mix_agc512(const float2 *__restrict__ x, size_t x_stride, int channels, int n1, int fftlen,
             const float *__restrict__ fhat, int vstride, const float *__restrict__ ckpt, float sens,
             int do_mix, float reference, const float4 *__restrict__ sine, cudaTextureObject_t sine_tex,
             float2 *__restrict__ out, size_t out_stride, const float2 *__restrict__ hist_in,
             float2 *__restrict__ hist_out)
{
    extern __shared__ float2 ys[];               // [kFPad]
    float *P = reinterpret_cast<float *>(ys + kFPad); // prefix maxima  [kFPad]
    float *S = P + kFPad;                             // suffix maxima  [kFPad]
    const int tid = threadIdx.x, lane = tid & 31;
    const int c = channel_index();
    if (c >= channels)
        return;
    const int t0 = blockIdx.x * kFOut;
    const int n0 = t0 - kFHalo + 16 * tid; // absolute index of this thread's first sample
    const float2 *xc = x + (size_t)c * x_stride;
    // stage the block's 4096 input samples through shared memory with coalesced loads (lane i
    // reads sample i of each 256-sample row), then every thread picks up its own 16
    {
        const int base = t0 - kFHalo;
        if (base >= 0 && base + kFSpan <= n1) { // interior block: no edge tests
            const float2 *xb = xc + base + tid;
#pragma unroll
            for (int k = 0; k < 16; k++)
                ys[k * 272 + tid + (tid >> 4)] = xb[k * 256]; // li + (li >> 4), li = 256 k + tid
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int li = k * 256 + tid, n = base + li;
                float2 v0 = make_float2(0.0f, 0.0f);
                if (n >= 0 && n < n1)
                    v0 = xc[n];
                else if (kHist && n < 0 && n >= -(kFHalo - 1)) // already mixed: the stream's AGC history
                    v0 = hist_in[(size_t)c * (kFHalo - 1) + (kFHalo - 1 + n)];
                ys[li + (li >> 4)] = v0;
            }
        }
    }
    __syncthreads();
    float2 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++)
        v[k] = ys[17 * tid + k];
    if (do_mix && n0 >= 0 && n0 < n1) { // n1 is a multiple of fftlen (a multiple of 16): whole segments
        const float ph0 = ckpt[(size_t)(n0 >> 4) * channels + c];
        const float inc = sens * fhat[(size_t)c * vstride + n0 / fftlen];
        const float F_PI = 3.14159265358979323846f;
        // straight-line fast path.  frequency_modulator_fc leaves d_phase = fmod(u, 2 pi) - pi in
        // (-3 pi, pi): below -pi whenever the frequency is negative (fmod keeps the sign).  There
        // float_to_fixed folds with d = floor(x / 2 pi + 0.5) = -1, i.e. x - (float)d * 2 pi =
        // x + 2 pi with one rounding.  d = -1 exactly for the floats in [RN(-1.5 * 2pi_f), -pi_f):
        // their float quotient x / 2pi_f lies in [-1.5, -0.5) (checked float by float at both
        // ends), and d_phase cannot go below RN(-1.5 * 2pi_f) = RN(nextabove(-2pi_f) - pi_f).
        // Anything else takes the general path below.
        const float F_2PI = 2.0f * F_PI;
        float ph = ph0;
        // One test for the whole segment: with |inc| <= pi_f and ph in [-3 pi_f, pi_f) a step has
        // u = (ph + inc) + pi_f in [-3 pi_f, 3 pi_f), the select fold of nco_step() is exact
        // there, and ph' = r - pi_f lands in (-3 pi_f, pi_f) again (r in (-2 pi_f, 2 pi_f)).
        // Anything else (also NaN) takes the general path.
        const bool bad = !(fabsf(inc) <= F_PI && ph0 >= -1.5f * F_2PI && ph0 < F_PI);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            ph = nco_step_inrange(ph, inc);
            const float folded = (ph < -F_PI) ? ph + F_2PI : ph;
            float sn, cs;
            fxpt_sincos4_tex(float_to_fixed_inrange(folded), sine_tex, &sn, &cs);
            v[k] = cmul_fma(v[k], make_float2(cs, sn));
        }
        if (bad) { // general path (fmod, fold, true division) from the raw samples still in ys
            ph = ph0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                ph = nco_step(ph, inc);
                float sn, cs;
                fxpt_sincos4(float_to_fixed(ph), sine, &sn, &cs);
                v[k] = cmul_fma(ys[17 * tid + k], make_float2(cs, sn));
            }
        }
    }
    if (kHist) { // the last 511 mixed items become the next call's AGC history
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int n = n0 + k;
            if (n >= n1 - (kFHalo - 1) && n < n1 && n >= -(kFHalo - 1))
                hist_out[(size_t)c * (kFHalo - 1) + (n - (n1 - (kFHalo - 1)))] = v[k];
        }
    }
    float pre[16], suf[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        ys[17 * tid + k] = v[k];
        pre[k] = agc_envelope(v[k].x, v[k].y);
        suf[k] = pre[k];
    }
#pragma unroll
    for (int k = 1; k < 16; k++)
        pre[k] = fmaxf(pre[k - 1], pre[k]);
#pragma unroll
    for (int k = 14; k >= 0; k--)
        suf[k] = fmaxf(suf[k + 1], suf[k]);
    // exclusive scans of the lane totals across the warp's 512-sample block (max is exact and
    // envelopes are >= 0, so 0 is the identity)
    float up = pre[15], dn = suf[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float a = __shfl_up_sync(0xffffffffu, up, o);
        const float b = __shfl_down_sync(0xffffffffu, dn, o);
        if (lane >= o)
            up = fmaxf(up, a);
        if (lane + o < 32)
            dn = fmaxf(dn, b);
    }
    float cup = __shfl_up_sync(0xffffffffu, up, 1);
    float cdn = __shfl_down_sync(0xffffffffu, dn, 1);
    if (lane == 0)
        cup = 0.0f;
    if (lane == 31)
        cdn = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        P[17 * tid + k] = fmaxf(cup, pre[k]);
        S[17 * tid + k] = fmaxf(cdn, suf[k]);
    }
    __syncthreads();
    // out[t] = y[t-511] * (reference / max(1e-4, max env(y[t-511 .. t]))); local index of
    // y[t-511] is i+1, of y[t] is i+512
    float2 *oc = out + (size_t)c * out_stride;
    const int nout = min(kFOut, n1 - t0);
    for (int i = tid; i < nout; i += 256) {
        const int a = i + 1, b = i + kFHalo;
        const float m = fmaxf(S[a + (a >> 4)], P[b + (b >> 4)]);
        const float gain = reference / fmaxf(1e-4f, m);
        const float2 y = ys[a + (a >> 4)];
        oc[t0 + i] = make_float2(gain * y.x, gain * y.y);
    }
}

} // namespace

int launch_mix_agc(const float2 *x, size_t x_stride, int channels, int n1, int fftlen,
                   const float *fhat, int vstride, const float *ckpt, int seg, float sens, int stages,
                   int agc_nsamples, float agc_reference, float2 *out, size_t out_stride,
                   const float2 *hist_in, float2 *hist_out, cudaStream_t s)
{
    if (n1 <= 0 || channels <= 0)
        return B200AIS_OK;
    if ((stages & B200AIS_STAGE_AGC) && (agc_nsamples < 1 || agc_nsamples > 2048)) {
        set_error("agc window must be in [1, 2048], got %d", agc_nsamples);
        return B200AIS_E_INVALID;
    }
    Tables tb;
    int rc = get_tables(&tb);
    if (rc)
        return rc;
    if ((stages & B200AIS_STAGE_AGC) && agc_nsamples == 512 && seg == 16 &&
        (!(stages & B200AIS_STAGE_FREQSYNC) || (fftlen % 16 == 0 && n1 % fftlen == 0))) {
        const size_t smem512 = (size_t)kFPad * (sizeof(float2) + 2 * sizeof(float));
        dim3 grid512 = channel_grid((n1 + kFOut - 1) / kFOut, channels);
        const int do_mix = (stages & B200AIS_STAGE_FREQSYNC) ? 1 : 0;
        const float4 *sine = reinterpret_cast<const float4 *>(tb.sine4);
        if (hist_in || hist_out) {
            if (!hist_in || !hist_out || hist_in == hist_out) {
                set_error("mix_agc: stream mode needs distinct history buffers in and out");
                return B200AIS_E_INVALID;
            }
#define B200_MIX(H, hi, ho)                                                                       \
    do {                                                                                          \
        B200_CU(cudaFuncSetAttribute(k_mix_agc512<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem512));                                              \
        k_mix_agc512<H><<<grid512, 256, smem512, s>>>(x, x_stride, channels, n1, fftlen, fhat,  \
                                                         vstride, ckpt, sens, do_mix, agc_reference, \
                                                         sine, tb.sine4_tex, out, out_stride, hi, ho); \
    } while (0)
            B200_MIX(true, hist_in, hist_out);
        } else {
            B200_MIX(false, nullptr, nullptr);
        }
#undef B200_MIX
        B200_LAUNCH_CHECK("k_mix_agc512");
        return B200AIS_OK;
    }
    if (hist_in || hist_out) {
        set_error("streaming needs the feedforward_agc_cc(512, .) fast path (16-sample checkpoints)");
        return B200AIS_E_INVALID;
    }
    int halo = (stages & B200AIS_STAGE_AGC) ? agc_nsamples - 1 : 0;
    size_t smem = (size_t)(kAgcTile + halo) * (sizeof(float2) + 2 * sizeof(float));
    B200_CU(cudaFuncSetAttribute(k_mix_agc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((kAgcTile + 2047) * (sizeof(float2) + 2 * sizeof(float)))));
    dim3 grid = channel_grid((n1 + kAgcTile - 1) / kAgcTile, channels);
    k_mix_agc<<<grid, kAgcThreads, smem, s>>>(x, x_stride, channels, n1, fftlen, fhat, vstride, ckpt, seg,
                                              sens, stages, agc_nsamples, agc_reference,
                                              reinterpret_cast<const float2 *>(tb.sine), out,
                                              out_stride);
    B200_LAUNCH_CHECK("k_mix_agc");
    return B200AIS_OK;
}

} // namespace b200ais
