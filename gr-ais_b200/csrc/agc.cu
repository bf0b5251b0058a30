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

#include <cuda.h> // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

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

// ---- the same fast path with the tile moved by the tensor-memory accelerator (TMA) ----
//
// What bounds k_mix_agc512 is the load/store pipe: to hand every thread 16 consecutive samples it
// transposes the tile through shared memory with 16 LDG + 16 STS + 16 LDS per thread, stores the
// mixed samples and the prefix / suffix maxima back (48 STS) and reads three arrays per output (11
// load/store instructions per sample, L1TEX 79 % busy).  Here
//   * the input tile arrives by ONE cp.async.bulk.tensor load (UTMALDG) of a [256 segments][16
//     samples] box with the 128-byte swizzle: segment r's 16-byte chunk j lands at chunk j ^ (r & 7),
//     so thread r reads its own 16 samples with eight conflict-free LDS.128 and no transposition;
//     samples in front of the record (the AGC's zero history) are the TMA's out-of-bounds zero fill;
//   * each warp owns one aligned 512-sample block: mixed samples, prefix and suffix maxima stay in
//     registers; the only exchange is the prefix maxima of the NEXT block (4 STS.128 + 4 LDS.128);
//   * the output row of a thread is [z1 .. z15, z0 of the next thread] (outputs trail the samples
//     by 511 = 512 - 1), written with eight swizzled STS.128 into the tile's own shared memory and
//     sent out by ONE cp.async.bulk.tensor store (UTMASTG).
// 2.6 load/store instructions per sample instead of 11, the same arithmetic in the same order.
constexpr int kTSeg = 16;                    // samples per segment = one 128-byte row
constexpr int kTRows = kFSpan / kTSeg;       // 256 rows in, one per thread
constexpr int kTOutRows = kFOut / kTSeg;     // 224 rows out
constexpr int kTPStride = 20;                // floats between two threads' 16 prefix maxima (80 B: conflict-free 16-byte accesses)
constexpr int kTSmem = 1024 /* alignment */ + kTRows * 128 + 8 * 32 * kTPStride * 4 + 8 * 8 + 64;

__device__ __forceinline__ void agc_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void agc_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void agc_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}

template <bool kHist>
__global__ void __launch_bounds__(256, 3)
k_mix_agc512_tma(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                 int channels, int n1, int fftlen, const float *__restrict__ fhat, int vstride,
                 const float *__restrict__ ckpt, float sens, int do_mix, float reference,
                 const float4 *__restrict__ sine, cudaTextureObject_t sine_tex,
                 const float2 *__restrict__ hist_in, float2 *__restrict__ hist_out)
{
    extern __shared__ unsigned char agc_smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = channel_index();
    if (c >= channels)
        return;
    // tile (1024-byte aligned: the swizzle pattern repeats every 8 rows), prefix maxima, z0 slots
    unsigned char *base = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(agc_smem_raw) + 1023) & ~(uintptr_t)1023);
    float *Ps = reinterpret_cast<float *>(base + kTRows * 128);            // [8][32][kTPStride]
    float2 *zfirst = reinterpret_cast<float2 *>(Ps + 8 * 32 * kTPStride);  // [8]
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(zfirst + 8);
    const unsigned tile_s = (unsigned)__cvta_generic_to_shared(base);
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar);

    const int t0 = blockIdx.x * kFOut;
    const int seg0 = t0 / kTSeg - kFHalo / kTSeg; // first segment of the tile (negative: zero fill)
    if (tid == 0) {
        agc_mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        agc_mbar_expect_tx(bar_s, kTRows * 128);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
                     "[%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(tile_s), "l"(&tm_in), "r"(0), "r"(seg0), "r"(c), "r"(bar_s)
                     : "memory");
    }
    const int n0 = t0 - kFHalo + 16 * tid; // absolute index of this thread's first sample
    const bool mix = do_mix && n0 >= 0 && n0 < n1;
    float ph0 = 0.0f, inc = 0.0f;
    if (mix) { // n1 is a multiple of fftlen (a multiple of 16): whole segments
        ph0 = ckpt[(size_t)(n0 >> 4) * channels + c];
        inc = sens * fhat[(size_t)c * vstride + n0 / fftlen];
    }
    // the NCO of this thread's 16 samples does not need the samples: it runs while the tile is
    // still in flight (phase recurrence, fixed-point angle, table fetch, sin / cos)
    float2 osc[16]; // (cos, sin) per sample
    if (mix) {
        const float F_PI = 3.14159265358979323846f;
        const float F_2PI = 2.0f * F_PI;
        // straight-line fast path and its one range test: see k_mix_agc512 above
        const bool bad = !(fabsf(inc) <= F_PI && ph0 >= -1.5f * F_2PI && ph0 < F_PI);
        if (!bad) {
            float ph = ph0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                ph = nco_step_inrange(ph, inc);
                const float folded = (ph < -F_PI) ? ph + F_2PI : ph;
                float sn, cs;
                fxpt_sincos4_tex(float_to_fixed_inrange(folded), sine_tex, &sn, &cs);
                osc[k] = make_float2(cs, sn);
            }
        } else { // general path (fmod, fold, true division)
            float ph = ph0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                ph = nco_step(ph, inc);
                float sn, cs;
                fxpt_sincos4(float_to_fixed(ph), sine, &sn, &cs);
                osc[k] = make_float2(cs, sn);
            }
        }
    }
    __syncthreads(); // the barrier's initialisation is visible before anyone polls it
    agc_mbar_wait(bar_s, 0);
    float2 v[16];
    {
        const unsigned row = tile_s + (unsigned)tid * 128u;
        const unsigned sw = (unsigned)(tid & 7) << 4;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float4 q;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                         : "r"(row + (((unsigned)j << 4) ^ sw)));
            v[2 * j] = make_float2(q.x, q.y);
            v[2 * j + 1] = make_float2(q.z, q.w);
        }
    }
    if (kHist && n0 < 0) { // already mixed: the stream's AGC history instead of the zero fill
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int n = n0 + k;
            if (n >= -(kFHalo - 1))
                v[k] = hist_in[(size_t)c * (kFHalo - 1) + (kFHalo - 1 + n)];
        }
    }
    if (mix) {
#pragma unroll
        for (int k = 0; k < 16; k++)
            v[k] = cmul_fma(v[k], osc[k]);
    }
    if (kHist) { // the last 511 mixed items become the next call's AGC history
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int n = n0 + k;
            if (n >= n1 - (kFHalo - 1) && n < n1 && n >= -(kFHalo - 1))
                hist_out[(size_t)c * (kFHalo - 1) + (n - (n1 - (kFHalo - 1)))] = v[k];
        }
    }
    // envelopes, prefix / suffix maxima of this warp's 512-sample block (van Herk / Gil-Werman)
    float pre[16], suf[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        pre[k] = agc_envelope(v[k].x, v[k].y);
        suf[k] = pre[k];
    }
#pragma unroll
    for (int k = 1; k < 16; k++)
        pre[k] = fmaxf(pre[k - 1], pre[k]);
#pragma unroll
    for (int k = 14; k >= 0; k--)
        suf[k] = fmaxf(suf[k + 1], suf[k]);
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
    // this block's prefix maxima for the warp in front of it; its first item's output (a window
    // that is exactly this block: max = S[0]) for that warp's last row
    {
        float *pr = Ps + (warp * 32 + lane) * kTPStride;
#pragma unroll
        for (int j = 0; j < 4; j++)
            *reinterpret_cast<float4 *>(pr + 4 * j) =
                make_float4(fmaxf(cup, pre[4 * j]), fmaxf(cup, pre[4 * j + 1]), fmaxf(cup, pre[4 * j + 2]),
                            fmaxf(cup, pre[4 * j + 3]));
        if (lane == 0) {
            const float g0 = reference / fmaxf(1e-4f, fmaxf(cdn, suf[0]));
            zfirst[warp] = make_float2(g0 * v[0].x, g0 * v[0].y);
        }
    }
    __syncthreads(); // every thread holds its samples in registers: the tile can be overwritten
    if (warp < 7) {
        // out[t] = y[t-511] * (reference / max(1e-4, max env(y[t-511 .. t]))): for item e of this
        // block the window is the block's suffix from e and the next block's prefix up to e - 1
        const float *pn = Ps + ((warp + 1) * 32 + lane) * kTPStride;
        float pnx[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 q = *reinterpret_cast<const float4 *>(pn + 4 * j);
            pnx[4 * j] = q.x;
            pnx[4 * j + 1] = q.y;
            pnx[4 * j + 2] = q.z;
            pnx[4 * j + 3] = q.w;
        }
        float pm1 = __shfl_up_sync(0xffffffffu, pnx[15], 1); // next block's prefix up to 16 lane - 1
        if (lane == 0)
            pm1 = 0.0f; // item 0: the window ends with this block
        float2 z[16];
        // Every window maximum of this thread is at most `bound`.  With it (and the reference
        // level) far from the ends of the exponent range, IEEE division is the straight-line
        // sequence below -- the very instructions nvcc emits for `a / b`, minus the range check
        // (FCHK), its branch and the call they guard: 6 instructions per sample instead of 12,
        // and the sixteen gains become one basic block.  Anything else takes the operator.
        const float bound = fmaxf(fmaxf(cdn, suf[0]), pnx[15]);
        const bool plain = bound <= 1e15f && reference >= 1e-10f && reference <= 1e10f;
        if (plain) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const float m = fmaxf(fmaxf(cdn, suf[k]), k ? pnx[k - 1] : pm1);
                const float gain = div_rn_inrange(reference, fmaxf(1e-4f, m));
                z[k] = make_float2(gain * v[k].x, gain * v[k].y);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const float m = fmaxf(fmaxf(cdn, suf[k]), k ? pnx[k - 1] : pm1);
                const float gain = reference / fmaxf(1e-4f, m);
                z[k] = make_float2(gain * v[k].x, gain * v[k].y);
            }
        }
        // row g = 32 warp + lane of the output tile: outputs 16 g .. 16 g + 15 = items 1 .. 15 of
        // this thread and item 0 of the next one
        float zx = __shfl_down_sync(0xffffffffu, z[0].x, 1);
        float zy = __shfl_down_sync(0xffffffffu, z[0].y, 1);
        if (lane == 31) {
            const float2 zf = zfirst[warp + 1];
            zx = zf.x;
            zy = zf.y;
        }
        const unsigned row = tile_s + (unsigned)tid * 128u;
        const unsigned sw = (unsigned)(tid & 7) << 4;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 a = z[2 * j + 1];
            const float2 b = (j == 7) ? make_float2(zx, zy) : z[2 * j + 2];
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"r"(row + (((unsigned)j << 4) ^ sw)), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)
                         : "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> async proxy
    __syncthreads();
    if (tid == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                     ::"l"(&tm_out), "r"(0), "r"(t0 / kTSeg), "r"(c), "r"(tile_s)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the tile is read before the CTA leaves
    }
}

// self-test: div_rn_inrange(a, b) against a / b for every float b with bits in [lo, hi]
__global__ void k_selftest_div(float a, uint32_t lo, uint32_t hi, unsigned long long *bad)
{
    unsigned long long mine = 0;
    const unsigned long long n = (unsigned long long)hi - lo + 1;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float b = __uint_as_float(lo + (uint32_t)i);
        const float q = a / b;
        const float f = div_rn_inrange(a, b);
        if (__float_as_uint(q) != __float_as_uint(f))
            mine++;
    }
    if (mine)
        atomicAdd(bad, mine);
}

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
encode_tiled_fn get_encoder()
{
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
    }
    return fn;
}

// [channels][items] complex rows as a 3-D tensor of 128-byte segments: {32 floats, items/16, channels}
bool make_segment_map(CUtensorMap *tm, const float2 *basep, size_t stride_items, int items, int channels,
                      int box_rows)
{
    encode_tiled_fn enc = get_encoder();
    if (!enc)
        return false;
    const cuuint64_t dims[3] = { 32, (cuuint64_t)(items / kTSeg), (cuuint64_t)channels };
    const cuuint64_t strides[2] = { 128, (cuuint64_t)stride_items * sizeof(float2) };
    const cuuint32_t box[3] = { 32, (cuuint32_t)box_rows, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2 *>(basep), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
        if ((hist_in || hist_out) && (!hist_in || !hist_out || hist_in == hist_out)) {
            set_error("mix_agc: stream mode needs distinct history buffers in and out");
            return B200AIS_E_INVALID;
        }
        // TMA path: whole 16-sample segments and 16-byte aligned rows on both sides
        static int no_tma = -1; // B200AIS_AGC_NO_TMA=1: the LDG/STS transposition everywhere (experiment)
        if (no_tma < 0) {
            const char *e = getenv("B200AIS_AGC_NO_TMA");
            no_tma = (e && atoi(e)) ? 1 : 0;
        }
        const bool aligned = n1 % kTSeg == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                             (reinterpret_cast<uintptr_t>(out) & 15) == 0 && x_stride % 2 == 0 &&
                             out_stride % 2 == 0;
        CUtensorMap tm_in, tm_out;
        if (!no_tma && aligned && make_segment_map(&tm_in, x, x_stride, n1, channels, kTRows) &&
            make_segment_map(&tm_out, out, out_stride, n1, channels, kTOutRows)) {
#define B200_MIX_TMA(H)                                                                           \
    do {                                                                                          \
        B200_CU(cudaFuncSetAttribute(k_mix_agc512_tma<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)kTSmem));                                               \
        k_mix_agc512_tma<H><<<grid512, 256, kTSmem, s>>>(tm_in, tm_out, channels, n1, fftlen, fhat, \
                                                            vstride, ckpt, sens, do_mix, agc_reference, \
                                                            sine, tb.sine4_tex, hist_in, hist_out); \
    } while (0)
            if (hist_in)
                B200_MIX_TMA(true);
            else
                B200_MIX_TMA(false);
#undef B200_MIX_TMA
            B200_LAUNCH_CHECK("k_mix_agc512_tma");
            return B200AIS_OK;
        }
        if (hist_in || hist_out) {
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

extern "C" int b200ais_selftest_div(float a, uint32_t b_lo_bits, uint32_t b_hi_bits,
                                    unsigned long long *mismatches)
{
    using namespace b200ais;
    if (!mismatches || b_hi_bits < b_lo_bits) {
        set_error("selftest_div: bad arguments");
        return B200AIS_E_INVALID;
    }
    unsigned long long *d = nullptr;
    B200_CU(cudaMalloc(&d, sizeof(*d)));
    B200_CU(cudaMemset(d, 0, sizeof(*d)));
    k_selftest_div<<<148 * 8, 256>>>(a, b_lo_bits, b_hi_bits, d);
    B200_LAUNCH_CHECK("k_selftest_div");
    cudaError_t e = cudaMemcpy(mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_error("selftest_div: %s", cudaGetErrorString(e));
        return B200AIS_E_CUDA;
    }
    return B200AIS_OK;
}
