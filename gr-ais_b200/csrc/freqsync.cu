// freqsync.cu -- square_and_fft_sync_cc front half and the freqest block.
//
// Replaces (paths relative to /root/reference):
//   python/gmsk_sync.py:22-24,30-31   multiply_cc(x,x) -> stream_to_vector -> fft_vcc(shift)
//   lib/freqest_impl.cc:57-88         freqest_impl::work
//   python/gmsk_sync.py:26-27         repeat -> frequency_modulator_fc (phase recurrence)
#include "device_math.cuh"
#include "internal.h"

namespace b200ais {

namespace {

constexpr int kFftThreads = 256;

struct Best {
    float e;
    int j;
};

__device__ __forceinline__ Best better(Best a, Best b)
{
    // strict '>' scan in ascending j: the largest energy wins, ties go to the smaller j
    if (b.e > a.e || (b.e == a.e && b.j < a.j))
        return b;
    return a;
}

__device__ __forceinline__ Best block_argmax(Best v, Best *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best w;
        w.e = __shfl_xor_sync(0xffffffffu, v.e, o);
        w.j = __shfl_xor_sync(0xffffffffu, v.j, o);
        v = better(v, w);
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int nw = (blockDim.x + 31) >> 5;
        Best w;
        w.e = 0.0f;
        w.j = 0x7fffffff;
        if (lane < nw)
            w = sh[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Best u;
            u.e = __shfl_xor_sync(0xffffffffu, w.e, o);
            u.j = __shfl_xor_sync(0xffffffffu, w.j, o);
            w = better(w, u);
        }
        v = w;
    }
    return v; // valid in warp 0
}

// One block per (vector b, channel c): x^2 -> radix-2 DIT FFT in shared memory ->
// |X| in fft-shifted order -> argmax_j |S[j]| + |S[j+offset]|.
__global__ void __launch_bounds__(kFftThreads)
k_sqfft_freqest(const float2 *__restrict__ x, size_t x_stride, int vstride, int n, int lg,
                const float2 *__restrict__ tw, int offset, int *__restrict__ raw, int channels)
{
    if (channel_index() >= channels)
        return;
    extern __shared__ float2 buf[];
    float *hs = reinterpret_cast<float *>(buf + n);
    __shared__ Best sh[kFftThreads / 32];
    const int b = blockIdx.x, c = channel_index();
    const float2 *src = x + (size_t)c * x_stride + (size_t)b * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float2 v = src[i];
        float2 sq = cmul_fma(v, v);
        buf[__brev((unsigned)i) >> (32 - lg)] = sq;
    }
    __syncthreads();
    for (int m = 2; m <= n; m <<= 1) {
        const int half = m >> 1, step = n / m;
        for (int q = threadIdx.x; q < (n >> 1); q += blockDim.x) {
            int j = q & (half - 1);
            int i0 = ((q - j) << 1) + j;
            int i1 = i0 + half;
            float2 w = tw[j * step];
            float2 a = buf[i0], bb = buf[i1];
            float2 t = cmul_fma(w, bb);
            buf[i0] = make_float2(a.x + t.x, a.y + t.y);
            buf[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float2 v = buf[(i + (n >> 1)) & (n - 1)]; // fft_vcc shift: out[i] = X[(i + n/2) mod n]
        hs[i] = hypot_canon(v.x, v.y);
    }
    __syncthreads();
    Best best;
    best.e = 0.0f;
    best.j = 0x7fffffff;
    for (int j = threadIdx.x; j < n - offset; j += blockDim.x) {
        float e = hs[j] + hs[j + offset];
        if (e > best.e) {
            best.e = e;
            best.j = j;
        }
    }
    best = block_argmax(best, sh);
    if (threadIdx.x == 0)
        raw[(size_t)c * vstride + b] = (best.j == 0x7fffffff) ? -1 : best.j + offset / 2;
}

// butterflies in packed FP32 (FFMA2 / FADD2 / FMUL2 on the (re, im) pair, device_math.cuh):
// the same roundings as the scalar forms, half the issue slots
__device__ __forceinline__ void bf(float2 &a, float2 &b, const float2 w)
{
    const float2 t = cmul_fma2(w, b), a0 = a;
    a = f2_add(a0, t);
    b = f2_sub(a0, t);
}
__device__ __forceinline__ void bf_one(float2 &a, float2 &b) // W = (1, 0)
{
    const float2 t = b, a0 = a;
    a = f2_add(a0, t);
    b = f2_sub(a0, t);
}
__device__ __forceinline__ void bf_mi(float2 &a, float2 &b) // W = (0, -1): t = (b.y, -b.x)
{
    const float2 a0 = a, b0 = b;
    a = f2_add(a0, make_float2(b0.y, -b0.x));
    b = f2_add(a0, make_float2(-b0.y, b0.x));
}

// ---- 1024-point path, one warp per transform: 32 values per thread, two register passes ----
//
// The same radix-2 decimation-in-time graph as the generic kernel and the oracle (every butterfly
// is the oracle's, with the oracle's twiddle: stage s combines X[e] and X[e + 2^(s-1)] with
// W[(e mod 2^(s-1)) * 1024/2^s]; W = 1 and W = -i, exact table entries, skip the multiply:
// fma(1,b,-0) = b and fma(0,x,y) = y are exact): stages 1-5 touch index bits 0-4 and stages 6-10 bits 5-9, so a thread that
// holds the 32 elements differing only in those bits runs five stages in registers and shared
// memory is crossed once.  No block-wide barrier: a transform belongs to one warp.
//   pass A: lane t holds e = 32 t + q, q = 0..31, loaded from x[bitrev10(e)] = x[bitrev5(q)*32 +
//           bitrev5(t)] (a warp's load covers 256 consecutive bytes).  The twiddles W_32^j do not
//           depend on the lane: constant memory (c_w32, filled from the table by get_twiddles).
//   pass B: lane lo holds e = 32 q + lo; stage 6..9 twiddles from the [slot][lo] copies behind the
//           table, stage 10 from the table itself (tw[lo + 32 j]); the result is X[k], k = 32 q + lo.
//   freqest: the |X| estimates stay in registers; only the partner bin k + offset comes through
//           shared memory.
constexpr int kTwWarp = 512; // first [slot][lo] entry of this kernel's copies (get_twiddles)
__constant__ float4 c_w32[16]; // W_1024^(32 j) = W_32^j as (re, re, im, im)

__device__ __forceinline__ void bf_c32(float2 &a, float2 &b, int j)
{
    const float4 w = c_w32[j];
    // cmul_fma2(w, b)
    const float2 t = f2_fma(b, make_float2(w.x, w.y), f2_mul(make_float2(-b.y, b.x), make_float2(w.z, w.w)));
    const float2 a0 = a;
    a = f2_add(a0, t);
    b = f2_sub(a0, t);
}

template <int S> __device__ __forceinline__ void warpfft_stage_a(float2 (&v)[32])
{
    constexpr int h = 1 << (S - 1);
#pragma unroll
    for (int q = 0; q < 32; q++) {
        if (q & h)
            continue;
        const int j = (q & (h - 1)) * (16 >> (S - 1)); // W_2^S^(q mod h) = W_32^j
        if (j == 0)
            bf_one(v[q], v[q + h]);
        else if (j == 8)
            bf_mi(v[q], v[q + h]);
        else
            bf_c32(v[q], v[q + h], j);
    }
}

template <int B> __device__ __forceinline__ void warpfft_stage_b(float2 (&u)[32], const float2 *__restrict__ twl)
{
    constexpr int h = 1 << B;
    float2 w[h];
#pragma unroll
    for (int j = 0; j < h; j++)
        w[j] = B < 4 ? twl[kTwWarp + ((h - 1) + j) * 32] : twl[32 * j];
#pragma unroll
    for (int q = 0; q < 32; q++)
        if (!(q & h))
            bf(u[q], u[q + h], w[q & (h - 1)]);
}

// v[q] for a warp-uniform q (a jump table over the 32 registers; the array stays in registers)
__device__ __forceinline__ float2 warpfft_pick(const float2 (&v)[32], int q)
{
    float2 r = v[0];
#define B200_PICK(Q) case Q: r = v[Q]; break;
    switch (q) {
        B200_PICK(1) B200_PICK(2) B200_PICK(3) B200_PICK(4) B200_PICK(5) B200_PICK(6) B200_PICK(7)
        B200_PICK(8) B200_PICK(9) B200_PICK(10) B200_PICK(11) B200_PICK(12) B200_PICK(13) B200_PICK(14)
        B200_PICK(15) B200_PICK(16) B200_PICK(17) B200_PICK(18) B200_PICK(19) B200_PICK(20) B200_PICK(21)
        B200_PICK(22) B200_PICK(23) B200_PICK(24) B200_PICK(25) B200_PICK(26) B200_PICK(27) B200_PICK(28)
        B200_PICK(29) B200_PICK(30) B200_PICK(31)
    default:
        break;
    }
#undef B200_PICK
    return r;
}

constexpr int kWarpCx = 1024 + 64; // 2 pad slots per 32 elements: 16-byte stores and 8-byte loads without conflicts

__global__ void __launch_bounds__(32, 16)
k_sqfft_freqest_1024(const float2 *__restrict__ x, size_t x_stride, int vstride,
                      const float2 *__restrict__ tw, int offset, int *__restrict__ raw, int channels)
{
    constexpr int N = 1024;
    if (channel_index() >= channels)
        return;
    __shared__ __align__(16) float2 cx[kWarpCx];
    const int lane = threadIdx.x;
    const int b = blockIdx.x, c = channel_index();
    const float2 *src = x + (size_t)c * x_stride + (size_t)b * N;
    const int r5 = (int)(__brev((unsigned)lane) >> 27); // bitrev5(lane)

    float2 v[32];
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const int rq = ((q & 1) << 4) | ((q & 2) << 2) | (q & 4) | ((q & 8) >> 2) | ((q & 16) >> 4);
        v[q] = src[rq * 32 + r5];
    }
#pragma unroll
    for (int q = 0; q < 32; q++)
        v[q] = cmul_fma2(v[q], v[q]); // blocks.multiply_cc(x, x)
    warpfft_stage_a<1>(v);
    warpfft_stage_a<2>(v);
    warpfft_stage_a<3>(v);
    warpfft_stage_a<4>(v);
    warpfft_stage_a<5>(v);
    {
        float4 *dst = reinterpret_cast<float4 *>(cx + 34 * lane);
#pragma unroll
        for (int q = 0; q < 32; q += 2)
            dst[q >> 1] = make_float4(v[q].x, v[q].y, v[q + 1].x, v[q + 1].y);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; q++)
        v[q] = cx[34 * q + lane];
    const float2 *twl = tw + lane;
    warpfft_stage_b<0>(v, twl);
    warpfft_stage_b<1>(v, twl);
    warpfft_stage_b<2>(v, twl);
    warpfft_stage_b<3>(v, twl);
    warpfft_stage_b<4>(v, twl);
    __syncwarp(); // every lane has read its pass-B inputs: the buffer now takes the |X| estimates
    // float estimates of |X[k]| for the pre-filter; hs in fft-shifted order: hs[j] = |X[(j + N/2) mod N]|
    float *hs = reinterpret_cast<float *>(cx);
    float est[32];
#pragma unroll
    for (int q = 0; q < 32; q++) {
        // MUFU.SQRT: 2^-22 relative, well inside the pre-filter's margin (below)
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(est[q]) : "f"(__fmaf_rn(v[q].x, v[q].x, v[q].y * v[q].y)));
        hs[((q * 32 + lane) + N / 2) & (N - 1)] = est[q];
    }
    __syncwarp();
    // ---- freqest: argmax_j |S[j]| + |S[j+offset]| with strict '>' in ascending j ----
    // The float estimates are within 5e-7 relative of the canonical (double-evaluated) sums
    // (sum of squares 1e-7, approximate root 2.4e-7, the add 6e-8), so only bins within 2e-6 of
    // the estimated maximum can be the canonical argmax; those few are re-evaluated canonically.
    // Degenerate spectra take the exact path for every bin.
    // Lane `lane` takes j = 32 i + lane: hs[j] is its own est[(i + 16) & 31].
    float m_est = 0.0f;
    float e_est[32];
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const int j = 32 * i + lane;
        const bool in = j < N - offset;
        e_est[i] = in ? est[(i + 16) & 31] + hs[in ? j + offset : 0] : -1.0f;
        m_est = fmaxf(m_est, e_est[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m_est = fmaxf(m_est, __shfl_xor_sync(0xffffffffu, m_est, o));
    const bool exact_all = !(m_est > 1e-12f && m_est < 1e30f);
    const float cut = exact_all ? 0.0f : m_est * (1.0f - 2e-6f); // valid sums are >= 0 (NaN: never the argmax)
    unsigned cand = 0;
#pragma unroll
    for (int i = 0; i < 32; i++)
        cand |= (e_est[i] >= cut ? 1u : 0u) << i;
    // The candidates (one or two per transform, every bin of a degenerate one), one at a time for
    // the whole warp in ascending (i, lane) order per lane: the spectrum is still in registers, so
    // the two lanes that own X[k0] and X[k1] pick them with a warp-uniform index.
    Best best;
    best.e = 0.0f;
    best.j = 0x7fffffff;
    if (exact_all) {
        // every bin is a candidate (an all-zero or non-finite spectrum): through shared memory,
        // every lane on its own bins
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 32; q++)
            cx[34 * q + lane] = v[q];
        __syncwarp();
        while (cand) { // ascending j within the lane
            const int i = __ffs(cand) - 1;
            cand &= cand - 1;
            const int j = 32 * i + lane;
            const int k0 = (j + N / 2) & (N - 1), k1 = (j + offset + N / 2) & (N - 1);
            const float2 p = cx[34 * (k0 >> 5) + (k0 & 31)];
            const float2 q2 = cx[34 * (k1 >> 5) + (k1 & 31)];
            const float e = hypot_canon(p.x, p.y) + hypot_canon(q2.x, q2.y);
            if (e > best.e) {
                best.e = e;
                best.j = j;
            }
        }
    }
    for (;;) {
        const unsigned who = __ballot_sync(0xffffffffu, cand != 0);
        if (!who)
            break;
        const int leader = __ffs(who) - 1;
        const int i = __shfl_sync(0xffffffffu, __ffs(cand) - 1, leader);
        if (lane == leader)
            cand &= cand - 1;
        const int j = 32 * i + leader;
        const int k0 = (j + N / 2) & (N - 1), k1 = (j + offset + N / 2) & (N - 1);
        const float2 p = warpfft_pick(v, k0 >> 5), q2 = warpfft_pick(v, k1 >> 5);
        const float h0 = __shfl_sync(0xffffffffu, hypot_canon(p.x, p.y), k0 & 31);
        const float h1 = __shfl_sync(0xffffffffu, hypot_canon(q2.x, q2.y), k1 & 31);
        const float e = h0 + h1;
        if (lane == leader && e > best.e) { // the leader's candidates come in ascending j
            best.e = e;
            best.j = j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best w;
        w.e = __shfl_xor_sync(0xffffffffu, best.e, o);
        w.j = __shfl_xor_sync(0xffffffffu, best.j, o);
        best = better(best, w);
    }
    if (lane == 0)
        raw[(size_t)c * vstride + b] = (best.j == 0x7fffffff) ? -1 : best.j + offset / 2;
}

// Stand-alone freqest on caller-supplied spectra (any fftlen).
__global__ void __launch_bounds__(kFftThreads)
k_freqest_spec(const float2 *__restrict__ spec, int nvec, int n, int offset, int *__restrict__ raw,
               int channels)
{
    if (channel_index() >= channels)
        return;
    __shared__ Best sh[kFftThreads / 32];
    const int b = blockIdx.x, c = channel_index();
    const float2 *in = spec + ((size_t)c * nvec + b) * (size_t)n;
    Best best;
    best.e = 0.0f;
    best.j = 0x7fffffff;
    for (int j = threadIdx.x; j < n - offset; j += blockDim.x) {
        float2 u = in[j], v = in[j + offset];
        float e = hypot_canon(u.x, u.y) + hypot_canon(v.x, v.y);
        if (e > best.e) {
            best.e = e;
            best.j = j;
        }
    }
    best = block_argmax(best, sh);
    if (threadIdx.x == 0)
        raw[(size_t)c * nvec + b] = (best.j == 0x7fffffff) ? -1 : best.j + offset / 2;
}

// lib/freqest_impl.cc:67-68,84: maxpos is a local of work(), zero at the start of the call
// and NOT reset per vector, so a vector with no energy repeats the previous estimate.
__device__ __forceinline__ float maxpos_to_hz(int maxpos, int n, float binsize)
{
    return ((float)(unsigned)maxpos - (float)((unsigned)n / 2u)) * binsize / 2.0f;
}

__global__ void k_freqest_resolve(const int *__restrict__ raw, int channels, int nvec, int n,
                                  float binsize, float *__restrict__ out)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    int maxpos = 0;
    for (int b = 0; b < nvec; b++) {
        int r = raw[(size_t)c * nvec + b];
        if (r >= 0)
            maxpos = r;
        out[(size_t)c * nvec + b] = maxpos_to_hz(maxpos, n, binsize);
    }
}

// One thread per channel walks the whole record: the frequency_modulator_fc phase is a
// float recurrence (one rounding per sample) and cannot be evaluated out of order.  Only
// the recurrence runs here; phases are checkpointed every `seg` samples so the mixing can
// be done by many threads in k_mix_agc.
template <int kSeg>
__global__ void k_nco_phase(const int *__restrict__ raw, int channels, int nvec, int vstride, int n,
                            float binsize, float sens, float *__restrict__ fhat,
                            float *__restrict__ ckpt, int seg, float *__restrict__ phase_state)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    float ph = phase_state ? phase_state[c] : 0.0f; // NCO phase carried by the stream (0 when fresh)
    int maxpos = 0;
    const int segs_per_vec = n / seg;
    float *cp = ckpt + c; // ckpt[segment * channels + c], walked with a running pointer
    for (int b = 0; b < nvec; b++) {
        int r = raw[(size_t)c * vstride + b];
        if (r >= 0)
            maxpos = r;
        float f = maxpos_to_hz(maxpos, n, binsize);
        fhat[(size_t)c * vstride + b] = f;
        const float inc = sens * f;
        for (int sgi = 0; sgi < segs_per_vec; sgi++) {
            *cp = ph;
            cp += channels;
            if (kSeg == 16) {
                const float ph0 = ph;
                bool bad = false;
#pragma unroll
                for (int i = 0; i < 16; i++)
                    ph = nco_step_nobranch(ph, inc, bad);
                if (bad) { // |phase| ran past 4 pi: take the general fmod path for this segment
                    ph = ph0;
                    for (int i = 0; i < 16; i++)
                        ph = nco_step(ph, inc);
                }
            } else {
                for (int i = 0; i < seg; i++)
                    ph = nco_step(ph, inc);
            }
        }
    }
    if (phase_state)
        phase_state[c] = ph;
}

} // namespace

// get_twiddles(1024): the lane-independent twiddles of k_sqfft_freqest_1024's first pass, from
// the table itself (so both sides of the parity contract use the same sixteen values)
int sqfft_set_w32(const float2 *tw_host)
{
    float4 h[16];
    for (int j = 0; j < 16; j++)
        h[j] = make_float4(tw_host[32 * j].x, tw_host[32 * j].x, tw_host[32 * j].y, tw_host[32 * j].y);
    B200_CU(cudaMemcpyToSymbol(c_w32, h, sizeof(h)));
    return B200AIS_OK;
}

static int ilog2_exact(int n)
{
    int lg = 0;
    while ((1 << lg) < n)
        lg++;
    return ((1 << lg) == n) ? lg : -1;
}

int launch_sqfft_freqest(const float2 *x, size_t x_stride, int channels, int nvec, int vstride,
                         int fftlen, int offset, int *raw, cudaStream_t s)
{
    int lg = ilog2_exact(fftlen);
    if (lg < 2 || fftlen > 4096) {
        set_error("fused freq sync needs a power-of-two fftlen in [4, 4096], got %d", fftlen);
        return B200AIS_E_INVALID;
    }
    if (nvec <= 0 || channels <= 0)
        return B200AIS_OK;
    const float2 *tw = nullptr;
    int rc = get_twiddles(fftlen, &tw);
    if (rc)
        return rc;
    size_t smem = (size_t)fftlen * (sizeof(float2) + sizeof(float));
    dim3 grid = channel_grid(nvec, channels);
    if (fftlen == 1024) {
        k_sqfft_freqest_1024<<<grid, 32, 0, s>>>(x, x_stride, vstride, tw, offset, raw, channels);
    } else {
        if (smem > 40 * 1024) // static shared memory counts towards the 48 KB default limit
            B200_CU(cudaFuncSetAttribute(k_sqfft_freqest, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        k_sqfft_freqest<<<grid, kFftThreads, smem, s>>>(x, x_stride, vstride, fftlen, lg, tw, offset, raw,
                                                        channels);
    }
    B200_LAUNCH_CHECK("k_sqfft_freqest");
    return B200AIS_OK;
}

int launch_freqest_spec(const float2 *spec, int channels, int nvec, int fftlen, int offset,
                        int *raw, cudaStream_t s)
{
    if (nvec <= 0 || channels <= 0)
        return B200AIS_OK;
    dim3 grid = channel_grid(nvec, channels);
    k_freqest_spec<<<grid, kFftThreads, 0, s>>>(spec, nvec, fftlen, offset, raw, channels);
    B200_LAUNCH_CHECK("k_freqest_spec");
    return B200AIS_OK;
}

int launch_freqest_resolve(const int *raw, int channels, int nvec, int fftlen, float binsize,
                           float *out, cudaStream_t s)
{
    if (nvec <= 0 || channels <= 0)
        return B200AIS_OK;
    int threads = 128;
    k_freqest_resolve<<<(channels + threads - 1) / threads, threads, 0, s>>>(raw, channels, nvec,
                                                                              fftlen, binsize, out);
    B200_LAUNCH_CHECK("k_freqest_resolve");
    return B200AIS_OK;
}

int launch_nco_phase(const int *raw, int channels, int nvec, int vstride, int fftlen, float binsize,
                     float sens, float *fhat, float *ckpt, int seg, float *phase_state, cudaStream_t s)
{
    if (nvec <= 0 || channels <= 0)
        return B200AIS_OK;
    int threads = 32; // latency-bound serial walk: spread the channels over as many SMs as possible
    const int blocks = (channels + threads - 1) / threads;
    if (seg == 16)
        k_nco_phase<16><<<blocks, threads, 0, s>>>(raw, channels, nvec, vstride, fftlen, binsize, sens,
                                                   fhat, ckpt, seg, phase_state);
    else
        k_nco_phase<0><<<blocks, threads, 0, s>>>(raw, channels, nvec, vstride, fftlen, binsize, sens,
                                                  fhat, ckpt, seg, phase_state);
    B200_LAUNCH_CHECK("k_nco_phase");
    return B200AIS_OK;
}

} // namespace b200ais
