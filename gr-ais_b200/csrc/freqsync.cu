// freqsync.cu -- square_and_fft_sync_cc front half and the freqest block.
//
// Replaces (paths relative to /root/reference):
//   python/gmsk_sync.py:22-24,30-31   multiply_cc(x,x) -> stream_to_vector -> fft_vcc(shift)
//   lib/freqest_impl.cc:57-88         freqest_impl::work
//   python/gmsk_sync.py:26-27         repeat -> frequency_modulator_fc (phase recurrence)
#include <cstdlib>

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

// ---- 1024-point fast path: 64 threads, 16 values per thread, three register passes ----
//
// Same radix-2 decimation-in-time graph as the generic kernel (and the oracle): stage s
// combines X[e] and X[e + 2^(s-1)] with W[(e mod 2^(s-1)) * 1024/2^s].  Stages 1-4 touch
// index bits 0-3, stages 5-7 bits 4-6, stages 8-10 bits 7-9, so a thread can hold all the
// elements that differ only in those bits and run the stages in registers; shared memory
// is crossed twice instead of ten times.  W = 1 and W = -i (exact table entries) skip the
// multiply: fma(1,b,-0) = b and fma(0,x,y) = y are exact.
constexpr int kFast = 64;

__device__ __forceinline__ int fphys(int e) { return e + (e >> 4); } // 1-in-16 padding

__device__ __forceinline__ void bf(float2 &a, float2 &b, const float2 w)
{
    const float2 t = cmul_fma(w, b), a0 = a;
    a = make_float2(a0.x + t.x, a0.y + t.y);
    b = make_float2(a0.x - t.x, a0.y - t.y);
}
__device__ __forceinline__ void bf_one(float2 &a, float2 &b) // W = (1, 0)
{
    const float2 t = b, a0 = a;
    a = make_float2(a0.x + t.x, a0.y + t.y);
    b = make_float2(a0.x - t.x, a0.y - t.y);
}
__device__ __forceinline__ void bf_mi(float2 &a, float2 &b) // W = (0, -1): t = (b.y, -b.x)
{
    const float2 a0 = a, b0 = b;
    a = make_float2(a0.x + b0.y, a0.y - b0.x);
    b = make_float2(a0.x - b0.y, a0.y + b0.x);
}

// three radix-2 stages on 8 register values whose pair distances are 1, 2, 4
__device__ __forceinline__ void pass8(float2 (&u)[8], const float2 w1, const float2 (&w2)[2],
                                      const float2 (&w3)[4])
{
#pragma unroll
    for (int q = 0; q < 8; q += 2)
        bf(u[q], u[q + 1], w1);
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (!(q & 2))
            bf(u[q], u[q + 2], w2[q & 1]);
#pragma unroll
    for (int q = 0; q < 4; q++)
        bf(u[q], u[q + 4], w3[q]);
}

__global__ void __launch_bounds__(kFast)
k_sqfft_freqest_1024(const float2 *__restrict__ x, size_t x_stride, int vstride,
                     const float2 *__restrict__ tw, int offset, int *__restrict__ raw, int channels)
{
    constexpr int N = 1024;
    if (channel_index() >= channels)
        return;
    __shared__ float2 cx[N + N / 16];
    __shared__ float hs[N];
    __shared__ Best sh[kFast / 32];
    __shared__ float s_max[kFast / 32];
    const int tid = threadIdx.x;
    const int b = blockIdx.x, c = channel_index();
    const float2 *src = x + (size_t)c * x_stride + (size_t)b * N;

    // ---- pass A: stages 1-4 on elements e = 16*tid + q, loaded from x[bitrev10(e)] ----
    {
        float2 v[16];
        const int r6 = (int)(__brev((unsigned)tid) >> 26); // bitrev6(tid)
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int r4 = ((q & 1) << 3) | ((q & 2) << 1) | ((q & 4) >> 1) | ((q & 8) >> 3);
            const float2 in = src[r4 * 64 + r6];
            v[q] = cmul_fma(in, in); // blocks.multiply_cc(x, x)
        }
        const float2 w128 = tw[128], w384 = tw[384];
        const float2 w64 = tw[64], w192 = tw[192], w320 = tw[320], w448 = tw[448];
#pragma unroll
        for (int q = 0; q < 16; q += 2)
            bf_one(v[q], v[q + 1]);
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
            bf_one(v[q], v[q + 2]);
            bf_mi(v[q + 1], v[q + 3]);
        }
#pragma unroll
        for (int q = 0; q < 16; q += 8) {
            bf_one(v[q], v[q + 4]);
            bf(v[q + 1], v[q + 5], w128);
            bf_mi(v[q + 2], v[q + 6]);
            bf(v[q + 3], v[q + 7], w384);
        }
        bf_one(v[0], v[8]);
        bf(v[1], v[9], w64);
        bf(v[2], v[10], w128);
        bf(v[3], v[11], w192);
        bf_mi(v[4], v[12]);
        bf(v[5], v[13], w320);
        bf(v[6], v[14], w384);
        bf(v[7], v[15], w448);
#pragma unroll
        for (int q = 0; q < 16; q++)
            cx[17 * tid + q] = v[q]; // fphys(16*tid + q)
    }
    __syncthreads();
    // ---- pass B: stages 5-7 on e = hi*128 + q*16 + lo ----
#pragma unroll 1
    for (int g = tid; g < 128; g += kFast) {
        const int hi = g >> 4, lo = g & 15;
        float2 u[8];
#pragma unroll
        for (int q = 0; q < 8; q++)
            u[q] = cx[fphys(hi * 128 + q * 16 + lo)];
        const float2 w1 = tw[lo << 5];
        const float2 w2[2] = { tw[lo << 4], tw[(lo + 16) << 4] };
        const float2 w3[4] = { tw[lo << 3], tw[(lo + 16) << 3], tw[(lo + 32) << 3], tw[(lo + 48) << 3] };
        pass8(u, w1, w2, w3);
#pragma unroll
        for (int q = 0; q < 8; q++)
            cx[fphys(hi * 128 + q * 16 + lo)] = u[q];
    }
    __syncthreads();
    // ---- pass C: stages 8-10 on e = q*128 + lo; output bin k = e (natural order) ----
#pragma unroll 1
    for (int lo = tid; lo < 128; lo += kFast) {
        float2 u[8];
#pragma unroll
        for (int q = 0; q < 8; q++)
            u[q] = cx[fphys(q * 128 + lo)];
        const float2 w1 = tw[lo << 2];
        const float2 w2[2] = { tw[lo << 1], tw[(lo + 128) << 1] };
        const float2 w3[4] = { tw[lo], tw[lo + 128], tw[lo + 256], tw[lo + 384] };
        pass8(u, w1, w2, w3);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k = q * 128 + lo;
            cx[fphys(k)] = u[q];
            // float estimate of |X[k]| for the pre-filter, stored in fft-shifted order
            hs[(k + N / 2) & (N - 1)] = sqrtf(__fmaf_rn(u[q].x, u[q].x, u[q].y * u[q].y));
        }
    }
    __syncthreads();
    // ---- freqest: argmax_j |S[j]| + |S[j+offset]| with strict '>' in ascending j ----
    // The float estimates are within 3e-7 relative of the canonical (double-evaluated) sums,
    // so only bins within 2e-6 of the estimated maximum can be the canonical argmax; those
    // few are re-evaluated canonically.  Degenerate spectra take the exact path for every bin.
    float m_est = 0.0f;
    for (int j = tid; j < N - offset; j += kFast)
        m_est = fmaxf(m_est, hs[j] + hs[j + offset]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m_est = fmaxf(m_est, __shfl_xor_sync(0xffffffffu, m_est, o));
    if ((tid & 31) == 0)
        s_max[tid >> 5] = m_est;
    __syncthreads();
    m_est = fmaxf(s_max[0], s_max[1]);
    const bool exact_all = !(m_est > 1e-12f && m_est < 1e30f);
    const float cut = m_est * (1.0f - 2e-6f);
    Best best;
    best.e = 0.0f;
    best.j = 0x7fffffff;
    for (int j = tid; j < N - offset; j += kFast) {
        if (exact_all || hs[j] + hs[j + offset] >= cut) {
            const float2 p = cx[fphys((j + N / 2) & (N - 1))];
            const float2 q2 = cx[fphys((j + offset + N / 2) & (N - 1))];
            const float e = hypot_canon(p.x, p.y) + hypot_canon(q2.x, q2.y);
            if (e > best.e) {
                best.e = e;
                best.j = j;
            }
        }
    }
    best = block_argmax(best, sh);
    if (tid == 0)
        raw[(size_t)c * vstride + b] = (best.j == 0x7fffffff) ? -1 : best.j + offset / 2;
}

// ---- 1024-point warp path: one warp per vector, 32 values per thread, two register passes ----
//
// The same decimation-in-time graph again, cut in two: stages 1-5 touch index bits 0-4, stages
// 6-10 bits 5-9.  Lane t first holds elements 32 t + q (q = 0..31: every combination of the low
// five bits), runs five stages in registers, the warp transposes through shared memory, lane t
// then holds elements 32 q + t and runs the other five.  One crossing instead of two, no block
// barrier (a warp owns its vector from the load to the argmax), and the twiddles of the first
// pass are the same sixteen for every lane (broadcast reads), those of the second are read from
// a per-stage table laid out [entry][lane] (conflict-free).
constexpr int kWarpsPerCta = 2;
constexpr int kVecPerWarp = 4; // vectors a warp walks (amortises the twiddle staging)

__device__ __forceinline__ Best warp_argmax(Best v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best w;
        w.e = __shfl_xor_sync(0xffffffffu, v.e, o);
        w.j = __shfl_xor_sync(0xffffffffu, v.j, o);
        v = better(v, w);
    }
    return v;
}

// stage ST (1..5) of the graph on the 32 register values of pass A: pairs q, q + 2^(ST-1),
// twiddle W[(q mod 2^(ST-1)) * 2^(10-ST)] = twA[(q mod 2^(ST-1)) << (5-ST)] (twA[k] = W[32 k])
template <int ST>
__device__ __forceinline__ void stage_a(float2 (&v)[32], const float2 *__restrict__ twA)
{
    constexpr int H = 1 << (ST - 1);
#pragma unroll
    for (int q = 0; q < 32; q++) {
        if (q & H)
            continue;
        const int k = (q & (H - 1)) << (5 - ST); // W[32 k]
        if (k == 0)
            bf_one(v[q], v[q + H]);
        else if (k == 8)
            bf_mi(v[q], v[q + H]);
        else
            bf(v[q], v[q + H], twA[k]);
    }
}

// stage ST (6..10) on pass B's values (element 32 q + t in slot q): pairs q, q + 2^(ST-6),
// twiddle W[(32 (q mod 2^(ST-6)) + t) << (10-ST)] = twB[(2^(ST-6) - 1 + q mod 2^(ST-6)) * 32 + t]
template <int ST>
__device__ __forceinline__ void stage_b(float2 (&v)[32], const float2 *__restrict__ twB_lane)
{
    constexpr int H = 1 << (ST - 6);
#pragma unroll
    for (int qq = 0; qq < H; qq++) {
        const float2 w = twB_lane[(H - 1 + qq) * 32];
#pragma unroll
        for (int q = qq; q < 32; q += 2 * H)
            bf(v[q], v[q + H], w);
    }
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
k_sqfft_freqest_1024w(const float2 *__restrict__ x, size_t x_stride, int vstride, int nvec,
                      const float2 *__restrict__ tw, int offset, int *__restrict__ raw, int channels)
{
    constexpr int N = 1024;
    if (channel_index() >= channels)
        return;
    __shared__ float2 s_twA[16];
    __shared__ float2 s_twB[31 * 32];
    __shared__ float2 s_cx[kWarpsPerCta][32 * 33];
    __shared__ float s_mag[kWarpsPerCta][N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = channel_index();
    for (int i = threadIdx.x; i < 16; i += blockDim.x)
        s_twA[i] = tw[32 * i];
    for (int i = threadIdx.x; i < 31 * 32; i += blockDim.x) {
        const int k1 = (i >> 5) + 1, t = i & 31; // k1 = 2^(ST-6) + qq
        const int sh = 31 - __clz(k1), qq = k1 - (1 << sh); // sh = ST - 6
        s_twB[i] = tw[(32 * qq + t) << (4 - sh)];
    }
    __syncthreads();
    float2 *cx = s_cx[warp];
    float *mag = s_mag[warp];
    const int r5l = (int)(__brev((unsigned)lane) >> 27); // bitrev5(lane)

#pragma unroll 1
    for (int it = 0; it < kVecPerWarp; it++) {
        const int b = (blockIdx.x * kWarpsPerCta + warp) * kVecPerWarp + it;
        if (b >= nvec)
            break; // warp-uniform
        const float2 *src = x + (size_t)c * x_stride + (size_t)b * N;
        float2 v[32];
        // ---- pass A: stages 1-5 on elements e = 32 lane + q, loaded from x[bitrev10(e)] ----
#pragma unroll
        for (int q = 0; q < 32; q++) {
            const int r5q = ((q & 1) << 4) | ((q & 2) << 2) | (q & 4) | ((q & 8) >> 2) | ((q & 16) >> 4);
            const float2 in = src[r5q * 32 + r5l];
            v[q] = cmul_fma(in, in); // blocks.multiply_cc(x, x)
        }
        stage_a<1>(v, s_twA);
        stage_a<2>(v, s_twA);
        stage_a<3>(v, s_twA);
        stage_a<4>(v, s_twA);
        stage_a<5>(v, s_twA);
        __syncwarp(); // the previous vector's readers are done with cx / mag
#pragma unroll
        for (int q = 0; q < 32; q++)
            cx[33 * lane + q] = v[q]; // element 32 lane + q
        __syncwarp();
        // ---- pass B: stages 6-10 on elements e = 32 q + lane ----
#pragma unroll
        for (int q = 0; q < 32; q++)
            v[q] = cx[33 * q + lane];
        stage_b<6>(v, s_twB + lane);
        stage_b<7>(v, s_twB + lane);
        stage_b<8>(v, s_twB + lane);
        stage_b<9>(v, s_twB + lane);
        stage_b<10>(v, s_twB + lane);
        __syncwarp();
        // float magnitudes (natural order, bin k = 32 q + lane) to shared memory; the spectrum
        // itself stays in registers
#pragma unroll
        for (int q = 0; q < 32; q++)
            mag[32 * q + lane] = sqrtf(__fmaf_rn(v[q].x, v[q].x, v[q].y * v[q].y));
        __syncwarp();
        // ---- freqest: argmax_j |S[j]| + |S[j+offset]|, S[j] = X[(j + 512) mod 1024] ----
        // float estimates first (within 3e-7 of the canonical sums), then the canonical
        // evaluation of the bins within 2e-6 of the estimated maximum (see the 64-thread kernel)
        float m_est = 0.0f;
        for (int j = lane; j < N - offset; j += 32)
            m_est = fmaxf(m_est, mag[(j + N / 2) & (N - 1)] + mag[(j + offset + N / 2) & (N - 1)]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            m_est = fmaxf(m_est, __shfl_xor_sync(0xffffffffu, m_est, o));
        const bool exact_all = !(m_est > 1e-12f && m_est < 1e30f);
        const float cut = m_est * (1.0f - 2e-6f);
        if (exact_all) { // an all-zero vector (a silent channel) has no maximum: skip the scan
            unsigned any = 0;
#pragma unroll
            for (int q = 0; q < 32; q++)
                any |= (__float_as_uint(v[q].x) | __float_as_uint(v[q].y)) << 1;
            if (!__any_sync(0xffffffffu, any != 0)) {
                if (lane == 0)
                    raw[(size_t)c * vstride + b] = -1;
                continue;
            }
        }
        Best best;
        best.e = 0.0f;
        best.j = 0x7fffffff;
        // bin k lives in slot k >> 5 of lane k & 31: a candidate pair is fetched with two
        // warp-wide picks (a select chain over the slots + a shuffle); candidates are few
        for (int j0 = 0; j0 < N - offset; j0 += 32) {
            const int j = j0 + lane;
            const int k1 = (j + N / 2) & (N - 1), k2 = (j + offset + N / 2) & (N - 1);
            const bool cand = j < N - offset && (exact_all || mag[k1] + mag[k2] >= cut);
            unsigned todo = __ballot_sync(0xffffffffu, cand);
            while (todo) {
                const int ow = __ffs(todo) - 1; // owner lane of this candidate
                todo &= todo - 1;
                const int a1 = __shfl_sync(0xffffffffu, k1, ow), a2 = __shfl_sync(0xffffffffu, k2, ow);
                float2 p = v[0], q2 = v[0];
#pragma unroll
                for (int q = 1; q < 32; q++) {
                    if ((a1 >> 5) == q)
                        p = v[q];
                    if ((a2 >> 5) == q)
                        q2 = v[q];
                }
                p.x = __shfl_sync(0xffffffffu, p.x, a1 & 31);
                p.y = __shfl_sync(0xffffffffu, p.y, a1 & 31);
                q2.x = __shfl_sync(0xffffffffu, q2.x, a2 & 31);
                q2.y = __shfl_sync(0xffffffffu, q2.y, a2 & 31);
                if (lane == ow) {
                    const float e = hypot_canon(p.x, p.y) + hypot_canon(q2.x, q2.y);
                    if (e > best.e) { // ascending j within a lane: strict '>' keeps the first
                        best.e = e;
                        best.j = j;
                    }
                }
            }
        }
        best = warp_argmax(best);
        if (lane == 0)
            raw[(size_t)c * vstride + b] = (best.j == 0x7fffffff) ? -1 : best.j + offset / 2;
    }
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
    static int variant = -1; // B200AIS_SQFFT=64: the 64-thread kernel (default: warp per vector)
    if (variant < 0) {
        const char *e = getenv("B200AIS_SQFFT");
        variant = (e && atoi(e) == 64) ? 64 : 32;
    }
    if (fftlen == 1024 && variant == 32) {
        const int per_cta = kWarpsPerCta * kVecPerWarp;
        dim3 gw = channel_grid((nvec + per_cta - 1) / per_cta, channels);
        k_sqfft_freqest_1024w<<<gw, 32 * kWarpsPerCta, 0, s>>>(x, x_stride, vstride, nvec, tw, offset,
                                                               raw, channels);
    } else if (fftlen == 1024) {
        k_sqfft_freqest_1024<<<grid, kFast, 0, s>>>(x, x_stride, vstride, tw, offset, raw, channels);
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
