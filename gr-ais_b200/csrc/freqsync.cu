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

// ---- 1024-point fast path: 64 threads, 16 values per thread, three register passes ----
//
// Same radix-2 decimation-in-time graph as the generic kernel (and the oracle): stage s
// combines X[e] and X[e + 2^(s-1)] with W[(e mod 2^(s-1)) * 1024/2^s].  Stages 1-4 touch
// index bits 0-3, stages 5-7 bits 4-6, stages 8-10 bits 7-9, so a thread can hold all the
// elements that differ only in those bits and run the stages in registers; shared memory
// is crossed twice instead of ten times.  W = 1 and W = -i (exact table entries) skip the
// multiply: fma(1,b,-0) = b and fma(0,x,y) = y are exact.
constexpr int kFast = 64;

// 1-in-16 padding plus one slot per 64 elements: conflict-free for the pass-A writers (thread
// tid owns the 16 elements of group bitrev6(tid), so that its loads are coalesced) and for the
// pass-B / pass-C accesses (16 consecutive `lo` per half-warp)
__device__ __forceinline__ int fphys(int e) { return e + (e >> 4) + (e >> 6); }

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
    __shared__ float2 cx[N + N / 16 + N / 64];
    __shared__ float hs[N];
    __shared__ Best sh[kFast / 32];
    __shared__ float s_max[kFast / 32];
    const int tid = threadIdx.x;
    const int b = blockIdx.x, c = channel_index();
    const float2 *src = x + (size_t)c * x_stride + (size_t)b * N;

    // ---- pass A: stages 1-4 on elements e = 16*t6 + q, loaded from x[bitrev10(e)] =
    // x[bitrev4(q)*64 + bitrev6(t6)]; thread tid takes the group t6 = bitrev6(tid), so that the
    // threads of a warp read consecutive items ----
    {
        float2 v[16];
        const int t6 = (int)(__brev((unsigned)tid) >> 26); // bitrev6(tid)
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int r4 = ((q & 1) << 3) | ((q & 2) << 1) | ((q & 4) >> 1) | ((q & 8) >> 3);
            const float2 in = src[r4 * 64 + tid];
            v[q] = cmul_fma2(in, in); // blocks.multiply_cc(x, x)
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
        const int base = 17 * t6 + (t6 >> 2); // fphys(16*t6 + q) = base + q
#pragma unroll
        for (int q = 0; q < 16; q++)
            cx[base + q] = v[q];
    }
    __syncthreads();
    // ---- pass B: stages 5-7 on e = hi*128 + q*16 + lo ----
    // tw[lo << 5], tw[(lo + 16 j) << 4], tw[(lo + 16 j) << 3] from the [slot][lo] copies
    // behind the table (get_twiddles): a warp's read covers 128 consecutive bytes.  Both of a
    // thread's groups (g = tid, tid + 64) have the same lo: one set of loads serves them.
    const int loB = tid & 15;
    const float2 wB1 = tw[512 + loB];
    const float2 wB2[2] = { tw[528 + loB], tw[544 + loB] };
    const float2 wB3[4] = { tw[560 + loB], tw[576 + loB], tw[592 + loB], tw[608 + loB] };
#pragma unroll 1
    for (int g = tid; g < 128; g += kFast) {
        const int hi = g >> 4, lo = g & 15;
        float2 u[8];
#pragma unroll
        for (int q = 0; q < 8; q++)
            u[q] = cx[fphys(hi * 128 + q * 16 + lo)];
        pass8(u, wB1, wB2, wB3);
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
        const float2 w1 = tw[624 + lo];                               // tw[lo << 2]
        const float2 w2[2] = { tw[752 + lo], tw[880 + lo] };          // tw[(lo + 128 j) << 1]
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
    if (fftlen == 1024) {
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
