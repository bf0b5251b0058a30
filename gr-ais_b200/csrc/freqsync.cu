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
                const float2 *__restrict__ tw, int offset, int *__restrict__ raw)
{
    extern __shared__ float2 buf[];
    float *hs = reinterpret_cast<float *>(buf + n);
    __shared__ Best sh[kFftThreads / 32];
    const int b = blockIdx.x, c = blockIdx.y;
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

// Stand-alone freqest on caller-supplied spectra (any fftlen).
__global__ void __launch_bounds__(kFftThreads)
k_freqest_spec(const float2 *__restrict__ spec, int nvec, int n, int offset, int *__restrict__ raw)
{
    __shared__ Best sh[kFftThreads / 32];
    const int b = blockIdx.x, c = blockIdx.y;
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
__global__ void k_nco_phase(const int *__restrict__ raw, int channels, int nvec, int vstride, int n,
                            float binsize, float sens, float *__restrict__ fhat,
                            float *__restrict__ ckpt, int seg)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    float ph = 0.0f;
    int maxpos = 0;
    const int segs_per_vec = n / seg;
    for (int b = 0; b < nvec; b++) {
        int r = raw[(size_t)c * vstride + b];
        if (r >= 0)
            maxpos = r;
        float f = maxpos_to_hz(maxpos, n, binsize);
        fhat[(size_t)c * vstride + b] = f;
        const float inc = sens * f;
        for (int sgi = 0; sgi < segs_per_vec; sgi++) {
            ckpt[((size_t)b * segs_per_vec + sgi) * channels + c] = ph;
            if (seg == 16) {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    ph = nco_step(ph, inc);
            } else {
                for (int i = 0; i < seg; i++)
                    ph = nco_step(ph, inc);
            }
        }
    }
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
    dim3 grid(nvec, channels);
    k_sqfft_freqest<<<grid, kFftThreads, smem, s>>>(x, x_stride, vstride, fftlen, lg, tw, offset, raw);
    B200_LAUNCH_CHECK("k_sqfft_freqest");
    return B200AIS_OK;
}

int launch_freqest_spec(const float2 *spec, int channels, int nvec, int fftlen, int offset,
                        int *raw, cudaStream_t s)
{
    if (nvec <= 0 || channels <= 0)
        return B200AIS_OK;
    dim3 grid(nvec, channels);
    k_freqest_spec<<<grid, kFftThreads, 0, s>>>(spec, nvec, fftlen, offset, raw);
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
                     float sens, float *fhat, float *ckpt, int seg, cudaStream_t s)
{
    if (nvec <= 0 || channels <= 0)
        return B200AIS_OK;
    int threads = 32; // latency-bound serial walk: spread the channels over as many SMs as possible
    k_nco_phase<<<(channels + threads - 1) / threads, threads, 0, s>>>(raw, channels, nvec, vstride,
                                                                        fftlen, binsize, sens, fhat,
                                                                        ckpt, seg);
    B200_LAUNCH_CHECK("k_nco_phase");
    return B200AIS_OK;
}

} // namespace b200ais
