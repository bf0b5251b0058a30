// corr_est.cu -- the corr_est_cc preamble correlator and its detector.
//
// Replaces (paths relative to /root/reference):
//   lib/corr_est_cc_impl.cc:186-191   fft_filter_ccc::filter + volk magnitude_squared
//   lib/corr_est_cc_impl.cc:193-271   threshold / peak climb / centre of mass / tags
//
// k_corr  : direct-form sliding correlation, register-tiled (8 outputs x 8 taps per
//           step per thread), taps and a padded input tile staged in shared memory.
//           It only emits one bit per sample (|corr|^2 > thresh); the correlator
//           stream is written only when the block's optional 2nd output is connected.
// k_detect: the serial scan of lib/corr_est_cc_impl.cc:193-271, one warp per channel.
//           It walks the bitmask with ballots and re-evaluates the correlation (same
//           operation order => same bits) only around the few samples it stops at.
#include "device_math.cuh"
#include "internal.h"

namespace b200ais {

namespace {

constexpr int kCorrThreads = 128;
constexpr int kCorrR = 8;                         // outputs per thread
constexpr int kCorrTile = kCorrThreads * kCorrR;  // outputs per block

// shared-memory index of logical sample i: 2 float2 of padding after every 8 keeps every
// thread's 16-byte loads (thread stride 8 samples -> 80 bytes) bank-conflict free.
__device__ __forceinline__ int phys(int i) { return i + 2 * (i >> 3); }

// canonical complex MAC, taps in time order g[m] pairing with sample t-L+1+m:
//   re = fma(gr, xr, re); re = fma(-gi, xi, re); im = fma(gr, xi, im); im = fma(gi, xr, im)
__device__ __forceinline__ void cmac(float2 &acc, const float2 g, const float2 x)
{
    acc.x = __fmaf_rn(g.x, x.x, acc.x);
    acc.x = __fmaf_rn(-g.y, x.y, acc.x);
    acc.y = __fmaf_rn(g.x, x.y, acc.y);
    acc.y = __fmaf_rn(g.y, x.x, acc.y);
}

__device__ __forceinline__ void load8(float2 (&dst)[8], const float2 *p)
{
    const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float4 v = q[k];
        dst[2 * k] = make_float2(v.x, v.y);
        dst[2 * k + 1] = make_float2(v.z, v.w);
    }
}

// 8 taps x 8 outputs: output r, tap j uses window sample j + r (lo = samples 0..7, hi = 8..15)
__device__ __forceinline__ void tile8(float2 (&acc)[8], const float2 (&lo)[8], const float2 (&hi)[8],
                                      const float2 (&g)[8])
{
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int k = j + r;
            cmac(acc[r], g[j], k < 8 ? lo[k] : hi[k - 8]);
        }
    }
}

__global__ void __launch_bounds__(kCorrThreads)
k_corr(const float2 *__restrict__ in, size_t in_stride, int n, int n_valid,
       const float2 *__restrict__ taps_time, int L, int Lp, float thresh,
       uint8_t *__restrict__ mask, size_t mask_stride, float2 *__restrict__ corr_out,
       size_t corr_stride)
{
    extern __shared__ float4 smem4[];
    float2 *ts = reinterpret_cast<float2 *>(smem4); // [Lp] taps, zero padded
    float2 *xs = ts + Lp;                           // padded input tile
    const int c = blockIdx.y;
    const int t0 = blockIdx.x * kCorrTile;
    const float2 *xc = in + (size_t)c * in_stride;
    const int nload = kCorrTile + Lp; // logical samples t0-(L-1) .. t0-(L-1)+nload-1
    for (int i = threadIdx.x; i < nload; i += kCorrThreads) {
        const int t = t0 - (L - 1) + i;
        float2 v = make_float2(0.0f, 0.0f);
        if (t >= -L && t < n_valid)
            v = xc[t];
        xs[phys(i)] = v;
    }
    for (int m = threadIdx.x; m < Lp; m += kCorrThreads)
        ts[m] = m < L ? taps_time[m] : make_float2(0.0f, 0.0f);
    __syncthreads();

    float2 acc[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
        acc[r] = make_float2(0.0f, 0.0f);
    const float2 *xb = xs + threadIdx.x * 10; // phys(tid*8)
    float2 wa[8], wb[8], g[8];
    load8(wa, xb);
    int mm = 0;
    for (; mm + 16 <= Lp; mm += 16) {
        load8(wb, xb + ((mm + 8) >> 3) * 10);
        load8(g, ts + mm);
        tile8(acc, wa, wb, g);
        load8(wa, xb + ((mm + 16) >> 3) * 10);
        load8(g, ts + mm + 8);
        tile8(acc, wb, wa, g);
    }
    if (mm < Lp) {
        load8(wb, xb + ((mm + 8) >> 3) * 10);
        load8(g, ts + mm);
        tile8(acc, wa, wb, g);
    }

    const int tbase = t0 + threadIdx.x * 8;
    unsigned bits = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        // volk_32fc_magnitude_squared_32f: re*re + im*im, two products and one add
        float mag = acc[r].x * acc[r].x + acc[r].y * acc[r].y;
        if (tbase + r < n && !(mag <= thresh))
            bits |= 1u << r;
    }
    const size_t byte = (size_t)(tbase >> 3);
    if (byte < mask_stride)
        mask[(size_t)c * mask_stride + byte] = (uint8_t)bits;
    if (corr_out) {
        float2 *oc = corr_out + (size_t)c * corr_stride;
#pragma unroll
        for (int r = 0; r < 8; r++)
            if (tbase + r < n)
                oc[tbase + r] = acc[r];
    }
}

// correlation at one output index, same operation order as k_corr
__device__ __forceinline__ float2 corr_at(const float2 *__restrict__ xc, int t,
                                          const float2 *__restrict__ g, int L)
{
    float2 acc = make_float2(0.0f, 0.0f);
    const float2 *p = xc + (t - L + 1);
    for (int m = 0; m < L; m++)
        cmac(acc, g[m], p[m]);
    return acc;
}

__device__ __forceinline__ void push_tag(b200ais_tag *tags, int max_tags, int &cnt, uint64_t off,
                                         int key, int port, double value)
{
    if (cnt < max_tags) {
        b200ais_tag tg;
        tg.offset = off;
        tg.key = key;
        tg.port = port;
        tg.value = value;
        tags[cnt] = tg;
    }
    cnt++;
}

__global__ void k_detect(const float2 *__restrict__ in, size_t in_stride, int channels, int n_total,
                         int chunk, int ns, const float2 *__restrict__ taps_time, int L,
                         float thresh, int isps, unsigned mark_delay,
                         const uint8_t *__restrict__ mask, size_t mask_stride, uint64_t base_offset,
                         int two_ports, const float *__restrict__ atan_tab,
                         b200ais_tag *__restrict__ tags, int max_tags, int *__restrict__ ntags,
                         int *__restrict__ n2_out, int *__restrict__ status)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= channels)
        return;
    const unsigned FULL = 0xffffffffu;
    const float2 *xc = in + (size_t)c * in_stride;
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mask + (size_t)c * mask_stride);
    b200ais_tag *tg = tags + (size_t)c * max_tags;
    int cnt = 0;
    int cs = 0;
    while (n_total - cs >= ns) {
        int nn = n_total - cs;
        nn = nn > chunk ? chunk : (nn / ns) * ns;
        const int ce = cs + nn;
        int t = cs;
        while (t < ce) {
            // next sample >= t (and < ce) whose |corr|^2 exceeds the threshold
            const int w0 = t >> 5;
            const int wbase = (w0 + lane) << 5;
            uint32_t word = 0;
            if (wbase < ce)
                word = mw[w0 + lane];
            if (lane == 0)
                word &= 0xffffffffu << (t & 31);
            if (wbase + 32 > ce)
                word &= (ce > wbase) ? ((ce - wbase >= 32) ? 0xffffffffu : ((1u << (ce - wbase)) - 1u)) : 0u;
            const unsigned ballot = __ballot_sync(FULL, word != 0);
            if (!ballot) {
                t = (w0 + 32) << 5;
                continue;
            }
            const int src = __ffs(ballot) - 1;
            const uint32_t wv = __shfl_sync(FULL, word, src);
            t = ((w0 + src) << 5) + __ffs(wv) - 1;
            // climb to the local peak; lane l holds the correlation at index wstart + l
            int wstart, pos;
            float2 cv;
            float mg;
            for (;;) {
                wstart = t - 1;
                pos = 1;
                const int tw = wstart + lane;
                cv = make_float2(0.0f, 0.0f);
                if (tw >= cs && tw < ce)
                    cv = corr_at(xc, tw, taps_time, L);
                mg = cv.x * cv.x + cv.y * cv.y;
                bool again = false;
                while (t < ce - 1) {
                    if (pos == 31) {
                        again = true;
                        break;
                    }
                    const float m0 = __shfl_sync(FULL, mg, pos);
                    const float m1 = __shfl_sync(FULL, mg, pos + 1);
                    if (m0 < m1) {
                        t++;
                        pos++;
                    } else {
                        break;
                    }
                }
                if (!again)
                    break;
            }
            const float m_prev = __shfl_sync(FULL, mg, pos - 1);
            const float m_here = __shfl_sync(FULL, mg, pos);
            const float m_next = __shfl_sync(FULL, mg, (pos + 1) & 31);
            const float c_re = __shfl_sync(FULL, cv.x, pos);
            const float c_im = __shfl_sync(FULL, cv.y, pos);
            double center = 0.0;
            if (t > cs && t < ce - 1) {
                // nom += (s+1)*mag: the int*float product is a float, the sums are double
                double nom = 0.0, den = 0.0;
                nom += (double)(1.0f * m_prev);
                den += (double)m_prev;
                nom += (double)(2.0f * m_here);
                den += (double)m_here;
                nom += (double)(3.0f * m_next);
                den += (double)m_next;
                center = nom / den - 2.0;
            }
            const float phase = fast_atan2f_tab(c_im, c_re, atan_tab);
            if (lane == 0) {
                const uint64_t o_i = base_offset + (uint64_t)t;
                const uint64_t o_m = o_i + mark_delay;
                push_tag(tg, max_tags, cnt, o_i, B200AIS_TAG_CORR_START, 0, (double)m_here);
                push_tag(tg, max_tags, cnt, o_m, B200AIS_TAG_PHASE_EST, 0, (double)phase);
                push_tag(tg, max_tags, cnt, o_m, B200AIS_TAG_TIME_EST, 0, center);
                push_tag(tg, max_tags, cnt, o_m, B200AIS_TAG_CORR_EST, 0, (double)m_here);
                if (two_ports) {
                    push_tag(tg, max_tags, cnt, o_i, B200AIS_TAG_PHASE_EST, 1, (double)phase);
                    push_tag(tg, max_tags, cnt, o_i, B200AIS_TAG_TIME_EST, 1, center);
                    push_tag(tg, max_tags, cnt, o_i, B200AIS_TAG_CORR_EST, 1, (double)m_here);
                }
            }
            t += isps;
        }
        cs = ce;
    }
    if (lane == 0) {
        ntags[c] = cnt;
        if (cnt > max_tags)
            atomicMin(status, (int)B200AIS_E_TAG_OVERFLOW);
        if (c == 0 && n2_out)
            *n2_out = cs;
    }
}

} // namespace

int launch_corr(const float2 *in, size_t in_stride, int channels, int n, int n_valid,
                const float2 *taps_time, int L, float thresh, uint8_t *mask, size_t mask_stride,
                float2 *corr_out, size_t corr_stride, cudaStream_t s)
{
    if (n <= 0 || channels <= 0)
        return B200AIS_OK;
    if (L < 1 || L > 4096) {
        set_error("corr_est supports 1..4096 taps, got %d", L);
        return B200AIS_E_INVALID;
    }
    const int Lp = (L + 7) & ~7;
    const int nload = kCorrTile + Lp;
    size_t smem = (size_t)Lp * sizeof(float2) + (size_t)(nload + 2 * (nload >> 3) + 16) * sizeof(float2);
    B200_CU(cudaFuncSetAttribute(k_corr, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    dim3 grid((n + kCorrTile - 1) / kCorrTile, channels);
    k_corr<<<grid, kCorrThreads, smem, s>>>(in, in_stride, n, n_valid, taps_time, L, Lp, thresh,
                                            mask, mask_stride, corr_out, corr_stride);
    B200_LAUNCH_CHECK("k_corr");
    return B200AIS_OK;
}

int launch_detect(const float2 *in, size_t in_stride, int channels, int n_total, int chunk,
                  int nsamples_mult, const float2 *taps_time, int L, float thresh, int isps,
                  unsigned mark_delay, const uint8_t *mask, size_t mask_stride, uint64_t base_offset,
                  int two_ports, b200ais_tag *tags, int max_tags, int *ntags, int *n2_out,
                  int *status, cudaStream_t s)
{
    if (channels <= 0)
        return B200AIS_OK;
    Tables tb;
    int rc = get_tables(&tb);
    if (rc)
        return rc;
    const int threads = 128; // 4 warps = 4 channels per block
    const int blocks = (channels * 32 + threads - 1) / threads;
    k_detect<<<blocks, threads, 0, s>>>(in, in_stride, channels, n_total, chunk, nsamples_mult,
                                        taps_time, L, thresh, isps, mark_delay, mask, mask_stride,
                                        base_offset, two_ports, tb.atan, tags, max_tags, ntags,
                                        n2_out, status);
    B200_LAUNCH_CHECK("k_detect");
    return B200AIS_OK;
}

int launch_copy_delay(const float2 *in, size_t in_stride, float2 *out, size_t out_stride,
                      int channels, int n, cudaStream_t s)
{
    if (n <= 0 || channels <= 0)
        return B200AIS_OK;
    // lib/corr_est_cc_impl.cc:184: memcpy(out, &in[0], n items) -- the history makes it a delay
    B200_CU(cudaMemcpy2DAsync(out, out_stride * sizeof(float2), in, in_stride * sizeof(float2),
                              (size_t)n * sizeof(float2), channels, cudaMemcpyDeviceToDevice, s));
    return B200AIS_OK;
}

} // namespace b200ais
