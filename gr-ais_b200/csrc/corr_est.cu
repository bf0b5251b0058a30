// corr_est.cu -- corr_est_cc's detector (the correlation filter itself is corr_fft.cu).
//
// Replaces (paths relative to /root/reference):
//   lib/corr_est_cc_impl.cc:193-271   threshold / peak climb / centre of mass / tags
//   lib/corr_est_cc_impl.cc:184       the delayed pass-through of output 0
//
// k_detect: the serial scan of :193-271, one warp per channel.  It walks the |corr|^2 > thresh
// bitmask written by k_corr_fft with ballots and only touches the correlator stream around the
// few samples it stops at.
#include "device_math.cuh"
#include "internal.h"

namespace b200ais {

namespace {

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

__global__ void k_detect(const float2 *__restrict__ corr, size_t corr_stride, int channels,
                         int n_total, int chunk, int ns, int isps, unsigned mark_delay,
                         const uint8_t *__restrict__ mask, size_t mask_stride, uint64_t base_offset,
                         int two_ports, const float *__restrict__ atan_tab,
                         b200ais_tag *__restrict__ tags, int max_tags, int *__restrict__ ntags,
                         int *__restrict__ status, int append)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= channels)
        return;
    const unsigned FULL = 0xffffffffu;
    const float2 *cc = corr + (size_t)c * corr_stride;
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mask + (size_t)c * mask_stride);
    b200ais_tag *tg = tags + (size_t)c * max_tags;
    int cnt = append ? ntags[c] : 0; // stream mode: the list persists from call to call
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
                    cv = cc[tw];
                mg = cv.x * cv.x + cv.y * cv.y; // volk_32fc_magnitude_squared_32f (:191)
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
    }
}

__global__ void k_tags_compact(b200ais_tag *__restrict__ tags, int max_tags, int *__restrict__ ntags,
                               const int *__restrict__ unconsumed, uint64_t written, int channels,
                               int *__restrict__ nold)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    b200ais_tag *tg = tags + (size_t)c * max_tags;
    const uint64_t read = written - (uint64_t)unconsumed[c]; // msk's nitems_read
    const int n = min(ntags[c], max_tags);
    int keep = 0;
    for (int k = 0; k < n; k++) {
        const b200ais_tag t = tg[k];
        if (t.offset >= read)
            tg[keep++] = t;
    }
    ntags[c] = keep;
    nold[c] = keep;
}

__global__ void k_tags_emit(const b200ais_tag *__restrict__ tags, int max_tags,
                            const int *__restrict__ ntags, const int *__restrict__ nold, int channels,
                            b200ais_tag *__restrict__ out_tags, int *__restrict__ out_ntags,
                            int *__restrict__ status)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    const int first = nold[c], cnt = ntags[c];
    const int last = min(cnt, max_tags);
    for (int k = first; k < last; k++)
        out_tags[(size_t)c * max_tags + (k - first)] = tags[(size_t)c * max_tags + k];
    if (out_ntags)
        out_ntags[c] = cnt - first;
    if (cnt > max_tags)
        atomicMin(status, (int)B200AIS_E_TAG_OVERFLOW);
}

__global__ void k_roll_rows(float2 *__restrict__ rows, size_t stride, int from, int len)
{
    extern __shared__ float2 hold[];
    float2 *r = rows + (size_t)blockIdx.x * stride;
    for (int i = threadIdx.x; i < len; i += blockDim.x)
        hold[i] = r[from + i];
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x)
        r[i] = hold[i];
}

} // namespace

int launch_tags_compact(b200ais_tag *tags, int max_tags, int *ntags, const int *unconsumed,
                        uint64_t written, int channels, int *nold, cudaStream_t s)
{
    if (channels <= 0)
        return B200AIS_OK;
    k_tags_compact<<<(channels + 127) / 128, 128, 0, s>>>(tags, max_tags, ntags, unconsumed, written,
                                                          channels, nold);
    B200_LAUNCH_CHECK("k_tags_compact");
    return B200AIS_OK;
}

int launch_tags_emit(const b200ais_tag *tags, int max_tags, const int *ntags, const int *nold,
                     int channels, b200ais_tag *out_tags, int *out_ntags, int *status, cudaStream_t s)
{
    if (channels <= 0)
        return B200AIS_OK;
    k_tags_emit<<<(channels + 127) / 128, 128, 0, s>>>(tags, max_tags, ntags, nold, channels, out_tags,
                                                       out_ntags, status);
    B200_LAUNCH_CHECK("k_tags_emit");
    return B200AIS_OK;
}

int launch_roll_rows(float2 *rows, size_t stride, int channels, int from, int len, cudaStream_t s)
{
    if (channels <= 0 || len <= 0 || from <= 0)
        return B200AIS_OK;
    const size_t smem = (size_t)len * sizeof(float2);
    if (smem > 200 * 1024) {
        set_error("roll_rows: %d items do not fit in shared memory", len);
        return B200AIS_E_INVALID;
    }
    if (smem > 48 * 1024)
        B200_CU(cudaFuncSetAttribute(k_roll_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_roll_rows<<<channels, 256, smem, s>>>(rows, stride, from, len);
    B200_LAUNCH_CHECK("k_roll_rows");
    return B200AIS_OK;
}

int launch_detect(const float2 *corr, size_t corr_stride, int channels, int n_total, int chunk,
                  int nsamples_mult, int isps, unsigned mark_delay, const uint8_t *mask,
                  size_t mask_stride, uint64_t base_offset, int two_ports, b200ais_tag *tags,
                  int max_tags, int *ntags, int *status, int append, cudaStream_t s)
{
    if (channels <= 0)
        return B200AIS_OK;
    Tables tb;
    int rc = get_tables(&tb);
    if (rc)
        return rc;
    const int threads = 128; // 4 warps = 4 channels per block
    const int blocks = (channels * 32 + threads - 1) / threads;
    k_detect<<<blocks, threads, 0, s>>>(corr, corr_stride, channels, n_total, chunk, nsamples_mult,
                                        isps, mark_delay, mask, mask_stride, base_offset, two_ports,
                                        tb.atan, tags, max_tags, ntags, status, append);
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
