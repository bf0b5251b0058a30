// msk.cu -- msk_timing_recovery_cc and the bit tail behind it.
//
// Replaces (paths relative to /root/reference):
//   lib/msk_timing_recovery_cc_impl.cc:107-206   general_work (the D'Andrea/Mengali/
//                                                Reggiannini loop at 2 x symbol rate)
//   python/ais_demod.py:48-52                    quadrature_demod_cf(pi/2) -> binary_slicer_fb
//                                                -> diff_decoder_bb(2) -> ais.invert
//   lib/invert_impl.cc:54-68                     invert_impl::work
//
// The loop carries mu / omega / the delay registers from one half-symbol to the next, so a
// channel is strictly serial: one thread per channel, thousands of channels in flight.
#include "device_math.cuh"
#include "internal.h"

namespace b200ais {

namespace {

// mmse_fir_interpolator_cc::interpolate: imu = rint(mu*128), dot product of in[0..7] with
// the reversed row.  Canonical summation: p_j = in[j]*t[j] (+) in[j+4]*t[j+4] fused,
// result (p0+p1)+(p2+p3).
__device__ __forceinline__ bool interp8(const float2 (&s)[8], float mu, const float *mmse,
                                        float2 *v)
{
    const int imu = (int)rintf(mu * 128.0f);
    if (imu < 0 || imu > 128)
        return false;
    const float *row = mmse + imu * 8;
    float pr[4], pi[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float t0 = row[7 - j], t1 = row[3 - j];
        pr[j] = __fmaf_rn(s[j + 4].x, t1, s[j].x * t0);
        pi[j] = __fmaf_rn(s[j + 4].y, t1, s[j].y * t0);
    }
    v->x = (pr[0] + pr[1]) + (pr[2] + pr[3]);
    v->y = (pi[0] + pi[1]) + (pi[2] + pi[3]);
    return true;
}

constexpr int kMskChunk = 64;             // samples per prefetch chunk
constexpr int kMskRing = 2 * kMskChunk;     // ring of two chunks per channel
constexpr int kMskPitch = kMskRing + 2;     // float2 per ring row (keeps rows 16-byte aligned)
constexpr int kMskLookahead = 40;           // issue the next chunk this many samples early

__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem, int src_bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem, int src_bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// One warp = 32 channels.  Lane l runs channel (warp*32 + l)'s loop; the input samples of
// all 32 channels are staged in a shared-memory ring (two 64-sample chunks per channel)
// that the warp fills cooperatively with cp.async, one coalesced 512-byte chunk per channel,
// issued kMskLookahead samples before the first lane needs it.
template <bool kTail>
__global__ void __launch_bounds__(32)
k_msk(const float2 *__restrict__ in, size_t in_stride, int channels,
      int noutput_items, const int *__restrict__ ninput_dev, int ninput_const,
      uint64_t nitems_read, const b200ais_tag *__restrict__ tags, int max_tags,
      const int *__restrict__ ntags, MskParams p, MskState *__restrict__ state,
      const float *__restrict__ g_mmse, const float *__restrict__ g_atan,
      float2 *__restrict__ out, float *__restrict__ out_err,
      float *__restrict__ out_mu, float *__restrict__ out_soft,
      uint8_t *__restrict__ bits, size_t out_stride, int *__restrict__ nproduced,
      int *__restrict__ nconsumed, int require_unbounded, int *__restrict__ status)
{
    __shared__ float s_mmse[129 * 8];
    __shared__ float s_atan[257];
    __shared__ __align__(16) float2 ring[32 * kMskPitch];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    for (int i = lane; i < 129 * 8; i += 32)
        s_mmse[i] = g_mmse[i];
    for (int i = lane; i < 257; i += 32)
        s_atan[i] = g_atan[i];
    const int c0 = blockIdx.x * 32;
    const int c = c0 + lane;
    const bool live = c < channels;
    const int cc = live ? c : channels - 1; // dead lanes shadow the last channel, never store

    const float2 *xin = in + (size_t)cc * in_stride;
    MskState st = state[cc];
    const int ninput_items = ninput_dev ? ninput_dev[0] : ninput_const;
    int oidx = 0, iidx = 0;
    const int ninp = (int)((double)ninput_items - 3.0 * (double)p.sps_half); // :119

    // ring preset: slots of the (virtual) chunk -1 are zero, in[-1] is the carried sample
    float2 *my = ring + lane * kMskPitch;
    for (int k = kMskChunk; k < kMskRing; k++)
        my[k] = make_float2(0.0f, 0.0f);
    my[kMskRing - 1] = make_float2(st.prev_re, st.prev_im);
    __syncwarp();

    // time_est tags inside [read, read+ninp), in offset order (:125-130)
    const b200ais_tag *tg = tags ? tags + (size_t)cc * max_tags : nullptr;
    const int nt = (tags && ntags && ninp > 0) ? min(ntags[cc], max_tags) : 0;
    auto next_tag = [&](int from) {
        int k = from;
        while (k < nt) {
            const b200ais_tag t = tg[k];
            if (t.key == B200AIS_TAG_TIME_EST && t.port == 0 && t.offset >= nitems_read &&
                t.offset < nitems_read + (uint64_t)ninp)
                break;
            k++;
        }
        return k;
    };
    int thead = next_tag(0);
    int tag_off = 0x7fffffff; // offset of the pending tag relative to the read pointer
    float tag_val = 0.0f;
    auto fetch_tag = [&]() {
        if (thead < nt) {
            const b200ais_tag t = tg[thead];
            tag_off = (int)(t.offset - nitems_read);
            tag_val = (float)t.value;
        } else {
            tag_off = 0x7fffffff;
        }
    };
    fetch_tag();

    // demod-tail state (fresh per call: the chain processes one record per call)
    float2 qprev = make_float2(0.0f, 0.0f);
    unsigned bprev = 0;
    const float qgain = 1.57079632679489661923f; // (float)(pi/2), python/ais_demod.py:48
    unsigned pack = 0;

    float2 *oc = out ? out + (size_t)cc * out_stride : nullptr;
    float *oe = out_err ? out_err + (size_t)cc * out_stride : nullptr;
    float *om = out_mu ? out_mu + (size_t)cc * out_stride : nullptr;
    float *os = out_soft ? out_soft + (size_t)cc * out_stride : nullptr;
    uint8_t *ob = bits ? bits + (size_t)cc * out_stride : nullptr;
    const bool word_ok = kTail && ((out_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(bits) & 3) == 0);

    // 16-byte cp.async needs the even samples of a row on 16-byte boundaries
    const bool row16 = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((in_stride & 1) == 0);
    const float2 *row0 = in + (size_t)c0 * in_stride;

    int next_chunk = 0;     // next chunk index to issue for this lane's channel
    bool inflight = false;  // warp-uniform: a cp.async group may still be pending
    int err_code = 0;
    bool active = live && ninp > 0 && noutput_items > 0;

    while (__any_sync(FULL, active)) {
        // (1) tag reset (:139-164)
        if (active && tag_off != 0x7fffffff) {
            if ((tag_off >= iidx) && ((float)tag_off < ((float)iidx + p.sps_half))) {
                if (tag_val != tag_val) {
                    thead = next_tag(thead + 1); // NaN: drop the tag, no reset (:144-147)
                } else {
                    st.mu = tag_val;
                    iidx = tag_off;
                    if (st.mu < 0) {
                        st.mu = st.mu + 1.0f;
                        iidx--;
                    }
                    st.div = 0;
                    st.omega = p.sps_half;
                    st.dly2_re = st.dly1_re;
                    st.dly2_im = st.dly1_im;
                    thead = next_tag(thead + 1);
                }
                fetch_tag();
            }
        }
        // (2) keep the ring ahead of every lane
        const int hi = iidx + 7;
        for (;;) {
            const bool want = active && ((hi + kMskLookahead) >> 6) >= next_chunk &&
                              (next_chunk << 6) < ninput_items;
            const unsigned wm = __ballot_sync(FULL, want);
            if (!wm)
                break;
            for (int ch = 0; ch < 32; ch++) {
                if (!((wm >> ch) & 1u))
                    continue;
                const int j = __shfl_sync(FULL, next_chunk, ch);
                const int s0 = (j << 6) + 2 * lane; // first of this lane's two samples
                const float2 *src = row0 + (size_t)ch * in_stride + s0;
                float2 *dst = ring + ch * kMskPitch + ((j & 1) << 6) + 2 * lane;
                int nb = (ninput_items - s0) * 8;
                nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
                if (row16) {
                    cp_async_16(dst, nb ? src : row0, nb);
                } else {
                    cp_async_8(dst, nb ? src : row0, nb >= 8 ? 8 : 0);
                    cp_async_8(dst + 1, nb > 8 ? src + 1 : row0, nb > 8 ? 8 : 0);
                }
            }
            cp_async_commit();
            inflight = true;
            if (want)
                next_chunk++;
        }
        if (inflight && __any_sync(FULL, active && (hi >> 6) >= next_chunk - 1)) {
            cp_async_wait_all();
            __syncwarp();
            inflight = false;
        }
        // (3) one half-symbol step
        if (active) {
            float2 s8[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
                s8[k] = my[(iidx + k) & (kMskRing - 1)];
            float2 v;
            if (!interp8(s8, st.mu, s_mmse, &v)) {
                err_code = B200AIS_E_INTERP;
                active = false;
                continue;
            }
            // std::complex arithmetic as GCC emits it: (ac - bd, ad + bc), no contraction
            const float sq_re = v.x * v.x - v.y * v.y, sq_im = v.x * v.y + v.y * v.x;
            const float d_re = st.dly2_re * st.dly2_re - st.dly2_im * st.dly2_im;
            const float d_im = -(st.dly2_re * st.dly2_im + st.dly2_im * st.dly2_re);
            const float nl_re = sq_re * d_re - sq_im * d_im;
            const float nl_im = sq_re * d_im + sq_im * d_re;
            float err_out = nl_re - st.diff1_re;
            if (st.div & 1) {
                err_out = branchless_clip(err_out, 3.0f);
                st.omega = st.omega + p.gain_omega * err_out;
                st.omega = p.sps_half + branchless_clip(st.omega - p.sps_half, p.limit);
                st.mu = st.mu + p.gain * err_out;
            }
            if (!(st.div & 1) || p.osps == 2) {
                if (oc)
                    oc[oidx] = v;
                if (oe)
                    oe[oidx] = err_out;
                if (om)
                    om[oidx] = st.mu;
                if (kTail) {
                    // quadrature_demod_cf: x[n]*conj(x[n-1]), VOLK multiply-conjugate FMA form
                    const float re = __fmaf_rn(v.x, qprev.x, v.y * qprev.y);
                    const float im = __fmaf_rn(v.y, qprev.x, -(v.x * qprev.y));
                    const float soft = qgain * fast_atan2f_tab(im, re, s_atan);
                    qprev = v;
                    const unsigned b = soft >= 0 ? 1u : 0u;   // binary_slicer_fb
                    const unsigned d = (b - bprev) % 2u;      // diff_decoder_bb(2)
                    bprev = b;
                    if (os)
                        os[oidx] = soft;
                    const unsigned bit = (d ^ 0x01u) & 0x01u; // lib/invert_impl.cc:63
                    if (word_ok) {
                        pack |= bit << (8 * (oidx & 3));
                        if ((oidx & 3) == 3) {
                            *reinterpret_cast<unsigned *>(ob + (oidx & ~3)) = pack;
                            pack = 0;
                        }
                    } else {
                        ob[oidx] = (uint8_t)bit;
                    }
                }
                oidx++;
            }
            st.div++;
            st.dly1_re = v.x;
            st.dly1_im = v.y;
            st.dly2_re = v.x;
            st.dly2_im = v.y;
            st.diff1_re = nl_re;
            st.diff1_im = nl_im;
            st.mu = st.mu + st.omega;
            const float fl = floorf(st.mu);
            iidx += (int)fl;
            st.mu = st.mu - fl;
            active = (oidx < noutput_items) && (iidx < ninp);
        }
    }
    cp_async_wait_all();
    if (!live)
        return;
    if (kTail && word_ok && (oidx & 3)) { // flush the partial word byte by byte
        for (int k = oidx & ~3; k < oidx; k++)
            ob[k] = (uint8_t)((pack >> (8 * (k & 3))) & 0xffu);
    }
    if (ninp > 0 && iidx > 0) {
        const float2 pv = xin[iidx - 1];
        st.prev_re = pv.x;
        st.prev_im = pv.y;
    }
    if (!err_code && require_unbounded && ninp > 0 && oidx >= noutput_items && iidx < ninp)
        err_code = B200AIS_E_OUT_OVERFLOW;
    if (err_code)
        atomicMin(status, err_code);
    if (ninp > 0)
        state[c] = st;
    nproduced[c] = oidx;
    nconsumed[c] = ninp > 0 ? iidx : 0;
}

__global__ void k_msk_reset(MskState *state, int channels, float sps_half)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    MskState st;
    st.mu = 0.5f; // lib/msk_timing_recovery_cc_impl.cc:51-55
    st.omega = sps_half;
    st.dly1_re = st.dly1_im = st.dly2_re = st.dly2_im = st.diff1_re = st.diff1_im = 0.0f;
    st.div = 0;
    st.prev_re = st.prev_im = 0.0f;
    st.pad = 0;
    state[c] = st;
}

__global__ void k_msk_set_omega(MskState *state, int channels, float omega)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < channels)
        state[c].omega = omega;
}

// lib/invert_impl.cc:63: out[i] = (in[i] ^ 0x01) & 0x01, 16 items per thread where aligned
__global__ void k_invert(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n,
                         int vec_ok)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    if (vec_ok) {
        const size_t nv = n / 16;
        const uint4 *vi = reinterpret_cast<const uint4 *>(in);
        uint4 *vo = reinterpret_cast<uint4 *>(out);
        for (size_t i = tid; i < nv; i += nthreads) {
            uint4 v = vi[i];
            v.x = (v.x ^ 0x01010101u) & 0x01010101u;
            v.y = (v.y ^ 0x01010101u) & 0x01010101u;
            v.z = (v.z ^ 0x01010101u) & 0x01010101u;
            v.w = (v.w ^ 0x01010101u) & 0x01010101u;
            vo[i] = v;
        }
        for (size_t i = nv * 16 + tid; i < n; i += nthreads)
            out[i] = (in[i] ^ 0x01) & 0x01;
    } else {
        for (size_t i = tid; i < n; i += nthreads)
            out[i] = (in[i] ^ 0x01) & 0x01;
    }
}

} // namespace

int launch_msk(const float2 *in, size_t in_stride, int channels, int noutput_items,
               const int *ninput_items_dev, int ninput_items_const, uint64_t nitems_read,
               const b200ais_tag *tags, int max_tags, const int *ntags, MskParams p, MskState *state,
               float2 *out, float *out_err, float *out_mu, float *out_soft, uint8_t *bits,
               size_t out_stride, int *nproduced, int *nconsumed, int require_unbounded,
               int *status, cudaStream_t s)
{
    if (channels <= 0)
        return B200AIS_OK;
    Tables tb;
    int rc = get_tables(&tb);
    if (rc)
        return rc;
    // serial per channel: one warp (32 channels) per block spreads the channels over the SMs
    const int threads = 32;
    const int blocks = (channels + threads - 1) / threads;
    if (bits)
        k_msk<true><<<blocks, threads, 0, s>>>(in, in_stride, channels, noutput_items,
                                               ninput_items_dev, ninput_items_const, nitems_read,
                                               tags, max_tags, ntags, p, state, tb.mmse, tb.atan,
                                               out, out_err, out_mu, out_soft, bits, out_stride,
                                               nproduced, nconsumed, require_unbounded, status);
    else
        k_msk<false><<<blocks, threads, 0, s>>>(in, in_stride, channels, noutput_items,
                                                ninput_items_dev, ninput_items_const, nitems_read,
                                                tags, max_tags, ntags, p, state, tb.mmse, tb.atan,
                                                out, out_err, out_mu, out_soft, bits, out_stride,
                                                nproduced, nconsumed, require_unbounded, status);
    B200_LAUNCH_CHECK("k_msk");
    return B200AIS_OK;
}

int launch_msk_reset(MskState *state, int channels, float sps_half, cudaStream_t s)
{
    if (channels <= 0)
        return B200AIS_OK;
    k_msk_reset<<<(channels + 127) / 128, 128, 0, s>>>(state, channels, sps_half);
    B200_LAUNCH_CHECK("k_msk_reset");
    return B200AIS_OK;
}

int launch_msk_set_omega(MskState *state, int channels, float omega, cudaStream_t s)
{
    if (channels <= 0)
        return B200AIS_OK;
    k_msk_set_omega<<<(channels + 127) / 128, 128, 0, s>>>(state, channels, omega);
    B200_LAUNCH_CHECK("k_msk_set_omega");
    return B200AIS_OK;
}

int launch_invert(const uint8_t *in, uint8_t *out, size_t n, cudaStream_t s)
{
    if (n == 0)
        return B200AIS_OK;
    const int vec_ok = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    size_t work = vec_ok ? (n + 15) / 16 : n;
    int blocks = (int)((work + 255) / 256);
    if (blocks > 148 * 16)
        blocks = 148 * 16;
    if (blocks < 1)
        blocks = 1;
    k_invert<<<blocks, 256, 0, s>>>(in, out, n, vec_ok);
    B200_LAUNCH_CHECK("k_invert");
    return B200AIS_OK;
}

} // namespace b200ais
