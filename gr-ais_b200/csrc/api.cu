// api.cu -- the C-ABI of libb200ais.so (include/b200ais.h): handle objects, host/device
// entry points, and the launch sequence of the fused ais_demod chain.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "internal.h"

using namespace b200ais;

namespace {

size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// kernel::fft_filter_ccc block size: fftsize = 2 * 2^ceil(log2 ntaps), nsamples = fftsize - ntaps + 1
int fft_filter_nsamples(int ntaps)
{
    int p = 1;
    while (p < ntaps)
        p <<= 1;
    return 2 * p - ntaps + 1;
}

int check_stream_status(cudaStream_t s, int *d_status)
{
    int st = 0;
    B200_CU(cudaMemcpyAsync(&st, d_status, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CU(cudaStreamSynchronize(s));
    if (st) {
        switch (st) {
        case B200AIS_E_TAG_OVERFLOW:
            set_error("a channel produced more tags than max_tags");
            break;
        case B200AIS_E_INTERP:
            set_error("mmse interpolator index out of [0,128] (non-finite input?)");
            break;
        case B200AIS_E_OUT_OVERFLOW:
            set_error("symbol output row too small for the record (raise max_bits)");
            break;
        default:
            set_error("kernel flagged status %d", st);
        }
    }
    return st;
}

} // namespace

// =========================================================== corr_est_cc

struct b200ais_corr_est {
    int channels = 0;
    int L = 0;
    float sps = 0;
    unsigned mark_delay = 0;
    float thresh = 0;
    int nsamples = 0;
    std::vector<float> taps; // stored FIR taps, interleaved (what symbols() returns)
    float2 *d_hbr = nullptr;  // transformed taps, bit-reversed order [fftsize]
    float2 *d_tail[2] = { nullptr, nullptr }; // fft_filter tail [channels][L-1], ping-pong
    int tail_cur = 0, tail_L = 0;
    int *d_status = nullptr;
    cudaStream_t stream = nullptr;
    DevBuf mask, corr, in, out0, out1, tags, ntags;
    std::mutex lock; // d_setlock (lib/corr_est_cc_impl.cc:135,169)
};

// kernel::fft_filter_ccc::set_taps: new transformed taps; the tail keeps its contents and is
// resized to ntaps - 1 (std::vector::resize)
static int corr_upload_taps(b200ais_corr_est *h)
{
    const int F = corr_fft_size(h->L);
    std::vector<float2> hbr((size_t)F);
    make_corr_spectrum(h->taps.data(), h->L, hbr.data());
    if (h->d_hbr)
        cudaFree(h->d_hbr);
    h->d_hbr = nullptr;
    B200_CU(cudaMalloc(&h->d_hbr, sizeof(float2) * hbr.size()));
    B200_CU(cudaMemcpy(h->d_hbr, hbr.data(), sizeof(float2) * hbr.size(), cudaMemcpyHostToDevice));
    const size_t tl = (size_t)std::max(h->L - 1, 1), C = (size_t)h->channels;
    float2 *nt[2] = { nullptr, nullptr };
    for (int k = 0; k < 2; k++) {
        B200_CU(cudaMalloc(&nt[k], sizeof(float2) * tl * C));
        B200_CU(cudaMemset(nt[k], 0, sizeof(float2) * tl * C));
    }
    if (h->d_tail[h->tail_cur] && h->tail_L > 1) {
        const size_t keep = (size_t)std::min(h->tail_L, h->L) - 1;
        if (keep)
            B200_CU(cudaMemcpy2D(nt[0], tl * sizeof(float2), h->d_tail[h->tail_cur],
                                 (size_t)(h->tail_L - 1) * sizeof(float2), keep * sizeof(float2), C,
                                 cudaMemcpyDeviceToDevice));
    }
    for (int k = 0; k < 2; k++) {
        if (h->d_tail[k])
            cudaFree(h->d_tail[k]);
        h->d_tail[k] = nt[k];
    }
    h->tail_cur = 0;
    h->tail_L = h->L;
    return B200AIS_OK;
}

extern "C" int b200ais_corr_est_create(b200ais_corr_est **out, const float *symbols_iq,
                                       int nsymbols, float sps, unsigned mark_delay,
                                       float threshold, int channels)
{
    if (!out || !symbols_iq || nsymbols < 5 || nsymbols > 2048 || channels < 1) {
        set_error("corr_est_create: need 5..2048 symbols and channels >= 1");
        return B200AIS_E_INVALID;
    }
    b200ais_corr_est *h = new (std::nothrow) b200ais_corr_est;
    if (!h)
        return B200AIS_E_NOMEM;
    h->channels = channels;
    h->L = nsymbols;
    h->sps = sps;
    // lib/corr_est_cc_impl.cc:59-63: conjugate, then reverse
    h->taps.resize(2 * (size_t)nsymbols);
    for (int i = 0; i < nsymbols; i++) {
        h->taps[2 * (nsymbols - 1 - i)] = symbols_iq[2 * i];
        h->taps[2 * (nsymbols - 1 - i) + 1] = -symbols_iq[2 * i + 1];
    }
    h->mark_delay = mark_delay >= (unsigned)nsymbols ? (unsigned)nsymbols - 1 : mark_delay; // :65-66
    float corr = 0; // :71-74, abs(z*conj(z)) = re*re + im*im
    for (int i = 0; i < nsymbols; i++) {
        volatile float re = h->taps[2 * i], im = h->taps[2 * i + 1];
        volatile float rr = re * re, ii = im * im;
        volatile float s = rr + ii;
        corr = corr + s;
    }
    volatile float tc = threshold * corr;
    h->thresh = tc * corr;
    h->nsamples = fft_filter_nsamples(nsymbols);
    int rc = corr_upload_taps(h);
    if (!rc) {
        cudaError_t e = cudaMalloc(&h->d_status, sizeof(int));
        if (e == cudaSuccess)
            e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess)
            rc = cuda_fail(e, "corr_est_create", __FILE__, __LINE__);
    }
    if (rc) {
        b200ais_corr_est_destroy(h);
        return rc;
    }
    *out = h;
    return B200AIS_OK;
}

extern "C" int b200ais_corr_est_destroy(b200ais_corr_est *h)
{
    if (!h)
        return B200AIS_OK;
    if (h->d_hbr)
        cudaFree(h->d_hbr);
    for (int k = 0; k < 2; k++)
        if (h->d_tail[k])
            cudaFree(h->d_tail[k]);
    if (h->d_status)
        cudaFree(h->d_status);
    if (h->stream)
        cudaStreamDestroy(h->stream);
    h->mask.release();
    h->corr.release();
    h->in.release();
    h->out0.release();
    h->out1.release();
    h->tags.release();
    h->ntags.release();
    delete h;
    return B200AIS_OK;
}

extern "C" int b200ais_corr_est_set_symbols(b200ais_corr_est *h, const float *symbols_iq,
                                            int nsymbols)
{
    if (!h || !symbols_iq || nsymbols < 5 || nsymbols > 2048) {
        set_error("set_symbols: need 5..2048 symbols");
        return B200AIS_E_INVALID;
    }
    std::lock_guard<std::mutex> lk(h->lock);
    // lib/corr_est_cc_impl.cc:137: d_symbols = symbols -- verbatim, threshold untouched
    h->taps.assign(symbols_iq, symbols_iq + 2 * (size_t)nsymbols);
    h->L = nsymbols;
    h->nsamples = fft_filter_nsamples(nsymbols);
    h->mark_delay = h->mark_delay >= (unsigned)nsymbols ? (unsigned)nsymbols - 1 : h->mark_delay;
    B200_CU(cudaStreamSynchronize(h->stream));
    return corr_upload_taps(h);
}

extern "C" int b200ais_corr_est_symbols(const b200ais_corr_est *h, float *out_iq, int cap, int *n)
{
    if (!h || !n) {
        set_error("symbols: null argument");
        return B200AIS_E_INVALID;
    }
    *n = h->L;
    if (out_iq) {
        if (cap < h->L) {
            set_error("symbols: buffer holds %d of %d items", cap, h->L);
            return B200AIS_E_INVALID;
        }
        memcpy(out_iq, h->taps.data(), sizeof(float) * 2 * (size_t)h->L);
    }
    return B200AIS_OK;
}

extern "C" int b200ais_corr_est_output_multiple(const b200ais_corr_est *h) { return h ? h->nsamples : 0; }
extern "C" int b200ais_corr_est_history(const b200ais_corr_est *h) { return h ? h->L + 1 : 0; }
extern "C" unsigned b200ais_corr_est_mark_delay(const b200ais_corr_est *h) { return h ? h->mark_delay : 0; }
extern "C" float b200ais_corr_est_threshold(const b200ais_corr_est *h) { return h ? h->thresh : 0.0f; }

extern "C" int b200ais_corr_est_work_dev(b200ais_corr_est *h, int noutput_items, const float *in,
                                         size_t in_stride, uint64_t nitems_written, float *out0,
                                         float *out1, size_t out_stride, b200ais_tag *tags,
                                         int max_tags, int *ntags, void *stream)
{
    if (!h || !in || !tags || !ntags || noutput_items < 0 || max_tags < 0 ||
        in_stride < (size_t)noutput_items + (size_t)h->L) {
        set_error("corr_est_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    std::lock_guard<std::mutex> lk(h->lock);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = noutput_items;
    if (n % h->nsamples) { // set_output_multiple(nsamples): the scheduler never does this
        set_error("corr_est_work: noutput_items (%d) is not a multiple of the output multiple (%d)", n,
                  h->nsamples);
        return B200AIS_E_INVALID;
    }
    const size_t ms = corr_mask_stride_bytes(h->L, n);
    int rc = h->mask.reserve(ms * (size_t)h->channels);
    if (rc)
        return rc;
    float2 *corr = reinterpret_cast<float2 *>(out1);
    size_t corr_stride = out_stride;
    if (!corr) { // the detector still needs the correlator stream
        corr_stride = round_up((size_t)std::max(n, 1), 2);
        if ((rc = h->corr.reserve(corr_stride * h->channels * sizeof(float2))))
            return rc;
        corr = h->corr.as<float2>();
    }
    const float2 *tw = nullptr;
    if ((rc = get_twiddles(corr_fft_size(h->L), &tw)))
        return rc;
    B200_CU(cudaMemsetAsync(h->d_status, 0, sizeof(int), s));
    const float2 *in2 = reinterpret_cast<const float2 *>(in);
    const float2 *in_eff = in2 + h->L; // &in[hist_len] (lib/corr_est_cc_impl.cc:188)
    rc = launch_corr_fft(in_eff, in_stride, h->channels, n, h->L, tw, h->d_hbr, h->thresh,
                         h->d_tail[h->tail_cur], h->d_tail[h->tail_cur ^ 1], h->mask.as<uint8_t>(), ms,
                         corr, corr_stride, out1 == nullptr, (int)(in_stride - (size_t)h->L), s);
    if (rc)
        return rc;
    if (n > 0)
        h->tail_cur ^= 1;
    const int isps = (int)(h->sps + 0.5f); // :193
    rc = launch_detect(corr, corr_stride, h->channels, n, n > 0 ? n : 1, 1, isps, h->mark_delay,
                       h->mask.as<uint8_t>(), ms, nitems_written, out1 != nullptr, tags, max_tags,
                       ntags, h->d_status, 0, s);
    if (rc)
        return rc;
    if (out0)
        rc = launch_copy_delay(in2, in_stride, reinterpret_cast<float2 *>(out0), out_stride,
                               h->channels, n, s);
    return rc;
}

extern "C" int b200ais_corr_est_work(b200ais_corr_est *h, int noutput_items, const float *in,
                                     size_t in_stride, uint64_t nitems_written, float *out0,
                                     float *out1, size_t out_stride, b200ais_tag *tags,
                                     int max_tags, int *ntags)
{
    if (!h || !in || !tags || !ntags || noutput_items < 0 || max_tags < 1) {
        set_error("corr_est_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    const int C = h->channels, n = noutput_items;
    const size_t row_in = round_up((size_t)n + h->L, 2), row_out = round_up((size_t)std::max(n, 1), 2);
    int rc;
    if ((rc = h->in.reserve(row_in * C * sizeof(float2))) ||
        (rc = h->out0.reserve(row_out * C * sizeof(float2))) ||
        (rc = h->out1.reserve(row_out * C * sizeof(float2))) ||
        (rc = h->tags.reserve((size_t)max_tags * C * sizeof(b200ais_tag))) ||
        (rc = h->ntags.reserve((size_t)C * sizeof(int))))
        return rc;
    cudaStream_t s = h->stream;
    B200_CU(cudaMemcpy2DAsync(h->in.p, row_in * sizeof(float2), in, in_stride * sizeof(float2),
                              ((size_t)n + h->L) * sizeof(float2), C, cudaMemcpyHostToDevice, s));
    rc = b200ais_corr_est_work_dev(h, n, h->in.as<float>(), row_in, nitems_written,
                                   out0 ? h->out0.as<float>() : nullptr,
                                   out1 ? h->out1.as<float>() : nullptr, row_out,
                                   h->tags.as<b200ais_tag>(), max_tags, h->ntags.as<int>(), s);
    if (rc)
        return rc;
    if (out0 && n > 0)
        B200_CU(cudaMemcpy2DAsync(out0, out_stride * sizeof(float2), h->out0.p,
                                  row_out * sizeof(float2), (size_t)n * sizeof(float2), C,
                                  cudaMemcpyDeviceToHost, s));
    if (out1 && n > 0)
        B200_CU(cudaMemcpy2DAsync(out1, out_stride * sizeof(float2), h->out1.p,
                                  row_out * sizeof(float2), (size_t)n * sizeof(float2), C,
                                  cudaMemcpyDeviceToHost, s));
    B200_CU(cudaMemcpyAsync(tags, h->tags.p, (size_t)max_tags * C * sizeof(b200ais_tag),
                            cudaMemcpyDeviceToHost, s));
    B200_CU(cudaMemcpyAsync(ntags, h->ntags.p, (size_t)C * sizeof(int), cudaMemcpyDeviceToHost, s));
    return check_stream_status(s, h->d_status);
}

// =============================================== msk_timing_recovery_cc

struct b200ais_msk {
    int channels = 0;
    float sps_arg = 0; // the sps handed to make()/set_sps()
    MskParams p{};
    MskState *d_state = nullptr;
    int *d_status = nullptr;
    cudaStream_t stream = nullptr;
    DevBuf in, tags, ntags, out, err, mu, nprod, ncons;
};

extern "C" int b200ais_msk_create(b200ais_msk **out, float sps, float gain, float limit, int osps,
                                  int channels)
{
    if (!out || channels < 1) {
        set_error("msk_create: bad arguments");
        return B200AIS_E_INVALID;
    }
    // lib/msk_timing_recovery_cc_impl.cc:45-62: set_sps, then set_gain (throws), then the osps check
    if (!(gain > 0)) {
        set_error("Gain must be positive");
        return B200AIS_E_RANGE;
    }
    if (osps != 1 && osps != 2) {
        set_error("osps must be 1 or 2");
        return B200AIS_E_RANGE;
    }
    b200ais_msk *h = new (std::nothrow) b200ais_msk;
    if (!h)
        return B200AIS_E_NOMEM;
    h->channels = channels;
    h->sps_arg = sps;
    h->p.sps_half = (float)((double)sps / 2.0);
    h->p.gain = gain;
    h->p.gain_omega = (float)((double)(gain * gain) * 0.25);
    h->p.limit = limit;
    h->p.osps = osps;
    cudaError_t e = cudaMalloc(&h->d_state, sizeof(MskState) * (size_t)channels);
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_status, sizeof(int));
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    int rc = e == cudaSuccess ? B200AIS_OK : cuda_fail(e, "msk_create", __FILE__, __LINE__);
    if (!rc)
        rc = launch_msk_reset(h->d_state, channels, h->p.sps_half, h->stream);
    if (!rc) {
        e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess)
            rc = cuda_fail(e, "msk_create sync", __FILE__, __LINE__);
    }
    if (rc) {
        b200ais_msk_destroy(h);
        return rc;
    }
    *out = h;
    return B200AIS_OK;
}

extern "C" int b200ais_msk_destroy(b200ais_msk *h)
{
    if (!h)
        return B200AIS_OK;
    if (h->d_state)
        cudaFree(h->d_state);
    if (h->d_status)
        cudaFree(h->d_status);
    if (h->stream)
        cudaStreamDestroy(h->stream);
    h->in.release();
    h->tags.release();
    h->ntags.release();
    h->out.release();
    h->err.release();
    h->mu.release();
    h->nprod.release();
    h->ncons.release();
    delete h;
    return B200AIS_OK;
}

extern "C" int b200ais_msk_set_gain(b200ais_msk *h, float gain)
{
    if (!h)
        return B200AIS_E_INVALID;
    h->p.gain = gain; // the reference stores first, then throws (:81-82)
    if (!(gain > 0)) {
        set_error("Gain must be positive");
        return B200AIS_E_RANGE;
    }
    h->p.gain_omega = (float)((double)(gain * gain) * 0.25);
    return B200AIS_OK;
}
extern "C" float b200ais_msk_get_gain(const b200ais_msk *h) { return h ? h->p.gain : 0.0f; }
extern "C" int b200ais_msk_set_limit(b200ais_msk *h, float limit)
{
    if (!h)
        return B200AIS_E_INVALID;
    h->p.limit = limit;
    return B200AIS_OK;
}
extern "C" float b200ais_msk_get_limit(const b200ais_msk *h) { return h ? h->p.limit : 0.0f; }
extern "C" float b200ais_msk_get_sps(const b200ais_msk *h) { return h ? h->p.sps_half : 0.0f; }

extern "C" int b200ais_msk_forecast(const b200ais_msk *h, int noutput_items)
{
    if (!h)
        return 0;
    // :103: (int)ceil((noutput_items*d_sps*2) + 3.0*d_sps + ntaps)
    volatile float a = (float)noutput_items * h->p.sps_half;
    volatile float b = a * 2.0f;
    return (int)std::ceil((double)b + 3.0 * (double)h->p.sps_half + 8.0);
}

extern "C" int b200ais_msk_reset(b200ais_msk *h)
{
    if (!h)
        return B200AIS_E_INVALID;
    int rc = launch_msk_reset(h->d_state, h->channels, h->p.sps_half, h->stream);
    if (rc)
        return rc;
    B200_CU(cudaStreamSynchronize(h->stream));
    return B200AIS_OK;
}

extern "C" int b200ais_msk_set_sps(b200ais_msk *h, float sps)
{
    if (!h)
        return B200AIS_E_INVALID;
    // :69-74: d_sps = sps/2; d_omega = d_sps (every channel's loop restarts at nominal rate)
    h->sps_arg = sps;
    h->p.sps_half = (float)((double)sps / 2.0);
    int rc = launch_msk_set_omega(h->d_state, h->channels, h->p.sps_half, h->stream);
    if (rc)
        return rc;
    B200_CU(cudaStreamSynchronize(h->stream));
    return B200AIS_OK;
}

extern "C" int b200ais_msk_general_work_dev(b200ais_msk *h, int noutput_items, int ninput_items,
                                            const float *in, size_t in_stride, uint64_t nitems_read,
                                            const b200ais_tag *tags, int max_tags, const int *ntags,
                                            float *out, float *out_err, float *out_mu,
                                            size_t out_stride, int *nproduced, int *nconsumed,
                                            void *stream)
{
    if (!h || !in || !out || !nproduced || !nconsumed || noutput_items < 0 || ninput_items < 0 ||
        out_stride < (size_t)noutput_items) {
        set_error("msk_general_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200_CU(cudaMemsetAsync(h->d_status, 0, sizeof(int), s));
    return launch_msk(reinterpret_cast<const float2 *>(in), in_stride, h->channels, noutput_items,
                      ninput_items, nitems_read, tags, max_tags, ntags, h->p, h->d_state,
                      reinterpret_cast<float2 *>(out), out_err, out_mu, out_stride, nproduced,
                      nconsumed, 0, h->d_status, nullptr, s);
}

extern "C" int b200ais_msk_general_work(b200ais_msk *h, int noutput_items, int ninput_items,
                                        const float *in, size_t in_stride, uint64_t nitems_read,
                                        const b200ais_tag *tags, int max_tags, const int *ntags,
                                        float *out, float *out_err, float *out_mu, size_t out_stride,
                                        int *nproduced, int *nconsumed)
{
    if (!h || !in || !out || !nproduced || !nconsumed || noutput_items < 0 || ninput_items < 0) {
        set_error("msk_general_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    const int C = h->channels;
    const size_t row_in = round_up((size_t)std::max(ninput_items, 1), 2);
    const size_t row_out = round_up((size_t)std::max(noutput_items, 1), 2);
    const int mt = (tags && ntags) ? std::max(max_tags, 0) : 0;
    int rc;
    if ((rc = h->in.reserve(row_in * C * sizeof(float2))) ||
        (rc = h->tags.reserve((size_t)std::max(mt, 1) * C * sizeof(b200ais_tag))) ||
        (rc = h->ntags.reserve((size_t)C * sizeof(int))) ||
        (rc = h->out.reserve(row_out * C * sizeof(float2))) ||
        (rc = h->err.reserve(row_out * C * sizeof(float))) ||
        (rc = h->mu.reserve(row_out * C * sizeof(float))) ||
        (rc = h->nprod.reserve((size_t)C * sizeof(int))) ||
        (rc = h->ncons.reserve((size_t)C * sizeof(int))))
        return rc;
    cudaStream_t s = h->stream;
    if (ninput_items > 0)
        B200_CU(cudaMemcpy2DAsync(h->in.p, row_in * sizeof(float2), in, in_stride * sizeof(float2),
                                  (size_t)ninput_items * sizeof(float2), C, cudaMemcpyHostToDevice, s));
    std::vector<b200ais_tag> sorted;
    if (mt > 0) {
        // get_tags_in_range hands the block its tags in offset order; keep insertion order on ties
        sorted.assign(tags, tags + (size_t)mt * C);
        for (int c = 0; c < C; c++) {
            int k = std::min(std::max(ntags[c], 0), mt);
            std::stable_sort(sorted.begin() + (size_t)c * mt, sorted.begin() + (size_t)c * mt + k,
                             [](const b200ais_tag &a, const b200ais_tag &b) { return a.offset < b.offset; });
        }
        B200_CU(cudaMemcpyAsync(h->tags.p, sorted.data(), sorted.size() * sizeof(b200ais_tag),
                                cudaMemcpyHostToDevice, s));
        B200_CU(cudaMemcpyAsync(h->ntags.p, ntags, (size_t)C * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    rc = b200ais_msk_general_work_dev(h, noutput_items, ninput_items, h->in.as<float>(), row_in,
                                      nitems_read, mt > 0 ? h->tags.as<b200ais_tag>() : nullptr, mt,
                                      mt > 0 ? h->ntags.as<int>() : nullptr, h->out.as<float>(),
                                      h->err.as<float>(), h->mu.as<float>(), row_out,
                                      h->nprod.as<int>(), h->ncons.as<int>(), s);
    if (rc)
        return rc;
    if (noutput_items > 0) {
        B200_CU(cudaMemcpy2DAsync(out, out_stride * sizeof(float2), h->out.p, row_out * sizeof(float2),
                                  (size_t)noutput_items * sizeof(float2), C, cudaMemcpyDeviceToHost, s));
        if (out_err)
            B200_CU(cudaMemcpy2DAsync(out_err, out_stride * sizeof(float), h->err.p,
                                      row_out * sizeof(float), (size_t)noutput_items * sizeof(float),
                                      C, cudaMemcpyDeviceToHost, s));
        if (out_mu)
            B200_CU(cudaMemcpy2DAsync(out_mu, out_stride * sizeof(float), h->mu.p,
                                      row_out * sizeof(float), (size_t)noutput_items * sizeof(float),
                                      C, cudaMemcpyDeviceToHost, s));
    }
    B200_CU(cudaMemcpyAsync(nproduced, h->nprod.p, (size_t)C * sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CU(cudaMemcpyAsync(nconsumed, h->ncons.p, (size_t)C * sizeof(int), cudaMemcpyDeviceToHost, s));
    return check_stream_status(s, h->d_status);
}

// ================================================================ freqest

struct b200ais_freqest {
    int channels = 0;
    int fftlen = 0;
    int offset = 0;
    float binsize = 0;
    cudaStream_t stream = nullptr;
    DevBuf spec, raw, out;
};

extern "C" int b200ais_freqest_create(b200ais_freqest **out, float sample_rate, int data_rate,
                                      int fftlen, int channels)
{
    if (!out || fftlen < 2 || channels < 1 || !(sample_rate > 0)) {
        set_error("freqest_create: bad arguments");
        return B200AIS_E_INVALID;
    }
    b200ais_freqest *h = new (std::nothrow) b200ais_freqest;
    if (!h)
        return B200AIS_E_NOMEM;
    h->channels = channels;
    h->fftlen = fftlen;
    // lib/freqest_impl.cc:46-47
    volatile float ratio = (float)data_rate / sample_rate;
    volatile float off = (float)fftlen * ratio;
    h->offset = (int)off;
    h->binsize = sample_rate / (float)fftlen;
    if (h->offset < 0 || h->offset >= fftlen) {
        delete h;
        set_error("freqest_create: data_rate/sample_rate puts the bin offset outside the FFT");
        return B200AIS_E_INVALID;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete h;
        return cuda_fail(e, "freqest_create", __FILE__, __LINE__);
    }
    *out = h;
    return B200AIS_OK;
}

extern "C" int b200ais_freqest_destroy(b200ais_freqest *h)
{
    if (!h)
        return B200AIS_OK;
    if (h->stream)
        cudaStreamDestroy(h->stream);
    h->spec.release();
    h->raw.release();
    h->out.release();
    delete h;
    return B200AIS_OK;
}

extern "C" int b200ais_freqest_work_dev(b200ais_freqest *h, int noutput_items, const float *spec,
                                        float *out, void *stream)
{
    if (!h || !spec || !out || noutput_items < 0) {
        set_error("freqest_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = h->raw.reserve((size_t)std::max(noutput_items, 1) * h->channels * sizeof(int));
    if (rc)
        return rc;
    rc = launch_freqest_spec(reinterpret_cast<const float2 *>(spec), h->channels, noutput_items,
                             h->fftlen, h->offset, h->raw.as<int>(), s);
    if (rc)
        return rc;
    return launch_freqest_resolve(h->raw.as<int>(), h->channels, noutput_items, h->fftlen,
                                  h->binsize, out, s);
}

extern "C" int b200ais_freqest_work(b200ais_freqest *h, int noutput_items, const float *spec,
                                    float *out)
{
    if (!h || !spec || !out || noutput_items < 0) {
        set_error("freqest_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    if (noutput_items == 0)
        return B200AIS_OK;
    const size_t items = (size_t)noutput_items * h->fftlen * h->channels;
    int rc;
    if ((rc = h->spec.reserve(items * sizeof(float2))) ||
        (rc = h->out.reserve((size_t)noutput_items * h->channels * sizeof(float))))
        return rc;
    cudaStream_t s = h->stream;
    B200_CU(cudaMemcpyAsync(h->spec.p, spec, items * sizeof(float2), cudaMemcpyHostToDevice, s));
    rc = b200ais_freqest_work_dev(h, noutput_items, h->spec.as<float>(), h->out.as<float>(), s);
    if (rc)
        return rc;
    B200_CU(cudaMemcpyAsync(out, h->out.p, (size_t)noutput_items * h->channels * sizeof(float),
                            cudaMemcpyDeviceToHost, s));
    B200_CU(cudaStreamSynchronize(s));
    return B200AIS_OK;
}

// ================================================================= invert

extern "C" int b200ais_invert_work_dev(const uint8_t *in, uint8_t *out, size_t nitems, void *stream)
{
    if ((!in || !out) && nitems) {
        set_error("invert_work: null buffer");
        return B200AIS_E_INVALID;
    }
    return launch_invert(in, out, nitems, static_cast<cudaStream_t>(stream));
}

extern "C" int b200ais_invert_work(const uint8_t *in, uint8_t *out, size_t nitems)
{
    if ((!in || !out) && nitems) {
        set_error("invert_work: null buffer");
        return B200AIS_E_INVALID;
    }
    if (!nitems)
        return B200AIS_OK;
    uint8_t *d = nullptr;
    B200_CU(cudaMalloc(&d, nitems));
    int rc = B200AIS_OK;
    cudaError_t e = cudaMemcpy(d, in, nitems, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = launch_invert(d, d, nitems, nullptr);
        if (!rc)
            e = cudaMemcpy(out, d, nitems, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (e != cudaSuccess)
        return cuda_fail(e, "invert_work", __FILE__, __LINE__);
    return rc;
}

// =============================================== the fused ais_demod chain

namespace {
constexpr int kKeep = 32;     // stream mode: corr_est output-0 items kept for the timing loop
constexpr int kSeg = 16;      // NCO phase checkpoint spacing (samples)
constexpr int kMaxGroups = 8; // channel groups pipelined over internal streams
constexpr long kPipeMaxChannels = 53248; // larger batches run enqueue_dev in stream order (see there)
} // namespace

struct b200ais_demod {
    b200ais_demod_config cfg{};
    int channels = 0, max_samples = 0, max_tags = 0;
    int L = 0, nsamples = 0, chunk = 0, isps = 0;
    unsigned mark_delay = 0;
    float thresh = 0;
    int offset = 0;
    float binsize = 0, sens = 0;
    MskParams mp{};
    int HP = 0;          // zero history items in front of every corr_est input row
    size_t a_stride = 0; // items per row of the corr_est input stream
    size_t mask_stride = 0;
    int nvec_max = 0;
    float2 *d_taps_time = nullptr; // transformed taps (bit-reversed order) [fftsize]
    float2 *d_corr = nullptr;      // correlator stream [channels][corr_stride]
    size_t corr_stride = 0;
    float2 *d_x = nullptr; // staging for the host variant [channels][max_samples]
    int16_t *d_x16 = nullptr; // staging of interleaved int16 I/Q (b200ais_demod_work_sc16)
    float2 *d_a = nullptr;
    uint8_t *d_mask = nullptr;
    int *d_raw = nullptr;
    float *d_fhat = nullptr, *d_ckpt = nullptr;
    b200ais_tag *d_tags = nullptr;
    int *d_ntags = nullptr, *d_nbits = nullptr, *d_ncons = nullptr;
    MskState *d_state = nullptr;
    uint8_t *d_bits = nullptr;
    size_t bits_cap = 0;
    int *d_status = nullptr; // [kMaxGroups + 1]
    cudaStream_t streams[kMaxGroups] = {};
    cudaEvent_t done[kMaxGroups] = {};
    cudaEvent_t ev_fork = nullptr;
    int overlap_groups = 1; // channel groups a *_dev call forks over internal streams (1 = none)
    bool taps_enabled = false;
    DevBuf t_sym, t_err, t_mu, t_soft, t_otags;
    bool profiling = false;
    std::vector<std::vector<cudaEvent_t>> ev_used, ev_free; // 7 events per profiled call
    double stage_ms[B200AIS_STAGE_T_COUNT] = { 0, 0, 0, 0, 0, 0, 0 };
    int prof_calls = 0;
    int last_n = 0, last_max_bits = 0, last_n1 = 0;
    // ---- pipelined submission (b200ais_demod_enqueue_dev): the timing loop + bit tail of call k
    // run on a high-priority side stream under the front half of call k+1 ----
    cudaStream_t back_stream = nullptr;
    float2 *d_a_alt = nullptr;        // second set of corr_est input rows (set 1)
    b200ais_tag *d_tags_alt = nullptr;
    int *d_ntags_alt = nullptr;
    cudaEvent_t ev_front[2] = { nullptr, nullptr }, ev_back[2] = { nullptr, nullptr };
    bool back_pending[2] = { false, false };
    bool inorder_pending = false; // records enqueued in stream order (full-wave batches) since the last join
    unsigned pipe_calls = 0;
    // ---- stream mode (b200ais_demod_stream_*): what every block keeps between calls ----
    bool st_ready = false;     // state allocated and reset
    bool pad_dirty = false;    // stream calls used the rows' history pads: batch calls re-zero them
    int st_nx = 0;             // input items waiting for a whole FFT vector (same for all channels)
    int st_na = 0;             // AGC outputs corr_est has not taken yet
    uint64_t st_written = 0;   // corr_est nitems_written
    float2 *d_xcarry = nullptr;            // [channels][fftlen]
    DevBuf xs;                             // assembled input rows [channels][xs_stride]
    size_t xs_stride = 0;
    float *d_phase = nullptr;              // NCO phase [channels]
    float2 *d_yhist[2] = { nullptr, nullptr }; // mixed AGC history [channels][511], ping-pong
    float2 *d_ctail[2] = { nullptr, nullptr }; // overlap-add tail [channels][L-1], ping-pong
    int yh_cur = 0, ct_cur = 0;
    int *d_unc = nullptr, *d_nold = nullptr;
    TailCarry *d_tcarry = nullptr;
};

extern "C" int b200ais_demod_default_config(b200ais_demod_config *cfg)
{
    if (!cfg)
        return B200AIS_E_INVALID;
    cfg->sample_rate = 48000.0f; // python/radio.py:47-48,62
    cfg->data_rate = 9600;
    cfg->fftlen = 1024;          // python/radio.py:61
    cfg->agc_nsamples = 512;     // python/ais_demod.py:35
    cfg->agc_reference = 2.0f;
    cfg->sps = 5.0f;
    cfg->mark_delay = 1;         // python/ais_demod.py:41
    cfg->threshold = 0.9f;       // python/ais_demod.py:42
    cfg->gain = 0.04f;           // python/radio.py:58
    cfg->limit = 0.01f;          // python/radio.py:59
    cfg->osps = 1;
    cfg->corr_chunk = 0;
    cfg->stages = B200AIS_STAGE_FREQSYNC | B200AIS_STAGE_AGC;
    return B200AIS_OK;
}

static int demod_max_bits(const b200ais_demod *h, int nsamples)
{
    // a multiple of 4: word-aligned bit rows let the timing loop write the bits itself
    return ((int)((double)nsamples / (double)h->cfg.sps * (double)h->cfg.osps * 1.05) + 64 + 3) & ~3;
}

extern "C" int b200ais_demod_max_bits(const b200ais_demod *h, int nsamples)
{
    return h ? demod_max_bits(h, nsamples) : 0;
}

extern "C" int b200ais_demod_create(b200ais_demod **out, const b200ais_demod_config *cfg,
                                    const float *symbols_iq, int nsymbols, int channels,
                                    int max_samples, int max_tags)
{
    if (!out || !cfg || !symbols_iq || nsymbols < 5 || nsymbols > 2048 || channels < 1 ||
        max_samples < 1 || max_tags < 4) {
        set_error("demod_create: bad arguments");
        return B200AIS_E_INVALID;
    }
    if (!(cfg->gain > 0)) {
        set_error("Gain must be positive");
        return B200AIS_E_RANGE;
    }
    if (cfg->osps != 1 && cfg->osps != 2) {
        set_error("osps must be 1 or 2");
        return B200AIS_E_RANGE;
    }
    if ((cfg->stages & B200AIS_STAGE_FREQSYNC) &&
        (cfg->fftlen < 16 || cfg->fftlen > 4096 || (cfg->fftlen & (cfg->fftlen - 1)) ||
         cfg->fftlen % kSeg)) {
        set_error("demod_create: fftlen must be a power of two in [16, 4096]");
        return B200AIS_E_INVALID;
    }
    b200ais_demod *h = new (std::nothrow) b200ais_demod;
    if (!h)
        return B200AIS_E_NOMEM;
    h->cfg = *cfg;
    h->channels = channels;
    h->max_samples = max_samples;
    h->max_tags = max_tags;
    h->L = nsymbols;
    h->nsamples = fft_filter_nsamples(nsymbols);
    int chunk = cfg->corr_chunk > 0 ? cfg->corr_chunk : (24576 / h->nsamples) * h->nsamples;
    chunk = (chunk / h->nsamples) * h->nsamples;
    if (chunk <= 0)
        chunk = h->nsamples;
    h->chunk = chunk;
    h->isps = (int)(cfg->sps + 0.5f);
    h->mark_delay = cfg->mark_delay >= (unsigned)nsymbols ? (unsigned)nsymbols - 1 : cfg->mark_delay;
    {
        // corr_est ctor arithmetic (lib/corr_est_cc_impl.cc:59-74) on the conj-reversed taps
        float corr = 0;
        for (int i = nsymbols - 1; i >= 0; i--) {
            volatile float re = symbols_iq[2 * i], im = -symbols_iq[2 * i + 1];
            volatile float rr = re * re, ii = im * im;
            volatile float s = rr + ii;
            corr = corr + s;
        }
        volatile float tc = cfg->threshold * corr;
        h->thresh = tc * corr;
    }
    {
        volatile float ratio = (float)cfg->data_rate / cfg->sample_rate; // lib/freqest_impl.cc:46-47
        volatile float off = (float)cfg->fftlen * ratio;
        h->offset = (int)off;
        h->binsize = cfg->sample_rate / (float)cfg->fftlen;
        h->sens = (float)(-1.0 / ((double)cfg->sample_rate / (2 * M_PI))); // python/gmsk_sync.py:27
    }
    h->mp.sps_half = (float)((double)cfg->sps / 2.0);
    h->mp.gain = cfg->gain;
    h->mp.gain_omega = (float)((double)(cfg->gain * cfg->gain) * 0.25);
    h->mp.limit = cfg->limit;
    h->mp.osps = cfg->osps;
    // rows are sized for a stream call: up to fftlen-1 waiting input items and nsamples-1
    // waiting AGC outputs come on top of the call's own samples
    const bool fsync = cfg->stages & B200AIS_STAGE_FREQSYNC;
    const size_t cap_n = (size_t)max_samples + (fsync ? cfg->fftlen : 0) + h->nsamples;
    h->HP = (int)round_up((size_t)nsymbols + kKeep, 4);
    h->a_stride = round_up((size_t)h->HP + cap_n + 16, 4);
    h->mask_stride = corr_mask_stride_bytes(nsymbols, (int)cap_n);
    h->corr_stride = round_up(cap_n, 2);
    h->nvec_max = fsync ? max_samples / cfg->fftlen + 1 : 0;

    const size_t C = (size_t)channels;
    // corr_est ctor: taps = reverse(conj(symbols)); fft_filter::set_taps transforms them once
    std::vector<float> fir(2 * (size_t)nsymbols);
    for (int i = 0; i < nsymbols; i++) {
        fir[2 * (size_t)(nsymbols - 1 - i)] = symbols_iq[2 * i];
        fir[2 * (size_t)(nsymbols - 1 - i) + 1] = -symbols_iq[2 * i + 1];
    }
    std::vector<float2> g((size_t)corr_fft_size(nsymbols));
    make_corr_spectrum(fir.data(), nsymbols, g.data());
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) {
        if (e == cudaSuccess)
            e = cudaMalloc(p, bytes ? bytes : 16);
    };
    alloc((void **)&h->d_taps_time, sizeof(float2) * g.size());
    alloc((void **)&h->d_a, sizeof(float2) * h->a_stride * C);
    alloc((void **)&h->d_corr, sizeof(float2) * h->corr_stride * C);
    alloc((void **)&h->d_mask, h->mask_stride * C);
    alloc((void **)&h->d_raw, sizeof(int) * (size_t)std::max(h->nvec_max, 1) * C);
    alloc((void **)&h->d_fhat, sizeof(float) * (size_t)std::max(h->nvec_max, 1) * C);
    alloc((void **)&h->d_ckpt, sizeof(float) * (size_t)std::max(h->nvec_max, 1) * (cfg->fftlen / kSeg + 1) * C);
    alloc((void **)&h->d_tags, sizeof(b200ais_tag) * (size_t)max_tags * C);
    alloc((void **)&h->d_ntags, sizeof(int) * C);
    alloc((void **)&h->d_nbits, sizeof(int) * C);
    alloc((void **)&h->d_ncons, sizeof(int) * C);
    alloc((void **)&h->d_state, sizeof(MskState) * C);
    alloc((void **)&h->d_status, sizeof(int) * (kMaxGroups + 1));
    if (e == cudaSuccess)
        e = cudaMemcpy(h->d_taps_time, g.data(), sizeof(float2) * g.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaMemset(h->d_a, 0, sizeof(float2) * h->a_stride * C); // the zero history pads
    if (e == cudaSuccess)
        e = cudaMemset(h->d_status, 0, sizeof(int) * (kMaxGroups + 1));
    if (e == cudaSuccess)
        e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    for (int g2 = 0; g2 < kMaxGroups && e == cudaSuccess; g2++) {
        e = cudaStreamCreateWithFlags(&h->streams[g2], cudaStreamNonBlocking);
        if (e == cudaSuccess)
            e = cudaEventCreateWithFlags(&h->done[g2], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "demod_create", __FILE__, __LINE__);
        b200ais_demod_destroy(h);
        return rc;
    }
    *out = h;
    return B200AIS_OK;
}

extern "C" int b200ais_demod_destroy(b200ais_demod *h)
{
    if (!h)
        return B200AIS_OK;
    void *ptrs[] = { h->d_taps_time, h->d_corr, h->d_x, h->d_x16, h->d_a, h->d_mask, h->d_raw, h->d_fhat, h->d_ckpt,
                     h->d_tags, h->d_ntags, h->d_nbits, h->d_ncons, h->d_state, h->d_bits,
                     h->d_status, h->d_xcarry, h->d_phase, h->d_yhist[0], h->d_yhist[1],
                     h->d_ctail[0], h->d_ctail[1], h->d_unc, h->d_nold, h->d_tcarry, h->d_a_alt,
                     h->d_tags_alt, h->d_ntags_alt };
    h->xs.release();
    if (h->back_stream)
        cudaStreamDestroy(h->back_stream);
    for (int k = 0; k < 2; k++) {
        if (h->ev_front[k])
            cudaEventDestroy(h->ev_front[k]);
        if (h->ev_back[k])
            cudaEventDestroy(h->ev_back[k]);
    }
    for (void *p : ptrs)
        if (p)
            cudaFree(p);
    if (h->ev_fork)
        cudaEventDestroy(h->ev_fork);
    for (int g = 0; g < kMaxGroups; g++) {
        if (h->streams[g])
            cudaStreamDestroy(h->streams[g]);
        if (h->done[g])
            cudaEventDestroy(h->done[g]);
    }
    h->t_sym.release();
    h->t_err.release();
    h->t_mu.release();
    h->t_soft.release();
    h->t_otags.release();
    for (auto *pool : { &h->ev_used, &h->ev_free })
        for (auto &set : *pool)
            for (auto e : set)
                cudaEventDestroy(e);
    delete h;
    return B200AIS_OK;
}

extern "C" int b200ais_demod_enable_taps(b200ais_demod *h, int enable)
{
    if (!h)
        return B200AIS_E_INVALID;
    h->taps_enabled = enable != 0;
    return B200AIS_OK;
}

static int demod_drain(b200ais_demod *h);

// Launch the chain for channels [c0, c0+cn) on stream s.  iq rows: iq + c*iq_stride.
// in_a != 0 means the samples already sit in the corr_est input rows (host variant with
// neither freq sync nor AGC enabled).
// back != nullptr: the timing loop and the bit tail go to that stream (after ev_front), reading
// the rows of a_base; everything up to the detector stays on s.
static int demod_launch_group(b200ais_demod *h, int c0, int cn, const float2 *iq, size_t iq_stride,
                              int n, int in_a, uint8_t *bits, int max_bits, int *nbits,
                              b200ais_tag *tags, int *ntags, int *d_status, cudaStream_t s,
                              float2 *a_base = nullptr, cudaStream_t back = nullptr,
                              cudaEvent_t ev_front = nullptr)
{
    const b200ais_demod_config &cfg = h->cfg;
    const bool fs = cfg.stages & B200AIS_STAGE_FREQSYNC;
    const int n1 = fs ? (n / cfg.fftlen) * cfg.fftlen : n;
    const int nvec = fs ? n1 / cfg.fftlen : 0;
    float2 *a_rows = (a_base ? a_base : h->d_a) + (size_t)c0 * h->a_stride + h->HP;
    uint8_t *mask = h->d_mask + (size_t)c0 * h->mask_stride;
    int *raw = h->d_raw + (size_t)c0 * std::max(h->nvec_max, 1);
    float *fhat = h->d_fhat + (size_t)c0 * std::max(h->nvec_max, 1);
    float *ckpt = h->d_ckpt + (size_t)c0 * std::max(h->nvec_max, 1) * (cfg.fftlen / kSeg + 1);
    const int vs = std::max(h->nvec_max, 1); // row pitch of raw / fhat
    int rc;
    std::vector<cudaEvent_t> *ev = nullptr;
    if (h->profiling && cn == h->channels) {
        if (h->ev_free.empty()) {
            std::vector<cudaEvent_t> set(B200AIS_STAGE_T_COUNT + 1);
            for (auto &e : set)
                B200_CU(cudaEventCreate(&e));
            h->ev_free.push_back(set);
        }
        h->ev_used.push_back(h->ev_free.back());
        h->ev_free.pop_back();
        ev = &h->ev_used.back();
        B200_CU(cudaEventRecord((*ev)[0], s));
    }
#define B200_MARK(k)                                                     \
    do {                                                                 \
        if (ev)                                                          \
            B200_CU(cudaEventRecord((*ev)[(k) + 1], s));                 \
    } while (0)
    if (fs) {
        if ((rc = launch_sqfft_freqest(iq, iq_stride, cn, nvec, vs, cfg.fftlen, h->offset, raw, s)))
            return rc;
    }
    B200_MARK(B200AIS_STAGE_T_SQFFT);
    if (fs) {
        if ((rc = launch_nco_phase(raw, cn, nvec, vs, cfg.fftlen, h->binsize, h->sens, fhat, ckpt, kSeg, nullptr, s)))
            return rc;
    }
    B200_MARK(B200AIS_STAGE_T_NCO);
    if (!in_a) {
        if ((rc = launch_mix_agc(iq, iq_stride, cn, n1, cfg.fftlen, fhat, vs, ckpt, kSeg, h->sens,
                                 cfg.stages, cfg.agc_nsamples, cfg.agc_reference, a_rows,
                                 h->a_stride, nullptr, nullptr, s)))
            return rc;
    }
    B200_MARK(B200AIS_STAGE_T_MIXAGC);
    // corr_est covers n2 = whole blocks of its output multiple (the scheduler's view); the filter
    // starts from a zero tail (freshly constructed block)
    const int n2 = (n1 / h->nsamples) * h->nsamples;
    float2 *corr = h->d_corr + (size_t)c0 * h->corr_stride;
    const float2 *tw = nullptr;
    if ((rc = get_twiddles(corr_fft_size(h->L), &tw)))
        return rc;
    if ((rc = launch_corr_fft(a_rows, h->a_stride, cn, n2, h->L, tw, h->d_taps_time, h->thresh,
                              nullptr, nullptr, mask, h->mask_stride, corr, h->corr_stride, 1,
                              (int)(h->a_stride - (size_t)h->HP), s)))
        return rc;
    B200_MARK(B200AIS_STAGE_T_CORR);
    if ((rc = launch_detect(corr, h->corr_stride, cn, n2, h->chunk, h->nsamples, h->isps,
                            h->mark_delay, mask, h->mask_stride, 0, 0, tags, h->max_tags, ntags,
                            d_status, 0, s)))
        return rc;
    B200_MARK(B200AIS_STAGE_T_DETECT);
    if (back) {
        B200_CU(cudaEventRecord(ev_front, s));
        B200_CU(cudaStreamWaitEvent(back, ev_front, 0));
        s = back;
    }
    if ((rc = launch_msk_reset(h->d_state + c0, cn, h->mp.sps_half, s)))
        return rc;
    static int no_fuse = -1; // B200AIS_NO_FUSE_TAIL=1: separate k_tail always (experiment)
    if (no_fuse < 0) {
        const char *e = getenv("B200AIS_NO_FUSE_TAIL");
        no_fuse = (e && *e && atoi(e)) ? 1 : 0;
    }
    // the loop writes the bits itself when nobody reads the symbols (no taps): no symbol stream in
    // HBM and no k_tail (65 536 channels: 50.5 -> 48.0 ms).  It makes every step a little longer,
    // which a batch too small to fill the schedulers pays in full (4096 channels: 5.2 -> 5.9 ms)
    static int fuse_min = -1; // B200AIS_FUSE_TAIL_MIN_CH: the smallest batch that fuses (tests, sanitizer runs)
    if (fuse_min < 0) {
        const char *e = getenv("B200AIS_FUSE_TAIL_MIN_CH");
        fuse_min = (e && *e) ? atoi(e) : 8192;
    }
    const bool fuse_tail = !h->taps_enabled && !no_fuse && cn >= fuse_min &&
                           (reinterpret_cast<uintptr_t>(bits) & 3) == 0 && (max_bits & 3) == 0;
    float2 *t_sym = h->t_sym.as<float2>() + (size_t)c0 * max_bits;
    float *t_err = h->taps_enabled ? h->t_err.as<float>() + (size_t)c0 * max_bits : nullptr;
    float *t_mu = h->taps_enabled ? h->t_mu.as<float>() + (size_t)c0 * max_bits : nullptr;
    float *t_soft = h->taps_enabled ? h->t_soft.as<float>() + (size_t)c0 * max_bits : nullptr;
    // msk reads corr_est output 0: out0[k] = in[k - L] (history delay), zeros for k < L
    rc = launch_msk(a_rows - h->L, h->a_stride, cn, max_bits, n2, 0, tags, h->max_tags, ntags,
                    h->mp, h->d_state + c0, t_sym, t_err, t_mu, (size_t)max_bits, nbits,
                    h->d_ncons + c0, 1, d_status, nullptr, s,
                    // pipelined submission: the loop shares the SMs with the next record's front
                    // kernels.  The small ring that lets them in costs a lone warp ~7 ms whatever
                    // the batch, so it only pays when that front is longer (whole chain, >= 12 k channels)
                    back != nullptr && cn >= 12288 && fs && (cfg.stages & B200AIS_STAGE_AGC),
                    // without taps nobody reads the symbols: the loop writes the bits itself
                    fuse_tail ? bits : nullptr, (size_t)max_bits);
    B200_MARK(B200AIS_STAGE_T_MSK);
    if (!rc && !fuse_tail)
        rc = launch_tail(t_sym, (size_t)max_bits, nbits, cn, max_bits, bits, (size_t)max_bits, t_soft, nullptr, s);
    B200_MARK(B200AIS_STAGE_T_TAIL);
#undef B200_MARK
    return rc;
}

// corr_est_cc::set_symbols on the chain's correlator (lib/corr_est_cc_impl.cc:132-162): the taps
// are replaced verbatim -- no conjugate, no reversal -- and d_thresh keeps the constructor's
// value.  The template length is fixed at create (it sizes every buffer).
extern "C" int b200ais_demod_set_symbols(b200ais_demod *h, const float *symbols_iq, int nsymbols,
                                         void *stream)
{
    if (!h || !symbols_iq) {
        set_error("demod_set_symbols: null argument");
        return B200AIS_E_INVALID;
    }
    if (nsymbols != h->L) {
        set_error("demod_set_symbols: the chain was created for %d symbols, got %d", h->L, nsymbols);
        return B200AIS_E_INVALID;
    }
    int rc = demod_drain(h);
    if (rc)
        return rc;
    B200_CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    for (int g = 0; g < kMaxGroups; g++)
        B200_CU(cudaStreamSynchronize(h->streams[g]));
    std::vector<float2> g((size_t)corr_fft_size(nsymbols));
    make_corr_spectrum(symbols_iq, nsymbols, g.data());
    B200_CU(cudaMemcpy(h->d_taps_time, g.data(), sizeof(float2) * g.size(), cudaMemcpyHostToDevice));
    return B200AIS_OK;
}

extern "C" int b200ais_demod_set_overlap(b200ais_demod *h, int groups)
{
    if (!h || groups < 1) {
        set_error("demod_set_overlap: groups must be >= 1");
        return B200AIS_E_INVALID;
    }
    h->overlap_groups = std::min(groups, kMaxGroups);
    return B200AIS_OK;
}

extern "C" int b200ais_demod_profile(b200ais_demod *h, int enable)
{
    if (!h)
        return B200AIS_E_INVALID;
    h->profiling = enable != 0;
    return B200AIS_OK;
}

extern "C" int b200ais_demod_stage_ms(b200ais_demod *h, double *stage_ms, int *calls)
{
    if (!h || !stage_ms) {
        set_error("demod_stage_ms: null argument");
        return B200AIS_E_INVALID;
    }
    for (auto &set : h->ev_used) {
        B200_CU(cudaEventSynchronize(set[B200AIS_STAGE_T_COUNT]));
        for (int k = 0; k < B200AIS_STAGE_T_COUNT; k++) {
            float ms = 0;
            B200_CU(cudaEventElapsedTime(&ms, set[k], set[k + 1]));
            h->stage_ms[k] += ms;
        }
        h->prof_calls++;
        h->ev_free.push_back(set);
    }
    h->ev_used.clear();
    for (int k = 0; k < B200AIS_STAGE_T_COUNT; k++) {
        stage_ms[k] = h->stage_ms[k];
        h->stage_ms[k] = 0;
    }
    if (calls)
        *calls = h->prof_calls;
    h->prof_calls = 0;
    return B200AIS_OK;
}

static int demod_join(b200ais_demod *h, cudaStream_t s);

// wait (on the host) for every back half still in flight: the paths that run on the handle's
// own streams cannot be ordered behind them any other way
static int demod_drain(b200ais_demod *h)
{
    h->inorder_pending = false;
    if (h->back_pending[0] || h->back_pending[1]) {
        B200_CU(cudaStreamSynchronize(h->back_stream));
        h->back_pending[0] = h->back_pending[1] = false;
    }
    return B200AIS_OK;
}

static int demod_prepare(b200ais_demod *h, int n, int max_bits)
{
    if (n < 1 || n > h->max_samples) {
        set_error("demod_work: nsamples %d outside [1, %d]", n, h->max_samples);
        return B200AIS_E_INVALID;
    }
    if (max_bits < 1) {
        set_error("demod_work: max_bits must be positive");
        return B200AIS_E_INVALID;
    }
    {
        int rc = h->t_sym.reserve((size_t)h->channels * max_bits * sizeof(float2));
        if (rc)
            return rc;
    }
    if (h->taps_enabled) {
        const size_t items = (size_t)h->channels * max_bits;
        int rc;
        if ((rc = h->t_err.reserve(items * sizeof(float))) ||
            (rc = h->t_mu.reserve(items * sizeof(float))) || (rc = h->t_soft.reserve(items * sizeof(float))))
            return rc;
    }
    if (h->pad_dirty) { // a stream left items in the history pads the batch path reads as zeros
        B200_CU(cudaDeviceSynchronize());
        B200_CU(cudaMemset2D(h->d_a, h->a_stride * sizeof(float2), 0, (size_t)h->HP * sizeof(float2),
                             (size_t)h->channels));
        B200_CU(cudaDeviceSynchronize());
        h->pad_dirty = false;
    }
    h->st_ready = false; // a batch call overwrites the rows a stream lives in: the next stream call starts fresh
    h->last_n = n;
    h->last_max_bits = max_bits;
    h->last_n1 = (h->cfg.stages & B200AIS_STAGE_FREQSYNC) ? (n / h->cfg.fftlen) * h->cfg.fftlen : n;
    return B200AIS_OK;
}

extern "C" int b200ais_demod_work_dev(b200ais_demod *h, const float *iq, int nsamples, uint8_t *bits,
                                      int max_bits, int *nbits, b200ais_tag *tags, int *ntags,
                                      void *stream)
{
    if (!h || !iq || !bits || !nbits) {
        set_error("demod_work: null argument");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = demod_join(h, s);
    if (rc)
        return rc;
    if ((rc = demod_prepare(h, nsamples, max_bits)))
        return rc;
    const float2 *iq2 = reinterpret_cast<const float2 *>(iq);
    B200_CU(cudaMemsetAsync(h->d_status, 0, sizeof(int) * (kMaxGroups + 1), s));
    int groups = h->profiling ? 1 : std::min(std::min(h->overlap_groups, kMaxGroups), h->channels / 64);
    if (groups <= 1) {
        rc = demod_launch_group(h, 0, h->channels, iq2, (size_t)nsamples, nsamples, 0, bits, max_bits,
                                nbits, h->d_tags, h->d_ntags, h->d_status, s);
        if (rc)
            return rc;
    } else {
        // The NCO-phase and timing-loop kernels are per-channel recurrences whose run time does
        // not depend on the channel count; forking channel groups over internal streams lets one
        // group's recurrences run under the other groups' throughput-bound kernels.
        B200_CU(cudaEventRecord(h->ev_fork, s));
        for (int g = 0; g < groups; g++) {
            const int c0 = (int)((long long)h->channels * g / groups);
            const int c1 = (int)((long long)h->channels * (g + 1) / groups);
            cudaStream_t gs = h->streams[g];
            B200_CU(cudaStreamWaitEvent(gs, h->ev_fork, 0));
            rc = demod_launch_group(h, c0, c1 - c0, iq2 + (size_t)c0 * nsamples, (size_t)nsamples,
                                    nsamples, 0, bits + (size_t)c0 * max_bits, max_bits, nbits + c0,
                                    h->d_tags + (size_t)c0 * h->max_tags, h->d_ntags + c0,
                                    h->d_status + 1 + g, gs);
            if (rc)
                return rc;
            B200_CU(cudaEventRecord(h->done[g], gs));
            B200_CU(cudaStreamWaitEvent(s, h->done[g], 0));
        }
    }
    if (tags)
        B200_CU(cudaMemcpyAsync(tags, h->d_tags, sizeof(b200ais_tag) * (size_t)h->max_tags * h->channels,
                                cudaMemcpyDeviceToDevice, s));
    if (ntags)
        B200_CU(cudaMemcpyAsync(ntags, h->d_ntags, sizeof(int) * (size_t)h->channels,
                                cudaMemcpyDeviceToDevice, s));
    return B200AIS_OK;
}

// make `s` wait for every back half still in flight
static int demod_join(b200ais_demod *h, cudaStream_t s)
{
    h->inorder_pending = false;
    for (int k = 0; k < 2; k++)
        if (h->back_pending[k]) {
            B200_CU(cudaStreamWaitEvent(s, h->ev_back[k], 0));
            h->back_pending[k] = false;
        }
    return B200AIS_OK;
}

extern "C" int b200ais_demod_join(b200ais_demod *h, void *stream)
{
    if (!h) {
        set_error("demod_join: null handle");
        return B200AIS_E_INVALID;
    }
    return demod_join(h, static_cast<cudaStream_t>(stream));
}

extern "C" int b200ais_demod_enqueue_dev(b200ais_demod *h, const float *iq, int nsamples, uint8_t *bits,
                                         int max_bits, int *nbits, b200ais_tag *tags, int *ntags,
                                         void *stream)
{
    if (!h || !iq || !bits || !nbits) {
        set_error("demod_enqueue: null argument");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t C = (size_t)h->channels;
    // A batch that fills the GPU on its own (the timing loop at 14 warps per SM) leaves the side
    // stream nothing to hide under: both halves then slow each other down and the record takes
    // longer than in stream order (65 536 channels: 52.4 ms against 50.3).  Such batches run the
    // whole record on the caller's stream; B200AIS_PIPE_MAX_CH overrides the limit (experiments).
    static long pipe_max = -1;
    if (pipe_max < 0) {
        const char *e = getenv("B200AIS_PIPE_MAX_CH");
        pipe_max = (e && *e) ? atol(e) : kPipeMaxChannels;
    }
    if ((long)h->channels > pipe_max) {
        if (h->back_pending[0] || h->back_pending[1]) {
            int rc = demod_join(h, s);
            if (rc)
                return rc;
        }
        if (h->t_sym.cap < C * (size_t)max_bits * sizeof(float2) || h->pad_dirty || h->taps_enabled)
            B200_CU(cudaStreamSynchronize(s));
        int rc = demod_prepare(h, nsamples, max_bits);
        if (rc)
            return rc;
        if (!h->inorder_pending) // the first record since the last join: the flags start clean
            B200_CU(cudaMemsetAsync(h->d_status, 0, sizeof(int) * (kMaxGroups + 1), s));
        h->inorder_pending = true;
        rc = demod_launch_group(h, 0, h->channels, reinterpret_cast<const float2 *>(iq), (size_t)nsamples,
                                nsamples, 0, bits, max_bits, nbits, h->d_tags, h->d_ntags, h->d_status, s);
        if (rc)
            return rc;
        h->pipe_calls++;
        if (tags)
            B200_CU(cudaMemcpyAsync(tags, h->d_tags, sizeof(b200ais_tag) * (size_t)h->max_tags * C,
                                    cudaMemcpyDeviceToDevice, s));
        if (ntags)
            B200_CU(cudaMemcpyAsync(ntags, h->d_ntags, sizeof(int) * C, cudaMemcpyDeviceToDevice, s));
        return B200AIS_OK;
    }
    if (!h->back_stream) {
        int lo = 0, hi = 0;
        B200_CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        B200_CU(cudaStreamCreateWithPriority(&h->back_stream, cudaStreamNonBlocking, hi));
        B200_CU(cudaMalloc(&h->d_a_alt, sizeof(float2) * h->a_stride * C));
        B200_CU(cudaMemset(h->d_a_alt, 0, sizeof(float2) * h->a_stride * C));
        B200_CU(cudaMalloc(&h->d_tags_alt, sizeof(b200ais_tag) * (size_t)h->max_tags * C));
        B200_CU(cudaMalloc(&h->d_ntags_alt, sizeof(int) * C));
        for (int k = 0; k < 2; k++) {
            B200_CU(cudaEventCreateWithFlags(&h->ev_front[k], cudaEventDisableTiming));
            B200_CU(cudaEventCreateWithFlags(&h->ev_back[k], cudaEventDisableTiming));
        }
    }
    const bool idle = !h->back_pending[0] && !h->back_pending[1];
    if (h->t_sym.cap < C * (size_t)max_bits * sizeof(float2) || h->pad_dirty || h->taps_enabled) {
        // scratch has to grow (or a stream left its items in the rows): drain the pipeline first
        int rc = demod_join(h, s);
        if (rc)
            return rc;
        B200_CU(cudaStreamSynchronize(s));
    }
    int rc = demod_prepare(h, nsamples, max_bits);
    if (rc)
        return rc;
    if (idle)
        B200_CU(cudaMemsetAsync(h->d_status, 0, sizeof(int) * (kMaxGroups + 1), s));
    const int set = (int)(h->pipe_calls & 1u);
    if (h->back_pending[set]) { // the back half of call k-2 read this set of rows and tags
        B200_CU(cudaStreamWaitEvent(s, h->ev_back[set], 0));
        h->back_pending[set] = false;
    }
    b200ais_tag *dtags = set ? h->d_tags_alt : h->d_tags;
    int *dntags = set ? h->d_ntags_alt : h->d_ntags;
    rc = demod_launch_group(h, 0, h->channels, reinterpret_cast<const float2 *>(iq), (size_t)nsamples,
                            nsamples, 0, bits, max_bits, nbits, dtags, dntags, h->d_status, s,
                            set ? h->d_a_alt : h->d_a, h->back_stream, h->ev_front[set]);
    if (rc)
        return rc;
    B200_CU(cudaEventRecord(h->ev_back[set], h->back_stream));
    h->back_pending[set] = true;
    h->pipe_calls++;
    // the tags are final once the detector has run: they go out on the caller's stream
    if (tags)
        B200_CU(cudaMemcpyAsync(tags, dtags, sizeof(b200ais_tag) * (size_t)h->max_tags * C,
                                cudaMemcpyDeviceToDevice, s));
    if (ntags)
        B200_CU(cudaMemcpyAsync(ntags, dntags, sizeof(int) * C, cudaMemcpyDeviceToDevice, s));
    return B200AIS_OK;
}

extern "C" int b200ais_demod_status(b200ais_demod *h)
{
    if (!h)
        return B200AIS_E_INVALID;
    int st[kMaxGroups + 1];
    B200_CU(cudaMemcpy(st, h->d_status, sizeof(st), cudaMemcpyDeviceToHost));
    for (int v : st)
        if (v) {
            set_error("demod kernel flagged status %d", v);
            return v;
        }
    return B200AIS_OK;
}

// host-buffer entry points: iq is complex64 rows, or (iq16 != nullptr) interleaved int16 I/Q rows
// converted on the device as (float)v * scale -- one IEEE multiply per component, exact for a
// power-of-two scale -- the format UHD / osmosdr sources carry on the wire before their host-side
// conversion (python/radio.py:151-203)
static int demod_work_host(b200ais_demod *h, const float *iq, const int16_t *iq16, float scale,
                           int nsamples, uint8_t *bits, int max_bits, int *nbits, b200ais_tag *tags,
                           int *ntags)
{
    if (!h || (!iq && !iq16) || !bits || !nbits) {
        set_error("demod_work: null argument");
        return B200AIS_E_INVALID;
    }
    int rc = demod_drain(h);
    if (rc)
        return rc;
    if ((rc = demod_prepare(h, nsamples, max_bits)))
        return rc;
    const int C = h->channels, n = nsamples;
    const bool direct = (h->cfg.stages & (B200AIS_STAGE_FREQSYNC | B200AIS_STAGE_AGC)) == 0;
    if ((!direct || iq16) && !h->d_x)
        B200_CU(cudaMalloc(&h->d_x, sizeof(float2) * (size_t)h->max_samples * C));
    if (iq16 && !h->d_x16)
        B200_CU(cudaMalloc(&h->d_x16, sizeof(int16_t) * 2 * (size_t)h->max_samples * C));
    if (h->bits_cap < (size_t)max_bits * C) {
        if (h->d_bits)
            cudaFree(h->d_bits);
        h->d_bits = nullptr;
        h->bits_cap = 0;
        B200_CU(cudaMalloc(&h->d_bits, (size_t)max_bits * C));
        h->bits_cap = (size_t)max_bits * C;
    }
    B200_CU(cudaMemset(h->d_status, 0, sizeof(int) * (kMaxGroups + 1)));
    // channel groups pipelined over streams: copy-in of group g+1 overlaps compute of group g
    const int ngroups = std::max(1, std::min(kMaxGroups, C / 64));
    for (int g = 0; g < ngroups; g++) {
        const int c0 = (int)((long long)C * g / ngroups), c1 = (int)((long long)C * (g + 1) / ngroups);
        const int cn = c1 - c0;
        if (cn <= 0)
            continue;
        cudaStream_t s = h->streams[g];
        const float2 *dev_iq;
        size_t dev_stride;
        int in_a = 0;
        if (iq16) {
            int16_t *d16 = h->d_x16 + (size_t)c0 * n * 2;
            B200_CU(cudaMemcpyAsync(d16, iq16 + (size_t)c0 * n * 2, sizeof(int16_t) * 2 * (size_t)n * cn,
                                    cudaMemcpyHostToDevice, s));
            float2 *dst = direct ? h->d_a + (size_t)c0 * h->a_stride + h->HP : h->d_x + (size_t)c0 * n;
            const size_t dst_stride = direct ? h->a_stride : (size_t)n;
            if ((rc = launch_sc16_to_fc(d16, (size_t)n, dst, dst_stride, cn, n, scale, s)))
                return rc;
            dev_iq = dst;
            dev_stride = dst_stride;
            in_a = direct ? 1 : 0;
        } else if (direct) {
            const float2 *src = reinterpret_cast<const float2 *>(iq) + (size_t)c0 * n;
            float2 *a_rows = h->d_a + (size_t)c0 * h->a_stride + h->HP;
            B200_CU(cudaMemcpy2DAsync(a_rows, h->a_stride * sizeof(float2), src, (size_t)n * sizeof(float2),
                                      (size_t)n * sizeof(float2), cn, cudaMemcpyHostToDevice, s));
            dev_iq = a_rows;
            dev_stride = h->a_stride;
            in_a = 1;
        } else {
            const float2 *src = reinterpret_cast<const float2 *>(iq) + (size_t)c0 * n;
            float2 *dx = h->d_x + (size_t)c0 * n;
            B200_CU(cudaMemcpyAsync(dx, src, sizeof(float2) * (size_t)n * cn, cudaMemcpyHostToDevice, s));
            dev_iq = dx;
            dev_stride = (size_t)n;
        }
        rc = demod_launch_group(h, c0, cn, dev_iq, dev_stride, n, in_a,
                                h->d_bits + (size_t)c0 * max_bits, max_bits, h->d_nbits + c0,
                                h->d_tags + (size_t)c0 * h->max_tags, h->d_ntags + c0,
                                h->d_status + 1 + g, s);
        if (rc)
            return rc;
        B200_CU(cudaMemcpyAsync(bits + (size_t)c0 * max_bits, h->d_bits + (size_t)c0 * max_bits,
                                (size_t)max_bits * cn, cudaMemcpyDeviceToHost, s));
        B200_CU(cudaMemcpyAsync(nbits + c0, h->d_nbits + c0, sizeof(int) * (size_t)cn,
                                cudaMemcpyDeviceToHost, s));
        if (tags)
            B200_CU(cudaMemcpyAsync(tags + (size_t)c0 * h->max_tags, h->d_tags + (size_t)c0 * h->max_tags,
                                    sizeof(b200ais_tag) * (size_t)h->max_tags * cn,
                                    cudaMemcpyDeviceToHost, s));
        if (ntags)
            B200_CU(cudaMemcpyAsync(ntags + c0, h->d_ntags + c0, sizeof(int) * (size_t)cn,
                                    cudaMemcpyDeviceToHost, s));
    }
    for (int g = 0; g < ngroups; g++)
        B200_CU(cudaStreamSynchronize(h->streams[g]));
    return b200ais_demod_status(h);
}

extern "C" int b200ais_demod_work(b200ais_demod *h, const float *iq, int nsamples, uint8_t *bits,
                                  int max_bits, int *nbits, b200ais_tag *tags, int *ntags)
{
    if (!iq) {
        set_error("demod_work: null argument");
        return B200AIS_E_INVALID;
    }
    return demod_work_host(h, iq, nullptr, 1.0f, nsamples, bits, max_bits, nbits, tags, ntags);
}

extern "C" int b200ais_demod_work_sc16(b200ais_demod *h, const int16_t *iq, float scale, int nsamples,
                                       uint8_t *bits, int max_bits, int *nbits, b200ais_tag *tags,
                                       int *ntags)
{
    if (!iq) {
        set_error("demod_work_sc16: null argument");
        return B200AIS_E_INVALID;
    }
    return demod_work_host(h, nullptr, iq, scale, nsamples, bits, max_bits, nbits, tags, ntags);
}


// =============================================== the chain as a stream
//
// b200ais_demod_work() restarts every block on each call (one record = one flowgraph run).
// The stream entry points keep what GNU Radio's scheduler and the blocks keep between work()
// calls, so a capture can be fed in pieces of any size:
//   - input items waiting for a whole FFT vector (stream_to_vector),
//   - the NCO phase (frequency_modulator_fc d_phase),
//   - the last 511 mixed items (feedforward_agc_cc history),
//   - AGC outputs waiting for a whole corr_est output multiple, the filter's overlap-add tail
//     and the L history items behind them (lib/corr_est_cc_impl.cc:77-78,188),
//   - corr_est output-0 items the timing loop has not consumed, its loop state and the tags it
//     may still meet (lib/msk_timing_recovery_cc_impl.cc:125-130,203),
//   - the previous symbol / slicer decision of quadrature_demod_cf / diff_decoder_bb.
// One call = one scheduler pass in which every block runs once over what is available to it.
// Row layout of d_a in stream mode: [HP items behind corr_est's read pointer | waiting AGC
// outputs | this call's AGC outputs]; corr_est's filter reads from row + HP, the timing loop
// from row + HP - L - unconsumed[c].

static int stream_alloc(b200ais_demod *h)
{
    if (h->d_unc)
        return B200AIS_OK;
    const size_t C = (size_t)h->channels;
    const bool fs = h->cfg.stages & B200AIS_STAGE_FREQSYNC;
    h->xs_stride = round_up((size_t)h->max_samples + (fs ? h->cfg.fftlen : 0) + 4, 4);
    B200_CU(cudaMalloc(&h->d_xcarry, sizeof(float2) * C * (size_t)std::max(h->cfg.fftlen, 1)));
    B200_CU(cudaMalloc(&h->d_phase, sizeof(float) * C));
    for (int k = 0; k < 2; k++) {
        B200_CU(cudaMalloc(&h->d_yhist[k], sizeof(float2) * C * 511));
        B200_CU(cudaMalloc(&h->d_ctail[k], sizeof(float2) * C * (size_t)(h->L - 1)));
    }
    B200_CU(cudaMalloc(&h->d_nold, sizeof(int) * C));
    B200_CU(cudaMalloc(&h->d_tcarry, sizeof(TailCarry) * C));
    B200_CU(cudaMalloc(&h->d_unc, sizeof(int) * C));
    return B200AIS_OK;
}

extern "C" int b200ais_demod_stream_reset(b200ais_demod *h, void *stream)
{
    if (!h) {
        set_error("demod_stream_reset: null handle");
        return B200AIS_E_INVALID;
    }
    if (!(h->cfg.stages & B200AIS_STAGE_AGC) || h->cfg.agc_nsamples != 512) {
        set_error("demod_stream: needs the AGC stage with the reference's 512-sample window");
        return B200AIS_E_INVALID;
    }
    if ((int)std::ceil(3.0 * h->mp.sps_half) + 6 > kKeep) {
        set_error("demod_stream: samples per symbol too large for the %d-item carry", kKeep);
        return B200AIS_E_INVALID;
    }
    int rc = stream_alloc(h);
    if (rc)
        return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t C = (size_t)h->channels;
    B200_CU(cudaMemset2DAsync(h->d_a, h->a_stride * sizeof(float2), 0,
                              (size_t)(h->HP + h->nsamples) * sizeof(float2), C, s));
    B200_CU(cudaMemsetAsync(h->d_phase, 0, sizeof(float) * C, s));
    for (int k = 0; k < 2; k++) {
        B200_CU(cudaMemsetAsync(h->d_yhist[k], 0, sizeof(float2) * C * 511, s));
        B200_CU(cudaMemsetAsync(h->d_ctail[k], 0, sizeof(float2) * C * (size_t)(h->L - 1), s));
    }
    B200_CU(cudaMemsetAsync(h->d_unc, 0, sizeof(int) * C, s));
    B200_CU(cudaMemsetAsync(h->d_nold, 0, sizeof(int) * C, s));
    B200_CU(cudaMemsetAsync(h->d_ntags, 0, sizeof(int) * C, s));
    B200_CU(cudaMemsetAsync(h->d_tcarry, 0, sizeof(TailCarry) * C, s));
    if ((rc = launch_msk_reset(h->d_state, h->channels, h->mp.sps_half, s)))
        return rc;
    h->st_nx = h->st_na = 0;
    h->st_written = 0;
    h->yh_cur = h->ct_cur = 0;
    h->st_ready = true;
    return B200AIS_OK;
}

extern "C" int b200ais_demod_stream_max_bits(const b200ais_demod *h, int nsamples)
{
    if (!h)
        return 0;
    const int extra = ((h->cfg.stages & B200AIS_STAGE_FREQSYNC) ? h->cfg.fftlen : 0) + h->nsamples + kKeep;
    return demod_max_bits(h, nsamples + extra);
}

extern "C" int b200ais_demod_stream_pending(const b200ais_demod *h, int *input_items, int *agc_items,
                                            uint64_t *corr_written)
{
    if (!h)
        return B200AIS_E_INVALID;
    if (input_items)
        *input_items = h->st_ready ? h->st_nx : 0;
    if (agc_items)
        *agc_items = h->st_ready ? h->st_na : 0;
    if (corr_written)
        *corr_written = h->st_ready ? h->st_written : 0;
    return B200AIS_OK;
}

// xin: this call's input rows on the device.  When items are waiting (st_nx > 0) or host_rows
// is set, the rows are [waiting | new] in h->xs and xin is ignored / already there.
static int stream_run(b200ais_demod *h, const float2 *xin, size_t xin_stride, bool in_xs, int n,
                      uint8_t *bits, int max_bits, int *nbits, b200ais_tag *tags, int *ntags,
                      cudaStream_t s)
{
    const b200ais_demod_config &cfg = h->cfg;
    const int C = h->channels;
    const bool fs = cfg.stages & B200AIS_STAGE_FREQSYNC;
    const int nx = h->st_nx, navail = nx + n;
    int rc;
    if (!in_xs && nx > 0) { // [waiting | new] into the assembly rows
        if ((rc = h->xs.reserve(h->xs_stride * (size_t)C * sizeof(float2))))
            return rc;
        float2 *xs = h->xs.as<float2>();
        B200_CU(cudaMemcpy2DAsync(xs + nx, h->xs_stride * sizeof(float2), xin, xin_stride * sizeof(float2),
                                  (size_t)n * sizeof(float2), C, cudaMemcpyDeviceToDevice, s));
        xin = xs;
        xin_stride = h->xs_stride;
        in_xs = true;
    }
    if (in_xs && nx > 0) {
        float2 *xs = h->xs.as<float2>();
        B200_CU(cudaMemcpy2DAsync(xs, h->xs_stride * sizeof(float2), h->d_xcarry,
                                  (size_t)cfg.fftlen * sizeof(float2), (size_t)nx * sizeof(float2), C,
                                  cudaMemcpyDeviceToDevice, s));
    }
    const int n1 = fs ? (navail / cfg.fftlen) * cfg.fftlen : navail;
    const int nvec = fs ? n1 / cfg.fftlen : 0;
    const int vs = std::max(h->nvec_max, 1);
    if (fs && nvec > 0) {
        if ((rc = launch_sqfft_freqest(xin, xin_stride, C, nvec, vs, cfg.fftlen, h->offset, h->d_raw, s)))
            return rc;
        if ((rc = launch_nco_phase(h->d_raw, C, nvec, vs, cfg.fftlen, h->binsize, h->sens, h->d_fhat,
                                   h->d_ckpt, kSeg, h->d_phase, s)))
            return rc;
    }
    const int na = h->st_na;
    if (n1 > 0) {
        if ((rc = launch_mix_agc(xin, xin_stride, C, n1, cfg.fftlen, h->d_fhat, vs, h->d_ckpt, kSeg,
                                 h->sens, cfg.stages, cfg.agc_nsamples, cfg.agc_reference,
                                 h->d_a + h->HP + na, h->a_stride, h->d_yhist[h->yh_cur],
                                 h->d_yhist[h->yh_cur ^ 1], s)))
            return rc;
        h->yh_cur ^= 1;
    }
    const int nx_new = navail - n1; // < fftlen
    if (nx_new > 0)
        B200_CU(cudaMemcpy2DAsync(h->d_xcarry, (size_t)cfg.fftlen * sizeof(float2), xin + n1,
                                  xin_stride * sizeof(float2), (size_t)nx_new * sizeof(float2), C,
                                  cudaMemcpyDeviceToDevice, s));
    // corr_est: whole output multiples, in work chunks, from the carried filter tail
    const int avail = na + n1;
    const int n2 = (avail / h->nsamples) * h->nsamples;
    const float2 *tw = nullptr;
    if ((rc = get_twiddles(corr_fft_size(h->L), &tw)))
        return rc;
    if (n2 > 0) {
        if ((rc = launch_corr_fft(h->d_a + h->HP, h->a_stride, C, n2, h->L, tw, h->d_taps_time,
                                  h->thresh, h->d_ctail[h->ct_cur], h->d_ctail[h->ct_cur ^ 1], h->d_mask,
                                  h->mask_stride, h->d_corr, h->corr_stride, 1,
                                  (int)(h->a_stride - (size_t)h->HP), s)))
            return rc;
        h->ct_cur ^= 1;
    }
    if ((rc = launch_tags_compact(h->d_tags, h->max_tags, h->d_ntags, h->d_unc, h->st_written, C,
                                  h->d_nold, s)))
        return rc;
    if ((rc = launch_detect(h->d_corr, h->corr_stride, C, n2, h->chunk, h->nsamples, h->isps,
                            h->mark_delay, h->d_mask, h->mask_stride, h->st_written, 0, h->d_tags,
                            h->max_tags, h->d_ntags, h->d_status, 1, s)))
        return rc;
    // the timing loop over everything corr_est has produced and it has not consumed
    float2 *t_sym = h->t_sym.as<float2>();
    float *t_err = h->taps_enabled ? h->t_err.as<float>() : nullptr;
    float *t_mu = h->taps_enabled ? h->t_mu.as<float>() : nullptr;
    float *t_soft = h->taps_enabled ? h->t_soft.as<float>() : nullptr;
    if ((rc = launch_msk(h->d_a + h->HP - h->L, h->a_stride, C, max_bits, n2, h->st_written, h->d_tags,
                         h->max_tags, h->d_ntags, h->mp, h->d_state, t_sym, t_err, t_mu,
                         (size_t)max_bits, nbits, h->d_ncons, 1, h->d_status, h->d_unc, s)))
        return rc;
    if ((rc = launch_tail(t_sym, (size_t)max_bits, nbits, C, max_bits, bits, (size_t)max_bits, t_soft,
                          h->d_tcarry, s)))
        return rc;
    if (tags || ntags) {
        if (!tags) {
            set_error("demod_stream_work: ntags without tags");
            return B200AIS_E_INVALID;
        }
        if ((rc = launch_tags_emit(h->d_tags, h->max_tags, h->d_ntags, h->d_nold, C, tags, ntags,
                                   h->d_status, s)))
            return rc;
    }
    // keep HP items behind corr_est's new read pointer and what it has not taken
    if ((rc = launch_roll_rows(h->d_a, h->a_stride, C, n2, h->HP + avail - n2, s)))
        return rc;
    h->st_nx = nx_new;
    h->st_na = avail - n2;
    h->st_written += (uint64_t)n2;
    h->pad_dirty = true;
    h->last_n = n;
    h->last_max_bits = max_bits;
    h->last_n1 = n1;
    return B200AIS_OK;
}

static int stream_prepare(b200ais_demod *h, int n, int max_bits, cudaStream_t s)
{
    if (n < 0 || n > h->max_samples) {
        set_error("demod_stream_work: nsamples %d outside [0, %d]", n, h->max_samples);
        return B200AIS_E_INVALID;
    }
    // a shorter output row would stop the timing loop on noutput_items and leave more than the
    // kKeep items the rows carry between calls unconsumed
    if (max_bits < b200ais_demod_stream_max_bits(h, n)) {
        set_error("demod_stream_work: max_bits %d below b200ais_demod_stream_max_bits(%d) = %d", max_bits, n,
                  b200ais_demod_stream_max_bits(h, n));
        return B200AIS_E_INVALID;
    }
    int rc = demod_drain(h);
    if (rc)
        return rc;
    if (!h->st_ready && (rc = b200ais_demod_stream_reset(h, s)))
        return rc;
    if ((rc = h->t_sym.reserve((size_t)h->channels * max_bits * sizeof(float2))))
        return rc;
    if (h->taps_enabled) {
        const size_t items = (size_t)h->channels * max_bits;
        if ((rc = h->t_err.reserve(items * sizeof(float))) ||
            (rc = h->t_mu.reserve(items * sizeof(float))) || (rc = h->t_soft.reserve(items * sizeof(float))))
            return rc;
    }
    return B200AIS_OK;
}

extern "C" int b200ais_demod_stream_work_dev(b200ais_demod *h, const float *iq, int nsamples,
                                             uint8_t *bits, int max_bits, int *nbits,
                                             b200ais_tag *tags, int *ntags, void *stream)
{
    if (!h || (!iq && nsamples > 0) || !bits || !nbits) {
        set_error("demod_stream_work: null argument");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = stream_prepare(h, nsamples, max_bits, s);
    if (rc)
        return rc;
    B200_CU(cudaMemsetAsync(h->d_status, 0, sizeof(int) * (kMaxGroups + 1), s));
    return stream_run(h, reinterpret_cast<const float2 *>(iq), (size_t)nsamples, false, nsamples, bits,
                      max_bits, nbits, tags, ntags, s);
}

extern "C" int b200ais_demod_stream_stage(b200ais_demod *h, int nsamples, int max_bits, float **dev_ptr,
                                          size_t *stride, void *stream)
{
    if (!h || !dev_ptr || !stride) {
        set_error("demod_stream_stage: null argument");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = stream_prepare(h, nsamples, max_bits, s);
    if (rc)
        return rc;
    if ((rc = h->xs.reserve(h->xs_stride * (size_t)h->channels * sizeof(float2))))
        return rc;
    *dev_ptr = reinterpret_cast<float *>(h->xs.as<float2>() + h->st_nx);
    *stride = h->xs_stride;
    return B200AIS_OK;
}

extern "C" int b200ais_demod_stream_work_staged(b200ais_demod *h, int nsamples, uint8_t *bits,
                                                int max_bits, int *nbits, b200ais_tag *tags,
                                                int *ntags, void *stream)
{
    if (!h || !bits || !nbits) {
        set_error("demod_stream_work_staged: null argument");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = stream_prepare(h, nsamples, max_bits, s);
    if (rc)
        return rc;
    if ((rc = h->xs.reserve(h->xs_stride * (size_t)h->channels * sizeof(float2))))
        return rc;
    B200_CU(cudaMemsetAsync(h->d_status, 0, sizeof(int) * (kMaxGroups + 1), s));
    return stream_run(h, h->xs.as<float2>(), h->xs_stride, true, nsamples, bits, max_bits, nbits, tags,
                      ntags, s);
}

extern "C" int b200ais_demod_stream_work(b200ais_demod *h, const float *iq, int nsamples, uint8_t *bits,
                                         int max_bits, int *nbits, b200ais_tag *tags, int *ntags)
{
    if (!h || (!iq && nsamples > 0) || !bits || !nbits) {
        set_error("demod_stream_work: null argument");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = h->streams[0];
    int rc = stream_prepare(h, nsamples, max_bits, s);
    if (rc)
        return rc;
    const int C = h->channels, n = nsamples;
    if ((rc = h->xs.reserve(h->xs_stride * (size_t)C * sizeof(float2))))
        return rc;
    if (h->bits_cap < (size_t)max_bits * C) {
        if (h->d_bits)
            cudaFree(h->d_bits);
        h->d_bits = nullptr;
        h->bits_cap = 0;
        B200_CU(cudaMalloc(&h->d_bits, (size_t)max_bits * C));
        h->bits_cap = (size_t)max_bits * C;
    }
    DevBuf &ot = h->t_otags; // this call's tags before they go to the host
    if (tags && (rc = ot.reserve(sizeof(b200ais_tag) * (size_t)h->max_tags * C)))
        return rc;
    B200_CU(cudaMemsetAsync(h->d_status, 0, sizeof(int) * (kMaxGroups + 1), s));
    float2 *xs = h->xs.as<float2>();
    if (n > 0)
        B200_CU(cudaMemcpy2DAsync(xs + h->st_nx, h->xs_stride * sizeof(float2), iq, (size_t)n * sizeof(float2),
                                  (size_t)n * sizeof(float2), C, cudaMemcpyHostToDevice, s));
    rc = stream_run(h, xs, h->xs_stride, true, n, h->d_bits, max_bits, h->d_nbits,
                    tags ? ot.as<b200ais_tag>() : nullptr, tags ? h->d_ncons : nullptr, s);
    if (rc)
        return rc;
    B200_CU(cudaMemcpyAsync(bits, h->d_bits, (size_t)max_bits * C, cudaMemcpyDeviceToHost, s));
    B200_CU(cudaMemcpyAsync(nbits, h->d_nbits, sizeof(int) * (size_t)C, cudaMemcpyDeviceToHost, s));
    if (tags) {
        B200_CU(cudaMemcpyAsync(tags, ot.as<b200ais_tag>(), sizeof(b200ais_tag) * (size_t)h->max_tags * C,
                                cudaMemcpyDeviceToHost, s));
        if (ntags)
            B200_CU(cudaMemcpyAsync(ntags, h->d_ncons, sizeof(int) * (size_t)C, cudaMemcpyDeviceToHost, s));
    }
    B200_CU(cudaStreamSynchronize(s));
    return b200ais_demod_status(h);
}

extern "C" int b200ais_demod_tap(b200ais_demod *h, int which, void **dev_ptr, size_t *row_items)
{
    if (!h || !dev_ptr || !row_items) {
        set_error("demod_tap: null argument");
        return B200AIS_E_INVALID;
    }
    const bool fs = h->cfg.stages & B200AIS_STAGE_FREQSYNC;
    switch (which) {
    case B200AIS_TAP_FHAT:
        *dev_ptr = h->d_fhat;
        *row_items = fs ? (size_t)(h->last_n1 / h->cfg.fftlen) : 0;
        return B200AIS_OK;
    case B200AIS_TAP_AGC:
        *dev_ptr = h->d_a + h->HP;
        *row_items = h->a_stride;
        return B200AIS_OK;
    case B200AIS_TAP_MASK:
        *dev_ptr = h->d_mask;
        *row_items = h->mask_stride;
        return B200AIS_OK;
    case B200AIS_TAP_SYM:
        *dev_ptr = h->t_sym.p;
        *row_items = (size_t)h->last_max_bits;
        return B200AIS_OK;
    case B200AIS_TAP_ERR:
    case B200AIS_TAP_MU:
    case B200AIS_TAP_SOFT: {
        if (!h->taps_enabled) {
            set_error("demod_tap: call b200ais_demod_enable_taps(h, 1) before the work call");
            return B200AIS_E_INVALID;
        }
        DevBuf *b = which == B200AIS_TAP_SYM ? &h->t_sym
                    : which == B200AIS_TAP_ERR ? &h->t_err
                    : which == B200AIS_TAP_MU  ? &h->t_mu
                                               : &h->t_soft;
        *dev_ptr = b->p;
        *row_items = (size_t)h->last_max_bits;
        return B200AIS_OK;
    }
    default:
        set_error("demod_tap: unknown tap %d", which);
        return B200AIS_E_INVALID;
    }
}

extern "C" int b200ais_demod_read_tap(b200ais_demod *h, int which, void *dst, size_t dst_bytes)
{
    void *p = nullptr;
    size_t row = 0;
    int rc = b200ais_demod_tap(h, which, &p, &row);
    if (rc)
        return rc;
    size_t item = which == B200AIS_TAP_AGC || which == B200AIS_TAP_SYM ? sizeof(float2)
                  : which == B200AIS_TAP_MASK                          ? 1
                                                                       : sizeof(float);
    const size_t C = (size_t)h->channels;
    if (which == B200AIS_TAP_FHAT) {
        const size_t nvec = row, pitch = (size_t)std::max(h->nvec_max, 1);
        if (dst_bytes < C * nvec * item) {
            set_error("demod_read_tap: destination too small");
            return B200AIS_E_INVALID;
        }
        B200_CU(cudaDeviceSynchronize());
        if (nvec)
            B200_CU(cudaMemcpy2D(dst, nvec * item, p, pitch * item, nvec * item, C, cudaMemcpyDeviceToHost));
        return B200AIS_OK;
    }
    if (dst_bytes < C * row * item) {
        set_error("demod_read_tap: destination too small (%zu < %zu)", dst_bytes, C * row * item);
        return B200AIS_E_INVALID;
    }
    B200_CU(cudaDeviceSynchronize());
    if (which == B200AIS_TAP_AGC) {
        // strided rows: row pitch a_stride, pointer already past the zero history
        B200_CU(cudaMemcpy2D(dst, row * item, p, row * item, (row - h->HP) * item, C, cudaMemcpyDeviceToHost));
        return B200AIS_OK;
    }
    B200_CU(cudaMemcpy(dst, p, C * row * item, cudaMemcpyDeviceToHost));
    return B200AIS_OK;
}
