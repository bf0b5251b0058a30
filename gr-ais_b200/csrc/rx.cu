// rx.cu -- the reference's ais_rx receiver (python/radio.py:39-72) for many wideband sources:
//   freq_xlating_fir_filter_ccf -> ais_demod -> hdlc_deframer_bp(11, 64) -> pdu_to_nmea
// composed from the C-ABI blocks of this library on one stream, every block keeping its
// state from call to call (a capture is fed in pieces of any size).  The channeliser writes
// straight into the demod chain's assembly rows; nothing between the wideband input and the
// NMEA sentences leaves the device.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <future>
#include <new>
#include <vector>

#include "internal.h"

using namespace b200ais;

struct b200ais_rx {
    b200ais_rx_config cfg;
    int D = 0, ntaps = 0, channels = 0, max_out = 0, max_bits = 0, slot = 0;
    b200ais_xlat *xl = nullptr;
    b200ais_demod *dm = nullptr;
    b200ais_hdlc *hd = nullptr;
    cudaStream_t stream = nullptr;
    float2 *d_x = nullptr; // [sources][x_stride]: history + leftover + this call's items
    size_t x_stride = 0;
    int carry = 0; // items of every row kept from the last call
    uint8_t *d_bits = nullptr;
    int *d_nbits = nullptr, *d_nframes = nullptr, *d_count = nullptr, *d_status = nullptr;
    uint64_t tag_overflows = 0; // calls in which a channel met more tags than its row holds
    b200ais_frame *d_frames = nullptr;
    char *d_des = nullptr;
    // device-side outputs of the host variant
    b200ais_frame *d_msgs = nullptr;
    char *d_sent = nullptr;
    int *d_lens = nullptr;
    int out_cap = 0;
    uint64_t items_in = 0, items_out = 0;
};

extern "C" int b200ais_rx_default_config(b200ais_rx_config *c)
{
    if (!c)
        return B200AIS_E_INVALID;
    memset(c, 0, sizeof(*c));
    c->rate = 250e3;                 // python/radio.py:120
    c->nfreqs = 2;                   // python/radio.py:88-89
    c->freqs[0] = 161.975e6 - 162.0e6;
    c->freqs[1] = 162.025e6 - 162.0e6;
    c->designators[0][0] = 'A';
    c->designators[1][0] = 'B';
    c->sources = 1;
    c->max_input_items = 1 << 18;
    c->max_frames = 64;
    c->bits_per_sec = 9600.0f;       // python/radio.py:47
    c->clockrec_gain = 0.04f;        // :58
    c->omega_relative_limit = 0.01f; // :59
    c->fftlen = 1024;                // :61
    c->lpf_cutoff = 11000.0;         // :49
    c->lpf_transition = 1000.0;
    c->hdlc_length_min = 11;         // :64
    c->hdlc_length_max = 64;
    return B200AIS_OK;
}

extern "C" int b200ais_rx_destroy(b200ais_rx *h)
{
    if (!h)
        return B200AIS_OK;
    b200ais_xlat_destroy(h->xl);
    b200ais_demod_destroy(h->dm);
    b200ais_hdlc_destroy(h->hd);
    void *bufs[] = {h->d_x, h->d_bits, h->d_nbits, h->d_nframes, h->d_count, h->d_status,
                    h->d_frames, h->d_des, h->d_msgs, h->d_sent, h->d_lens};
    for (void *b : bufs)
        if (b)
            cudaFree(b);
    if (h->stream)
        cudaStreamDestroy(h->stream);
    delete h;
    return B200AIS_OK;
}

extern "C" int b200ais_rx_decimation(const b200ais_rx *h) { return h ? h->D : 0; }
extern "C" int b200ais_rx_channels(const b200ais_rx *h) { return h ? h->channels : 0; }
extern "C" int b200ais_rx_sentence_slot(const b200ais_rx *h) { return h ? h->slot : 0; }
extern "C" float b200ais_rx_samples_per_symbol(const b200ais_rx *h)
{
    return h ? (float)((h->cfg.rate / h->D) / h->cfg.bits_per_sec) : 0.0f;
}

extern "C" int b200ais_rx_create(b200ais_rx **out, const b200ais_rx_config *cfg,
                                 const float *symbols_iq, int nsymbols)
{
    if (!out || !cfg || !symbols_iq || nsymbols < 1 || cfg->nfreqs < 1 || cfg->nfreqs > 16 ||
        cfg->sources < 1 || cfg->max_input_items < 1 || cfg->max_frames < 1 || !(cfg->rate > 0) ||
        !(cfg->bits_per_sec > 0)) {
        set_error("rx_create: bad configuration");
        return B200AIS_E_INVALID;
    }
    b200ais_rx *h = new (std::nothrow) b200ais_rx;
    if (!h)
        return B200AIS_E_NOMEM;
    h->cfg = *cfg;
    // python/radio.py:50: int(rate / (bits_per_sec * samples_per_symbol)), samples_per_symbol = 5
    h->D = (int)(cfg->rate / ((double)cfg->bits_per_sec * 5.0));
    if (h->D < 1) {
        delete h;
        set_error("rx_create: rate below one 48 ksps channel");
        return B200AIS_E_INVALID;
    }
    int rc, nt = 0;
    if ((rc = b200ais_firdes_low_pass(1.0, cfg->rate, cfg->lpf_cutoff, cfg->lpf_transition, nullptr, 0,
                                      &nt))) {
        delete h;
        return rc;
    }
    std::vector<float> taps((size_t)nt);
    rc = b200ais_firdes_low_pass(1.0, cfg->rate, cfg->lpf_cutoff, cfg->lpf_transition, taps.data(), nt, &nt);
    h->ntaps = nt;
    h->channels = cfg->sources * cfg->nfreqs;
    h->max_out = (cfg->max_input_items + h->D - 1) / h->D + 1;
    if (!rc)
        rc = b200ais_xlat_create(&h->xl, h->D, taps.data(), nt, cfg->freqs, cfg->nfreqs, cfg->rate,
                                 cfg->sources);
    if (!rc) {
        b200ais_demod_config dc;
        b200ais_demod_default_config(&dc);
        const float sps = b200ais_rx_samples_per_symbol(h); // python/radio.py:57
        dc.sps = sps;
        dc.data_rate = (int)cfg->bits_per_sec;
        dc.sample_rate = sps * cfg->bits_per_sec;           // python/ais_demod.py:30
        dc.fftlen = cfg->fftlen;
        dc.gain = cfg->clockrec_gain;
        dc.limit = cfg->omega_relative_limit;
        // corr_est_cc adds four tags per detection and at most one detection per isps items
        // (lib/corr_est_cc_impl.cc:213-256,270); a busy channel carries 37.5 slots/s and about six
        // detections per burst.  The row holds the worst case up to 4096 tags; beyond that the
        // call counts an overflow (b200ais_rx_tag_overflows) and goes on.
        const int isps = (int)(sps + 0.5f);
        const int max_tags = std::min(4096, std::max(256, 4 * (h->max_out / std::max(isps, 1) + 2)));
        rc = b200ais_demod_create(&h->dm, &dc, symbols_iq, nsymbols, h->channels, h->max_out, max_tags);
    }
    if (!rc)
        rc = b200ais_hdlc_create(&h->hd, cfg->hdlc_length_min, cfg->hdlc_length_max, h->channels);
    if (!rc) {
        h->max_bits = b200ais_demod_stream_max_bits(h->dm, h->max_out);
        h->max_bits = (h->max_bits + 15) / 16 * 16; // rows of 16-byte words for k_hdlc
        h->slot = b200ais_nmea_slot_bytes(cfg->hdlc_length_max, "12345678");
        h->x_stride = ((size_t)(nt - 1) + h->D + cfg->max_input_items + 3) / 2 * 2;
        const size_t C = (size_t)h->channels;
        cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_x, sizeof(float2) * h->x_stride * cfg->sources);
        if (e == cudaSuccess) e = cudaMemset(h->d_x, 0, sizeof(float2) * h->x_stride * cfg->sources);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_bits, (size_t)h->max_bits * C);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_nbits, sizeof(int) * C);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_nframes, sizeof(int) * C);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_count, sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&h->d_status, sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(h->d_status, 0, sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&h->d_frames, sizeof(b200ais_frame) * C * cfg->max_frames);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_des, 16 * 8);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_des, cfg->designators, 16 * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess)
            rc = cuda_fail(e, "rx_create", __FILE__, __LINE__);
    }
    if (rc) {
        b200ais_rx_destroy(h);
        return rc;
    }
    h->carry = nt - 1; // freq_xlating's history starts as zeros
    *out = h;
    return B200AIS_OK;
}

extern "C" int b200ais_rx_reset(b200ais_rx *h)
{
    if (!h)
        return B200AIS_E_INVALID;
    B200_CU(cudaDeviceSynchronize());
    B200_CU(cudaMemset(h->d_x, 0, sizeof(float2) * h->x_stride * h->cfg.sources));
    h->carry = h->ntaps - 1;
    h->items_in = h->items_out = 0;
    int rc;
    if ((rc = b200ais_xlat_reset(h->xl)) || (rc = b200ais_demod_stream_reset(h->dm, nullptr)) ||
        (rc = b200ais_hdlc_reset(h->hd)))
        return rc;
    B200_CU(cudaDeviceSynchronize());
    return B200AIS_OK;
}

// everything after this call's items are in d_x at [carry, carry + n)
static int rx_run(b200ais_rx *h, int n, b200ais_frame *msgs, char *sentences, int slot, int *lens,
                  int max_msgs, int *count, cudaStream_t s)
{
    const int S = h->cfg.sources;
    const int avail = h->carry + n;
    const int nout = avail >= h->ntaps - 1 + h->D ? (avail - (h->ntaps - 1)) / h->D : 0;
    int rc;
    float *stage = nullptr;
    size_t sstride = 0;
    if ((rc = b200ais_demod_stream_stage(h->dm, nout, h->max_bits, &stage, &sstride, s)))
        return rc;
    if (nout > 0) {
        if ((rc = b200ais_xlat_work_dev(h->xl, nout, reinterpret_cast<const float *>(h->d_x), h->x_stride,
                                        stage, sstride, s)))
            return rc;
        if ((rc = launch_roll_rows(h->d_x, h->x_stride, S, nout * h->D, avail - nout * h->D, s)))
            return rc;
    }
    h->carry = avail - nout * h->D;
    h->items_in += (uint64_t)n;
    h->items_out += (uint64_t)nout;
    if ((rc = b200ais_demod_stream_work_staged(h->dm, nout, h->d_bits, h->max_bits, h->d_nbits, nullptr,
                                               nullptr, s)))
        return rc;
    if ((rc = b200ais_hdlc_work_dev(h->hd, h->d_bits, (size_t)h->max_bits, h->d_nbits, 0, h->d_frames,
                                    h->cfg.max_frames, h->d_nframes, s)))
        return rc;
    B200_CU(cudaMemsetAsync(count, 0, sizeof(int), s));
    if ((rc = launch_gather_frames(h->d_frames, h->d_nframes, h->channels, h->cfg.max_frames, msgs,
                                   max_msgs, count, h->d_status, s)))
        return rc;
    return launch_nmea_dense(msgs, count, max_msgs, h->d_des, h->cfg.nfreqs, sentences, slot, lens, s);
}

extern "C" int b200ais_rx_work_dev(b200ais_rx *h, const float *iq, size_t iq_stride, int nitems,
                                   b200ais_frame *msgs, char *sentences, int slot, int *lens,
                                   int max_msgs, int *nmsgs, void *stream)
{
    if (!h || (!iq && nitems > 0) || !msgs || !sentences || !lens || !nmsgs || max_msgs < 1 ||
        nitems < 0 || nitems > h->cfg.max_input_items || slot < h->slot) {
        set_error("rx_work: bad arguments (nitems <= max_input_items, slot >= b200ais_rx_sentence_slot)");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (nitems > 0)
        B200_CU(cudaMemcpy2DAsync(h->d_x + h->carry, h->x_stride * sizeof(float2), iq,
                                  iq_stride * sizeof(float2), (size_t)nitems * sizeof(float2),
                                  (size_t)h->cfg.sources, cudaMemcpyDeviceToDevice, s));
    return rx_run(h, nitems, msgs, sentences, slot, lens, max_msgs, nmsgs, s);
}

extern "C" int b200ais_rx_status(b200ais_rx *h)
{
    if (!h)
        return B200AIS_E_INVALID;
    int st = 0;
    B200_CU(cudaMemcpy(&st, h->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (st) {
        B200_CU(cudaMemset(h->d_status, 0, sizeof(int)));
        set_error("rx: more messages than max_msgs in one call");
        return st;
    }
    int rc = b200ais_hdlc_status(h->hd);
    if (rc)
        return rc;
    rc = b200ais_demod_status(h->dm);
    if (rc == B200AIS_E_TAG_OVERFLOW) { // tags beyond the row were dropped: the messages stand
        h->tag_overflows++;
        rc = B200AIS_OK;
    }
    return rc;
}

extern "C" uint64_t b200ais_rx_tag_overflows(const b200ais_rx *h) { return h ? h->tag_overflows : 0; }

extern "C" int b200ais_rx_work(b200ais_rx *h, const float *iq, size_t iq_stride, int nitems,
                               b200ais_frame *msgs, char *sentences, int slot, int *lens,
                               int max_msgs, int *nmsgs)
{
    if (!h || (!iq && nitems > 0) || !msgs || !sentences || !lens || !nmsgs || max_msgs < 1 ||
        nitems < 0 || nitems > h->cfg.max_input_items || slot < h->slot) {
        set_error("rx_work: bad arguments (nitems <= max_input_items, slot >= b200ais_rx_sentence_slot)");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = h->stream;
    if (h->out_cap < max_msgs) {
        if (h->d_msgs) cudaFree(h->d_msgs);
        if (h->d_sent) cudaFree(h->d_sent);
        if (h->d_lens) cudaFree(h->d_lens);
        h->d_msgs = nullptr;
        h->d_sent = nullptr;
        h->d_lens = nullptr;
        h->out_cap = 0;
        B200_CU(cudaMalloc(&h->d_msgs, sizeof(b200ais_frame) * (size_t)max_msgs));
        B200_CU(cudaMalloc(&h->d_sent, (size_t)h->slot * max_msgs));
        B200_CU(cudaMalloc(&h->d_lens, sizeof(int) * (size_t)max_msgs));
        h->out_cap = max_msgs;
    }
    if (nitems > 0)
        B200_CU(cudaMemcpy2DAsync(h->d_x + h->carry, h->x_stride * sizeof(float2), iq,
                                  iq_stride * sizeof(float2), (size_t)nitems * sizeof(float2),
                                  (size_t)h->cfg.sources, cudaMemcpyHostToDevice, s));
    int rc = rx_run(h, nitems, h->d_msgs, h->d_sent, h->slot, h->d_lens, max_msgs, h->d_count, s);
    if (rc)
        return rc;
    int n = 0;
    B200_CU(cudaMemcpyAsync(&n, h->d_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CU(cudaStreamSynchronize(s));
    if ((rc = b200ais_rx_status(h)))
        return rc;
    *nmsgs = n;
    if (n > 0) {
        B200_CU(cudaMemcpyAsync(msgs, h->d_msgs, sizeof(b200ais_frame) * (size_t)n, cudaMemcpyDeviceToHost, s));
        B200_CU(cudaMemcpyAsync(lens, h->d_lens, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
        B200_CU(cudaMemcpy2DAsync(sentences, (size_t)slot, h->d_sent, (size_t)h->slot, (size_t)h->slot,
                                  (size_t)n, cudaMemcpyDeviceToHost, s));
        B200_CU(cudaStreamSynchronize(s));
    }
    return B200AIS_OK;
}

// ------------------------------------------------------------ recorded-IQ replay

namespace {

// one capture to every source row: dst[s][i] = src[i]
__global__ void __launch_bounds__(256)
k_fanout(const float2 *__restrict__ src, int n, float2 *__restrict__ dst, size_t dst_stride, int sources)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float2 v = src[i];
    for (int s = blockIdx.y; s < sources; s += gridDim.y)
        dst[(size_t)s * dst_stride + i] = v;
}

// a source of raw interleaved float32 IQ: fills dst with up to `want` items, returns how many it
// got (0 = end of the capture)
struct ItemReader {
    virtual ~ItemReader() {}
    virtual size_t read(float2 *dst, size_t want) = 0;
};

struct FileReader : ItemReader {
    FILE *f = nullptr;
    ~FileReader() override
    {
        if (f)
            fclose(f);
    }
    size_t read(float2 *dst, size_t want) override { return fread(dst, sizeof(float2), want, f); }
};

// blocks.udp_source(gr.sizeof_gr_complex, ip, port, payload_size, eof) [G]: datagram payloads
// are a byte stream of items; a zero-length datagram is the end-of-stream mark (eof = True);
// idle_ms without traffic also ends the capture.  Bytes short of a whole item wait for the
// next datagram.
struct UdpReader : ItemReader {
    int fd = -1;
    int idle_ms = 1000;
    bool eof = false;
    std::vector<unsigned char> pend; // bytes beyond the items already handed out
    ~UdpReader() override
    {
        if (fd >= 0)
            close(fd);
    }
    size_t read(float2 *dst, size_t want) override
    {
        unsigned char *out = reinterpret_cast<unsigned char *>(dst);
        const size_t want_b = want * sizeof(float2);
        size_t have = std::min(pend.size(), want_b);
        memcpy(out, pend.data(), have);
        pend.erase(pend.begin(), pend.begin() + (long)have);
        unsigned char dgram[65536];
        while (have < want_b && !eof) {
            pollfd pf = {fd, POLLIN, 0};
            const int pr = poll(&pf, 1, idle_ms);
            if (pr < 0 && errno == EINTR)
                continue; // a signal, not the end of the capture
            if (pr <= 0)
                break; // idle (or error): the capture is over
            const ssize_t n = recv(fd, dgram, sizeof(dgram), 0);
            if (n < 0 && (errno == EINTR || errno == EAGAIN))
                continue;
            if (n < 0)
                break;
            if (n == 0) {
                eof = true;
                break;
            }
            const size_t take = std::min((size_t)n, want_b - have);
            memcpy(out + have, dgram, take);
            have += take;
            pend.insert(pend.end(), dgram + take, dgram + n);
        }
        const size_t items = have / sizeof(float2), rem = have % sizeof(float2);
        if (rem) // keep the split item for the next call
            pend.insert(pend.begin(), out + items * sizeof(float2), out + have);
        return items;
    }
};

size_t reader_read(ItemReader *r, float2 *dst, size_t want) { return r->read(dst, want); }

} // namespace

// blocks.file_source(gr.sizeof_gr_complex, filename) (python/radio.py:211-213) feeding every
// source of the receiver with the same recorded capture (replay fan-out): raw interleaved
// float32 IQ, read in chunks through two pinned buffers -- the read of chunk k+1 runs on a
// host thread while chunk k is copied and processed -- copied to the device once and
// replicated there.
// Chunks from `rd` through two pinned buffers: chunk k+1 is read on a host thread while chunk k
// is copied (once) to the device, replicated to every source row and processed.
static int rx_pump(b200ais_rx *h, ItemReader *rd, int chunk_items, int max_msgs, b200ais_rx_sink sink,
                   void *user, uint64_t *items_read, uint64_t max_items)
{
    float2 *pin[2] = {nullptr, nullptr}, *d_stage = nullptr;
    std::vector<b200ais_frame> msgs((size_t)max_msgs);
    std::vector<char> sent((size_t)max_msgs * h->slot);
    std::vector<int> lens((size_t)max_msgs);
    int rc = B200AIS_OK;
    uint64_t total = 0;
    cudaError_t e = cudaMallocHost(&pin[0], sizeof(float2) * (size_t)chunk_items);
    if (e == cudaSuccess) e = cudaMallocHost(&pin[1], sizeof(float2) * (size_t)chunk_items);
    if (e == cudaSuccess) e = cudaMalloc(&d_stage, sizeof(float2) * (size_t)chunk_items);
    if (e == cudaSuccess && h->out_cap < max_msgs) {
        if (h->d_msgs) cudaFree(h->d_msgs);
        if (h->d_sent) cudaFree(h->d_sent);
        if (h->d_lens) cudaFree(h->d_lens);
        h->d_msgs = nullptr;
        h->d_sent = nullptr;
        h->d_lens = nullptr;
        h->out_cap = 0;
        e = cudaMalloc(&h->d_msgs, sizeof(b200ais_frame) * (size_t)max_msgs);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_sent, (size_t)h->slot * max_msgs);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_lens, sizeof(int) * (size_t)max_msgs);
        if (e == cudaSuccess) h->out_cap = max_msgs;
    }
    if (e != cudaSuccess)
        rc = cuda_fail(e, "rx_pump", __FILE__, __LINE__);
    cudaStream_t s = h->stream;
    const int S = h->cfg.sources;
    auto want = [&](uint64_t done) {
        return (size_t)std::min<uint64_t>((uint64_t)chunk_items, max_items - done);
    };
    size_t have = rc ? 0 : rd->read(pin[0], want(0));
    uint64_t asked = have;
    for (int k = 0; !rc && have > 0; k++) {
        float2 *cur = pin[k & 1];
        std::future<size_t> next = std::async(std::launch::async, reader_read, rd, pin[(k + 1) & 1],
                                              want(asked));
        const int n = (int)have;
        e = cudaMemcpyAsync(d_stage, cur, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            dim3 grid((unsigned)((n + 255) / 256), (unsigned)std::min(S, 64));
            k_fanout<<<grid, 256, 0, s>>>(d_stage, n, h->d_x + h->carry, h->x_stride, S);
            count_launch();
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "rx_pump", __FILE__, __LINE__);
        } else {
            rc = rx_run(h, n, h->d_msgs, h->d_sent, h->slot, h->d_lens, max_msgs, h->d_count, s);
        }
        int cnt = 0;
        if (!rc) {
            e = cudaMemcpyAsync(&cnt, h->d_count, sizeof(int), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess)
                rc = cuda_fail(e, "rx_pump", __FILE__, __LINE__);
        }
        if (!rc)
            rc = b200ais_rx_status(h);
        if (!rc && cnt > 0) {
            e = cudaMemcpy(msgs.data(), h->d_msgs, sizeof(b200ais_frame) * (size_t)cnt, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess)
                e = cudaMemcpy(lens.data(), h->d_lens, sizeof(int) * (size_t)cnt, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess)
                e = cudaMemcpy(sent.data(), h->d_sent, (size_t)h->slot * cnt, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess)
                rc = cuda_fail(e, "rx_pump", __FILE__, __LINE__);
            else if (sink)
                sink(user, msgs.data(), sent.data(), h->slot, lens.data(), cnt);
        }
        total += (uint64_t)n;
        have = next.get(); // always joined, also on errors
        asked += have;
    }
    if (items_read)
        *items_read = total;
    if (pin[0]) cudaFreeHost(pin[0]);
    if (pin[1]) cudaFreeHost(pin[1]);
    if (d_stage) cudaFree(d_stage);
    return rc;
}

extern "C" int b200ais_rx_replay_file(b200ais_rx *h, const char *path, int chunk_items, int max_msgs,
                                      b200ais_rx_sink sink, void *user, uint64_t *items_read)
{
    if (!h || !path || chunk_items < 1 || chunk_items > h->cfg.max_input_items || max_msgs < 1) {
        set_error("rx_replay_file: bad arguments (chunk_items <= max_input_items)");
        return B200AIS_E_INVALID;
    }
    FileReader rd;
    rd.f = fopen(path, "rb");
    if (!rd.f) {
        set_error("rx_replay_file: cannot open %s", path);
        return B200AIS_E_INVALID;
    }
    return rx_pump(h, &rd, chunk_items, max_msgs, sink, user, items_read, ~(uint64_t)0);
}

// blocks.udp_source(gr.sizeof_gr_complex, ip, port) (python/radio.py:204-210)
extern "C" int b200ais_rx_serve_udp(b200ais_rx *h, const char *bind_ip, int port, int chunk_items,
                                    int max_msgs, uint64_t max_items, int idle_ms,
                                    b200ais_rx_sink sink, void *user, uint64_t *items_read)
{
    if (!h || !bind_ip || port < 1 || port > 65535 || chunk_items < 1 ||
        chunk_items > h->cfg.max_input_items || max_msgs < 1 || idle_ms < 1) {
        set_error("rx_serve_udp: bad arguments (chunk_items <= max_input_items)");
        return B200AIS_E_INVALID;
    }
    UdpReader rd;
    rd.idle_ms = idle_ms;
    rd.fd = socket(AF_INET, SOCK_DGRAM, 0);
    sockaddr_in addr;
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)port);
    const int buf = 8 << 20;
    if (rd.fd >= 0)
        setsockopt(rd.fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof(buf));
    if (rd.fd < 0 || inet_pton(AF_INET, bind_ip, &addr.sin_addr) != 1 ||
        bind(rd.fd, reinterpret_cast<sockaddr *>(&addr), sizeof(addr)) != 0) {
        set_error("rx_serve_udp: cannot bind %s:%d", bind_ip, port);
        return B200AIS_E_INVALID;
    }
    return rx_pump(h, &rd, chunk_items, max_msgs, sink, user, items_read,
                   max_items ? max_items : ~(uint64_t)0);
}
