// framing.cu -- the step after the demod path in the reference's receiver
// (python/radio.py:64-72): digital.hdlc_deframer_bp(11, 64) and gr::ais::pdu_to_nmea
// (lib/pdu_to_nmea_impl.cc:63-131), batched over channels, with their C-ABI.
//
// Byte/integer work on a 9600 bit/s stream per channel: one lane per channel walks its row of
// unpacked bits 16 at a time (one 128-bit load), the partial frame lives in the lane's local
// memory between delimiters, and the per-channel deframer state is carried in HBM between calls.
#include <cstring>
#include <new>

#include "internal.h"

using namespace b200ais;

namespace {

struct HdlcState {
    int ones, bitctr, bytectr;
    unsigned crcs; // running CRC register (low half) and its value at the last byte boundary
                   // (high half), both ^ 0xFFFF so that a zeroed state means "fresh"
    unsigned long long nitems_read;
    unsigned char pktbuf[B200AIS_FRAME_MAX + 8];
};

// A delimiter closed a frame whose CRC is good: publish it.  Out of line and by value, so the
// bit loop's state stays in registers.  Returns the new frame count, or -1 when the row of
// frames is full.
__device__ __noinline__ int hdlc_emit(const unsigned char *buf, int bytectr, b200ais_frame *frames,
                                      int nf, int max_frames, int channel, unsigned long long end_bit)
{
    const int len = bytectr - 2;
    if (nf >= max_frames)
        return -1;
    b200ais_frame *f = frames + nf;
    f->end_bit = end_bit;
    f->len = len;
    f->channel = channel;
    for (int k = 0; k < len; k++)
        f->data[k] = buf[k];
    for (int k = len; k < B200AIS_FRAME_MAX; k++)
        f->data[k] = 0;
    return nf + 1;
}

// hdlc_deframer_bp_impl::work [G], one bit.
//  * The partial byte lives in a register (`cur`): the reference shifts it in place in
//    d_pktbuf, and stale high bits leave it the same way.
//  * The CRC runs with the bits instead of over the buffer at every delimiter (noise closes a
//    "frame" every ~250 bits, and a lane walking 20-odd bytes stalls its whole warp): frame
//    bits arrive LSB first, which is the reflected CRC's own bit order, so the register takes
//    one shift/xor per stored bit; `crcb` is its value at the last byte boundary (the bits after
//    it are the closing flag's).  crc_ccitt(data) == the two bytes that follow, the reference's
//    test, holds exactly when the register over data + those two bytes is the CRC-16/X.25
//    residue 0xF0B8 (for a given prefix the last 16 bits map one-to-one onto the register).
// The step is written with selects: lanes of a warp sit in different states, and as branches
// every path would run for every bit.  Only a good frame (rare) leaves the straight line.
#define HDLC_STEP(bit_, i_)                                                                  \
    do {                                                                                     \
        const unsigned b__ = (bit_);                                                         \
        const bool five__ = ones >= 5;                                                       \
        const bool delim__ = five__ && b__;             /* six ones: frame delimiter */      \
        const bool over__ = !five__ && bytectr > length_max;                                 \
        const bool shift__ = !five__ && !over__;        /* else: stuffed zero, dropped */    \
        if (delim__ && bytectr >= length_min && crcb == 0xF0B8u) {                           \
            const int r__ = hdlc_emit(buf, bytectr, myframes, nf, max_frames, c,             \
                                      base + (unsigned long long)(i_));                      \
            overflow |= r__ < 0;                                                             \
            nf = r__ < 0 ? nf : r__;                                                         \
        }                                                                                    \
        const unsigned cur_n__ = (cur >> 1) | (b__ << 7);                                    \
        const unsigned crc_n__ = (crc >> 1) ^ (((crc ^ b__) & 1u) ? 0x8408u : 0u);           \
        cur = shift__ ? cur_n__ : cur;                                                       \
        crc = shift__ ? crc_n__ : crc;                                                       \
        bitctr += shift__ ? 1 : 0;                                                           \
        const bool full__ = shift__ && bitctr == 8;                                          \
        if (full__)                                                                          \
            buf[bytectr] = (unsigned char)cur;                                               \
        bytectr += full__ ? 1 : 0;                                                           \
        bitctr = full__ ? 0 : bitctr;                                                        \
        crcb = full__ ? crc : crcb;                                                          \
        const bool rst__ = delim__ || over__;                                                \
        bytectr = rst__ ? 0 : bytectr;                                                       \
        bitctr = rst__ ? 0 : bitctr;                                                         \
        crc = rst__ ? 0xFFFFu : crc;                                                         \
        crcb = rst__ ? 0xFFFFu : crcb;                                                       \
        ones = b__ ? ones + 1 : 0;                                                           \
    } while (0)

// sixteen unpacked bits, one per byte of w, as a 16-bit mask (bit k = k-th bit of the stream)
__device__ __forceinline__ unsigned pack16(uint4 w)
{
    return (((w.x & 0x01010101u) * 0x01020408u) >> 24) |
           ((((w.y & 0x01010101u) * 0x01020408u) >> 24) << 4) |
           ((((w.z & 0x01010101u) * 0x01020408u) >> 24) << 8) |
           ((((w.w & 0x01010101u) * 0x01020408u) >> 24) << 12);
}

// One lane per channel.  Rows are read 16 bits (one 128-bit load) at a time, the next load
// issued before the current bits are walked.
__global__ void __launch_bounds__(32)
k_hdlc(const uint8_t *__restrict__ bits, size_t bits_stride, const int *__restrict__ nbits,
       int nbits_all, int channels, HdlcState *__restrict__ state, int length_min, int length_max,
       b200ais_frame *__restrict__ frames, int max_frames, int *__restrict__ nframes, int *status)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    HdlcState *st = state + c;
    unsigned char buf[B200AIS_FRAME_MAX + 8];
    int ones = st->ones, bitctr = st->bitctr, bytectr = st->bytectr, nf = 0;
    bool overflow = false;
    const unsigned long long base = st->nitems_read;
    b200ais_frame *myframes = frames + (size_t)c * max_frames;
    for (int k = 0; k < bytectr; k++)
        buf[k] = st->pktbuf[k];
    unsigned cur = st->pktbuf[bytectr];
    unsigned crc = (st->crcs & 0xFFFFu) ^ 0xFFFFu, crcb = (st->crcs >> 16) ^ 0xFFFFu;

    const int n = nbits ? nbits[c] : nbits_all;
    const uint8_t *row = bits + (size_t)c * bits_stride;
    int i = 0;
    while (i < n && ((reinterpret_cast<uintptr_t>(row + i)) & 15)) {
        HDLC_STEP(row[i] & 1u, i);
        i++;
    }
    if (i + 16 <= n) {
        uint4 w = *reinterpret_cast<const uint4 *>(row + i);
        for (; i + 32 <= n; i += 16) {
            const uint4 nx = *reinterpret_cast<const uint4 *>(row + i + 16);
            const unsigned m = pack16(w);
#pragma unroll
            for (int k = 0; k < 16; k++)
                HDLC_STEP((m >> k) & 1u, i + k);
            w = nx;
        }
        const unsigned m = pack16(w);
#pragma unroll 1
        for (int k = 0; k < 16; k++)
            HDLC_STEP((m >> k) & 1u, i + k);
        i += 16;
    }
#pragma unroll 1
    for (; i < n; i++)
        HDLC_STEP(row[i] & 1u, i);

    st->ones = ones;
    st->bitctr = bitctr;
    st->bytectr = bytectr;
    st->crcs = ((crc ^ 0xFFFFu) & 0xFFFFu) | ((crcb ^ 0xFFFFu) << 16);
    st->nitems_read = base + (unsigned long long)(n > 0 ? n : 0);
    for (int k = 0; k < bytectr; k++)
        st->pktbuf[k] = buf[k];
    st->pktbuf[bytectr] = (unsigned char)cur;
    nframes[c] = nf;
    if (overflow)
        atomicExch(status, B200AIS_E_FRAME_OVERFLOW);
}
#undef HDLC_STEP

__device__ __forceinline__ int put_int(char *o, int v)
{
    char tmp[12];
    int n = 0;
    if (v == 0)
        tmp[n++] = '0';
    while (v > 0) {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    }
    for (int k = 0; k < n; k++)
        o[k] = tmp[n - 1 - k];
    return n;
}

// pdu_to_nmea_impl::msg_to_sentence (lib/pdu_to_nmea_impl.cc:63-131), one lane per frame.
__global__ void __launch_bounds__(128)
k_nmea(const b200ais_frame *__restrict__ frames, const int *__restrict__ nframes, int channels,
       int max_frames, const char *__restrict__ designators, int des_mod,
       char *__restrict__ sentences, int slot, int *__restrict__ lens)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= channels * max_frames)
        return;
    const int c = idx / max_frames, f = idx - c * max_frames;
    if (f >= min(nframes[c], max_frames)) {
        lens[idx] = 0;
        return;
    }
    const b200ais_frame *fr = frames + idx;
    char *out = sentences + (size_t)idx * slot;
    const int len = fr->len;
    if (len < 1 || len > B200AIS_FRAME_MAX) {
        lens[idx] = 0;
        return;
    }
    char des[8];
    int dl = 0;
    if (designators) {
        const int row = des_mod > 0 ? fr->channel % des_mod : c;
        for (; dl < 8 && designators[row * 8 + dl]; dl++)
            des[dl] = designators[row * 8 + dl];
    } else {
        des[0] = 'A';
        dl = 1;
    }
    const int nbits = len * 8;
    const int npad = (6 - (nbits % 6)) % 6;          // :67
    const int nchar = (nbits + npad) / 6;
    const int num_frags = 1 + ((nchar - 1) / 56);    // :103-104
    if (num_frags * (22 + dl) + nchar > slot) {      // would not fit the caller's slot
        lens[idx] = -1;
        return;
    }
    int pos = 0, ch = 0;
    for (int frag = 1; frag <= num_frags; frag++) {
        if (frag > 1)
            out[pos++] = '\n';
        const int start = pos;
        const char head[7] = {'!', 'A', 'I', 'V', 'D', 'M', ','};
        for (int k = 0; k < 7; k++)
            out[pos++] = head[k];
        pos += put_int(out + pos, num_frags);
        out[pos++] = ',';
        pos += put_int(out + pos, frag);
        out[pos++] = ',';
        out[pos++] = ',';
        for (int k = 0; k < dl; k++)
            out[pos++] = des[k];
        out[pos++] = ',';
        const int fl = min(56, nchar - ch);
        for (int k = 0; k < fl; k++, ch++) {
            // unpack_bits :63-79: six bits MSB first starting at bit 6*ch
            unsigned v = 0;
            for (int b = 0; b < 6; b++) {
                const int i = 6 * ch + b;
                if (i < nbits)
                    v |= ((fr->data[i >> 3] >> (7 - (i & 7))) & 1u) << (5 - b);
            }
            if (npad && ch == nbits / 6)
                v = (v << npad) & 0xFFu; // :75-77 shifts the left-aligned group again (uint8_t)
            // to_ascii :81-88 on (signed) char
            int sc = (int)(signed char)v;
            if (sc > 39)
                sc = (int)(signed char)(sc + 8);
            sc = (int)(signed char)(sc + 48);
            out[pos++] = (char)sc;
        }
        out[pos++] = ',';
        pos += put_int(out + pos, npad);
        unsigned sum = 0; // get_checksum :90-96
        for (int k = start + 1; k < pos; k++)
            sum ^= (unsigned char)out[k];
        const char hex[17] = "0123456789ABCDEF";
        out[pos++] = '*';
        out[pos++] = hex[(sum >> 4) & 15];
        out[pos++] = hex[sum & 15];
    }
    if (pos < slot)
        out[pos] = 0;
    lens[idx] = pos;
}

} // namespace

// =========================================================== hdlc_deframer_bp

struct b200ais_hdlc {
    int channels = 0, length_min = 0, length_max = 0;
    cudaStream_t stream = nullptr;
    HdlcState *d_state = nullptr;
    int *d_status = nullptr;
    DevBuf bits, nbits, frames, nframes;
};

extern "C" int b200ais_hdlc_create(b200ais_hdlc **out, int length_min, int length_max, int channels)
{
    if (!out || channels < 1 || length_min < 2 || length_max < length_min ||
        length_max > B200AIS_FRAME_MAX - 2) {
        set_error("hdlc_create: need channels >= 1 and 2 <= length_min <= length_max <= %d",
                  B200AIS_FRAME_MAX - 2);
        return B200AIS_E_INVALID;
    }
    b200ais_hdlc *h = new (std::nothrow) b200ais_hdlc;
    if (!h)
        return B200AIS_E_NOMEM;
    h->channels = channels;
    h->length_min = length_min;
    h->length_max = length_max;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_state, sizeof(HdlcState) * (size_t)channels);
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_status, sizeof(int));
    if (e == cudaSuccess)
        e = cudaMemset(h->d_state, 0, sizeof(HdlcState) * (size_t)channels);
    if (e == cudaSuccess)
        e = cudaMemset(h->d_status, 0, sizeof(int));
    if (e != cudaSuccess) {
        b200ais_hdlc_destroy(h);
        return cuda_fail(e, "hdlc_create", __FILE__, __LINE__);
    }
    *out = h;
    return B200AIS_OK;
}

extern "C" int b200ais_hdlc_destroy(b200ais_hdlc *h)
{
    if (!h)
        return B200AIS_OK;
    if (h->stream)
        cudaStreamDestroy(h->stream);
    if (h->d_state)
        cudaFree(h->d_state);
    if (h->d_status)
        cudaFree(h->d_status);
    h->bits.release();
    h->nbits.release();
    h->frames.release();
    h->nframes.release();
    delete h;
    return B200AIS_OK;
}

extern "C" int b200ais_hdlc_reset(b200ais_hdlc *h)
{
    if (!h)
        return B200AIS_E_INVALID;
    B200_CU(cudaDeviceSynchronize());
    B200_CU(cudaMemset(h->d_state, 0, sizeof(HdlcState) * (size_t)h->channels));
    B200_CU(cudaMemset(h->d_status, 0, sizeof(int)));
    return B200AIS_OK;
}

extern "C" int b200ais_hdlc_work_dev(b200ais_hdlc *h, const uint8_t *bits, size_t bits_stride,
                                     const int *nbits, int nbits_all, b200ais_frame *frames,
                                     int max_frames, int *nframes, void *stream)
{
    if (!h || !bits || !frames || !nframes || max_frames < 1 || (!nbits && nbits_all < 0)) {
        set_error("hdlc_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int threads = 32;
    k_hdlc<<<(h->channels + threads - 1) / threads, threads, 0, s>>>(
        bits, bits_stride, nbits, nbits_all, h->channels, h->d_state, h->length_min, h->length_max,
        frames, max_frames, nframes, h->d_status);
    B200_LAUNCH_CHECK("k_hdlc");
    return B200AIS_OK;
}

extern "C" int b200ais_hdlc_status(b200ais_hdlc *h)
{
    if (!h)
        return B200AIS_E_INVALID;
    int st = 0;
    B200_CU(cudaMemcpy(&st, h->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (st) {
        B200_CU(cudaMemset(h->d_status, 0, sizeof(int)));
        set_error("a channel produced more HDLC frames than max_frames");
    }
    return st;
}

extern "C" int b200ais_hdlc_work(b200ais_hdlc *h, const uint8_t *bits, size_t bits_stride,
                                 const int *nbits, int nbits_all, b200ais_frame *frames,
                                 int max_frames, int *nframes)
{
    if (!h || !bits || !frames || !nframes || max_frames < 1 || (!nbits && nbits_all < 0)) {
        set_error("hdlc_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    const int C = h->channels;
    int maxn = nbits_all;
    if (nbits) {
        maxn = 0;
        for (int c = 0; c < C; c++) {
            if (nbits[c] < 0 || (size_t)nbits[c] > bits_stride) {
                set_error("hdlc_work: nbits[%d] outside the row", c);
                return B200AIS_E_INVALID;
            }
            maxn = nbits[c] > maxn ? nbits[c] : maxn;
        }
    } else if ((size_t)nbits_all > bits_stride) {
        set_error("hdlc_work: nbits_all outside the row");
        return B200AIS_E_INVALID;
    }
    const size_t dstride = ((size_t)maxn + 15) / 16 * 16 + 16;
    int rc;
    if ((rc = h->bits.reserve(dstride * C)) || (rc = h->nbits.reserve(sizeof(int) * C)) ||
        (rc = h->frames.reserve(sizeof(b200ais_frame) * (size_t)C * max_frames)) ||
        (rc = h->nframes.reserve(sizeof(int) * C)))
        return rc;
    cudaStream_t s = h->stream;
    if (maxn > 0)
        B200_CU(cudaMemcpy2DAsync(h->bits.p, dstride, bits, bits_stride, (size_t)maxn, (size_t)C,
                                  cudaMemcpyHostToDevice, s));
    if (nbits)
        B200_CU(cudaMemcpyAsync(h->nbits.p, nbits, sizeof(int) * C, cudaMemcpyHostToDevice, s));
    rc = b200ais_hdlc_work_dev(h, h->bits.as<uint8_t>(), dstride, nbits ? h->nbits.as<int>() : nullptr,
                               nbits_all, h->frames.as<b200ais_frame>(), max_frames,
                               h->nframes.as<int>(), s);
    if (rc)
        return rc;
    B200_CU(cudaMemcpyAsync(nframes, h->nframes.p, sizeof(int) * C, cudaMemcpyDeviceToHost, s));
    B200_CU(cudaStreamSynchronize(s));
    // copy only the frames that exist (row by row while few rows have any, else all at once)
    int busy = 0;
    for (int c = 0; c < C; c++)
        busy += nframes[c] > 0;
    if (busy > 64) {
        B200_CU(cudaMemcpyAsync(frames, h->frames.p, sizeof(b200ais_frame) * (size_t)C * max_frames,
                                cudaMemcpyDeviceToHost, s));
    } else {
        for (int c = 0; c < C; c++)
            if (nframes[c] > 0)
                B200_CU(cudaMemcpyAsync(frames + (size_t)c * max_frames,
                                        h->frames.as<b200ais_frame>() + (size_t)c * max_frames,
                                        sizeof(b200ais_frame) * (size_t)nframes[c],
                                        cudaMemcpyDeviceToHost, s));
    }
    B200_CU(cudaStreamSynchronize(s));
    return b200ais_hdlc_status(h);
}

// ================================================================ pdu_to_nmea

extern "C" int b200ais_nmea_slot_bytes(int max_len, const char *designator)
{
    if (max_len < 1 || max_len > B200AIS_FRAME_MAX)
        return B200AIS_E_INVALID;
    const int dl = designator ? (int)strnlen(designator, 8) : 1;
    const int nchar = (max_len * 8 + 5) / 6;
    const int frags = 1 + (nchar - 1) / 56;
    // "\n!AIVDM," + 2 x up to 2 digits + ",,," + designator + "," + npad + "*XX"
    const int per = 1 + 7 + 2 + 1 + 2 + 2 + dl + 1 + 1 + 1 + 3;
    return (frags * per + nchar + 1 + 15) / 16 * 16;
}

extern "C" int b200ais_nmea_format_dev(const b200ais_frame *frames, const int *nframes, int channels,
                                       int max_frames, const char *designators, char *sentences,
                                       int slot, int *lens, void *stream)
{
    if (!frames || !nframes || !sentences || !lens || channels < 1 || max_frames < 1 ||
        slot < 32) {
        set_error("nmea_format: bad arguments (slot: see b200ais_nmea_slot_bytes)");
        return B200AIS_E_INVALID;
    }
    const long total = (long)channels * max_frames;
    k_nmea<<<(unsigned)((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        frames, nframes, channels, max_frames, designators, 0, sentences, slot, lens);
    B200_LAUNCH_CHECK("k_nmea");
    return B200AIS_OK;
}

namespace b200ais {

// ais_rx (rx.cu): gather every channel's frames into one dense list (order of channels is
// whatever the atomics give; a channel's own frames stay in order) ...
__global__ void __launch_bounds__(128)
k_gather_frames(const b200ais_frame *__restrict__ frames, const int *__restrict__ nframes,
                int channels, int max_frames, b200ais_frame *__restrict__ dense, int max_msgs,
                int *__restrict__ count, int *status)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    const int nf = nframes[c];
    if (nf <= 0)
        return;
    const int base = atomicAdd(count, nf);
    if (base + nf > max_msgs) {
        atomicExch(status, B200AIS_E_FRAME_OVERFLOW);
        return;
    }
    static_assert(sizeof(b200ais_frame) % sizeof(uint2) == 0, "frames are copied as 8-byte words");
    const uint2 *src = reinterpret_cast<const uint2 *>(frames + (size_t)c * max_frames);
    uint2 *dst = reinterpret_cast<uint2 *>(dense + base);
    const int words = nf * (int)(sizeof(b200ais_frame) / sizeof(uint2));
    for (int k = 0; k < words; k++)
        dst[k] = src[k];
}

int launch_gather_frames(const b200ais_frame *frames, const int *nframes, int channels,
                         int max_frames, b200ais_frame *dense, int max_msgs, int *count,
                         int *status, cudaStream_t s)
{
    k_gather_frames<<<(channels + 127) / 128, 128, 0, s>>>(frames, nframes, channels, max_frames,
                                                           dense, max_msgs, count, status);
    B200_LAUNCH_CHECK("k_gather_frames");
    return B200AIS_OK;
}

// ... and format the dense list; designator row = frame.channel % des_mod
int launch_nmea_dense(const b200ais_frame *dense, const int *count, int max_msgs,
                      const char *designators, int des_mod, char *sentences, int slot, int *lens,
                      cudaStream_t s)
{
    k_nmea<<<(max_msgs + 127) / 128, 128, 0, s>>>(dense, count, 1, max_msgs, designators, des_mod,
                                                  sentences, slot, lens);
    B200_LAUNCH_CHECK("k_nmea");
    return B200AIS_OK;
}

} // namespace b200ais

extern "C" int b200ais_nmea_format(const b200ais_frame *frames, const int *nframes, int channels,
                                   int max_frames, const char *designators, char *sentences,
                                   int slot, int *lens)
{
    if (!frames || !nframes || !sentences || !lens || channels < 1 || max_frames < 1) {
        set_error("nmea_format: bad arguments");
        return B200AIS_E_INVALID;
    }
    const size_t nf = (size_t)channels * max_frames;
    void *d_frames = nullptr, *d_nframes = nullptr, *d_des = nullptr, *d_sent = nullptr,
         *d_lens = nullptr;
    int rc = B200AIS_OK;
    cudaError_t e = cudaMalloc(&d_frames, nf * sizeof(b200ais_frame));
    if (e == cudaSuccess) e = cudaMalloc(&d_nframes, sizeof(int) * channels);
    if (e == cudaSuccess) e = cudaMalloc(&d_sent, nf * (size_t)slot);
    if (e == cudaSuccess) e = cudaMemset(d_sent, 0, nf * (size_t)slot);
    if (e == cudaSuccess) e = cudaMalloc(&d_lens, nf * sizeof(int));
    if (e == cudaSuccess && designators) e = cudaMalloc(&d_des, (size_t)channels * 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_frames, frames, nf * sizeof(b200ais_frame), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_nframes, nframes, sizeof(int) * channels, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && designators)
        e = cudaMemcpy(d_des, designators, (size_t)channels * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = b200ais_nmea_format_dev((const b200ais_frame *)d_frames, (const int *)d_nframes, channels,
                                     max_frames, (const char *)d_des, (char *)d_sent, slot,
                                     (int *)d_lens, nullptr);
        if (!rc) {
            e = cudaMemcpy(sentences, d_sent, nf * (size_t)slot, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess)
                e = cudaMemcpy(lens, d_lens, nf * sizeof(int), cudaMemcpyDeviceToHost);
        }
    }
    cudaFree(d_frames);
    cudaFree(d_nframes);
    cudaFree(d_des);
    cudaFree(d_sent);
    cudaFree(d_lens);
    if (e != cudaSuccess)
        return cuda_fail(e, "nmea_format", __FILE__, __LINE__);
    return rc;
}
