/*
 * ais_oracle_rx.c -- CPU restatement ("oracle") of the blocks either side of the demod
 * path in the reference's receiver (python/radio.py:39-72):
 *
 *   firdes.low_pass + freq_xlating_fir_filter_ccf   python/radio.py:49-54      [G]
 *   digital.hdlc_deframer_bp(11, 64)                python/radio.py:64         [G]
 *   gr::ais::pdu_to_nmea                            lib/pdu_to_nmea_impl.cc:63-131 [R]
 *
 * TEST INFRASTRUCTURE ONLY (see ais_oracle.h).  PARITY UNPINNED for the [G] parts: GNU
 * Radio 3.8 is absent from this image, so they are restated from the published algorithms;
 * the [R] part follows the reference source line by line and is pinned by public AIVDM
 * sentences (tests/golden/aivdm_kat.json: each one carries its own NMEA checksum).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ais_oracle.h"

/* ------------------------------------------------------------------ CRC */

/* hdlc_deframer_bp_impl::crc_ccitt [G]: CRC-16/X.25, reflected poly 0x8408, init and
 * final xor 0xFFFF */
unsigned ao_crc_ccitt(const uint8_t *data, size_t len)
{
    unsigned short crc = 0xFFFF;
    for (size_t i = 0; i < len; i++) {
        crc ^= (unsigned short)data[i];
        for (int j = 0; j < 8; j++)
            crc = (crc & 1) ? (unsigned short)((crc >> 1) ^ 0x8408) : (unsigned short)(crc >> 1);
    }
    return (unsigned)(crc ^ 0xFFFF) & 0xFFFFu;
}

/* ---------------------------------------------------- hdlc_deframer_bp */

void ao_hdlc_init(ao_hdlc *h, int length_min, int length_max)
{
    memset(h, 0, sizeof(*h));
    h->length_min = length_min;
    h->length_max = length_max;
}

/* hdlc_deframer_bp_impl::work [G]: one unpacked bit per item.  A 1 after five or more 1s is
 * a frame delimiter: a frame of d_bytectr >= length_min bytes whose last two bytes equal the
 * CRC of the others is published (without the CRC); a 0 after five 1s is a stuffed bit and is
 * dropped; any other bit is shifted into the current byte LSB first; a frame that grows past
 * length_max bytes is dropped (that bit is lost).  Returns the number of frames written
 * (frames beyond max_frames are counted in *dropped). */
int ao_hdlc_work(ao_hdlc *h, const uint8_t *bits, int n, ao_frame *frames, int max_frames,
                 int *dropped)
{
    int nf = 0;
    if (dropped)
        *dropped = 0;
    for (int i = 0; i < n; i++) {
        unsigned char bit = bits[i];
        if (h->ones >= 5) {
            if (bit) { /* six ones: frame delimiter */
                if (h->bytectr >= h->length_min) {
                    int len = h->bytectr - 2;
                    unsigned crc = ao_crc_ccitt(h->pktbuf, (size_t)len);
                    unsigned got = (unsigned)h->pktbuf[len] | ((unsigned)h->pktbuf[len + 1] << 8);
                    if (crc == got) {
                        if (nf < max_frames) {
                            frames[nf].end_bit = h->nitems_read + (uint64_t)i;
                            frames[nf].len = len;
                            frames[nf].channel = 0;
                            memset(frames[nf].data, 0, sizeof(frames[nf].data));
                            memcpy(frames[nf].data, h->pktbuf, (size_t)len);
                            nf++;
                        } else if (dropped) {
                            (*dropped)++;
                        }
                    }
                }
                h->bitctr = 0;
                h->bytectr = 0;
            } /* else: unstuff */
        } else {
            if (h->bytectr > h->length_max) {
                h->bytectr = 0;
                h->bitctr = 0;
            } else {
                h->pktbuf[h->bytectr] >>= 1;
                if (bit)
                    h->pktbuf[h->bytectr] |= 0x80;
                h->bitctr++;
                if (h->bitctr == 8) {
                    h->bitctr = 0;
                    h->bytectr++;
                }
            }
        }
        h->ones = bit ? h->ones + 1 : 0;
    }
    h->nitems_read += (uint64_t)n;
    return nf;
}

/* ----------------------------------------------------------- pdu_to_nmea */

/* pdu_to_nmea_impl::msg_to_sentence (lib/pdu_to_nmea_impl.cc:63-131) [R].
 * unpack_bits :63-79: bytes MSB first into 6-bit groups, npad = (6 - nbits % 6) % 6, and the
 *   last group is then shifted left npad MORE times in a uint8_t although its bits are already
 *   left-aligned (":75-77; TODO: test with padding more thoroughly") -- bits fall off the top;
 * to_ascii :81-88: on `char` (signed here): > 39 adds 8, then adds 48;
 * to_sentence :99-125: fragments of 56 characters, "!AIVDM,<n>,<i>,,<designator>,<frag>,<npad>",
 *   checksum = xor of everything after '!' (:90-96), "*%02X", fragments joined by '\n'.
 * Returns the sentence length (no terminator counted), or -1 if it does not fit / len < 1. */
int ao_pdu_to_nmea(const char *designator, const uint8_t *data, int len, char *out, int cap)
{
    if (len < 1)
        return -1;
    int nbits = len * 8;
    int npad = (6 - (nbits % 6)) % 6;
    int nchar = (nbits + npad) / 6;
    uint8_t *up = (uint8_t *)calloc((size_t)nchar, 1);
    for (int i = 0; i < nbits; i++) {
        uint8_t bit = (uint8_t)((data[i / 8] >> (7 - (i % 8))) & 1);
        up[i / 6] |= (uint8_t)(bit << (5 - (i % 6)));
    }
    for (int i = 0; i < npad; i++)
        up[nbits / 6] = (uint8_t)(up[nbits / 6] << 1);
    for (int i = 0; i < nchar; i++) {
        signed char c = (signed char)up[i];
        if (c > 39)
            c = (signed char)(c + 8);
        c = (signed char)(c + 48);
        up[i] = (uint8_t)c;
    }
    const int nmea_max = 56;
    const int num_frags = 1 + ((nchar - 1) / nmea_max);
    int pos = 0, frag_offset = 0;
    for (int frag_id = 1; frag_id <= num_frags; frag_id++) {
        char head[64];
        int hl = snprintf(head, sizeof(head), "!AIVDM,%d,%d,,%s,", num_frags, frag_id, designator);
        int fl = nchar - frag_offset < nmea_max ? nchar - frag_offset : nmea_max;
        char tailb[16];
        int tl = snprintf(tailb, sizeof(tailb), ",%d", npad);
        int need = (frag_id > 1) + hl + fl + tl + 3;
        if (hl < 0 || hl >= (int)sizeof(head) || pos + need > cap) {
            free(up);
            return -1;
        }
        if (frag_id > 1)
            out[pos++] = '\n';
        int start = pos;
        memcpy(out + pos, head, (size_t)hl);
        pos += hl;
        memcpy(out + pos, up + frag_offset, (size_t)fl);
        pos += fl;
        frag_offset += fl;
        memcpy(out + pos, tailb, (size_t)tl);
        pos += tl;
        uint8_t sum = 0;
        for (int i = start + 1; i < pos; i++)
            sum ^= (uint8_t)out[i];
        static const char hex[] = "0123456789ABCDEF";
        out[pos++] = '*';
        out[pos++] = hex[sum >> 4];
        out[pos++] = hex[sum & 15];
    }
    free(up);
    if (pos < cap)
        out[pos] = 0;
    return pos;
}

/* ------------------------------------------------------ firdes::low_pass */

/* gr::filter::firdes::low_pass(gain, fs, fc, tw, WIN_HAMMING) [G]: ntaps =
 * (int)(53 * fs / (22 * tw)) made odd; Hamming window 0.54 - 0.46 cos(2 pi n / (ntaps-1))
 * held in float; taps = sinc * window evaluated in double and stored as float; normalised
 * to unit DC gain with the sum accumulated in double over the float taps.
 * Returns ntaps (taps may be NULL to query the length), or -1 if cap is too small. */
int ao_firdes_low_pass(double gain, double fs, double cutoff, double tw, float *taps, int cap)
{
    int ntaps = (int)(53.0 * fs / (22.0 * tw));
    if ((ntaps & 1) == 0)
        ntaps++;
    if (!taps)
        return ntaps;
    if (cap < ntaps)
        return -1;
    int M = (ntaps - 1) / 2;
    double fwT0 = 2 * M_PI * cutoff / fs;
    for (int n = -M; n <= M; n++) {
        float w = (float)(0.54 - 0.46 * cos((2 * M_PI * (n + M)) / (ntaps - 1)));
        if (n == 0)
            taps[n + M] = (float)(fwT0 / M_PI * w);
        else
            taps[n + M] = (float)(sin(n * fwT0) / (n * M_PI) * w);
    }
    double fmax = taps[0 + M];
    for (int n = 1; n <= M; n++)
        fmax += 2 * taps[n + M];
    gain /= fmax;
    for (int i = 0; i < ntaps; i++)
        taps[i] = (float)(taps[i] * gain);
    return ntaps;
}

/* ------------------------------------------- freq_xlating_fir_filter_ccf */

/* freq_xlating_fir_filter_ccf(decimation, taps, center_freq, sampling_freq) [G]:
 * build_composite_fir: fwT0 = (float)(2 pi center_freq / sampling_freq);
 *   ctaps[i] = taps[i] * exp(j * (i * fwT0))   (band-pass at center_freq);
 *   rotator increment exp(-j * fwT0 * decimation), normalised by its magnitude;
 * work: out[j] = rotator.rotate(fir.filter(&in[j * decimation])), history = ntaps. */
int ao_xlat_init(ao_xlat *x, int decimation, const float *taps, int ntaps, double center_freq,
                 double sampling_freq)
{
    memset(x, 0, sizeof(*x));
    if (decimation < 1 || ntaps < 1)
        return -1;
    x->decim = decimation;
    x->ntaps = ntaps;
    x->ctaps = (float *)malloc(sizeof(float) * 2 * (size_t)ntaps);
    if (!x->ctaps)
        return -1;
    float fwT0 = (float)(2 * M_PI * center_freq / sampling_freq);
    for (int i = 0; i < ntaps; i++) {
        float th = (float)i * fwT0;
        x->ctaps[2 * i] = taps[i] * cosf(th);
        x->ctaps[2 * i + 1] = taps[i] * sinf(th);
    }
    float th = -fwT0 * (float)decimation;
    float ir = cosf(th), ii = sinf(th);
    float a = ao_hypotf(ir, ii);
    x->incr_re = ir / a;
    x->incr_im = ii / a;
    x->phase_re = 1.0f;
    x->phase_im = 0.0f;
    x->counter = 0;
    return 0;
}

void ao_xlat_free(ao_xlat *x)
{
    free(x->ctaps);
    x->ctaps = NULL;
}

/* One work() call: in holds ntaps-1 history items followed by noutput*decim new ones.
 * Canonical dot product (VOLK's order is machine dependent): the taps
 * are visited polyphase-major -- for p in [0, D): for k = p, p+D, p+2D, ... -- on four
 * real fused-multiply-add chains
 *     Pr += cr*xr   Pi += cr*xi   Qr += ci*xr   Qi += ci*xi,     y = (Pr - Qi, Pi + Qr)
 * with x = in[j*D + ntaps-1-k].  gr::blocks::rotator [G]: z = y * phase (std::complex
 * product, separate multiplies and adds); phase *= incr; every 512th item phase /= |phase|.
 * fir_out (nullable): the filter output before the rotator. */
int ao_xlat_work(ao_xlat *x, int noutput, const float *in, float *out, float *fir_out)
{
    const int D = x->decim, nt = x->ntaps;
    for (int j = 0; j < noutput; j++) {
        const float *base = in + 2 * ((size_t)j * D + nt - 1);
        float Pr = 0, Pi = 0, Qr = 0, Qi = 0;
        for (int p = 0; p < D; p++)
            for (int k = p; k < nt; k += D) {
                float cr = x->ctaps[2 * k], ci = x->ctaps[2 * k + 1];
                float xr = base[-2 * k], xi = base[-2 * k + 1];
                Pr = fmaf(cr, xr, Pr);
                Pi = fmaf(cr, xi, Pi);
                Qr = fmaf(ci, xr, Qr);
                Qi = fmaf(ci, xi, Qi);
            }
        float yr = Pr - Qi, yi = Pi + Qr;
        if (fir_out) {
            fir_out[2 * j] = yr;
            fir_out[2 * j + 1] = yi;
        }
        x->counter++;
        float pr = x->phase_re, pi = x->phase_im;
        out[2 * j] = yr * pr - yi * pi;
        out[2 * j + 1] = yr * pi + yi * pr;
        float nr = pr * x->incr_re - pi * x->incr_im;
        float ni = pr * x->incr_im + pi * x->incr_re;
        if ((x->counter % 512) == 0) {
            float a = ao_hypotf(nr, ni);
            nr = nr / a;
            ni = ni / a;
        }
        x->phase_re = nr;
        x->phase_im = ni;
    }
    return noutput;
}

/* float64 truth of the same filter for tolerance tests: direct sum in tap order, the float
 * fwT0 GNU Radio uses, exact rotator angle -fwT0*D*j */
void ao_xlat_f64(int decimation, const float *taps, int ntaps, double center_freq,
                 double sampling_freq, uint64_t first_output, int noutput, const float *in,
                 double *out)
{
    double w = (double)(float)(2 * M_PI * center_freq / sampling_freq);
    for (int j = 0; j < noutput; j++) {
        const float *base = in + 2 * ((size_t)j * decimation + ntaps - 1);
        double re = 0, im = 0;
        for (int k = 0; k < ntaps; k++) {
            double cr = taps[k] * cos(k * w), ci = taps[k] * sin(k * w);
            double xr = base[-2 * k], xi = base[-2 * k + 1];
            re += cr * xr - ci * xi;
            im += cr * xi + ci * xr;
        }
        double th = -w * decimation * (double)(first_output + (uint64_t)j);
        double c = cos(th), s = sin(th);
        out[2 * j] = re * c - im * s;
        out[2 * j + 1] = re * s + im * c;
    }
}
