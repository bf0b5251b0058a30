/*
 * ais_oracle.h -- CPU restatement ("oracle") of the gr-ais IQ-demod hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gr-ais_b200/, the
 * C-ABI library, bench.py's GPU arm) may include, link or call this.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / the CPU arm.
 *
 * PARITY STATUS.  The reference (bistromath/gr-ais @ 2162103) ships no tests,
 * fixtures or golden vectors (lib/qa_ais.cc:30-36 is an empty suite).  This file restates
 *   [R] the gr-ais C++:  lib/corr_est_cc_impl.cc:48-117,164-279,
 *       lib/msk_timing_recovery_cc_impl.cc:45-105,107-206,
 *       lib/freqest_impl.cc:41-48,57-88, lib/invert_impl.cc:54-68,
 *       lib/pdu_to_nmea_impl.cc:63-131,
 *       wiring python/ais_demod.py:28-56, python/gmsk_sync.py:22-37;
 *   [G] the GNU Radio 3.8 / VOLK kernels those sources call and the stock blocks between
 *       them, from their published algorithms -- GNU Radio itself
 *       is not installed here.
 * The [R] part is PINNED TO THE REFERENCE'S OWN CODE: oracle/ref_build compiles the five
 * reference sources unmodified (against a stub of the GNU Radio runtime whose [G] kernels
 * call the restatements below) into oracle/_ref/libais_ref.so, and tests/test_ref_pin.py
 * asserts block by block, and over the whole chain, that it and this file agree bit for
 * bit.  The [G] part stays unpinned against a real GNU Radio (no vectors exist): it is
 * held by first-principles known-answer tests and float64 truth versions of each stage.
 *
 * Canonical arithmetic (DESIGN.md "Canonical arithmetic"): IEEE-754 binary32,
 * round-to-nearest-even, no implicit contraction (-ffp-contract=off); fused
 * multiply-adds appear only where written as fmaf().
 */
#ifndef AIS_ORACLE_H
#define AIS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { AO_TAG_CORR_START = 0, AO_TAG_PHASE_EST = 1, AO_TAG_TIME_EST = 2, AO_TAG_CORR_EST = 3 };

typedef struct ao_tag {
    uint64_t offset; /* absolute item offset in corr_est output 0 */
    int32_t key;     /* AO_TAG_* */
    int32_t port;    /* output port the tag was added to (0, or 1 for the debug copies) */
    double value;    /* pmt::from_double payload */
} ao_tag;

/* ---- tables (regenerated GNU Radio data, the .inc files under oracle/tables) ---- */
const float *ao_mmse_taps(void);   /* [129][8] */
const float *ao_atan_table(void);  /* [257]    */
const float *ao_sine_table(void);  /* [1024][2] */

/* ---- scalar helpers ---- */
float ao_fast_atan2f(float y, float x);
float ao_hypotf(float re, float im);
float ao_branchless_clip(float x, float clip);
int32_t ao_float_to_fixed(float x);
void ao_fxpt_sincos(int32_t angle, float *s, float *c);
float ao_agc_envelope(float re, float im);

/* ---- A0 / G7: preamble template ---- */
/* gmsk_mod(sps, bt) applied to packed bytes (MSB first), as modulate_vector_bc
 * does (python/ais_demod.py:36-38).  out_iq holds 8*nbytes*sps complex. */
int ao_gmsk_template_packed(const uint8_t *bytes, int nbytes, int sps, float bt, float *out_iq);
/* same modulator fed with unpacked bits (north-star 24-bit template). */
int ao_gmsk_template_bits(const uint8_t *bits, int nbits, int sps, float bt, float *out_iq);
void ao_firdes_gaussian(double gain, double spb, double bt, int ntaps, float *taps);

/* ---- G1: square, FFT, shift ---- */
void ao_square(const float *x, float *out, int n);
int ao_fft_forward(const float *in, float *out, int n); /* radix-2 DIT, canonical order */
void ao_fft_shift(const float *in, float *out, int n);

/* ---- A8: freqest ---- */
typedef struct ao_freqest {
    int offset;
    float binsize;
    int fftlen;
} ao_freqest;
void ao_freqest_init(ao_freqest *f, float sample_rate, int data_rate, int fftlen);
/* one work() call over nvec input vectors; returns nvec.  maxpos_out (optional) */
int ao_freqest_work(const ao_freqest *f, const float *spec, int nvec, float *out, int *maxpos_out);

/* ---- G2: repeat + frequency_modulator_fc + mix ---- */
/* phase: NCO state carried across calls. freq[b] applies to samples [b*rep,(b+1)*rep) */
void ao_nco_mix(float *phase, float sensitivity, const float *freq, int rep, const float *x, int n,
                float *out);

/* ---- G3: feedforward_agc_cc ---- */
/* in holds n + nsamples - 1 items (history first), out n items. */
void ao_agc_work(const float *in, int n, int nsamples, float reference, float *out);

/* ---- filter::kernel::fft_filter_ccc [G] (the kernel behind corr_est_cc_impl.cc:77,84,188) ---- */
typedef struct ao_fftfilt {
    int ntaps;
    int fftsize;  /* 2 * 2^ceil(log2 ntaps) */
    int nsamples; /* fftsize - ntaps + 1: items per block = what set_taps() returns */
    float *H;     /* [fftsize] complex: transformed taps/fftsize, bit-reversed order */
    float *tail;  /* [ntaps-1] complex overlap-add tail, zero at construction, kept across calls */
    float *tw;    /* [fftsize/2] complex twiddles */
} ao_fftfilt;
void ao_fftfilt_init(ao_fftfilt *f);                                  /* empty kernel */
int ao_fftfilt_set_taps(ao_fftfilt *f, const float *taps, int ntaps); /* returns nsamples */
/* nitems must be a multiple of nsamples (the caller's output multiple); returns nitems */
int ao_fftfilt_filter(ao_fftfilt *f, int nitems, const float *in, float *out);
void ao_fftfilt_free(ao_fftfilt *f);
/* volk_32fc_magnitude_squared_32f [G] */
void ao_mag_squared(const float *in, int n, float *out);
/* mmse_fir_interpolator_cc::interpolate [G] on in[0..7]; -1 when rint(mu*128) is outside
 * [0,128] (GNU Radio throws std::runtime_error) */
int ao_mmse_interpolate(const float *in8, float mu, float *vr, float *vi);

/* ---- A1-A4: corr_est_cc ---- */
typedef struct ao_corr_est {
    float *taps; /* [L] complex: ctor stores reverse(conj(symbols)); set_symbols stores verbatim */
    int L;
    float sps;
    unsigned mark_delay;
    float thresh;
    ao_fftfilt f; /* d_filter; f.nsamples = the block's output multiple */
} ao_corr_est;
int ao_corr_est_init(ao_corr_est *c, const float *symbols, int L, float sps, unsigned mark_delay,
                     float threshold);
void ao_corr_est_set_symbols(ao_corr_est *c, const float *symbols, int L);
void ao_corr_est_free(ao_corr_est *c);
/* the correlation of the whole call by a float64 direct sum (test truth, not the canonical path) */
void ao_corr_direct_f64(const ao_corr_est *c, int n, const float *in, double *corr_out);
/* forward DIF (natural in, bit-reversed out) / inverse DIT (bit-reversed in, natural out,
 * unnormalised) of the correlator, in place on n complex values */
int ao_fft_dif_inplace(float *x, int n);
int ao_ifft_dit_inplace(float *x, int n);
/* One work() call.  in holds n+L items (history first).  corr/mag may be NULL.
 * two_ports != 0 also emits the port-1 debug tags.  Returns n; *ntags = tags written
 * (tags beyond max_tags are counted but dropped). */
int ao_corr_est_work(ao_corr_est *c, int n, const float *in, uint64_t nitems_written,
                     float *out0, float *corr, float *mag, int two_ports, ao_tag *tags,
                     int max_tags, int *ntags);

/* ---- A5-A7: msk_timing_recovery_cc ---- */
typedef struct ao_msk {
    float sps; /* d_sps = sps/2 */
    float gain, gain_omega, limit;
    float mu, omega;
    float dly1_re, dly1_im, dly2_re, dly2_im, diff1_re, diff1_im;
    int div;
    int osps;
    float prev_re, prev_im; /* in[-1]: last consumed item of the previous call (0 at start) */
} ao_msk;
int ao_msk_init(ao_msk *m, float sps, float gain, float limit, int osps); /* <0 on out_of_range */
int ao_msk_forecast(const ao_msk *m, int noutput_items);
/* One general_work() call.  tags: time_est tags (any order of keys is filtered; offsets absolute),
 * nitems_read: absolute offset of in[0].  Returns items produced; *consumed = consume_each arg. */
int ao_msk_general_work(ao_msk *m, int noutput_items, int ninput_items, const float *in,
                        uint64_t nitems_read, const ao_tag *tags, int ntags, float *out,
                        float *out_err, float *out_mu, int *consumed);

/* ---- G4-G6 + A9: demod tail ---- */
/* prev: v[-1] (2 floats, updated).  Returns soft output y[k] = gain*fast_atan2f(...) */
void ao_quad_demod(float *prev, const float *in, int n, float gain, float *out);
void ao_binary_slicer(const float *in, int n, uint8_t *out);
void ao_diff_decoder(uint8_t *prev, const uint8_t *in, int n, unsigned modulus, uint8_t *out);
void ao_invert(const uint8_t *in, int n, uint8_t *out);

/* ---- the ais_demod chain over one record from fresh state ---- */
enum { AO_STAGE_FREQSYNC = 1, AO_STAGE_AGC = 2 };
typedef struct ao_chain_cfg {
    float sample_rate;   /* 48000 */
    int data_rate;       /* 9600 */
    int fftlen;          /* 1024 */
    int agc_nsamples;    /* 512 */
    float agc_reference; /* 2 */
    float sps;           /* 5 */
    unsigned mark_delay; /* 1 */
    float threshold;     /* 0.9 */
    float gain;          /* 0.04 */
    float limit;         /* 0.01 */
    int osps;            /* 1 */
    int corr_chunk;      /* corr_est work-chunk (0 = largest multiple of nsamples <= 24576) */
    int stages;          /* AO_STAGE_* mask; corr_est, msk and the bit tail always run */
} ao_chain_cfg;

typedef struct ao_chain_out {
    /* required */
    uint8_t *bits;
    int max_bits;
    int nbits;
    ao_tag *tags;
    int max_tags;
    int ntags;
    /* optional taps into the chain (NULL to skip) */
    float *fhat;  /* [N/fftlen] */
    float *mixed; /* [N1] complex */
    float *agc;   /* [N1] complex */
    float *corr;  /* [N2] complex */
    float *mag;   /* [N2] */
    float *sym;   /* [nbits] complex */
    float *err;   /* [nbits] */
    float *mu;    /* [nbits] */
    float *soft;  /* [nbits] */
    int n1, n2;   /* samples out of the AGC / processed by corr_est */
    int consumed; /* msk consume_each */
} ao_chain_out;

/* Who implements the gr-ais blocks [R] inside the chain.  ao_blocks_oracle() = the
 * restatements in this file; oracle/_ref/libais_ref.so exports ref_blocks() = the reference's
 * own classes (its lib/ sources compiled unmodified).  Everything between the blocks (the stock
 * GNU Radio blocks [G] and the schedule) is shared, so the two chains differ exactly by
 * "restated" against "the reference's code". */
typedef struct ao_blocks {
    const char *name;
    void *(*corr_new)(const float *symbols, int L, float sps, unsigned mark_delay, float threshold);
    void (*corr_delete)(void *h);
    int (*corr_output_multiple)(void *h);
    int (*corr_set_symbols)(void *h, const float *symbols, int L);
    int (*corr_work)(void *h, int n, const float *in, uint64_t nitems_written, float *out0,
                     float *corr, float *mag, int two_ports, ao_tag *tags, int max_tags, int *ntags);
    void *(*msk_new)(float sps, float gain, float limit, int osps, int *status);
    void (*msk_delete)(void *h);
    float (*msk_get_sps)(void *h); /* d_sps = sps / 2 */
    int (*msk_general_work)(void *h, int noutput_items, int ninput_items, const float *in,
                            uint64_t nitems_read, const ao_tag *tags, int ntags, float *out,
                            float *out_err, float *out_mu, int *consumed);
    void *(*freqest_new)(float sample_rate, int data_rate, int fftlen);
    void (*freqest_delete)(void *h);
    int (*freqest_work)(void *h, const float *spec, int nvec, float *out);
    void (*invert_work)(const uint8_t *in, int n, uint8_t *out);
} ao_blocks;
const ao_blocks *ao_blocks_oracle(void);

int ao_default_corr_chunk(int L);
/* ..._with: the same chain over another block provider (NULL = ao_blocks_oracle()) */
int ao_demod_chain_with(const ao_blocks *blk, const ao_chain_cfg *cfg, const float *symbols, int L,
                        const float *x, int n, ao_chain_out *out);
int ao_demod_chain_batch_with(const ao_blocks *blk, const ao_chain_cfg *cfg, const float *symbols,
                              int L, const float *x, int channels, int n, uint8_t *bits,
                              int max_bits, int *nbits, ao_tag *tags, int max_tags, int *ntags,
                              int nthreads);
int ao_demod_chain(const ao_chain_cfg *cfg, const float *symbols, int L, const float *x, int n,
                   ao_chain_out *out);
/* batch over channels with OpenMP; x is [C][n] complex, bits [C][max_bits], nbits [C],
 * tags [C][max_tags], ntags [C].  Returns 0 or the first negative status. */
int ao_demod_chain_batch(const ao_chain_cfg *cfg, const float *symbols, int L, const float *x,
                         int channels, int n, uint8_t *bits, int max_bits, int *nbits,
                         ao_tag *tags, int max_tags, int *ntags, int nthreads);

/* ---- the same chain as a stream: blocks keep their state from call to call ----
 * One call = one scheduler pass in which every block runs once, in flowgraph order, over all
 * the items available to it: whole FFT vectors (the remainder waits), every mixed item through
 * the AGC, whole corr_est output multiples in work chunks (the remainder waits), everything
 * corr_est has produced so far through one msk general_work(), then the bit tail. */
typedef struct ao_stream {
    ao_chain_cfg cfg;
    const ao_blocks *blk;
    void *ce, *mk, *fe; /* block handles of the provider */
    int ns;             /* corr_est output multiple */
    int L;
    float nco_phase;
    float *xcarry;   /* input items waiting for a whole FFT vector */
    int nxcarry;
    float *agc_hist; /* last agc_nsamples-1 mixed items */
    float *acarry;   /* L history items + AGC outputs corr_est has not taken yet */
    int nacarry;     /* items after the L history */
    float *ocarry;   /* corr_est output-0 items msk has not consumed */
    int nocarry;
    ao_tag *tags;    /* tags msk may still need */
    int ntags, captags;
    uint64_t written; /* corr_est nitems_written */
    uint64_t read;    /* msk nitems_read */
    float qprev[2];
    uint8_t dprev;
} ao_stream;
int ao_stream_init(ao_stream *s, const ao_blocks *blk, const ao_chain_cfg *cfg, const float *symbols, int L);
ao_stream *ao_stream_new(const ao_chain_cfg *cfg, const float *symbols, int L);
ao_stream *ao_stream_new_with(const ao_blocks *blk, const ao_chain_cfg *cfg, const float *symbols, int L);
void ao_stream_delete(ao_stream *s);
void ao_stream_free(ao_stream *s);
int ao_stream_set_symbols(ao_stream *s, const float *symbols, int L);
/* bits/tags of this call only; returns 0 or a negative status */
int ao_stream_work(ao_stream *s, const float *x, int n, uint8_t *bits, int max_bits, int *nbits,
                   ao_tag *tags_out, int max_tags, int *ntags_out);

/* ======== the blocks either side of the path (python/radio.py:39-72), ais_oracle_rx.c ======== */

/* ---- digital.hdlc_deframer_bp [G] ---- */
#define AO_FRAME_MAX 248
typedef struct ao_frame {
    uint64_t end_bit; /* absolute index of the bit that completed the closing flag */
    int32_t len;      /* payload bytes (CRC removed) */
    int32_t channel;  /* unused by the single-stream oracle (0) */
    uint8_t data[AO_FRAME_MAX];
} ao_frame;
typedef struct ao_hdlc {
    int length_min, length_max;
    int ones, bitctr, bytectr, pad;
    uint64_t nitems_read;
    uint8_t pktbuf[AO_FRAME_MAX + 8];
} ao_hdlc;
unsigned ao_crc_ccitt(const uint8_t *data, size_t len);
void ao_hdlc_init(ao_hdlc *h, int length_min, int length_max);
int ao_hdlc_work(ao_hdlc *h, const uint8_t *bits, int n, ao_frame *frames, int max_frames,
                 int *dropped);

/* ---- gr::ais::pdu_to_nmea (lib/pdu_to_nmea_impl.cc:63-131) [R] ---- */
int ao_pdu_to_nmea(const char *designator, const uint8_t *data, int len, char *out, int cap);

/* ---- firdes.low_pass + freq_xlating_fir_filter_ccf (python/radio.py:49-54) [G] ---- */
int ao_firdes_low_pass(double gain, double fs, double cutoff, double tw, float *taps, int cap);
typedef struct ao_xlat {
    int decim, ntaps;
    float *ctaps; /* [ntaps] complex band-pass taps */
    float incr_re, incr_im, phase_re, phase_im;
    unsigned counter;
} ao_xlat;
int ao_xlat_init(ao_xlat *x, int decimation, const float *taps, int ntaps, double center_freq,
                 double sampling_freq);
void ao_xlat_free(ao_xlat *x);
int ao_xlat_work(ao_xlat *x, int noutput, const float *in, float *out, float *fir_out);
void ao_xlat_f64(int decimation, const float *taps, int ntaps, double center_freq,
                 double sampling_freq, uint64_t first_output, int noutput, const float *in,
                 double *out);

#ifdef __cplusplus
}
#endif
#endif
