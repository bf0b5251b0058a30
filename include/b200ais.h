/*
 * b200ais.h -- C-ABI of libb200ais.so: the gr-ais IQ-demod hot path on B200 (sm_100a).
 *
 * This is the drop-in boundary.  Every entry point takes plain pointers and
 * sizes (no C++ or torch types) and replaces one interface of the reference
 * (bistromath/gr-ais @ 2162103; paths below are relative to /root/reference):
 *
 *   b200ais_corr_est_*   gr::ais::corr_est_cc            include/ais/corr_est_cc.h:85-107,
 *                                                        lib/corr_est_cc_impl.cc:48-117 (ctor),
 *                                                        :132-162 (set_symbols), :164-279 (work)
 *   b200ais_msk_*        gr::ais::msk_timing_recovery_cc include/ais/msk_timing_recovery_cc.h:46-70,
 *                                                        lib/msk_timing_recovery_cc_impl.cc:45-96
 *                                                        (ctor/setters), :98-105 (forecast),
 *                                                        :107-206 (general_work)
 *   b200ais_freqest_*    gr::ais::freqest                include/ais/freqest.h:37-50,
 *                                                        lib/freqest_impl.cc:41-48, :57-88 (work)
 *   b200ais_invert_work  gr::ais::invert                 include/ais/invert.h:37-50,
 *                                                        lib/invert_impl.cc:54-68 (work)
 *   b200ais_demod_*      ais_demod hier-block            python/ais_demod.py:21-56 with
 *                        + square_and_fft_sync_cc        python/gmsk_sync.py:14-37 fused in
 *
 * and, either side of that path inside the reference's ais_rx receiver (python/radio.py:39-72):
 *
 *   b200ais_firdes_low_pass  filter.firdes.low_pass      python/radio.py:49
 *   b200ais_xlat_*       filter.freq_xlating_fir_filter_ccf  python/radio.py:51-54
 *   b200ais_hdlc_*       digital.hdlc_deframer_bp(11,64) python/radio.py:64
 *   b200ais_nmea_*       gr::ais::pdu_to_nmea            include/ais/pdu_to_nmea.h:37-54,
 *                                                        lib/pdu_to_nmea_impl.cc:63-131
 *   b200ais_rx_*         ais_rx hier-block (all of the above chained)  python/radio.py:39-72
 *
 * The *_work functions mirror one GNU Radio work()/general_work() call, batched
 * over `channels` independent streams (the GR adapter calls with channels = 1).
 * Host-pointer variants copy to/from the device inside the call; *_dev variants
 * take device pointers and a cudaStream_t (as void*) and do not synchronise.
 *
 * Layout: channel-major [channels][stride] rows of interleaved (re, im) float32
 * (= gr_complex); strides are in items (complex samples / bytes), rows 16-byte
 * aligned for the device variants.
 *
 * All functions return B200AIS_OK (0) or a negative B200AIS_E_* code; the text of
 * the last error on the calling thread is b200ais_last_error().  There is no CPU
 * fallback: without a usable sm_100 device every compute call fails with
 * B200AIS_E_CUDA.
 */
#ifndef B200AIS_H
#define B200AIS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B200AIS_API
#else
#define B200AIS_API __attribute__((visibility("default")))
#endif

enum {
    B200AIS_OK = 0,
    B200AIS_E_INVALID = -1,      /* bad argument */
    B200AIS_E_RANGE = -2,        /* the reference throws std::out_of_range here */
    B200AIS_E_CUDA = -3,         /* CUDA runtime error / no device */
    B200AIS_E_NOMEM = -4,
    B200AIS_E_TAG_OVERFLOW = -5, /* more tags than max_tags on some channel */
    B200AIS_E_INTERP = -6,       /* mmse interpolator index out of [0,128] (reference: runtime_error) */
    B200AIS_E_OUT_OVERFLOW = -7, /* an output row was too small */
    B200AIS_E_FRAME_OVERFLOW = -8 /* more HDLC frames than max_frames on some channel */
};

/* stream-tag keys emitted by corr_est_cc (lib/corr_est_cc_impl.cc:213-256) */
enum {
    B200AIS_TAG_CORR_START = 0,
    B200AIS_TAG_PHASE_EST = 1,
    B200AIS_TAG_TIME_EST = 2,
    B200AIS_TAG_CORR_EST = 3
};

/* POD stand-in for gr::tag_t {offset, key, pmt::from_double(value), srcid} */
typedef struct b200ais_tag {
    uint64_t offset; /* absolute item offset on corr_est output `port` */
    int32_t key;     /* B200AIS_TAG_* */
    int32_t port;    /* 0, or 1 for the debug copies on the optional 2nd output */
    double value;
} b200ais_tag;

/* ---------------------------------------------------------------- library */
B200AIS_API int b200ais_version(void);
B200AIS_API const char *b200ais_last_error(void);
B200AIS_API int b200ais_device_count(int *count);
B200AIS_API int b200ais_set_device(int device);
/* pinned host memory, so the host-pointer variants overlap copies with kernels */
B200AIS_API int b200ais_host_alloc(void **ptr, size_t bytes);
B200AIS_API int b200ais_host_free(void *ptr);
/* number of kernels this library has launched on the calling process (bench bookkeeping) */
B200AIS_API uint64_t b200ais_launch_count(void);

/* ------------------------------------------------------------ corr_est_cc */
typedef struct b200ais_corr_est b200ais_corr_est;

/* corr_est_cc::make(symbols, sps, mark_delay, threshold) for `channels` streams */
B200AIS_API int b200ais_corr_est_create(b200ais_corr_est **h, const float *symbols_iq, int nsymbols,
                                        float sps, unsigned mark_delay, float threshold,
                                        int channels);
B200AIS_API int b200ais_corr_est_destroy(b200ais_corr_est *h);
/* set_symbols(): taps replaced verbatim, threshold unchanged (lib/corr_est_cc_impl.cc:132-162) */
B200AIS_API int b200ais_corr_est_set_symbols(b200ais_corr_est *h, const float *symbols_iq,
                                             int nsymbols);
/* symbols(): the stored taps (conj-reversed after the ctor), *n items */
B200AIS_API int b200ais_corr_est_symbols(const b200ais_corr_est *h, float *out_iq, int cap, int *n);
B200AIS_API int b200ais_corr_est_output_multiple(const b200ais_corr_est *h); /* fft_filter nsamples */
B200AIS_API int b200ais_corr_est_history(const b200ais_corr_est *h);         /* nsymbols + 1 */
B200AIS_API unsigned b200ais_corr_est_mark_delay(const b200ais_corr_est *h);
B200AIS_API float b200ais_corr_est_threshold(const b200ais_corr_est *h);     /* d_thresh */

/* One work() call.  in: [channels][in_stride], each row noutput_items + nsymbols items
 * (history first, exactly what GNU Radio hands work()).  out0 (nullable): the delayed
 * pass-through; out1 (nullable): the correlator output, both [channels][out_stride].
 * tags: [channels][max_tags], ntags: [channels]. */
B200AIS_API int b200ais_corr_est_work(b200ais_corr_est *h, int noutput_items, const float *in,
                                      size_t in_stride, uint64_t nitems_written, float *out0,
                                      float *out1, size_t out_stride, b200ais_tag *tags,
                                      int max_tags, int *ntags);
B200AIS_API int b200ais_corr_est_work_dev(b200ais_corr_est *h, int noutput_items, const float *in,
                                          size_t in_stride, uint64_t nitems_written, float *out0,
                                          float *out1, size_t out_stride, b200ais_tag *tags,
                                          int max_tags, int *ntags, void *stream);

/* ------------------------------------------------ msk_timing_recovery_cc */
typedef struct b200ais_msk b200ais_msk;

B200AIS_API int b200ais_msk_create(b200ais_msk **h, float sps, float gain, float limit, int osps,
                                   int channels);
B200AIS_API int b200ais_msk_destroy(b200ais_msk *h);
B200AIS_API int b200ais_msk_set_gain(b200ais_msk *h, float gain);
B200AIS_API float b200ais_msk_get_gain(const b200ais_msk *h);
B200AIS_API int b200ais_msk_set_limit(b200ais_msk *h, float limit);
B200AIS_API float b200ais_msk_get_limit(const b200ais_msk *h);
B200AIS_API int b200ais_msk_set_sps(b200ais_msk *h, float sps);
B200AIS_API float b200ais_msk_get_sps(const b200ais_msk *h); /* returns sps/2, as the reference */
B200AIS_API int b200ais_msk_forecast(const b200ais_msk *h, int noutput_items);
B200AIS_API int b200ais_msk_reset(b200ais_msk *h); /* back to the freshly constructed loop state */

/* One general_work() call.  in: [channels][in_stride] with ninput_items valid items per
 * row; tags: [channels][max_tags] (time_est tags on port 0 are used, others ignored),
 * ntags: [channels]; out: [channels][out_stride] complex, out_err/out_mu (nullable) float.
 * nproduced / nconsumed: [channels] (the return value and consume_each argument). */
B200AIS_API int b200ais_msk_general_work(b200ais_msk *h, int noutput_items, int ninput_items,
                                         const float *in, size_t in_stride, uint64_t nitems_read,
                                         const b200ais_tag *tags, int max_tags, const int *ntags,
                                         float *out, float *out_err, float *out_mu,
                                         size_t out_stride, int *nproduced, int *nconsumed);
B200AIS_API int b200ais_msk_general_work_dev(b200ais_msk *h, int noutput_items, int ninput_items,
                                             const float *in, size_t in_stride,
                                             uint64_t nitems_read, const b200ais_tag *tags,
                                             int max_tags, const int *ntags, float *out,
                                             float *out_err, float *out_mu, size_t out_stride,
                                             int *nproduced, int *nconsumed, void *stream);

/* ---------------------------------------------------------------- freqest */
typedef struct b200ais_freqest b200ais_freqest;

B200AIS_API int b200ais_freqest_create(b200ais_freqest **h, float sample_rate, int data_rate,
                                       int fftlen, int channels);
B200AIS_API int b200ais_freqest_destroy(b200ais_freqest *h);
/* One work() call: spec [channels][noutput_items*fftlen] complex, out [channels][noutput_items] */
B200AIS_API int b200ais_freqest_work(b200ais_freqest *h, int noutput_items, const float *spec,
                                     float *out);
B200AIS_API int b200ais_freqest_work_dev(b200ais_freqest *h, int noutput_items, const float *spec,
                                         float *out, void *stream);

/* ----------------------------------------------------------------- invert */
B200AIS_API int b200ais_invert_work(const uint8_t *in, uint8_t *out, size_t nitems);
B200AIS_API int b200ais_invert_work_dev(const uint8_t *in, uint8_t *out, size_t nitems,
                                        void *stream);

/* ---------------------------------------------- the fused ais_demod chain */
enum { B200AIS_STAGE_FREQSYNC = 1, B200AIS_STAGE_AGC = 2 };

typedef struct b200ais_demod_config {
    float sample_rate;   /* samples_per_symbol * bits_per_sec (python/ais_demod.py:30) */
    int data_rate;       /* bits_per_sec */
    int fftlen;          /* options["fftlen"] (python/radio.py:61) */
    int agc_nsamples;    /* feedforward_agc_cc(512, 2) (python/ais_demod.py:35) */
    float agc_reference;
    float sps;
    unsigned mark_delay; /* 1 (python/ais_demod.py:41) */
    float threshold;     /* 0.9 (python/ais_demod.py:42) */
    float gain;          /* clockrec_gain */
    float limit;         /* omega_relative_limit */
    int osps;            /* 1 */
    int corr_chunk;      /* corr_est work-chunk; 0 = largest multiple of nsamples <= 24576 */
    int stages;          /* B200AIS_STAGE_* mask; corr_est, msk and the bit tail always run */
} b200ais_demod_config;

typedef struct b200ais_demod b200ais_demod;

B200AIS_API int b200ais_demod_default_config(b200ais_demod_config *cfg);
B200AIS_API int b200ais_demod_create(b200ais_demod **h, const b200ais_demod_config *cfg,
                                     const float *symbols_iq, int nsymbols, int channels,
                                     int max_samples, int max_tags);
B200AIS_API int b200ais_demod_destroy(b200ais_demod *h);
/* bits a record of nsamples can produce per channel (5 % over the nominal rate + 64), rounded up
 * to a multiple of 4: with word-aligned bit rows the timing loop writes the bits itself */
B200AIS_API int b200ais_demod_max_bits(const b200ais_demod *h, int nsamples);
/* corr_est_cc::set_symbols (lib/corr_est_cc_impl.cc:132-162) on the chain's correlator: the taps
 * are replaced verbatim (no conjugate / reversal, unlike the constructor) and d_thresh keeps its
 * value; nsymbols must equal the count given at create.  Synchronises `stream`; applies to every
 * later batch and stream call. */
B200AIS_API int b200ais_demod_set_symbols(b200ais_demod *h, const float *symbols_iq, int nsymbols,
                                          void *stream);
/* One record per channel, every block freshly constructed (ais_demod on a new stream).
 * iq: [channels][nsamples] complex; bits: [channels][max_bits] unpacked 0/1 bytes;
 * nbits: [channels]; tags (nullable): [channels][max_tags]; ntags (nullable): [channels]. */
B200AIS_API int b200ais_demod_work(b200ais_demod *h, const float *iq, int nsamples, uint8_t *bits,
                                   int max_bits, int *nbits, b200ais_tag *tags, int *ntags);
/* The same call with the IQ delivered as interleaved int16 I/Q (the SDR wire format: UHD "sc16",
 * osmosdr; the reference's sources convert it to gr_complex on the host, python/radio.py:151-203).
 * Every component becomes (float)v * scale on the device (exact for a power-of-two scale), then
 * the chain runs as in b200ais_demod_work: half the bytes cross PCIe.
 * iq: [channels][nsamples][2] int16. */
B200AIS_API int b200ais_demod_work_sc16(b200ais_demod *h, const int16_t *iq, float scale, int nsamples,
                                        uint8_t *bits, int max_bits, int *nbits, b200ais_tag *tags,
                                        int *ntags);
B200AIS_API int b200ais_demod_work_dev(b200ais_demod *h, const float *iq, int nsamples,
                                       uint8_t *bits, int max_bits, int *nbits, b200ais_tag *tags,
                                       int *ntags, void *stream);
/* after a *_dev call has completed: 0 or the B200AIS_E_* a kernel flagged */
B200AIS_API int b200ais_demod_status(b200ais_demod *h);

/* Pipelined submission of independent records (same arguments and results as work_dev).  The
 * two per-channel recurrences (NCO phase, timing loop) take the same time for any batch size,
 * so a strictly ordered call always exposes one full timing-loop pass.  enqueue_dev orders
 * everything up to the corr_est detector on `stream` and hands msk_timing_recovery + the bit
 * tail to an internal high-priority stream, where they run under the front half of the NEXT
 * enqueue (corr_est input rows and tags are double-buffered).  tags / ntags are complete in
 * `stream` order; bits / nbits are complete once a later b200ais_demod_join(h, stream) has
 * been reached in `stream` order (any other b200ais_demod_* work call joins by itself).
 * A handle with more than 53 248 channels fills the GPU with its timing loop alone; its
 * records run entirely on `stream` (the contract above still holds, join is then a no-op). */
B200AIS_API int b200ais_demod_enqueue_dev(b200ais_demod *h, const float *iq, int nsamples,
                                          uint8_t *bits, int max_bits, int *nbits,
                                          b200ais_tag *tags, int *ntags, void *stream);
B200AIS_API int b200ais_demod_join(b200ais_demod *h, void *stream);

/* The same chain fed as a stream: a capture arrives in pieces of any size (0..max_samples
 * items per call, the same count on every channel) and every block keeps, from call to
 * call, what it keeps between work() calls under the GNU Radio scheduler:
 *   stream_to_vector's partial vector           python/gmsk_sync.py:23
 *   frequency_modulator_fc's phase              python/gmsk_sync.py:27
 *   feedforward_agc_cc's 511-item history       python/ais_demod.py:35
 *   corr_est_cc's history, output multiple and
 *     fft_filter tail                           lib/corr_est_cc_impl.cc:77-78,105-117,188
 *   msk_timing_recovery_cc's loop state, its
 *     unconsumed input and the tags in it       lib/msk_timing_recovery_cc_impl.cc:125-130,203
 *   quadrature_demod_cf / diff_decoder_bb history   python/ais_demod.py:48-51
 * One call is one scheduler pass: every block runs once over what is available to it.  bits /
 * nbits / tags / ntags hold THIS call's output (tag offsets are absolute corr_est item
 * offsets).  Needs the AGC stage with the reference's 512-sample window.  The first stream
 * call after create / a batch call / stream_reset starts from freshly constructed blocks.
 * max_bits per call: b200ais_demod_stream_max_bits(h, nsamples). */
B200AIS_API int b200ais_demod_stream_reset(b200ais_demod *h, void *stream);
B200AIS_API int b200ais_demod_stream_max_bits(const b200ais_demod *h, int nsamples);
B200AIS_API int b200ais_demod_stream_work(b200ais_demod *h, const float *iq, int nsamples,
                                          uint8_t *bits, int max_bits, int *nbits,
                                          b200ais_tag *tags, int *ntags);
B200AIS_API int b200ais_demod_stream_work_dev(b200ais_demod *h, const float *iq, int nsamples,
                                              uint8_t *bits, int max_bits, int *nbits,
                                              b200ais_tag *tags, int *ntags, void *stream);
/* The same pass with the input produced on the device by the caller (the channeliser writes
 * straight into the chain's assembly rows): stage() returns where this call's `nsamples` new
 * items of channel c go (*dev_ptr + c * *stride complex items); work_staged() then runs the
 * pass.  Nothing else may touch the handle between the two calls. */
B200AIS_API int b200ais_demod_stream_stage(b200ais_demod *h, int nsamples, int max_bits,
                                           float **dev_ptr, size_t *stride, void *stream);
B200AIS_API int b200ais_demod_stream_work_staged(b200ais_demod *h, int nsamples, uint8_t *bits,
                                                 int max_bits, int *nbits, b200ais_tag *tags,
                                                 int *ntags, void *stream);
/* items waiting inside the stream (all nullable): input items short of an FFT vector, AGC
 * outputs short of a corr_est output multiple, corr_est's nitems_written */
B200AIS_API int b200ais_demod_stream_pending(const b200ais_demod *h, int *input_items,
                                             int *agc_items, uint64_t *corr_written);

/* Per-stage device timing of the *_dev chain (CUDA events around every launch, on the
 * caller's stream).  stage_ms: sums over the work_dev calls since profiling was enabled
 * or last read; index = B200AIS_STAGE_T_*.  Reading synchronises the recorded events. */
enum {
    B200AIS_STAGE_T_SQFFT = 0,  /* x^2 -> FFT -> freqest argmax */
    B200AIS_STAGE_T_NCO = 1,    /* NCO phase recurrence */
    B200AIS_STAGE_T_MIXAGC = 2, /* NCO mix + feedforward AGC */
    B200AIS_STAGE_T_CORR = 3,   /* corr_est correlator (the dominant kernel) */
    B200AIS_STAGE_T_DETECT = 4, /* corr_est detector */
    B200AIS_STAGE_T_MSK = 5,    /* msk_timing_recovery recurrence */
    B200AIS_STAGE_T_TAIL = 6,   /* quadrature demod -> slicer -> diff decoder -> invert */
    B200AIS_STAGE_T_COUNT = 7
};
B200AIS_API int b200ais_demod_profile(b200ais_demod *h, int enable);
/* Number of channel groups a work_dev call forks over internal CUDA streams (default 1 = run
 * every kernel on the caller's stream; measured: no gain device-resident).  The caller's stream
 * still orders the whole call.  The host-buffer variant always pipelines 8 channel groups so
 * that copies overlap kernels. */
B200AIS_API int b200ais_demod_set_overlap(b200ais_demod *h, int groups);
B200AIS_API int b200ais_demod_stage_ms(b200ais_demod *h, double *stage_ms, int *calls);

/* Device pointers to the chain's intermediate streams of the last work call (parity
 * tests and profiling).  which: */
enum {
    B200AIS_TAP_FHAT = 0,  /* float  [channels][nsamples/fftlen]  freqest output (Hz) */
    B200AIS_TAP_AGC = 1,   /* complex[channels][row]              corr_est input stream */
    B200AIS_TAP_SYM = 2,   /* complex[channels][max_bits]         msk out0 */
    B200AIS_TAP_ERR = 3,   /* float  [channels][max_bits]         msk out1 */
    B200AIS_TAP_MU = 4,    /* float  [channels][max_bits]         msk out2 */
    B200AIS_TAP_SOFT = 5,  /* float  [channels][max_bits]         quadrature_demod output */
    B200AIS_TAP_MASK = 6   /* uint8  [channels][row/8]            corr_est |corr|^2 > thresh bitmask */
};
B200AIS_API int b200ais_demod_enable_taps(b200ais_demod *h, int enable);
B200AIS_API int b200ais_demod_tap(b200ais_demod *h, int which, void **dev_ptr, size_t *row_items);
/* copy a tap to host: dst holds channels*row_items items of the tap's type */
B200AIS_API int b200ais_demod_read_tap(b200ais_demod *h, int which, void *dst, size_t dst_bytes);

/* ------------------- channeliser: firdes.low_pass + freq_xlating_fir_filter_ccf */
/* firdes::low_pass(gain, sampling_freq, cutoff, transition_width) with the default Hamming
 * window (python/radio.py:49); init-time, host.  taps == NULL returns the tap count in *ntaps. */
B200AIS_API int b200ais_firdes_low_pass(double gain, double sampling_freq, double cutoff_freq,
                                        double transition_width, float *taps, int cap, int *ntaps);

typedef struct b200ais_xlat b200ais_xlat;
/* `nfreqs` freq_xlating_fir_filter_ccf(decimation, taps, center_freqs[k], sampling_freq) blocks
 * fed by the same input (python/radio.py:86-91: the A and B rx paths share one source), for
 * `sources` independent wideband inputs.  Output channel s*nfreqs + k is source s translated
 * by center_freqs[k]. */
B200AIS_API int b200ais_xlat_create(b200ais_xlat **h, int decimation, const float *taps, int ntaps,
                                    const double *center_freqs, int nfreqs, double sampling_freq,
                                    int sources);
B200AIS_API int b200ais_xlat_destroy(b200ais_xlat *h);
B200AIS_API int b200ais_xlat_history(const b200ais_xlat *h);    /* ntaps */
B200AIS_API int b200ais_xlat_decimation(const b200ais_xlat *h);
B200AIS_API int b200ais_xlat_set_center_freq(b200ais_xlat *h, int k, double center_freq);
B200AIS_API int b200ais_xlat_set_taps(b200ais_xlat *h, const float *taps, int ntaps);
/* back to freshly constructed rotators (phase 1, counter 0) */
B200AIS_API int b200ais_xlat_reset(b200ais_xlat *h);
/* One work() call.  in: [sources][in_stride] complex, each row ntaps-1 history items followed
 * by noutput_items*decimation new ones (what GNU Radio hands work()); out:
 * [sources*nfreqs][out_stride] complex, noutput_items per row.  The rotators advance. */
B200AIS_API int b200ais_xlat_work(b200ais_xlat *h, int noutput_items, const float *in,
                                  size_t in_stride, float *out, size_t out_stride);
B200AIS_API int b200ais_xlat_work_dev(b200ais_xlat *h, int noutput_items, const float *in,
                                      size_t in_stride, float *out, size_t out_stride,
                                      void *stream);

/* -------------------------------------------------- hdlc_deframer_bp + CRC */
#define B200AIS_FRAME_MAX 248
/* one PDU published by hdlc_deframer_bp: pmt::cons(PMT_NIL, blob(data, len)) */
typedef struct b200ais_frame {
    uint64_t end_bit; /* absolute index, in the channel's bit stream, of the bit that completed
                         the closing flag */
    int32_t len;      /* payload bytes, CRC removed */
    int32_t channel;  /* row of the bit stream it came from */
    uint8_t data[B200AIS_FRAME_MAX];
} b200ais_frame;

typedef struct b200ais_hdlc b200ais_hdlc;
/* hdlc_deframer_bp(length_min, length_max) for `channels` bit streams (python/radio.py:64) */
B200AIS_API int b200ais_hdlc_create(b200ais_hdlc **h, int length_min, int length_max, int channels);
B200AIS_API int b200ais_hdlc_destroy(b200ais_hdlc *h);
B200AIS_API int b200ais_hdlc_reset(b200ais_hdlc *h);
/* One work() call.  bits: [channels][bits_stride] unpacked 0/1 bytes, nbits[c] valid items in
 * row c (nbits == NULL: nbits_all on every row); frames: [channels][max_frames]; nframes:
 * [channels].  The deframer state (run of ones, partial frame) carries to the next call. */
B200AIS_API int b200ais_hdlc_work(b200ais_hdlc *h, const uint8_t *bits, size_t bits_stride,
                                  const int *nbits, int nbits_all, b200ais_frame *frames,
                                  int max_frames, int *nframes);
B200AIS_API int b200ais_hdlc_work_dev(b200ais_hdlc *h, const uint8_t *bits, size_t bits_stride,
                                      const int *nbits, int nbits_all, b200ais_frame *frames,
                                      int max_frames, int *nframes, void *stream);
/* after a *_dev call has completed: 0 or B200AIS_E_FRAME_OVERFLOW */
B200AIS_API int b200ais_hdlc_status(b200ais_hdlc *h);

/* ------------------------------------------------------------ pdu_to_nmea */
/* bytes one frame's sentence(s) can take for payloads up to max_len bytes and this designator */
B200AIS_API int b200ais_nmea_slot_bytes(int max_len, const char *designator);
/* pdu_to_nmea::msg_to_sentence (lib/pdu_to_nmea_impl.cc:127-131) for every frame:
 * frames [channels][max_frames], nframes [channels]; designators: [channels][8] NUL-padded
 * strings (NULL: "A" everywhere); sentences: [channels][max_frames][slot] characters (fragments
 * joined by '\n', NUL-terminated when room), lens: [channels][max_frames]. */
B200AIS_API int b200ais_nmea_format(const b200ais_frame *frames, const int *nframes, int channels,
                                    int max_frames, const char *designators, char *sentences,
                                    int slot, int *lens);
B200AIS_API int b200ais_nmea_format_dev(const b200ais_frame *frames, const int *nframes,
                                        int channels, int max_frames, const char *designators,
                                        char *sentences, int slot, int *lens, void *stream);

/* ------------------------------------------- ais_rx: the whole receiver path */
/* python/radio.py:39-72 for `sources` wideband inputs:
 *   freq_xlating_fir_filter_ccf(int(rate/48000), firdes.low_pass(1, rate, 11000, 1000), freq, rate)
 *   -> ais_demod(options) -> hdlc_deframer_bp(11, 64) -> pdu_to_nmea(designator),
 * one such path per entry of `freqs` sharing each source (python/radio.py:86-91).  Channel
 * s*nfreqs + k is source s at freqs[k].  A capture is fed in pieces; every block keeps its state. */
typedef struct b200ais_rx_config {
    double rate;               /* options.rate (python/radio.py:120), default 250e3 */
    int nfreqs;                /* 1..16 */
    double freqs[16];          /* offsets from the tuned centre (python/radio.py:88-89) */
    char designators[16][8];   /* NUL-padded, "A"/"B" */
    int sources;
    int max_input_items;       /* per source per call */
    int max_frames;            /* per channel per call */
    float bits_per_sec;        /* 9600 */
    float clockrec_gain;       /* 0.04 */
    float omega_relative_limit; /* 0.01 */
    int fftlen;                /* 1024 */
    double lpf_cutoff, lpf_transition; /* 11000, 1000 */
    int hdlc_length_min, hdlc_length_max; /* 11, 64 */
} b200ais_rx_config;

typedef struct b200ais_rx b200ais_rx;
B200AIS_API int b200ais_rx_default_config(b200ais_rx_config *cfg);
/* symbols: the corr_est template, digital.gmsk_mod(samples_per_symbol, 0.4) of the preamble
 * (python/ais_demod.py:36-38) at b200ais_rx_samples_per_symbol() */
B200AIS_API int b200ais_rx_create(b200ais_rx **h, const b200ais_rx_config *cfg,
                                  const float *symbols_iq, int nsymbols);
B200AIS_API int b200ais_rx_destroy(b200ais_rx *h);
B200AIS_API int b200ais_rx_reset(b200ais_rx *h);
B200AIS_API int b200ais_rx_decimation(const b200ais_rx *h);
B200AIS_API int b200ais_rx_channels(const b200ais_rx *h);
B200AIS_API float b200ais_rx_samples_per_symbol(const b200ais_rx *h); /* python/radio.py:57 */
B200AIS_API int b200ais_rx_sentence_slot(const b200ais_rx *h);
/* Feed the next `nitems` wideband items of every source (iq: [sources][iq_stride] complex) and
 * collect the messages completed in this call: msgs [max_msgs] (channel = s*nfreqs + k; end_bit
 * = absolute position in that channel's bit stream), their NMEA sentence(s) in
 * sentences [max_msgs][slot] with lens [max_msgs], *nmsgs of them.  Messages of one channel are
 * in stream order; the order across channels is unspecified. */
B200AIS_API int b200ais_rx_work(b200ais_rx *h, const float *iq, size_t iq_stride, int nitems,
                                b200ais_frame *msgs, char *sentences, int slot, int *lens,
                                int max_msgs, int *nmsgs);
/* device pointers everywhere (nmsgs too), asynchronous on `stream`; b200ais_rx_status after it
 * has completed */
B200AIS_API int b200ais_rx_work_dev(b200ais_rx *h, const float *iq, size_t iq_stride, int nitems,
                                    b200ais_frame *msgs, char *sentences, int slot, int *lens,
                                    int max_msgs, int *nmsgs, void *stream);
B200AIS_API int b200ais_rx_status(b200ais_rx *h);
/* calls in which some channel met more corr_est tags than its row holds (the extra tags were
 * dropped, i.e. some timing resets were missed; the call's messages are delivered all the same) */
B200AIS_API uint64_t b200ais_rx_tag_overflows(const b200ais_rx *h);

/* ---- self-test hooks (tests/ only) ----
 * The AGC kernel's straight-line IEEE division (device_math.cuh: div_rn_inrange) against the `/`
 * operator on the device, for the dividend a and every float b whose bit pattern lies in
 * [b_lo_bits, b_hi_bits]: *mismatches = the number of b with different quotient bits. */
B200AIS_API int b200ais_selftest_div(float a, uint32_t b_lo_bits, uint32_t b_hi_bits,
                                     unsigned long long *mismatches);
/* Recorded-IQ replay: blocks.file_source(gr.sizeof_gr_complex, path) (python/radio.py:211-213)
 * feeding every source of the receiver with the same capture.  The file (raw interleaved
 * float32 IQ) is read in chunks of chunk_items through two pinned buffers, the read of the next
 * chunk overlapping the copy and the processing of the current one; each chunk crosses PCIe
 * once and is replicated on the device.  sink (nullable) is called once per chunk that
 * completed messages, on the calling thread. */
typedef void (*b200ais_rx_sink)(void *user, const b200ais_frame *msgs, const char *sentences,
                                int slot, const int *lens, int nmsgs);
B200AIS_API int b200ais_rx_replay_file(b200ais_rx *h, const char *path, int chunk_items,
                                       int max_msgs, b200ais_rx_sink sink, void *user,
                                       uint64_t *items_read);
/* The same pump fed by blocks.udp_source(gr.sizeof_gr_complex, ip, port) (python/radio.py:
 * 204-210): datagram payloads are a byte stream of raw float32 IQ items.  Returns after a
 * zero-length datagram (GNU Radio's end-of-stream mark), after max_items items (0 = no limit), or
 * when no datagram arrives for idle_ms. */
B200AIS_API int b200ais_rx_serve_udp(b200ais_rx *h, const char *bind_ip, int port, int chunk_items,
                                     int max_msgs, uint64_t max_items, int idle_ms,
                                     b200ais_rx_sink sink, void *user, uint64_t *items_read);

#ifdef __cplusplus
}
#endif
#endif /* B200AIS_H */
