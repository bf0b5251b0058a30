// internal.h -- shared declarations of libb200ais.so (not part of the C-ABI).
//
// Canonical arithmetic: every kernel is compiled with -fmad=false, so
// a multiply-add is fused only where the source says fmaf()/__fmaf_rn(); division,
// sqrt and float<->double conversions are the IEEE round-to-nearest CUDA defaults.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "b200ais.h"

namespace b200ais {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
void count_launch(int n = 1);

#define B200_CU(x)                                                              \
    do {                                                                        \
        cudaError_t e__ = (x);                                                  \
        if (e__ != cudaSuccess)                                                 \
            return ::b200ais::cuda_fail(e__, #x, __FILE__, __LINE__);           \
    } while (0)

#define B200_LAUNCH_CHECK(name)                                                 \
    do {                                                                        \
        ::b200ais::count_launch();                                              \
        cudaError_t e__ = cudaGetLastError();                                   \
        if (e__ != cudaSuccess)                                                 \
            return ::b200ais::cuda_fail(e__, name, __FILE__, __LINE__);         \
    } while (0)

// ---- small device-buffer helper: grow-only scratch owned by a handle ----
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap)
            return B200AIS_OK;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        B200_CU(cudaMalloc(&p, bytes));
        cap = bytes;
        return B200AIS_OK;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// Channels ride on grid.y, and on grid.z beyond the 65535 limit of grid.y:
// c = blockIdx.z * gridDim.y + blockIdx.y, CTAs with c >= channels return at once.
inline dim3 channel_grid(unsigned x, int channels)
{
    const unsigned z = ((unsigned)channels + 65534u) / 65535u;
    const unsigned y = ((unsigned)channels + z - 1) / z;
    return dim3(x, y, z);
}
#ifdef __CUDACC__
__device__ __forceinline__ int channel_index() { return (int)(blockIdx.z * gridDim.y + blockIdx.y); }
#endif

// Regenerated GNU Radio data tables, resident in global memory of the current device.
struct Tables {
    const float *mmse; // [129][8]  mmse_fir_interpolator taps
    const float *atan; // [257]     gr::fast_atan2f
    const float *sine; // [1024][2] gr::fxpt sine table {slope, intercept}
    const float *sine4; // [1024][4] the same, entry i = {sine[i], sine[(i + 256) % 1024]}: the
                        // sine and cosine segments of one angle in a single 16-byte load
    cudaTextureObject_t sine4_tex; // sine4 as a 1-D float4 texture (point fetch: the same bits
                                   // through the texture pipe instead of the load/store pipe)
};
int get_tables(Tables *t);
// FFT twiddles for length n on the current device: n/2 complex, W[k] = exp(-2 pi i k/n)
// rounded from double, k = 0 and k = n/4 exact.  Cached per (device, n).
int get_twiddles(int n, const float2 **tw);

// status words written by kernels (device int, 0 = ok, else a B200AIS_E_* code)
struct MskParams {
    float sps_half; // d_sps
    float gain, gain_omega, limit;
    int osps;
    int pair_fetch; // experiments: B200AIS_MSK_NO_PAIR_FETCH=1
    int no_slide; // experiments: B200AIS_MSK_NO_SLIDE=1 turns the sliding-window round off
};

// Per-channel loop state of msk_timing_recovery_cc (lib/msk_timing_recovery_cc_impl.h:37-46)
struct MskState {
    float mu, omega;
    float dly1_re, dly1_im, dly2_re, dly2_im, diff1_re, diff1_im;
    int div;
    float prev_re, prev_im; // in[-1]
    int pad;
};

// Stream state of the bit tail: the last two symbols and the last slicer decision.
struct TailCarry {
    float m2x, m2y, m1x, m1y;
    int bprev, pad;
};

// ---- launch wrappers (each returns B200AIS_OK or an error code) ----

// G1 + A8 first half: square -> FFT -> shifted |.| -> argmax, one block per (vector, channel).
// vstride: row pitch of the per-vector arrays (raw, fhat).
// raw[c*vstride+b] = maxpos (j + offset/2) or -1 when no bin pair had energy > 0.
int launch_sqfft_freqest(const float2 *x, size_t x_stride, int channels, int nvec, int vstride,
                         int fftlen, int offset, int *raw, cudaStream_t s);
// A8 on caller-supplied spectra (stand-alone freqest block)
int launch_freqest_spec(const float2 *spec, int channels, int nvec, int fftlen, int offset,
                        int *raw, cudaStream_t s);
// A8 second half for the stand-alone block: maxpos carry-over + Hz conversion
int launch_freqest_resolve(const int *raw, int channels, int nvec, int fftlen, float binsize,
                           float *out, cudaStream_t s);
// G2 serial part: maxpos carry-over, Hz conversion and the NCO phase recurrence, one thread
// per channel; phase checkpoints every `seg` samples, ckpt[(n/seg)*channels + c].
// phase_state (nullable): [channels] NCO phase carried across calls (read, then updated).
int launch_nco_phase(const int *raw, int channels, int nvec, int vstride, int fftlen, float binsize,
                     float sens, float *fhat, float *ckpt, int seg, float *phase_state, cudaStream_t s);
// G2 parallel part + G3: mix with the NCO and apply feedforward_agc_cc.
// stages: B200AIS_STAGE_* mask.  out rows have `out_stride` items; out[c*out_stride + t].
int launch_mix_agc(const float2 *x, size_t x_stride, int channels, int n1, int fftlen,
                   const float *fhat, int vstride, const float *ckpt, int seg, float sens, int stages,
                   int agc_nsamples, float agc_reference, float2 *out, size_t out_stride,
                   const float2 *hist_in, float2 *hist_out, cudaStream_t s);
// A3: the correlation filter (GNU Radio's fft_filter_ccc: FFT overlap-add), |.|^2 > thresh
// bitmask and the correlator stream.  in rows: in[c*in_stride + t], t in [0, n), n a multiple
// of the block size fftsize - L + 1.  tw: fftsize/2 twiddles, hbr: fftsize transformed taps in
// bit-reversed order (make_corr_spectrum).  tail_in/tail_out: [channels][L-1] filter state
// (nullptr = zero / discard); they must not alias.
int corr_fft_size(int L);
int corr_blocks_per_cta(int L);
size_t corr_mask_stride_bytes(int L, int n);
int make_corr_spectrum(const float *taps_iq, int L, float2 *hbr_host /* [fftsize] */);
// the kernel's immediates for the 16th roots of unity against a host twiddle table of length n
int corr_check_w16(const float2 *tw_host, int n);
int sqfft_set_w32(const float2 *tw_host);
// sparse != 0: corr_out only has to serve k_detect (blocks without a sample above the threshold
// write their first and last item only).  in_readable: items readable from `in` in every row
// (>= n; the bulk-copy path rounds an odd n up to the next 16 bytes).
int launch_corr_fft(const float2 *in, size_t in_stride, int channels, int n, int L,
                    const float2 *tw, const float2 *hbr, float thresh, const float2 *tail_in,
                    float2 *tail_out, uint8_t *mask, size_t mask_stride, float2 *corr_out,
                    size_t corr_stride, int sparse, int in_readable, cudaStream_t s);
// A4: the serial detector, one warp per channel, on the correlator stream.
int launch_detect(const float2 *corr, size_t corr_stride, int channels, int n_total, int chunk,
                  int nsamples_mult, int isps, unsigned mark_delay, const uint8_t *mask,
                  size_t mask_stride, uint64_t base_offset, int two_ports, b200ais_tag *tags,
                  int max_tags, int *ntags, int *status, int append, cudaStream_t s);
// A7: the timing-loop recurrence, one lane per channel (symbols out).
// share_sm: the kernel will run beside other kernels of the chain (prefer the smallest ring).
// bits (nullable): the bit tail of a fresh chain (quadrature demod -> slicer -> diff decoder ->
// invert, k_tail's arithmetic) fused into the loop: one byte per symbol to bits[c * bits_stride + k]
// and NO symbol output (out is not written; out_err / out_mu must be null).
// unconsumed (nullable, stream mode): [channels] items in front of `in` that the last call left
// unconsumed (read, then updated); the rows must hold them and one more item in front.
int launch_msk(const float2 *in, size_t in_stride, int channels, int noutput_items,
               int ninput_items, uint64_t nitems_read, const b200ais_tag *tags, int max_tags,
               const int *ntags, MskParams p, MskState *state, float2 *out, float *out_err,
               float *out_mu, size_t out_stride, int *nproduced, int *nconsumed,
               int require_unbounded, int *status, int *unconsumed, cudaStream_t s, int share_sm = 0,
               uint8_t *bits = nullptr, size_t bits_stride = 0);
// G4-G6 + A9: quadrature demod -> slicer -> diff decoder -> invert on the symbol stream.
int launch_tail(const float2 *sym, size_t sym_stride, const int *nsym, int channels, int max_sym,
                uint8_t *bits, size_t bits_stride, float *soft, TailCarry *carry, cudaStream_t s);
// stream mode of the tag list: drop the tags the timing loop can no longer see (offset <
// written - unconsumed[c]); nold[c] = tags kept.  k_detect then appends (append != 0).
int launch_tags_compact(b200ais_tag *tags, int max_tags, int *ntags, const int *unconsumed,
                        uint64_t written, int channels, int *nold, cudaStream_t s);
// copy the tags appended since launch_tags_compact to the caller's rows
int launch_tags_emit(const b200ais_tag *tags, int max_tags, const int *ntags, const int *nold,
                     int channels, b200ais_tag *out_tags, int *out_ntags, int *status, cudaStream_t s);
// move items [from, from+len) of every row to the front of the row (regions may overlap)
int launch_roll_rows(float2 *rows, size_t stride, int channels, int from, int len, cudaStream_t s);
int launch_msk_reset(MskState *state, int channels, float sps_half, cudaStream_t s);
int launch_msk_set_omega(MskState *state, int channels, float omega, cudaStream_t s);
int launch_invert(const uint8_t *in, uint8_t *out, size_t n, cudaStream_t s);
// interleaved int16 I/Q rows -> complex float rows: out = (float)in * scale per component
int launch_sc16_to_fc(const int16_t *in, size_t in_stride, float2 *out, size_t out_stride, int channels,
                      int n, float scale, cudaStream_t s);
// ais_rx output (framing.cu): dense message list + its sentences
int launch_gather_frames(const b200ais_frame *frames, const int *nframes, int channels,
                         int max_frames, b200ais_frame *dense, int max_msgs, int *count,
                         int *status, cudaStream_t s);
int launch_nmea_dense(const b200ais_frame *dense, const int *count, int max_msgs,
                      const char *designators, int des_mod, char *sentences, int slot, int *lens,
                      cudaStream_t s);
int launch_copy_delay(const float2 *in, size_t in_stride, float2 *out, size_t out_stride,
                      int channels, int n, cudaStream_t s);

} // namespace b200ais
