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
#include <cstdlib>

#include "device_math.cuh"
#include "internal.h"

namespace b200ais {

namespace {

// Ring geometry.  A warp of the loop is one dependent instruction stream (an instruction every
// ~6 cycles whatever else the SM holds), so the kernel's throughput is the number of warps an SM
// holds, and what bounds that is the shared memory of the per-lane rings; but a small ring leaves
// less room to request samples ahead of the DRAM latency (~2 us under load = ~2 rounds).  Three
// geometries, picked by the channel count (launch_msk):
//   kind 0: 128-sample ring, 32-sample chunks, 8-step rounds, 43 KB/warp ->  4 warps/SM (18 944 ch)
//   kind 1:  96-sample ring, 16-sample chunks, 8-step rounds, 29 KB/warp ->  7 warps/SM (33 152 ch)
//   kind 2:  48-sample ring, 12-sample chunks, 4-step rounds, 15 KB/warp -> 14 warps/SM (66 304 ch)
// The warps of a CTA are on their own; they share the 4 KB interpolator table, which has to sit
// in shared memory (a table row read through L1 misses often enough to show on the critical
// path).  Small batches run one warp per CTA so that every SM gets its share of the warps; the
// large ones pack 7 warps behind one table (7 or 14 warps per SM).
#ifndef MSK2_CHUNK
#define MSK2_CHUNK 12
#endif
template <int KIND> struct MskCfg;
template <> struct MskCfg<0> {
    static constexpr int Chunk = 32, Ring = 128, Fast = 8, Need = 32, TagCap = 32;
};
template <> struct MskCfg<1> {
    static constexpr int Chunk = 16, Ring = 96, Fast = 8, Need = 32, TagCap = 8;
};
template <> struct MskCfg<2> {
    // 12-sample chunks (96 bytes, four to the ring): a 4-step round reads ~10 samples, so one
    // chunk per round keeps up (8-sample chunks fell behind and spent a quarter of the loop's
    // iterations waiting), and a lane may ask for the next one 0.4 to 1.5 rounds before it reads it
    // (16-sample chunks: 0.2 to 1.1).  65 536 channels: 10.3 / 10.0 / 10.2 ms with 8 / 12 / 16.
    static constexpr int Chunk = MSK2_CHUNK, Ring = 48, Fast = 4, Need = 20, TagCap = 4;
};
constexpr int kMskMirror = 8;   // samples 0..7 repeated after the ring: 10-sample reads never wrap
constexpr int kMskInner = 4;    // half-symbol steps per careful round (fewer when sps is large)
constexpr int kMskTable = 2 * 132 * 16; // the interpolator table (two arrays of half rows) in front of the rings
template <int KIND> __host__ __device__ constexpr int msk_warp_smem()
{
    return (MskCfg<KIND>::Ring + kMskMirror) / 2 * 32 * 16 + MskCfg<KIND>::TagCap * 32 * 8;
}

__device__ __forceinline__ void cp_async_8(unsigned smem, const void *gmem, int src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem), "l"(gmem), "r"(src_bytes));
}
// .cg: every item is read once, so it goes past L1 (which keeps the interpolator table);
// .L2::128B: the first 16 bytes a lane asks of a line bring the whole line into L2, so DRAM sees
// one 128-byte request per line instead of eight partial ones (each lane streams its own row)
__device__ __forceinline__ void cp_async_16(unsigned smem, const void *gmem, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_last() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

// binary_slicer_fb(quadrature_demod_cf): (pi/2) * fast_atan2f(y, x) >= 0.  Only the sign reaches
// the bit stream, and for finite arguments fast_atan2f's sign is y's -- its result is -base_angle,
// base_angle - pi or -pi/2 +- base_angle (base_angle in [0, pi/4]) for y < 0 and non-negative for
// y >= 0 (also -0 and (0, 0)) -- with one exception: y < 0 < x and |y| / |x| rounding to zero
// gives -0, which passes `>= 0`.  That corner (|y| < |x| * 2^-100 is a superset of it) and
// non-finite arguments take the table path.
// the corner cases, out of line: the loop's straight-line rounds stay short (their cost is their
// length, and a lone warp cannot hide an instruction-cache miss)
static __device__ __noinline__ unsigned slicer_bit_rare(float y, float x, const float *__restrict__ tab)
{
    return (1.57079632679489661923f * fast_atan2f_tab(y, x, tab)) >= 0 ? 1u : 0u;
}
__device__ __forceinline__ unsigned slicer_bit(float y, float x, const float *__restrict__ tab)
{
    const float ya = fabsf(y), xa = fabsf(x);
    const bool plain = xa <= 3.402823466e+38f && ya <= 3.402823466e+38f &&
                       !(y < 0.0f && x > 0.0f && ya < xa * 7.8886090522101181e-31f);
    if (plain)
        return y >= 0.0f ? 1u : 0u;
    return slicer_bit_rare(y, x, tab);
}

// Loop registers of one channel (lib/msk_timing_recovery_cc_impl.h:37-46), in registers.
struct MskLane {
    float mu, omega;
    float psq_re, psq_im;     // previous interpolant squared (conj(dly2^2) = conj of this)
    float diff1_re, diff1_im;
    float2 vlast;
    int div, iidx, oidx;
    int rpos;                 // iidx mod kMskRing
    float2 *op;               // next symbol slot
    float2 hold;              // a symbol waiting for its neighbour: pairs go out as one 16-byte store
    bool have_hold;
    bool bad_imu;
    // fused bit tail (FUSE): the symbol in front of the next one, the last slicer decision, the
    // bytes of the 32-bit word being filled, the next bit slot
    float2 tprev;
    unsigned tb, pack;
    uint8_t *bp;
    const float *atab;
};

// G4-G6 + A9 on one symbol (the arithmetic of k_tail below): quadrature_demod_cf(pi/2) ->
// binary_slicer_fb -> diff_decoder_bb(2) -> invert, one byte per symbol, four bytes per store
__device__ __forceinline__ void msk_emit_bit(MskLane &L, float2 v)
{
    const float2 cur = L.tprev;
    const float re = __fmaf_rn(v.x, cur.x, v.y * cur.y);
    const float im = __fmaf_rn(v.y, cur.x, -(v.x * cur.y));
    L.tprev = v;
    const unsigned b = slicer_bit(im, re, L.atab);
    const unsigned d = (b - L.tb) % 2u; // diff_decoder_bb(2)
    L.tb = b;
    // rows start on a 32-bit word (launch_msk checks): a word goes out when its fourth byte is in
    const unsigned sh = ((unsigned)reinterpret_cast<uintptr_t>(L.bp) & 3u) * 8u;
    L.pack |= ((d ^ 0x01u) & 0x01u) << sh; // lib/invert_impl.cc:63
    if (sh == 24u)
        *reinterpret_cast<unsigned *>(L.bp - 3) = L.pack;
    L.pack = sh == 24u ? 0u : L.pack;
    L.bp++;
}
__device__ __forceinline__ void msk_flush_bits(MskLane &L)
{
    const int nb = (int)(reinterpret_cast<uintptr_t>(L.bp) & 3u); // bytes of an unfinished word
    for (int q = 0; q < nb; q++)
        (L.bp - nb)[q] = (uint8_t)((L.pack >> (8 * q)) & 0xffu);
    L.pack = 0;
}

template <int RING> __device__ __forceinline__ int ring_pos(int iidx)
{
    int r = iidx % RING;
    return r < 0 ? r + RING : r;
}

// One half-symbol step from the row index imu (:170-201 without the tag test): interpolate,
// error detector, loop filter on odd steps, output on even steps, advance.  ring4: this lane's
// column of the 16-byte-unit ring.  Returns x = mu + omega before the floor.
// WIDE: the eight samples come as five 16-byte loads and an odd / even select (fewest shared-
// memory wavefronts: the choice when many warps share an SM); otherwise as eight 8-byte loads at
// the samples' own addresses (13 fewer instructions on a step whose cost is its instruction
// count when a warp has a scheduler to itself).  The arithmetic is packed FP32 on the (re, im)
// pairs (device_math.cuh): same roundings as the scalar forms in the comments.
// The eight samples in[rpos .. rpos+7] of a lane's ring.
// WIDE: five 16-byte loads and an odd / even select (fewest shared-memory wavefronts: the choice
// when many warps share an SM); otherwise eight 8-byte loads at the samples' own addresses (13
// fewer instructions on a step whose cost is its instruction count when a warp has a scheduler
// to itself).
template <bool WIDE>
__device__ __forceinline__ void msk_load_window(const float4 *__restrict__ ring4, int rpos, float2 (&s)[8])
{
    if (WIDE) {
        const int u0 = rpos >> 1;
        const bool par = rpos & 1;
        const float4 U0 = ring4[(u0 + 0) * 32], U1 = ring4[(u0 + 1) * 32], U2 = ring4[(u0 + 2) * 32];
        const float4 U3 = ring4[(u0 + 3) * 32], U4 = ring4[(u0 + 4) * 32];
        s[0] = par ? make_float2(U0.z, U0.w) : make_float2(U0.x, U0.y);
        s[1] = par ? make_float2(U1.x, U1.y) : make_float2(U0.z, U0.w);
        s[2] = par ? make_float2(U1.z, U1.w) : make_float2(U1.x, U1.y);
        s[3] = par ? make_float2(U2.x, U2.y) : make_float2(U1.z, U1.w);
        s[4] = par ? make_float2(U2.z, U2.w) : make_float2(U2.x, U2.y);
        s[5] = par ? make_float2(U3.x, U3.y) : make_float2(U2.z, U2.w);
        s[6] = par ? make_float2(U3.z, U3.w) : make_float2(U3.x, U3.y);
        s[7] = par ? make_float2(U4.x, U4.y) : make_float2(U3.z, U3.w);
    } else {
        // sample k of the lane sits at byte ((k >> 1) * 32) * 16 + (k & 1) * 8 of its column
        const unsigned char *col = reinterpret_cast<const unsigned char *>(ring4);
        const unsigned char *q0 = col + (rpos >> 1) * 512 + (rpos & 1) * 8; // sample rpos
        const unsigned char *q1 = col + ((rpos + 1) >> 1) * 512 + ((rpos + 1) & 1) * 8; // sample rpos + 1
        s[0] = *reinterpret_cast<const float2 *>(q0);
        s[1] = *reinterpret_cast<const float2 *>(q1);
        s[2] = *reinterpret_cast<const float2 *>(q0 + 512);
        s[3] = *reinterpret_cast<const float2 *>(q1 + 512);
        s[4] = *reinterpret_cast<const float2 *>(q0 + 1024);
        s[5] = *reinterpret_cast<const float2 *>(q1 + 1024);
        s[6] = *reinterpret_cast<const float2 *>(q0 + 1536);
        s[7] = *reinterpret_cast<const float2 *>(q1 + 1536);
    }
}

// One half-symbol step from the row index imu (:170-201 without the tag test) on the eight
// samples s[] = in[iidx .. iidx+7]: interpolate, error detector, loop filter on odd steps, output
// on even steps.  Returns x = mu + omega before the floor.  The arithmetic is packed FP32 on the
// (re, im) pairs (device_math.cuh): same roundings as the scalar forms in the comments.
// DEFER: the caller stores the symbols (the straight-line round: one symbol per two steps).
template <bool kDebug, bool PAIR, bool DEFER = false, bool FUSE = false>
__device__ __forceinline__ float msk_core(MskLane &L, int imu_c, const float2 (&s)[8],
                                          const float *__restrict__ s_mmse, const MskParams &p,
                                          float *oe, float *om)
{
    // mmse_fir_interpolator_cc::interpolate: in[0..7] . reversed row.
    // the table is kept as two arrays of half rows (16 bytes each): a row index then spreads the
    // lanes over all eight 16-byte bank groups instead of four
    const float4 ta = reinterpret_cast<const float4 *>(s_mmse)[imu_c];
    const float4 tb = reinterpret_cast<const float4 *>(s_mmse)[132 + imu_c];
    const float2 s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3], s4 = s[4], s5 = s[5], s6 = s[6], s7 = s[7];
    // p_j = in[j]*T[7-j] (+fused) in[j+4]*T[3-j]; v = (p0+p1)+(p2+p3), re and im side by side
    const float2 p0 = f2_fma(s4, make_float2(ta.w, ta.w), f2_mul(s0, make_float2(tb.w, tb.w)));
    const float2 p1 = f2_fma(s5, make_float2(ta.z, ta.z), f2_mul(s1, make_float2(tb.z, tb.z)));
    const float2 p2 = f2_fma(s6, make_float2(ta.y, ta.y), f2_mul(s2, make_float2(tb.y, tb.y)));
    const float2 p3 = f2_fma(s7, make_float2(ta.x, ta.x), f2_mul(s3, make_float2(tb.x, tb.x)));
    const float2 v = f2_add(f2_add(p0, p1), f2_add(p2, p3));
    // std::complex arithmetic as GCC emits it: (ac - bd, ad + bc), no contraction
    //   sq = v*v:            (v.x*v.x - v.y*v.y, v.x*v.y + v.y*v.x)
    //   nl = sq*conj(dly2^2): (sq_re*d_re - sq_im*d_im, sq_re*d_im + sq_im*d_re)
    const float2 m = f2_mul(v, v);
    const float vxy = v.x * v.y;
    const float sq_re = m.x - m.y, sq_im = vxy + vxy;
    const float d_re = L.psq_re, d_im = -L.psq_im;
    const float2 A = f2_mul(make_float2(d_re, d_im), make_float2(sq_re, sq_re));
    const float2 Bp = f2_mul(make_float2(d_im, d_re), make_float2(sq_im, sq_im));
    // scalar adds on purpose: ptxas 12.9 contracts mul.rn.f32x2 -> add.rn.f32x2 into one FFMA2
    // (even under -fmad=false), which would drop the rounding of the product
    const float nl_re = A.x - Bp.x, nl_im = A.y + Bp.y;
    const float err_raw = nl_re - L.diff1_re;
    // odd half-steps run the loop filter (:179-184); evaluated always, selected by parity
    const bool odd = L.div & 1;
    const float err_c = branchless_clip(err_raw, 3.0f);
    const float om_t = L.omega + p.gain_omega * err_c;
    const float om_n = p.sps_half + branchless_clip(om_t - p.sps_half, p.limit);
    const float mu_n = L.mu + p.gain * err_c;
    L.omega = odd ? om_n : L.omega;
    L.mu = odd ? mu_n : L.mu;
    if (!DEFER && (!odd || p.osps == 2)) {
        // a lane's symbols go to its own row, one 32-byte sector per store whatever its size: two
        // symbols per store halve the requests the SM sends to L2
        if (FUSE) {
            msk_emit_bit(L, v);
        } else if (PAIR) {
            const bool al = (reinterpret_cast<uintptr_t>(L.op) & 15) == 0;
            const bool st4 = L.have_hold, st2 = !L.have_hold && !al;
            if (st4)
                *reinterpret_cast<float4 *>(L.op - 1) = make_float4(L.hold.x, L.hold.y, v.x, v.y);
            if (st2)
                *L.op = v;
            L.hold = v;
            L.have_hold = !st4 && !st2;
        } else {
            *L.op = v;
        }
        L.op++;
        if (kDebug) {
            if (oe)
                oe[L.oidx] = odd ? err_c : err_raw;
            if (om)
                om[L.oidx] = L.mu;
        }
        L.oidx++;
    }
    L.div++;
    L.vlast = v;
    L.psq_re = sq_re;
    L.psq_im = sq_im;
    L.diff1_re = nl_re;
    L.diff1_im = nl_im;
    return L.mu + L.omega;
}

// ring4: this lane's column of the 16-byte-unit ring (conflict-free: the lane picks the banks).
template <bool kDebug, bool WIDE, bool PAIR, bool FUSE>
__device__ __forceinline__ float msk_step(MskLane &L, int imu_c, const float4 *__restrict__ ring4,
                                          const float *__restrict__ s_mmse, const MskParams &p,
                                          float *oe, float *om)
{
    float2 s[8];
    msk_load_window<WIDE>(ring4, L.rpos, s);
    return msk_core<kDebug, PAIR, false, FUSE>(L, imu_c, s, s_mmse, p, oe, om);
}

// The serial core of msk_timing_recovery_cc: one lane per channel.  A channel's loop is a
// recurrence on (mu, omega, iidx, div, previous interpolant): a warp is one dependent instruction
// stream, a step costs (its instructions) x ~6 cycles however many channels run, and the kernel's
// throughput is the number of warps an SM can hold.  What bounds that is shared memory: each lane
// streams its own channel through a private ring (stored as 16-byte units interleaved across
// lanes so that a lane's reads never meet another lane's banks, 8 samples mirrored behind it so
// that reads never wrap) filled by cp.async; the smallest geometry is 15 KB per warp -> 14 warps
// per SM, 2072 warps = 66 304 channels in flight on 148 SMs (BASELINE configs[2]'s 65 536 in one
// wave).  A lane requests a chunk a round before it is needed.  Up to TagCap of the lane's
// time_est tags are staged in shared memory and refilled from the list when they run out (a tag
// fetched on the loop's critical path costs a DRAM round trip; a burst carries about six).
// FUSE: the bit tail (quadrature demod -> slicer -> diff decoder -> invert) runs on every symbol
// as it is produced and the kernel writes bits instead of symbols.
template <bool kDebug, int KIND, int WARPS, bool FUSE>
__global__ void __launch_bounds__(32 * WARPS)
k_msk(const float2 *__restrict__ in, size_t in_stride, int channels, int noutput_items,
      int ninput_items, uint64_t nitems_read, const b200ais_tag *__restrict__ tags, int max_tags,
      const int *__restrict__ ntags, MskParams p, MskState *__restrict__ state,
      const float *__restrict__ g_mmse, float2 *__restrict__ out, float *__restrict__ out_err,
      float *__restrict__ out_mu, size_t out_stride, int *__restrict__ nproduced,
      int *__restrict__ nconsumed, int require_unbounded, int *__restrict__ status,
      int *__restrict__ unconsumed, uint8_t *__restrict__ bits, size_t bits_stride,
      const float *__restrict__ g_atan)
{
    using Cfg = MskCfg<KIND>;
    constexpr int kMskChunk = Cfg::Chunk, kMskRing = Cfg::Ring, kMskFast = Cfg::Fast, kMskNeed = Cfg::Need;
    constexpr int kMskTagCap = Cfg::TagCap;
    constexpr int kMskUnits = (kMskRing + kMskMirror) / 2; // 16-byte units (2 samples) per lane
    // request the next chunk once fewer than this many samples are requested ahead: the chunk it
    // replaces then ends at or before iidx - 2, so in[iidx - 1] (negative centre) stays
    constexpr int kMskAhead = kMskRing - kMskChunk - 1;
    constexpr int kMskWarpSmem = msk_warp_smem<KIND>();
    constexpr int kMskWarps = WARPS;
    extern __shared__ __align__(16) unsigned char msk_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *s_mmse = reinterpret_cast<float *>(msk_smem);
    float4 *ring = reinterpret_cast<float4 *>(msk_smem + kMskTable + warp * kMskWarpSmem);
    int2 *s_tags = reinterpret_cast<int2 *>(msk_smem + kMskTable + warp * kMskWarpSmem + kMskUnits * 32 * 16);
    for (int i = threadIdx.x; i < 129 * 8; i += 32 * kMskWarps) { // row r, tap k -> half (k >> 2), row r
        const int r = i >> 3, k = i & 7;
        s_mmse[((k >> 2) * 132 + r) * 4 + (k & 3)] = g_mmse[i];
    }
    __syncthreads(); // the only block-wide step: from here on every warp is on its own
    const int c = (blockIdx.x * kMskWarps + warp) * 32 + lane;
    if (c >= channels)
        return;

    MskState st = state[c];
    // Stream mode (unconsumed != nullptr): `in` points at the first NEW item of every row and
    // the unc items the last call left unconsumed sit right in front of it, so this channel's
    // read pointer is in - unc.  The row is then moved back by one more item where that makes
    // it 16-byte aligned (the item in front of the read pointer is in[-1] of the stream).
    const int unc = unconsumed ? unconsumed[c] : 0;
    const float2 *row = in + (size_t)c * in_stride - unc;
    const int mis = (unconsumed && (reinterpret_cast<uintptr_t>(row) & 15)) ? 1 : 0;
    row -= mis;
    const int navail = ninput_items + unc; // ninput_items[0] of this channel
    ninput_items = navail + mis;
    nitems_read -= (uint64_t)(unc + mis);
    const int ninp0 = (int)((double)navail - 3.0 * (double)p.sps_half); // :119
    if (ninp0 <= 0 || noutput_items <= 0) {
        nproduced[c] = 0;
        nconsumed[c] = 0;
        if (unconsumed)
            unconsumed[c] = navail;
        return;
    }
    const int ninp = ninp0 + mis;

    // ring preset: samples before 0 are zero, in[-1] is the item carried from the last call
    float4 *ring4 = ring + lane;
    for (int u = 0; u < kMskUnits; u++)
        ring4[u * 32] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    ring4[(kMskRing / 2 - 1) * 32] = make_float4(0.0f, 0.0f, st.prev_re, st.prev_im);
    const unsigned my_s = (unsigned)__cvta_generic_to_shared(ring4);
    const bool row16 = (reinterpret_cast<uintptr_t>(row) & 15) == 0;

    // time_est tags inside [read, read+ninp), in offset order (:125-130): up to kMskTagCap of
    // them wait in shared memory; the list is scanned on from gpos when they have been used
    const b200ais_tag *tg = tags ? tags + (size_t)c * max_tags : nullptr;
    const int nt = (tags && ntags) ? min(ntags[c], max_tags) : 0;
    auto matches = [&](const b200ais_tag &t) {
        return t.key == B200AIS_TAG_TIME_EST && t.port == 0 && t.offset >= nitems_read &&
               t.offset < nitems_read + (uint64_t)ninp;
    };
    int gpos = 0;             // next entry of tg[] to look at
    int nstaged = 0, thead = 0;
    int tag_off = 0x7fffffff; // pending tag, relative to the read pointer
    float tag_val = 0.0f;
    auto fetch_tag = [&]() { // the next tag in list order
        if (thead >= nstaged) {
            nstaged = 0;
            thead = 0;
            // four list entries per trip: their loads are independent, so a refill costs one or
            // two DRAM round trips instead of one per entry
            while (gpos < nt && nstaged < kMskTagCap) {
                b200ais_tag t[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    t[j] = tg[min(gpos + j, nt - 1)];
                int used = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (gpos + j < nt && nstaged < kMskTagCap) {
                        used = j + 1;
                        if (matches(t[j])) {
                            s_tags[nstaged * 32 + lane] =
                                make_int2((int)(t[j].offset - nitems_read), __float_as_int((float)t[j].value));
                            nstaged++;
                        }
                    }
                }
                gpos += used;
            }
        }
        tag_off = 0x7fffffff;
        if (thead < nstaged) {
            const int2 e = s_tags[thead * 32 + lane];
            tag_off = e.x;
            tag_val = __int_as_float(e.y);
        }
    };
    fetch_tag();
    const int tag_span = (int)ceilf(p.sps_half) + 1; // integer pre-test before the float compare

    float2 *oc = out + (size_t)c * out_stride;
    float *oe = kDebug && out_err ? out_err + (size_t)c * out_stride : nullptr;
    float *om = kDebug && out_mu ? out_mu + (size_t)c * out_stride : nullptr;

    // conj(dly2^2) of the reference equals conj(previous v^2): d_dly_conj_1 and _2 are always
    // assigned together (:194-195), so only the squared previous interpolant is carried.
    MskLane L;
    L.psq_re = st.dly2_re * st.dly2_re - st.dly2_im * st.dly2_im;
    L.psq_im = st.dly2_re * st.dly2_im + st.dly2_im * st.dly2_re;
    L.vlast = make_float2(st.dly1_re, st.dly1_im);
    L.mu = st.mu;
    L.omega = st.omega;
    L.diff1_re = st.diff1_re;
    L.diff1_im = st.diff1_im;
    L.div = st.div;
    L.iidx = mis;
    L.rpos = mis;
    L.oidx = 0;
    L.op = oc;
    L.hold = make_float2(0.0f, 0.0f);
    L.have_hold = false;
    L.bad_imu = false;
    L.tprev = make_float2(0.0f, 0.0f); // quadrature_demod / diff_decoder history of a fresh chain
    L.tb = 0;
    L.pack = 0;
    L.bp = FUSE ? bits + (size_t)c * bits_stride : nullptr;
    L.atab = g_atan;

    int issue_end = 0; // samples [0, issue_end) of this lane's channel have been requested
    int issue_u = 0;   // ring unit the next chunk goes to (0, 8, 16)
    int ready_end = 0; // samples [0, ready_end) are known to have landed
    bool in_last = false; // this lane has a chunk in the newest cp.async group
    int err_code = 0;
    bool active = true;
    const unsigned FULL = __activemask(); // the lanes that run the loop
    const bool pair_ok = p.pair_fetch && ((FULL >> (lane ^ 1)) & 1u); // the neighbour lane runs it too
    const unsigned pair_m = __ballot_sync(FULL, pair_ok);

    // One step advances the read index by floor(mu + omega) <= advmax items.  A round may touch
    // kMskNeed samples: the straight-line round needs kMskFast steps' worth of them, the careful
    // round runs as many steps as fit (launch_msk rejects rates for which not even one does).
    const int advmax = (int)floorf(1.0f + 3.0f * p.gain + p.sps_half + fabsf(p.limit) + 1e-3f);
    const int fast_in = kMskFast * advmax;
    const bool fast_ok = advmax >= 1 && fast_in + 8 <= kMskNeed && 3.0f * p.gain < 0.5f;
    const int inner = max(1, min(kMskInner, (kMskNeed - 8) / max(advmax, 1)));
    // every straight-line step advances by exactly 2 or 3 items (x = mu + omega after the loop
    // filter lies in [sps/2 - limit - 3 gain, 1 + 3 gain + sps/2 + limit)): the sliding-window round
    const bool slide_ok = fast_ok && advmax == 3 && p.sps_half - fabsf(p.limit) - 3.0f * p.gain >= 2.001f &&
                          p.osps == 1 && !p.no_slide;

    // cp.async groups are tracked per warp, not per lane, so requests and waits happen in
    // warp-wide rounds, decided by one warp-wide OR per round: a round first waits (only for
    // chunks requested in EARLIER rounds, a microsecond ago), then requests the next chunk for
    // every lane that is within kMskAhead samples of the end of what it has requested, then
    // runs its steps.
    //
    // When no lane can meet a tag, the end of its input or the end of its output row within
    // kMskFast steps, the steps run as one straight-line block: no per-step tests, and the
    // interpolator row of the next step comes from rint(128 x) - 128 floor(x), x = mu + omega
    // -- equal to rint(128 (x - floor(x))) because x - floor(x) is exact and 128 floor(x) is an
    // even integer -- so that conversion runs beside the floor instead of after it.
    for (;;) {
        const bool need = active && (L.iidx + kMskNeed > ready_end);
        const bool want = active && (issue_end <= L.iidx + kMskAhead);
        const bool easy = fast_ok && (L.iidx + fast_in < ninp) && (L.oidx + kMskFast <= noutput_items) &&
                          ((unsigned)(tag_off - L.iidx) >= (unsigned)(fast_in + tag_span));
        const unsigned code = (active ? 1u : 0u) | (need ? 2u : 0u) | (want ? 4u : 0u) |
                              ((active && easy) ? 0u : 8u);
        const unsigned any = __reduce_or_sync(FULL, code);
        if (!(any & 1u))
            break;
        bool starved = false;
        if (any & 2u) {
            // all but the newest group (the last round's requests, which may still be on their
            // way from DRAM): a lane asks for a chunk a round before it reads it, so that is
            // nearly always enough
            cp_async_wait_but_last();
            ready_end = in_last ? issue_end - kMskChunk : issue_end;
            if (__any_sync(FULL, active && (L.iidx + kMskNeed > ready_end))) {
                cp_async_wait_all();
                ready_end = issue_end;
                in_last = false;
                starved = __any_sync(FULL, active && (L.iidx + kMskNeed > ready_end));
            }
        }
        if (any & 4u) {
            // Lanes 2k and 2k+1 fetch their chunks together: one instruction moves one 32-byte
            // sector of the even lane's chunk (16 bytes per lane), the next one of the odd
            // lane's, so the SM sends one request per sector to L2 instead of two -- the
            // L1 -> crossbar request port is what a full wave of this kernel fills.
            const bool coop = want && row16 && issue_end + kMskChunk <= ninput_items;
            const unsigned coop_m = __ballot_sync(FULL, coop);
            const bool mc = pair_ok && coop;
            if (p.pair_fetch && (coop_m & pair_m)) {
                const float2 *my_src = row + issue_end;
                const unsigned long long sp = __shfl_xor_sync(FULL, (unsigned long long)(uintptr_t)my_src, 1);
                const int pu = __shfl_xor_sync(FULL, issue_u, 1);
                const bool pc = pair_ok && ((coop_m >> (lane ^ 1)) & 1u);
                const int half = lane & 1;
#pragma unroll
                for (int w = 0; w < 2; w++) { // w = 0: the even lane's chunk, 1: the odd lane's
                    const bool mine = half == w;
                    if (mine ? mc : pc) {
                        const float2 *src = mine ? my_src : reinterpret_cast<const float2 *>((uintptr_t)sp);
                        const int u0 = mine ? issue_u : pu;
                        const unsigned col = mine ? my_s : (half ? my_s - 16 : my_s + 16);
#pragma unroll
                        for (int i = 0; i < kMskChunk / 4; i++)
                            cp_async_16(col + (u0 + 2 * i + half) * 512, src + 2 * (2 * i + half), 16);
                        if (u0 == 0) {
#pragma unroll
                            for (int i = 0; i < kMskMirror / 4; i++)
                                cp_async_16(col + (kMskRing / 2 + 2 * i + half) * 512, src + 2 * (2 * i + half), 16);
                        }
                    }
                }
            }
            if (want) {
                const unsigned dst = my_s + issue_u * 512;
                const float2 *src = row + issue_end;
                const bool first = issue_u == 0; // also feeds the mirror units
                if (mc) {
                    // fetched with the partner lane below
                } else if (row16 && issue_end + kMskChunk <= ninput_items) {
#pragma unroll
                    for (int u = 0; u < kMskChunk / 2; u++)
                        cp_async_16(dst + 512 * u, src + 2 * u, 16);
                    if (first) {
#pragma unroll
                        for (int u = 0; u < kMskMirror / 2; u++)
                            cp_async_16(my_s + (kMskRing / 2 + u) * 512, src + 2 * u, 16);
                    }
                } else {
                    for (int k = 0; k < kMskChunk; k++) {
                        const int nb = issue_end + k < ninput_items ? 8 : 0;
                        const unsigned d = dst + 512 * (k >> 1) + 8 * (k & 1);
                        cp_async_8(d, nb ? src + k : row, nb);
                        if (first && k < kMskMirror)
                            cp_async_8(my_s + (kMskRing / 2 + (k >> 1)) * 512 + 8 * (k & 1), nb ? src + k : row, nb);
                    }
                }
                issue_end += kMskChunk;
                issue_u = issue_u + kMskChunk / 2 == kMskRing / 2 ? 0 : issue_u + kMskChunk / 2;
            }
            in_last = want;
            cp_async_commit();
        }
        if (starved)
            continue; // start-up only: go around and wait for what was just requested
        if (!(any & 8u)) {
            // ---- straight-line round: every lane runs kMskFast steps ----
            int imu = __float2int_rn(L.mu * 128.0f);
            if (slide_ok) {
                // Every step advances by 2 or 3 items here, so the eight samples stay in
                // registers: a step shifts the window by its advance and takes the two or three
                // new samples from loads issued BEFORE the step (they depend on the old index
                // only).  The ring is read once per sample instead of ~3 times, and the
                // advance -> address -> load round trip leaves the loop's critical path.
                float2 W[8];
                msk_load_window<KIND != 0>(ring4, L.rpos, W);
                const unsigned char *col = reinterpret_cast<const unsigned char *>(ring4);
                // Of two consecutive steps exactly one gives a symbol (the one whose count is
                // even), so the round stores one symbol per two steps, picked by the lane's
                // parity: no test per step.  (Debug builds keep the per-step form: they also
                // write the error and mu streams.)
                constexpr bool kDefer = !kDebug;
                const bool odd0 = L.div & 1;
                float2 vprev = make_float2(0.0f, 0.0f), e0 = vprev;
#pragma unroll
                for (int it = 0; it < kMskFast; it++) {
                    float2 N0 = W[7], N1 = W[7], N2 = W[7];
                    if (it + 1 < kMskFast) {
                        int p8 = L.rpos + 8;
                        p8 = p8 >= kMskRing ? p8 - kMskRing : p8; // p8 + 2 stays inside the mirror
                        const unsigned char *q0 = col + (p8 >> 1) * 512 + (p8 & 1) * 8;
                        const unsigned char *q1 = col + ((p8 + 1) >> 1) * 512 + ((p8 + 1) & 1) * 8;
                        N0 = *reinterpret_cast<const float2 *>(q0);
                        N1 = *reinterpret_cast<const float2 *>(q1);
                        N2 = *reinterpret_cast<const float2 *>(q0 + 512);
                    }
                    const unsigned imu_c = min((unsigned)imu, 128u); // mu in [0, 1): never clamps
                    L.bad_imu |= (imu_c != (unsigned)imu);
                    const float x = msk_core<kDebug, KIND == 2, kDefer, FUSE>(L, (int)imu_c, W, s_mmse, p, oe, om);
                    if (kDefer) {
                        if (it & 1) {
                            // the symbol of steps it-1, it
                            const float2 e = odd0 ? L.vlast : vprev;
                            if (FUSE) {
                                msk_emit_bit(L, e);
                            } else if (KIND != 2) {
                                *L.op = e;
                                L.op++;
                            } else if ((it & 3) == 1) {
                                e0 = e;
                            } else {
                                // two symbols, one 16-byte store where the row allows it:
                                //   holding one:      (hold, e0) -> op - 1, keep e
                                //   aligned, no hold: (e0, e)    -> op
                                //   neither:          e0 -> op (8 bytes), keep e
                                const bool al = (reinterpret_cast<uintptr_t>(L.op) & 15) == 0;
                                const bool hv = L.have_hold;
                                const bool wide = hv || al;
                                const float2 a = hv ? L.hold : e0, b = hv ? e0 : e;
                                float2 *dst = hv ? L.op - 1 : L.op;
                                if (wide)
                                    *reinterpret_cast<float4 *>(dst) = make_float4(a.x, a.y, b.x, b.y);
                                else
                                    *dst = a;
                                L.hold = e;
                                L.have_hold = hv || !al;
                                L.op += 2;
                            }
                        } else {
                            vprev = L.vlast;
                        }
                    }
                    const int fl_i = __float2int_rd(x);
                    imu = __float2int_rn(x * 128.0f) - 128 * fl_i;
                    L.iidx += fl_i;
                    const int rp = L.rpos + fl_i;
                    L.rpos = rp >= kMskRing ? rp - kMskRing : rp;
                    L.mu = x - floorf(x);
                    if (it + 1 < kMskFast) {
                        const bool three = fl_i == 3;
                        L.bad_imu |= ((unsigned)(fl_i - 2) > 1u); // finite input never gets here
#pragma unroll
                        for (int k = 0; k < 5; k++)
                            W[k] = three ? W[k + 3] : W[k + 2];
                        W[5] = three ? N0 : W[7];
                        W[6] = three ? N1 : N0;
                        W[7] = three ? N2 : N1;
                    }
                }
                if (kDefer)
                    L.oidx += kMskFast / 2;
                active = (L.oidx < noutput_items) && (L.iidx < ninp);
                continue;
            }
#pragma unroll
            for (int it = 0; it < kMskFast; it++) {
                const unsigned imu_c = min((unsigned)imu, 128u); // mu in [0, 1): never clamps
                L.bad_imu |= (imu_c != (unsigned)imu);
                const float x = msk_step<kDebug, KIND != 0, KIND == 2, FUSE>(L, (int)imu_c, ring4, s_mmse, p, oe, om);
                const int fl_i = __float2int_rd(x);
                imu = __float2int_rn(x * 128.0f) - 128 * fl_i;
                L.iidx += fl_i;
                const int rp = L.rpos + fl_i; // 0 <= fl_i <= advmax < kMskRing
                L.rpos = rp >= kMskRing ? rp - kMskRing : rp;
                L.mu = x - floorf(x);
            }
            active = (L.oidx < noutput_items) && (L.iidx < ninp);
            continue;
        }
        // ---- careful round: up to kMskInner steps with every test ----
#pragma unroll
        for (int it = 0; it < kMskInner; it++) {
            if (active && it < inner) {
                // tag reset (:139-164); rare: an integer window test guards the float compare
                if (((unsigned)tag_off - (unsigned)L.iidx) < (unsigned)tag_span) {
                    if ((float)tag_off < ((float)L.iidx + p.sps_half)) {
                        if (tag_val == tag_val) { // NaN: drop the tag, no reset (:144-147)
                            L.mu = tag_val;
                            L.iidx = tag_off;
                            if (L.mu < 0) {
                                L.mu = L.mu + 1.0f;
                                L.iidx--;
                            }
                            L.rpos = ring_pos<kMskRing>(L.iidx);
                            L.div = 0;
                            L.omega = p.sps_half;
                        }
                        thead++;
                        fetch_tag();
                    }
                }
                // imu = rint(mu*128); the reference's interpolator throws outside [0, 128]
                const int imu = __float2int_rn(L.mu * 128.0f);
                const int imu_c = min(max(imu, 0), 128);
                L.bad_imu |= (imu != imu_c);
                const float x = msk_step<kDebug, KIND != 0, KIND == 2, FUSE>(L, imu_c, ring4, s_mmse, p, oe, om);
                const float fl = floorf(x);
                L.iidx += (int)fl;
                L.rpos = ring_pos<kMskRing>(L.iidx);
                L.mu = x - fl;
                active = (L.oidx < noutput_items) && (L.iidx < ninp);
            }
        }
    }
    if (L.have_hold)
        L.op[-1] = L.hold;
    if (FUSE)
        msk_flush_bits(L);
    if (L.bad_imu)
        err_code = B200AIS_E_INTERP;
    cp_async_wait_all();
    st.mu = L.mu;
    st.omega = L.omega;
    st.div = L.div;
    st.diff1_re = L.diff1_re;
    st.diff1_im = L.diff1_im;
    st.dly1_re = st.dly2_re = L.vlast.x;
    st.dly1_im = st.dly2_im = L.vlast.y;
    if (L.iidx > 0) {
        const float2 pv = row[L.iidx - 1];
        st.prev_re = pv.x;
        st.prev_im = pv.y;
    }
    state[c] = st;
    if (!err_code && require_unbounded && L.oidx >= noutput_items && L.iidx < ninp)
        err_code = B200AIS_E_OUT_OVERFLOW;
    if (err_code)
        atomicMin(status, err_code);
    nproduced[c] = L.oidx;
    nconsumed[c] = L.iidx - mis;
    if (unconsumed)
        unconsumed[c] = navail - (L.iidx - mis);
}

// G4-G6 + A9 on the symbol stream: quadrature_demod_cf(pi/2) -> binary_slicer_fb ->
// diff_decoder_bb(2) -> invert.  bit[k] depends on sym[k], sym[k-1], sym[k-2] only, so every
// thread produces four consecutive bits (one 32-bit store) from six symbols.
__global__ void __launch_bounds__(128)
k_tail(const float2 *__restrict__ sym, size_t sym_stride, const int *__restrict__ nsym,
       int channels, const float *__restrict__ g_atan, uint8_t *__restrict__ bits,
       size_t bits_stride, float *__restrict__ soft_out, const TailCarry *__restrict__ carry)
{
    const float *s_atan = g_atan; // read through L1: only the soft output and odd corners need it
    const int c = channel_index();
    if (c >= channels)
        return;
    const int n = nsym[c];
    const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (k0 >= n)
        return;
    const float2 *sc = sym + (size_t)c * sym_stride;
    const float qgain = 1.57079632679489661923f; // (float)(pi/2), python/ais_demod.py:48
    const float2 zero = make_float2(0.0f, 0.0f);
    // symbols -2 and -1 and the last slicer decision come from the previous call of a stream
    TailCarry tc;
    tc.m2x = tc.m2y = tc.m1x = tc.m1y = 0.0f;
    tc.bprev = 0;
    if (carry && k0 < 2)
        tc = carry[c];
    float2 prev = k0 >= 2 ? sc[k0 - 2] : (k0 == 1 ? make_float2(tc.m1x, tc.m1y) : make_float2(tc.m2x, tc.m2y));
    float2 cur = k0 >= 1 ? sc[k0 - 1] : make_float2(tc.m1x, tc.m1y);
    // slicer decision of symbol k0-1 (0 before the first symbol: diff_decoder history)
    unsigned bprev = (unsigned)tc.bprev;
    if (k0 >= 1) {
        const float re = __fmaf_rn(cur.x, prev.x, cur.y * prev.y);
        const float im = __fmaf_rn(cur.y, prev.x, -(cur.x * prev.y));
        bprev = slicer_bit(im, re, s_atan);
    }
    unsigned pack = 0;
    const int kend = min(4, n - k0);
    for (int q = 0; q < kend; q++) {
        const float2 v = sc[k0 + q];
        // quadrature_demod_cf: x[n]*conj(x[n-1]), VOLK multiply-conjugate FMA form
        const float re = __fmaf_rn(v.x, cur.x, v.y * cur.y);
        const float im = __fmaf_rn(v.y, cur.x, -(v.x * cur.y));
        cur = v;
        unsigned b;                                 // binary_slicer_fb
        if (soft_out) {
            const float soft = qgain * fast_atan2f_tab(im, re, s_atan);
            soft_out[(size_t)c * sym_stride + k0 + q] = soft;
            b = soft >= 0 ? 1u : 0u;
        } else {
            b = slicer_bit(im, re, s_atan);
        }
        const unsigned d = (b - bprev) % 2u;        // diff_decoder_bb(2)
        bprev = b;
        pack |= ((d ^ 0x01u) & 0x01u) << (8 * q);   // lib/invert_impl.cc:63
    }
    uint8_t *ob = bits + (size_t)c * bits_stride + k0;
    if (kend == 4 && ((reinterpret_cast<uintptr_t>(ob) & 3) == 0)) {
        *reinterpret_cast<unsigned *>(ob) = pack;
    } else {
        for (int q = 0; q < kend; q++)
            ob[q] = (uint8_t)((pack >> (8 * q)) & 0xffu);
    }
}

// after k_tail of a stream call: remember the last two symbols and the last slicer decision
__global__ void k_tail_carry(const float2 *__restrict__ sym, size_t sym_stride,
                             const int *__restrict__ nsym, int channels,
                             const float *__restrict__ atan_tab, TailCarry *__restrict__ carry)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels)
        return;
    const int n = nsym[c];
    if (n <= 0)
        return;
    TailCarry tc = carry[c];
    const float2 *sc = sym + (size_t)c * sym_stride;
    if (n >= 2) {
        tc.m2x = sc[n - 2].x;
        tc.m2y = sc[n - 2].y;
    } else {
        tc.m2x = tc.m1x;
        tc.m2y = tc.m1y;
    }
    tc.m1x = sc[n - 1].x;
    tc.m1y = sc[n - 1].y;
    const float re = __fmaf_rn(tc.m1x, tc.m2x, tc.m1y * tc.m2y);
    const float im = __fmaf_rn(tc.m1y, tc.m2x, -(tc.m1x * tc.m2y));
    tc.bprev = (1.57079632679489661923f * fast_atan2f_tab(im, re, atan_tab)) >= 0 ? 1 : 0;
    carry[c] = tc;
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

// interleaved int16 I/Q -> complex float, four items (16 bytes in, 32 bytes out) per thread
__global__ void __launch_bounds__(256)
k_sc16_to_fc(const int16_t *__restrict__ in, size_t in_stride, float2 *__restrict__ out,
             size_t out_stride, int channels, int n, float scale)
{
    const int c = channel_index();
    if (c >= channels)
        return;
    const int16_t *ic = in + (size_t)c * in_stride * 2;
    float2 *oc = out + (size_t)c * out_stride;
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n)
        return;
    const bool vec = i0 + 4 <= n && ((reinterpret_cast<uintptr_t>(ic + 2 * (size_t)i0) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(oc + i0) & 15) == 0);
    if (vec) {
        const int4 w = *reinterpret_cast<const int4 *>(ic + 2 * (size_t)i0);
        const int v[4] = { w.x, w.y, w.z, w.w };
        float4 o[2];
        float *of = reinterpret_cast<float *>(o);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            of[2 * k] = (float)(short)(v[k] & 0xffff) * scale;
            of[2 * k + 1] = (float)(short)(v[k] >> 16) * scale;
        }
        float4 *o4 = reinterpret_cast<float4 *>(oc + i0);
        o4[0] = o[0];
        o4[1] = o[1];
    } else {
        for (int k = i0; k < min(i0 + 4, n); k++)
            oc[k] = make_float2((float)ic[2 * k] * scale, (float)ic[2 * k + 1] * scale);
    }
}

} // namespace

int launch_sc16_to_fc(const int16_t *in, size_t in_stride, float2 *out, size_t out_stride, int channels,
                      int n, float scale, cudaStream_t s)
{
    if (channels <= 0 || n <= 0)
        return B200AIS_OK;
    dim3 grid = channel_grid((unsigned)((n + 1023) / 1024), channels);
    k_sc16_to_fc<<<grid, 256, 0, s>>>(in, in_stride, out, out_stride, channels, n, scale);
    B200_LAUNCH_CHECK("k_sc16_to_fc");
    return B200AIS_OK;
}

int launch_msk(const float2 *in, size_t in_stride, int channels, int noutput_items,
               int ninput_items, uint64_t nitems_read, const b200ais_tag *tags, int max_tags,
               const int *ntags, MskParams p, MskState *state, float2 *out, float *out_err,
               float *out_mu, size_t out_stride, int *nproduced, int *nconsumed,
               int require_unbounded, int *status, int *unconsumed, cudaStream_t s, int share_sm,
               uint8_t *bits, size_t bits_stride)
{
    if (channels <= 0)
        return B200AIS_OK;
    Tables tb;
    int rc = get_tables(&tb);
    if (rc)
        return rc;
    // ring geometry by occupancy need (see MskCfg): the smallest number of warps per SM that
    // runs every channel in one wave; B200AIS_MSK_KIND=0/1/2 overrides (experiments)
    const int advmax = (int)floorf(1.0f + 3.0f * p.gain + p.sps_half + fabsf(p.limit) + 1e-3f);
    const int warps = (channels + 31) / 32;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess)
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int per_sm = (warps + sms - 1) / sms;
    // (kind, warps per CTA): 4 x 1, 6 x 1, 1 x 7 or 2 x 7 warps per SM
    int kind = per_sm <= 4 ? 0 : (per_sm <= 7 ? 1 : 2);
    bool packed = per_sm > 6;
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("B200AIS_MSK_KIND");
        forced = (e && *e) ? atoi(e) : -1;
    }
    // share_sm: the loop runs on a side stream under the next record's front kernels (pipelined
    // submission): the smallest ring leaves them the shared memory they need, and what the small
    // ring costs the loop itself (exposed DRAM latency at low occupancy) is hidden under them
    if (share_sm && kind < 2 && advmax + 8 <= MskCfg<2>::Need)
        kind = 2;
    if (forced >= 0 && forced <= 2) {
        kind = forced;
        packed = kind == 2 || (kind == 1 && per_sm > 6);
    }
    const bool small_spread = kind == 2 && per_sm <= 7; // one warp per CTA: every SM gets its share
    // a round must be able to make one step inside the ring's ready window
    if (kind == 2 && !(advmax + 8 <= MskCfg<2>::Need)) {
        kind = 1;
        packed = true;
    }
    if (!(advmax + 8 <= MskCfg<1>::Need)) {
        set_error("msk_timing_recovery: sps %g with gain %g / limit %g advances up to %d items per "
                  "half-symbol step; this build supports up to %d", 2.0 * p.sps_half, p.gain, p.limit,
                  advmax, MskCfg<1>::Need - 8);
        return B200AIS_E_INVALID;
    }
    {
        static int no_slide = -1;
        if (no_slide < 0) {
            const char *e = getenv("B200AIS_MSK_NO_SLIDE");
            no_slide = (e && *e && atoi(e)) ? 1 : 0;
        }
        p.no_slide = no_slide;
        static int no_pf = -1;
        if (no_pf < 0) {
            const char *e = getenv("B200AIS_MSK_NO_PAIR_FETCH");
            no_pf = (e && *e && atoi(e)) ? 1 : 0;
        }
        p.pair_fetch = (no_pf || kind == 0) ? 0 : 1; // a lone warp per scheduler: the extra instructions cost more
    }
    const bool dbg = out_err || out_mu;
#define B200_MSK(DBG, KIND, W, FUSE)                                                               \
    do {                                                                                           \
        const int per_cta = 32 * W;                                                                \
        const int blocks = (channels + per_cta - 1) / per_cta;                                     \
        const size_t smem = (size_t)kMskTable + (size_t)W * msk_warp_smem<KIND>();                 \
        static bool attr_set = false;                                                              \
        if (!attr_set) { /* all the shared memory the SM has: occupancy is the throughput */       \
            B200_CU(cudaFuncSetAttribute(k_msk<DBG, KIND, W, FUSE>,                                \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            B200_CU(cudaFuncSetAttribute(k_msk<DBG, KIND, W, FUSE>,                                \
                                         cudaFuncAttributePreferredSharedMemoryCarveout, 100));    \
            attr_set = true;                                                                       \
        }                                                                                          \
        k_msk<DBG, KIND, W, FUSE><<<blocks, per_cta, smem, s>>>(                                   \
            in, in_stride, channels, noutput_items, ninput_items, nitems_read, tags, max_tags, ntags, \
            p, state, tb.mmse, out, out_err, out_mu, out_stride, nproduced, nconsumed,             \
            require_unbounded, status, unconsumed, bits, bits_stride, tb.atan);                    \
    } while (0)
#define B200_MSK_KIND(DBG, FUSE)                                                                   \
    do {                                                                                           \
        if (kind == 0)                                                                             \
            B200_MSK(DBG, 0, 1, FUSE);                                                             \
        else if (kind == 1 && !packed)                                                             \
            B200_MSK(DBG, 1, 1, FUSE);                                                             \
        else if (kind == 1)                                                                        \
            B200_MSK(DBG, 1, 7, FUSE);                                                             \
        else if (small_spread)                                                                     \
            B200_MSK(DBG, 2, 1, FUSE);                                                             \
        else                                                                                       \
            B200_MSK(DBG, 2, 7, FUSE);                                                             \
    } while (0)
    if (bits && (dbg || (reinterpret_cast<uintptr_t>(bits) & 3) || (bits_stride & 3))) {
        set_error("msk: the fused bit tail needs word-aligned rows and has no error / mu outputs");
        return B200AIS_E_INVALID;
    }
    if (dbg)
        B200_MSK_KIND(true, false);
    else if (bits)
        B200_MSK_KIND(false, true);
    else
        B200_MSK_KIND(false, false);
#undef B200_MSK_KIND
#undef B200_MSK
    B200_LAUNCH_CHECK("k_msk");
    return B200AIS_OK;
}

int launch_tail(const float2 *sym, size_t sym_stride, const int *nsym, int channels, int max_sym,
                uint8_t *bits, size_t bits_stride, float *soft, TailCarry *carry, cudaStream_t s)
{
    if (channels <= 0 || max_sym <= 0)
        return B200AIS_OK;
    Tables tb;
    int rc = get_tables(&tb);
    if (rc)
        return rc;
    dim3 grid = channel_grid((max_sym + 4 * 128 - 1) / (4 * 128), channels);
    k_tail<<<grid, 128, 0, s>>>(sym, sym_stride, nsym, channels, tb.atan, bits, bits_stride, soft,
                                carry);
    B200_LAUNCH_CHECK("k_tail");
    if (carry) {
        k_tail_carry<<<(channels + 127) / 128, 128, 0, s>>>(sym, sym_stride, nsym, channels, tb.atan,
                                                            carry);
        B200_LAUNCH_CHECK("k_tail_carry");
    }
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
