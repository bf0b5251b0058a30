// corr_fft.cu -- corr_est_cc's correlation filter as GNU Radio runs it: kernel::fft_filter_ccc,
// an FFT overlap-add filter (reference lib/corr_est_cc_impl.cc:77,84,188), followed by
// volk_32fc_magnitude_squared_32f and the threshold compare (:191,197).
//
// Canonical arithmetic (the oracle does the same):
//   fftsize F = 2 * 2^ceil(log2 L), block ns = F - L + 1 items;
//   per block: zero-pad to F -> radix-2 DIF forward (natural in, bit-reversed out) ->
//   multiply by the transformed taps (volk product order) -> radix-2 DIT inverse with
//   conjugated twiddles (bit-reversed in, natural out, unnormalised; the taps carry 1/F) ->
//   the first L-1 outputs get the previous block's tail added, the last L-1 become the tail.
//
// Mapping: F/16 threads own one block transform, 16 values per thread.  A pass runs up to four
// consecutive radix-2 stages in registers on the values that differ only in those index bits;
// between passes the values cross a padded shared-memory buffer.  DIF leaves the spectrum in
// bit-reversed positions, which is exactly what the DIT inverse wants, so the taps spectrum is
// stored bit-reversed and nothing is permuted.  A CTA walks NB consecutive blocks of one channel
// (NB * ns is a multiple of 32, so it owns whole words of the detector bitmask), GROUPS blocks at
// a time, handing each block's tail to the next through shared memory; it recomputes the block
// before its range only for that tail.
//
// Blackwell specifics (sm_100a):
//  * Every complex add / subtract / twiddle product is issued as packed FP32 (FADD2 / FMUL2 /
//    FFMA2 on the (re, im) register pair, device_math.cuh): the kernel is bound by instruction
//    issue, not by HBM, and a radix-2 butterfly drops from 8 issue slots to 4 with identical bits.
//    The twiddles of the lowest pass are the 16th roots of unity, whatever F is: they are
//    immediates; the upper passes read (wr, wr, wi, wi) entries, one 16-byte load per butterfly.
//  * Input: when the rows are 16-byte aligned a CTA streams its blocks through a shared-memory
//    buffer filled by the bulk-copy engine (cp.async.bulk.shared::cluster.global + mbarrier
//    complete_tx; UBLKCP / SYNCS in the SASS).  A "full" / "empty" pair of mbarriers hands the
//    buffer back and forth: the threads copy the round's items into registers and arrive on
//    "empty", the elected thread then issues the copy of round r+1, which has the whole of
//    round r to land.  One buffer, not two: the 9 KB it frees (with the transformed taps read
//    through L1 instead of staged) make room for a fifth CTA per SM.  No load instruction,
//    address arithmetic or staging register is spent on the IQ.  Unaligned rows take the
//    register-prefetch path (LDG.64).
//  * The correlator stream (8 B/sample) is only needed by the detector around samples above
//    the threshold: in `sparse` mode a block is written only when it holds such a sample (plus
//    the first and last item of every block, the detector's neighbours across a block edge).
#include <cstdlib>

#include "device_math.cuh"
#include "internal.h"

namespace b200ais {

namespace {

__device__ __forceinline__ int xphys(int e) { return e + (e >> 4); }

template <int LOGF> struct Plan {
    static constexpr int F = 1 << LOGF;
    static constexpr int NT = F / 16;                       // threads per transform
    static constexpr int THREADS = NT > 128 ? NT : 128;     // CTA size
    static constexpr int GROUPS = THREADS / NT;             // transforms in flight per CTA
    static constexpr int NPASS = (LOGF + 3) / 4;
    static constexpr int REM = LOGF % 4;                    // width of the lowest pass (0 = full)
    // Twiddles of the passes above the lowest one, re-ordered per stage so that the threads of a
    // transform read consecutive entries.  Pass over bits [S0, S0+4): stage bq uses
    // W[((jq << S0) | lo) << (LOGF - S0 - bq - 1)], jq < 2^bq, lo < 2^S0; stored at
    // ((2^bq - 1 + jq) << S0) + lo, 15 << S0 entries per pass, each as (wr, wr, wi, wi).
    static constexpr int TOP = LOGF - 4;
    // Two-pass transforms (F = 32 .. 256): the top pass's 15 twiddles depend on the thread only,
    // not on the block, so they live in registers for the whole CTA (no table, no loads).
    static constexpr bool TWREG = NPASS == 2;
    static constexpr int TWP_TOP = (LOGF > 4 && !TWREG) ? (15 << TOP) : 0;
    static constexpr int TWP_MID = (NPASS >= 3) ? (15 << (TOP - 4)) : 0;
    static constexpr int TWP = TWP_TOP + TWP_MID;
};

// element index of slot (g, q) of thread t for a pass over index bits [S0, S0+R)
template <int S0, int R> __host__ __device__ constexpr __forceinline__ int elem(int t, int g, int q)
{
    const int u = t * (16 >> R) + g;
    const int lo = u & ((1 << S0) - 1), hi = u >> S0;
    return (hi << (S0 + R)) | (q << S0) | lo;
}

// W_16^k = exp(-2 pi i k / 16), k = 0..7, rounded from double exactly as get_twiddles() rounds
// (checked against the table on the host, corr_check_w16): the twiddles of a pass over the
// lowest index bits, for every transform length.
#define B200_C8 0.92387953251128674f  // cos(pi/8)
#define B200_S8 0.38268343236508978f  // sin(pi/8)
#define B200_R2 0.70710678118654752f  // sqrt(1/2)
__device__ __forceinline__ float2 w16(int k)
{
    switch (k & 7) {
    case 1:
        return make_float2(B200_C8, -B200_S8);
    case 2:
        return make_float2(B200_R2, -B200_R2);
    case 3:
        return make_float2(B200_S8, -B200_C8);
    case 5:
        return make_float2(-B200_S8, -B200_C8);
    case 6:
        return make_float2(-B200_R2, -B200_R2);
    case 7:
        return make_float2(-B200_C8, -B200_S8);
    default:
        return make_float2(1.0f, 0.0f); // k = 0 and k = 4 never come here (add / swap only)
    }
}

// one pass: R radix-2 stages on index bits [S0, S0+R), forward DIF or inverse DIT
// twr: the pass's 15 twiddles of this thread in registers (TWREG: they do not depend on the block)
// ZQ: slots ZQ .. 15 hold zero padding on entry (forward top pass with a compile-time tap count):
// the first stage's butterflies on them are a' = a + 0, b' = W (a - 0), so only the product is
// computed (a' keeps a: the value is the same, and a zero keeps its own sign instead of +0)
template <int LOGF, int S0, int R, bool INV, bool TWREG = false, int ZQ = 16>
__device__ __forceinline__ void run_pass(float2 (&v)[16], int t, const float4 *__restrict__ twp,
                                         const float2 (&twr)[15])
{
    constexpr int G = 16 >> R, Q = 1 << R;
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int lo = (t * G + g) & ((1 << S0) - 1);
#pragma unroll
        for (int st = 0; st < R; st++) {
            const int bq = INV ? st : (R - 1 - st);  // bit of q this stage pairs on
#pragma unroll
            for (int q = 0; q < Q; q++) {
                if (q & (1 << bq))
                    continue;
                const int q1 = q | (1 << bq);
                const int jq = q & ((1 << bq) - 1);
                // S0 == 0: W = exp(-2 pi i jq / 2^(bq+1)) = W_16^(jq << (3 - bq))
                const int k16 = jq << (3 - bq);
                float2 &a = v[g * Q + q], &b = v[g * Q + q1];
                const float2 a0 = a, b0 = b;
                if (!INV && st == 0 && g * Q + q1 >= ZQ && TWREG) { // b0 == 0: b' = W a, a' = a
                    b = cmul_fma2(twr[((1 << bq) - 1) + jq], a0);
                    continue;
                }
                if (S0 == 0 && k16 == 0) { // W = 1
                    a = f2_add(a0, b0);
                    b = f2_sub(a0, b0);
                } else if (S0 == 0 && k16 == 4) { // W = -i (conj: +i)
                    if (INV) { // t = i*b = (-b.y, b.x)
                        a = f2_add(a0, make_float2(-b0.y, b0.x));
                        b = f2_add(a0, make_float2(b0.y, -b0.x));
                    } else { // (a-b) * (-i) = (d.y, -d.x)
                        const float2 d = f2_sub(a0, b0);
                        a = f2_add(a0, b0);
                        b = make_float2(d.y, -d.x);
                    }
                } else if (S0 == 0) {
                    const float2 w = w16(k16);
                    if (INV) {
                        const float2 tt = cmul_fma2_conj(w, b0);
                        a = f2_add(a0, tt);
                        b = f2_sub(a0, tt);
                    } else {
                        a = f2_add(a0, b0);
                        b = cmul_fma2(w, f2_sub(a0, b0));
                    }
                } else if (TWREG) {
                    const float2 w = twr[((1 << bq) - 1) + jq];
                    if (INV) {
                        const float2 tt = cmul_fma2_conj(w, b0);
                        a = f2_add(a0, tt);
                        b = f2_sub(a0, tt);
                    } else {
                        a = f2_add(a0, b0);
                        b = cmul_fma2(w, f2_sub(a0, b0));
                    }
                } else {
                    const float4 e = twp[((((1 << bq) - 1) + jq) << S0) + lo];
                    if (INV) {
                        const float2 tt = cmul_tw4_conj(e, b0);
                        a = f2_add(a0, tt);
                        b = f2_sub(a0, tt);
                    } else {
                        a = f2_add(a0, b0);
                        b = cmul_tw4(e, f2_sub(a0, b0));
                    }
                }
            }
        }
    }
}

template <int S0, int R>
__device__ __forceinline__ void to_smem(const float2 (&v)[16], int t, float2 *xb)
{
#pragma unroll
    for (int s = 0; s < 16; s++)
        xb[xphys(elem<S0, R>(t, s >> R, s & ((1 << R) - 1)))] = v[s];
}
template <int S0, int R>
__device__ __forceinline__ void from_smem(float2 (&v)[16], int t, const float2 *xb)
{
#pragma unroll
    for (int s = 0; s < 16; s++)
        v[s] = xb[xphys(elem<S0, R>(t, s >> R, s & ((1 << R) - 1)))];
}

// forward transform, multiply by the taps spectrum, inverse transform; v enters and leaves in
// the top-pass mapping (element t + NT*q in slot q)
template <int LOGF, int ZQ>
__device__ __forceinline__ void block_filter(float2 (&v)[16], int t, const float4 *__restrict__ twp,
                                             const float2 (&twr)[15], const float4 *__restrict__ h4,
                                             float2 *xb)
{
    constexpr int NP = Plan<LOGF>::NPASS, REM = Plan<LOGF>::REM, NT = Plan<LOGF>::NT;
    constexpr int TOP = LOGF - 4;
    constexpr bool TR = Plan<LOGF>::TWREG;
    const float4 *twp_mid = twp + Plan<LOGF>::TWP_TOP;
    // the exchange buffer belongs to one transform: only its own threads have to meet
    auto xsync = [&]() {
        if constexpr (Plan<LOGF>::NT <= 32)
            __syncwarp();
        else if constexpr (Plan<LOGF>::GROUPS == 1)
            __syncthreads();
        else
            asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / Plan<LOGF>::NT)),
                         "r"(Plan<LOGF>::NT));
    };
    // ---- forward, top bits first ----
    run_pass<LOGF, TOP, 4, false, TR, ZQ>(v, t, twp, twr);
    if constexpr (NP >= 2) {
        to_smem<TOP, 4>(v, t, xb);
        xsync();
        constexpr int S1 = (NP == 2) ? 0 : TOP - 4;
        constexpr int R1 = (NP == 2 && REM) ? REM : 4;
        from_smem<S1, R1>(v, t, xb);
        xsync();
        run_pass<LOGF, S1, R1, false>(v, t, twp_mid, twr);
        if constexpr (NP >= 3) {
            to_smem<S1, R1>(v, t, xb);
            xsync();
            constexpr int R2 = REM ? REM : 4;
            from_smem<0, R2>(v, t, xb);
            xsync();
            run_pass<LOGF, 0, R2, false>(v, t, nullptr, twr);
        }
    }
    // ---- pointwise product with the (bit-reversed) transformed taps: volk multiply(X, H) ----
    {
        constexpr int SL = 0;
        constexpr int RL = (NP == 1) ? 4 : (REM ? REM : 4);
        // slot s of thread t holds spectrum element 16 t + s in every lowest-pass mapping; the
        // taps are staged as float4 {H[16t+2k], H[16t+2k+1]} at [k][t]: conflict-free 16-byte reads
        static_assert(elem<SL, RL>(3, 5 >> RL, 5 & ((1 << RL) - 1)) == 16 * 3 + 5, "slot order");
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 hh = __ldg(h4 + k * NT + t);
            v[2 * k] = cmul_fma2(v[2 * k], make_float2(hh.x, hh.y));
            v[2 * k + 1] = cmul_fma2(v[2 * k + 1], make_float2(hh.z, hh.w));
        }
        // ---- inverse, low bits first ----
        run_pass<LOGF, SL, RL, true>(v, t, nullptr, twr);
    }
    if constexpr (NP >= 3) {
        constexpr int R2 = REM ? REM : 4;
        constexpr int S1 = TOP - 4;
        to_smem<0, R2>(v, t, xb);
        xsync();
        from_smem<S1, 4>(v, t, xb);
        xsync();
        run_pass<LOGF, S1, 4, true>(v, t, twp_mid, twr);
        to_smem<S1, 4>(v, t, xb);
        xsync();
        from_smem<TOP, 4>(v, t, xb);
        xsync();
        run_pass<LOGF, TOP, 4, true>(v, t, twp, twr);
    } else if constexpr (NP == 2) {
        constexpr int R1 = REM ? REM : 4;
        to_smem<0, R1>(v, t, xb);
        xsync();
        from_smem<TOP, 4>(v, t, xb);
        xsync();
        run_pass<LOGF, TOP, 4, true, TR>(v, t, twp, twr);
    }
}

// ---- bulk-copy (TMA) input ring: mbarrier + cp.async.bulk, one elected thread ----
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// group-wide OR of a predicate over the NT threads of one transform
template <int LOGF> __device__ __forceinline__ bool group_any(bool p)
{
    using P = Plan<LOGF>;
    if constexpr (P::NT <= 32) {
        const unsigned bal = __ballot_sync(0xffffffffu, p);
        if constexpr (P::NT == 32)
            return bal != 0u;
        const unsigned sh = ((threadIdx.x & 31) / P::NT) * P::NT;
        return ((bal >> sh) & ((1u << P::NT) - 1u)) != 0u;
    } else if constexpr (P::GROUPS == 1) {
        return __syncthreads_or(p ? 1 : 0) != 0;
    } else {
        unsigned r;
        asm volatile("{\n"
                     ".reg .pred pi, po;\n"
                     "setp.ne.u32 pi, %1, 0;\n"
                     "bar.red.or.pred po, %2, %3, pi;\n"
                     "selp.u32 %0, 1, 0, po;\n"
                     "}"
                     : "=r"(r)
                     : "r"(p ? 1u : 0u), "r"(8 + (int)(threadIdx.x / P::NT)), "r"(P::NT));
        return r != 0u;
    }
}

// LT: the tap count when it is known at compile time (0 = run-time L): with F = 256 and the
// north-star's 120 taps every `element < ns` test on a slot other than q = 8 folds away.
// TMA: input through the bulk-copy ring (rows 16-byte aligned), else register prefetch.
template <int LOGF, int LT, bool TMA>
__global__ void __launch_bounds__(Plan<LOGF>::THREADS, (LOGF <= 8 ? 5 : (LOGF <= 10 ? 3 : 1)))
k_corr_fft(const float2 *__restrict__ in, size_t in_stride, int nblocks, int L_rt, int nb_per_cta,
           const float2 *__restrict__ tw, const float2 *__restrict__ hbr, float thresh,
           const float2 *__restrict__ tail_in, float2 *__restrict__ tail_out,
           uint32_t *__restrict__ mask, size_t mask_stride_words, float2 *__restrict__ corr_out,
           size_t corr_stride, int channels, int sparse)
{
    using P = Plan<LOGF>;
    if (channel_index() >= channels)
        return;
    constexpr int F = P::F, NT = P::NT, GROUPS = P::GROUPS;
    const int L = LT ? LT : L_rt;
    const int ns = F - L + 1, tl = L - 1;
    const int stage_items = (GROUPS * ns + 3) & ~1; // one round of input, +1 item of alignment slack, even
    extern __shared__ float4 smem_raw[];
    float4 *s_twp = smem_raw;                                               // [TWP] (wr, wr, wi, wi)
    float2 *s_in = reinterpret_cast<float2 *>(s_twp + P::TWP);              // [stage_items] (TMA)
    float2 *s_x = s_in + (TMA ? stage_items : 0);                           // [GROUPS][F + F/16]
    // block tails: groups 0 .. GROUPS-2 hand theirs to the next group of the same round (two
    // rotating sets), the last group to group 0 of the next round (three rotating slots)
    float2 *s_tail = s_x + GROUPS * (F + F / 16);                           // [2][GROUPS-1][L-1]
    float2 *s_last = s_tail + 2 * (GROUPS - 1) * (L > 1 ? L - 1 : 1);       // [3][L-1]
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_last + 3 * (L > 1 ? L - 1 : 1));
    const int nwords = (nb_per_cta * ns) >> 5;
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<uintptr_t>(s_mask + nwords + 1) + 7 & ~(uintptr_t)7); // [2] mbarriers (TMA)

    const int c = channel_index();
    const int b0 = blockIdx.x * nb_per_cta;
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const float2 *xc = in + (size_t)c * in_stride;
    const int first = b0 > 0 ? b0 - 1 : 0;                   // lead block: only its tail is used
    const int last = min(b0 + nb_per_cta, nblocks);          // exclusive
    const int rounds = (last - first + GROUPS - 1) / GROUPS;

    // round r of the ring: items [i0 - (i0 & 1), end) of the row, end rounded up to even
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(s_bar);
    const unsigned in0 = (unsigned)__cvta_generic_to_shared(s_in);
    auto issue = [&](int r) {
        const int i0 = (first + r * GROUPS) * ns;
        const int i1 = min(first + (r + 1) * GROUPS, last) * ns;
        const int lo = i0 & ~1, hi = (i1 + 1) & ~1;
        const unsigned bytes = (unsigned)(hi - lo) * 8u;
        mbar_expect_tx(bar0, bytes);
        bulk_g2s(in0, xc + lo, bytes, bar0);
    };
    if constexpr (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(bar0, 1);               // "full": the round's bytes have landed
            mbar_init(bar0 + 8, blockDim.x);  // "empty": every thread has copied its items out
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < P::TWP; i += blockDim.x) {
        const bool mid = i >= P::TWP_TOP;
        const int j = mid ? i - P::TWP_TOP : i;
        const int s0 = mid ? P::TOP - 4 : P::TOP;
        const int k1 = (j >> s0) + 1, lo = j & ((1 << s0) - 1); // k1 = 2^bq + jq, 1..15
        const int bq = 31 - __clz(k1), jq = k1 - (1 << bq);
        const float2 w = tw[((jq << s0) | lo) << (LOGF - s0 - bq - 1)];
        s_twp[i] = make_float4(w.x, w.x, w.y, w.y);
    }
    for (int i = threadIdx.x; i < nwords; i += blockDim.x)
        s_mask[i] = 0u;
    __syncthreads();
    if constexpr (TMA) {
        if (threadIdx.x == 0)
            issue(0);
    }

    float2 *xb = s_x + g * (F + F / 16);
    float2 twr[15];
    if constexpr (P::TWREG) {
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) { // k1 = 2^bq + jq: W[((jq << TOP) | t) << (LOGF - TOP - bq - 1)]
            const int bq = 31 - __clz(k1), jq = k1 - (1 << bq);
            twr[k1 - 1] = tw[((jq << P::TOP) | t) << (LOGF - P::TOP - bq - 1)];
        }
    }
    // register-prefetch path: the input block of round r+1 is requested before round r's
    // transform (the slots at or beyond ns are zero padding and never loaded)
    auto fetch = [&](float2 (&d)[16], int bb) {
        const bool ok = bb < last;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int e = t + NT * q;
            d[q] = (ok && e < ns) ? xc[(size_t)bb * ns + e] : make_float2(0.0f, 0.0f);
        }
    };
    constexpr bool PREFETCH = !TMA && !P::TWREG; // (the register twiddles use what the prefetch would)
    float2 vnext[PREFETCH ? 16 : 1];
    if constexpr (PREFETCH)
        fetch(vnext, first + g);
    for (int r = 0; r < rounds; r++) {
        const int b = first + r * GROUPS + g;
        const bool valid = b < last;
        float2 v[16];
        if constexpr (TMA) {
            mbar_wait(bar0, (unsigned)(r & 1));
            const float2 *src = s_in + (((first + r * GROUPS) * ns) & 1) + g * ns;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int e = t + NT * q;
                v[q] = (valid && e < ns) ? src[e] : make_float2(0.0f, 0.0f);
            }
            // the round's items are in registers: hand the buffer back; the elected thread
            // refills it with the next round as soon as everybody has, a whole round ahead of use
            mbar_arrive(bar0 + 8);
            if (threadIdx.x == 0 && r + 1 < rounds) {
                mbar_wait(bar0 + 8, (unsigned)(r & 1));
                issue(r + 1);
            }
        } else if constexpr (PREFETCH) {
#pragma unroll
            for (int q = 0; q < 16; q++)
                v[q] = vnext[q];
            if (r + 1 < rounds)
                fetch(vnext, b + GROUPS);
        } else {
            fetch(v, b);
        }
        // slots whose smallest element t + NT q is zero padding for every thread: q >= ns / NT
        constexpr int ZQ = LT ? ((P::F - LT + 1) + NT - 1) / NT : 16;
        block_filter<LOGF, ZQ>(v, t, s_twp, twr, reinterpret_cast<const float4 *>(hbr), xb);
        // stash this block's tail (outputs ns .. F-1) for the next block
        float2 *my_tail = g < GROUPS - 1 ? s_tail + ((r & 1) * (GROUPS - 1) + g) * tl : s_last + (r % 3) * tl;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int e = t + NT * q;
            if (valid && e >= ns)
                my_tail[e - ns] = v[q];
        }
        if (valid && b == nblocks - 1 && tail_out) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int e = t + NT * q;
                if (e >= ns)
                    tail_out[(size_t)c * tl + (e - ns)] = v[q];
            }
        }
        __syncthreads();
        const bool emit = valid && b >= b0;
        bool hit = false;
        if (emit) {
            const float2 *prev = nullptr; // tail of block b-1
            if (g > 0)
                prev = s_tail + ((r & 1) * (GROUPS - 1) + g - 1) * tl;
            else if (r > 0)
                prev = s_last + ((r - 1) % 3) * tl;
            else if (b == 0 && tail_in)
                prev = tail_in + (size_t)c * tl; // state carried from the previous work() call
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int e = t + NT * q;
                if (e < ns) {
                    if (e < tl && prev)
                        v[q] = f2_add(v[q], prev[e]);
                    // volk_32fc_magnitude_squared_32f, then `mag <= thresh` skips (:191,197)
                    const float2 sq = f2_mul(v[q], v[q]);
                    const float mag = sq.x + sq.y;
                    if (!(mag <= thresh)) {
                        const int bit = (b - b0) * ns + e;
                        atomicOr(&s_mask[bit >> 5], 1u << (bit & 31));
                        hit = true;
                    }
                }
            }
        }
        if (corr_out) {
            // the detector reads the stream at samples above the threshold and their two
            // neighbours: a block without such a sample only has to supply its edge items
            const bool all = !sparse || group_any<LOGF>(hit);
            if (emit) {
                float2 *orow = corr_out + (size_t)c * corr_stride + (size_t)b * ns;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int e = t + NT * q;
                    if (e < ns && (all || e == 0 || e == ns - 1))
                        orow[e] = v[q];
                }
            }
        }
        // a slot of s_tail read above is next written in round r+2, after round r+1's barrier,
        // which its reader has passed; the last group's slot is read one round later: three
    }
    __syncthreads();
    uint32_t *mrow = mask + (size_t)c * mask_stride_words + (((size_t)b0 * ns) >> 5);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x)
        mrow[i] = s_mask[i];
}

template <int LOGF, int LT, bool TMA>
int launch_one(const float2 *in, size_t in_stride, int channels, int nblocks, int L, int nb,
               const float2 *tw, const float2 *hbr, float thresh, const float2 *tail_in,
               float2 *tail_out, uint32_t *mask, size_t msw, float2 *corr_out, size_t corr_stride,
               int sparse, cudaStream_t s)
{
    using P = Plan<LOGF>;
    const int ns = P::F - L + 1;
    const int stage_items = (P::GROUPS * ns + 3) & ~1;
    size_t smem = sizeof(float4) * (size_t)P::TWP +
                  sizeof(float2) * (size_t)((TMA ? stage_items : 0) +
                                            P::GROUPS * (P::F + P::F / 16) +
                                            (2 * (P::GROUPS - 1) + 3) * (L > 1 ? L - 1 : 1)) +
                  sizeof(uint32_t) * (size_t)(((nb * ns) >> 5) + 1) + 32;
    if (smem > 200 * 1024) {
        set_error("corr_est: %d taps need %zu bytes of shared memory", L, smem);
        return B200AIS_E_INVALID;
    }
    B200_CU(cudaFuncSetAttribute(k_corr_fft<LOGF, LT, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    dim3 grid = channel_grid((nblocks + nb - 1) / nb, channels);
    k_corr_fft<LOGF, LT, TMA><<<grid, P::THREADS, smem, s>>>(in, in_stride, nblocks, L, nb, tw, hbr,
                                                         thresh, tail_in, tail_out, mask, msw,
                                                         corr_out, corr_stride, channels, sparse);
    B200_LAUNCH_CHECK("k_corr_fft");
    return B200AIS_OK;
}

} // namespace

int corr_fft_size(int L)
{
    int p = 1;
    while (p < L)
        p <<= 1;
    return 2 * p;
}

// blocks per CTA: a multiple of 32/gcd(ns, 32) so the CTA's outputs cover whole bitmask words
int corr_blocks_per_cta(int L)
{
    const int F = corr_fft_size(L), ns = F - L + 1;
    int g = 32, a = ns;
    while (a) {
        int tmp = g % a;
        g = a;
        a = tmp;
    }
    // long enough that the CTA's start-up (taps, twiddles, mbarriers) and the block recomputed
    // in front of its range for the tail are a few percent of its work
    int nb0 = 32 / g;
    int nb = nb0;
    while (nb < 128 && nb * ns < 16384)
        nb += nb0;
    return nb;
}

size_t corr_mask_stride_bytes(int L, int n)
{
    const int F = corr_fft_size(L), ns = F - L + 1, nb = corr_blocks_per_cta(L);
    const int nblocks = n / ns;
    const size_t ctas = (size_t)(nblocks + nb - 1) / nb;
    size_t words = (ctas ? ctas : 1) * (((size_t)nb * ns) >> 5);
    return words * 4;
}

// The immediates of the lowest pass must be the table's own values (both sides of the parity
// contract derive the 16th roots of unity from get_twiddles()/make_twiddles()).
int corr_check_w16(const float2 *tw_host, int n)
{
    if (n < 16)
        return B200AIS_OK;
    const float want[8][2] = { { 1.0f, 0.0f },         { B200_C8, -B200_S8 }, { B200_R2, -B200_R2 },
                               { B200_S8, -B200_C8 },  { 0.0f, -1.0f },       { -B200_S8, -B200_C8 },
                               { -B200_R2, -B200_R2 }, { -B200_C8, -B200_S8 } };
    for (int k = 0; k < 8; k++) {
        const float2 w = tw_host[k * (n / 16)];
        if (w.x != want[k][0] || w.y != want[k][1]) {
            set_error("twiddle table and the kernel's W_16^%d immediates disagree (n = %d)", k, n);
            return B200AIS_E_INVALID;
        }
    }
    return B200AIS_OK;
}

int launch_corr_fft(const float2 *in, size_t in_stride, int channels, int n, int L,
                    const float2 *tw, const float2 *hbr, float thresh, const float2 *tail_in,
                    float2 *tail_out, uint8_t *mask, size_t mask_stride, float2 *corr_out,
                    size_t corr_stride, int sparse, int in_readable, cudaStream_t s)
{
    if (n <= 0 || channels <= 0)
        return B200AIS_OK;
    const int F = corr_fft_size(L), ns = F - L + 1;
    if (n % ns) {
        set_error("corr_est: noutput_items (%d) must be a multiple of the output multiple (%d)", n, ns);
        return B200AIS_E_INVALID;
    }
    int lg = 0;
    while ((1 << lg) < F)
        lg++;
    const int nblocks = n / ns, nb = corr_blocks_per_cta(L);
    uint32_t *m32 = reinterpret_cast<uint32_t *>(mask);
    const size_t msw = mask_stride / 4;
    // bulk copies move whole 16-byte units: the rows must be 16-byte aligned and an odd item
    // count needs one readable item behind the last block
    static int no_tma = -1; // B200AIS_CORR_NO_TMA=1: register-prefetch input everywhere (experiment)
    if (no_tma < 0) {
        const char *e = getenv("B200AIS_CORR_NO_TMA");
        no_tma = (e && atoi(e)) ? 1 : 0;
    }
    // ... and the round buffer must fit beside the rest (tested up to fftsize 2048)
    const bool tma = !no_tma && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (in_stride & 1) == 0 &&
                     ((n & 1) == 0 || in_readable > n) && F <= 2048;
#define B200_CASE(LG)                                                                             \
    case LG:                                                                                      \
        return tma ? launch_one<LG, 0, true>(in, in_stride, channels, nblocks, L, nb, tw, hbr,    \
                                             thresh, tail_in, tail_out, m32, msw, corr_out,       \
                                             corr_stride, sparse, s)                              \
                   : launch_one<LG, 0, false>(in, in_stride, channels, nblocks, L, nb, tw, hbr,   \
                                              thresh, tail_in, tail_out, m32, msw, corr_out,      \
                                              corr_stride, sparse, s);
    // the three templates of python/ais_demod.py:36-38: tap count folded in
    static int generic = -1; // B200AIS_CORR_GENERIC=1: run-time tap count everywhere (experiment)
    if (generic < 0) {
        const char *e = getenv("B200AIS_CORR_GENERIC");
        generic = (e && atoi(e)) ? 1 : 0;
    }
#define B200_KNOWN(LG, LEN)                                                                       \
    if (L == LEN && !generic)                                                                     \
        return tma ? launch_one<LG, LEN, true>(in, in_stride, channels, nblocks, L, nb, tw, hbr,  \
                                               thresh, tail_in, tail_out, m32, msw, corr_out,     \
                                               corr_stride, sparse, s)                            \
                   : launch_one<LG, LEN, false>(in, in_stride, channels, nblocks, L, nb, tw, hbr, \
                                                thresh, tail_in, tail_out, m32, msw, corr_out,    \
                                                corr_stride, sparse, s);
    B200_KNOWN(8, 120)
    B200_KNOWN(9, 140)
    B200_KNOWN(12, 1120)
#undef B200_KNOWN
    switch (lg) {
        B200_CASE(4)
        B200_CASE(5)
        B200_CASE(6)
        B200_CASE(7)
        B200_CASE(8)
        B200_CASE(9)
        B200_CASE(10)
        B200_CASE(11)
        B200_CASE(12)
    default:
        set_error("corr_est supports 5..2048 taps (fft size 16..4096), got %d taps", L);
        return B200AIS_E_INVALID;
    }
#undef B200_CASE
}

} // namespace b200ais
