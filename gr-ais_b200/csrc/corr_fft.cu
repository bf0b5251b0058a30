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
    // transform read consecutive entries (the plain table is read with strides of 2, 4, 8 ...
    // entries there: 2- to 8-way bank conflicts).  Pass over bits [S0, S0+4): stage bq uses
    // W[((jq << S0) | lo) << (LOGF - S0 - bq - 1)], jq < 2^bq, lo < 2^S0; stored at
    // ((2^bq - 1 + jq) << S0) + lo, 15 << S0 entries per pass.
    static constexpr int TOP = LOGF - 4;
    static constexpr int TWP_TOP = (LOGF > 4) ? (15 << TOP) : 0;
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

// one pass: R radix-2 stages on index bits [S0, S0+R), forward DIF or inverse DIT
template <int LOGF, int S0, int R, bool INV>
__device__ __forceinline__ void run_pass(float2 (&v)[16], int t, const float2 *__restrict__ tw,
                                         const float2 *__restrict__ twp)
{
    constexpr int G = 16 >> R, Q = 1 << R, F = 1 << LOGF;
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int lo = (t * G + g) & ((1 << S0) - 1);
#pragma unroll
        for (int st = 0; st < R; st++) {
            const int bq = INV ? st : (R - 1 - st);  // bit of q this stage pairs on
            const int beta = S0 + bq;                // index bit
#pragma unroll
            for (int q = 0; q < Q; q++) {
                if (q & (1 << bq))
                    continue;
                const int q1 = q | (1 << bq);
                const int jq = (q & ((1 << bq) - 1)) << S0;
                float2 &a = v[g * Q + q], &b = v[g * Q + q1];
                const float2 a0 = a, b0 = b;
                if (S0 == 0 && (jq << (LOGF - beta - 1)) == 0) { // W = 1
                    a = make_float2(a0.x + b0.x, a0.y + b0.y);
                    b = make_float2(a0.x - b0.x, a0.y - b0.y);
                } else if (S0 == 0 && (jq << (LOGF - beta - 1)) == F / 4) { // W = -i (conj: +i)
                    if (INV) { // t = i*b = (-b.y, b.x)
                        a = make_float2(a0.x - b0.y, a0.y + b0.x);
                        b = make_float2(a0.x + b0.y, a0.y - b0.x);
                    } else { // (a-b) * (-i) = (d.y, -d.x)
                        const float dx = a0.x - b0.x, dy = a0.y - b0.y;
                        a = make_float2(a0.x + b0.x, a0.y + b0.y);
                        b = make_float2(dy, -dx);
                    }
                } else {
                    float2 w;
                    if constexpr (S0 > 0) // same table entry, conflict-free position
                        w = twp[((((1 << bq) - 1) + (q & ((1 << bq) - 1))) << S0) + lo];
                    else
                        w = tw[(jq | lo) << (LOGF - beta - 1)];
                    if (INV) {
                        w.y = -w.y;
                        const float2 tt = cmul_fma(w, b0);
                        a = make_float2(a0.x + tt.x, a0.y + tt.y);
                        b = make_float2(a0.x - tt.x, a0.y - tt.y);
                    } else {
                        a = make_float2(a0.x + b0.x, a0.y + b0.y);
                        b = cmul_fma(w, make_float2(a0.x - b0.x, a0.y - b0.y));
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
template <int LOGF>
__device__ __forceinline__ void block_filter(float2 (&v)[16], int t, const float2 *__restrict__ tw,
                                             const float2 *__restrict__ twp,
                                             const float4 *__restrict__ h4, float2 *xb)
{
    constexpr int NP = Plan<LOGF>::NPASS, REM = Plan<LOGF>::REM, NT = Plan<LOGF>::NT;
    constexpr int TOP = LOGF - 4;
    const float2 *twp_mid = twp + Plan<LOGF>::TWP_TOP;
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
    run_pass<LOGF, TOP, 4, false>(v, t, tw, twp);
    if constexpr (NP >= 2) {
        to_smem<TOP, 4>(v, t, xb);
        xsync();
        constexpr int S1 = (NP == 2) ? 0 : TOP - 4;
        constexpr int R1 = (NP == 2 && REM) ? REM : 4;
        from_smem<S1, R1>(v, t, xb);
        xsync();
        run_pass<LOGF, S1, R1, false>(v, t, tw, twp_mid);
        if constexpr (NP >= 3) {
            to_smem<S1, R1>(v, t, xb);
            xsync();
            constexpr int R2 = REM ? REM : 4;
            from_smem<0, R2>(v, t, xb);
            xsync();
            run_pass<LOGF, 0, R2, false>(v, t, tw, nullptr);
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
            const float4 hh = h4[k * NT + t];
            v[2 * k] = cmul_fma(v[2 * k], make_float2(hh.x, hh.y));
            v[2 * k + 1] = cmul_fma(v[2 * k + 1], make_float2(hh.z, hh.w));
        }
        // ---- inverse, low bits first ----
        run_pass<LOGF, SL, RL, true>(v, t, tw, nullptr);
    }
    if constexpr (NP >= 3) {
        constexpr int R2 = REM ? REM : 4;
        constexpr int S1 = TOP - 4;
        to_smem<0, R2>(v, t, xb);
        xsync();
        from_smem<S1, 4>(v, t, xb);
        xsync();
        run_pass<LOGF, S1, 4, true>(v, t, tw, twp_mid);
        to_smem<S1, 4>(v, t, xb);
        xsync();
        from_smem<TOP, 4>(v, t, xb);
        xsync();
        run_pass<LOGF, TOP, 4, true>(v, t, tw, twp);
    } else if constexpr (NP == 2) {
        constexpr int R1 = REM ? REM : 4;
        to_smem<0, R1>(v, t, xb);
        xsync();
        from_smem<TOP, 4>(v, t, xb);
        xsync();
        run_pass<LOGF, TOP, 4, true>(v, t, tw, twp);
    }
}

// LT: the tap count when it is known at compile time (0 = run-time L): with F = 256 and the
// north-star's 120 taps every `element < ns` test on a slot other than q = 8 folds away.
template <int LOGF, int LT>
__global__ void __launch_bounds__(Plan<LOGF>::THREADS, (LOGF <= 8 ? 4 : (LOGF <= 10 ? 3 : 1)))
k_corr_fft(const float2 *__restrict__ in, size_t in_stride, int nblocks, int L_rt, int nb_per_cta,
           const float2 *__restrict__ tw, const float2 *__restrict__ hbr, float thresh,
           const float2 *__restrict__ tail_in, float2 *__restrict__ tail_out,
           uint32_t *__restrict__ mask, size_t mask_stride_words, float2 *__restrict__ corr_out,
           size_t corr_stride, int channels)
{
    using P = Plan<LOGF>;
    if (channel_index() >= channels)
        return;
    constexpr int F = P::F, NT = P::NT, GROUPS = P::GROUPS;
    const int L = LT ? LT : L_rt;
    extern __shared__ float4 smem_raw[];
    float2 *s_tw = reinterpret_cast<float2 *>(smem_raw);     // [F/2]
    float2 *s_h = s_tw + F / 2;                              // [F] as float4 pairs [8][NT]
    float2 *s_twp = s_h + F;                                 // [TWP] per-stage twiddles of the upper passes
    float2 *s_x = s_twp + P::TWP;                            // [GROUPS][F + F/16]
    float2 *s_tail = s_x + GROUPS * (F + F / 16);            // [3][GROUPS][L-1], round r uses r % 3
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_tail + 3 * GROUPS * (L > 1 ? L - 1 : 1));

    const int ns = F - L + 1, tl = L - 1;
    const int c = channel_index();
    const int b0 = blockIdx.x * nb_per_cta;
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const int nwords = (nb_per_cta * ns) >> 5;
    for (int i = threadIdx.x; i < F / 2; i += blockDim.x)
        s_tw[i] = tw[i];
    for (int i = threadIdx.x; i < F; i += blockDim.x) { // H[16 t + s] -> pair k = s/2 of thread t
        const int tt = i >> 4, sl = i & 15;
        s_h[(((sl >> 1) * NT + tt) << 1) | (sl & 1)] = hbr[i];
    }
    for (int i = threadIdx.x; i < P::TWP; i += blockDim.x) {
        const bool mid = i >= P::TWP_TOP;
        const int j = mid ? i - P::TWP_TOP : i;
        const int s0 = mid ? P::TOP - 4 : P::TOP;
        const int k1 = (j >> s0) + 1, lo = j & ((1 << s0) - 1); // k1 = 2^bq + jq, 1..15
        const int bq = 31 - __clz(k1), jq = k1 - (1 << bq);
        s_twp[i] = tw[((jq << s0) | lo) << (LOGF - s0 - bq - 1)];
    }
    for (int i = threadIdx.x; i < nwords; i += blockDim.x)
        s_mask[i] = 0u;
    __syncthreads();

    const float2 *xc = in + (size_t)c * in_stride;
    float2 *xb = s_x + g * (F + F / 16);
    const int first = b0 > 0 ? b0 - 1 : 0;                   // lead block: only its tail is used
    const int last = min(b0 + nb_per_cta, nblocks);          // exclusive
    const int rounds = (last - first + GROUPS - 1) / GROUPS;
    // the input block of round r+1 is requested before round r's transform, so that HBM latency
    // runs under it (the slots at or beyond ns are zero padding and never loaded)
    auto fetch = [&](float2 (&d)[16], int bb) {
        const bool ok = bb < last;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int e = t + NT * q;
            d[q] = (ok && e < ns) ? xc[(size_t)bb * ns + e] : make_float2(0.0f, 0.0f);
        }
    };
    float2 vnext[16];
    fetch(vnext, first + g);
    for (int r = 0; r < rounds; r++) {
        const int b = first + r * GROUPS + g;
        const bool valid = b < last;
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; q++)
            v[q] = vnext[q];
        if (r + 1 < rounds)
            fetch(vnext, b + GROUPS);
        block_filter<LOGF>(v, t, s_tw, s_twp, reinterpret_cast<const float4 *>(s_h), xb);
        // stash this block's tail (outputs ns .. F-1) for the next block
        float2 *my_tail = s_tail + ((r % 3) * GROUPS + g) * tl;
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
        if (valid && b >= b0) {
            const float2 *prev = nullptr; // tail of block b-1
            if (g > 0)
                prev = s_tail + ((r % 3) * GROUPS + g - 1) * tl;
            else if (r > 0)
                prev = s_tail + (((r - 1) % 3) * GROUPS + GROUPS - 1) * tl;
            else if (b == 0 && tail_in)
                prev = tail_in + (size_t)c * tl; // state carried from the previous work() call
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int e = t + NT * q;
                if (e < ns) {
                    float2 y = v[q];
                    if (e < tl && prev) {
                        const float2 pt = prev[e];
                        y.x += pt.x;
                        y.y += pt.y;
                    }
                    const size_t idx = (size_t)b * ns + e;
                    if (corr_out)
                        corr_out[(size_t)c * corr_stride + idx] = y;
                    // volk_32fc_magnitude_squared_32f, then `mag <= thresh` skips (:191,197)
                    const float mag = y.x * y.x + y.y * y.y;
                    if (!(mag <= thresh)) {
                        const int bit = (b - b0) * ns + e;
                        atomicOr(&s_mask[bit >> 5], 1u << (bit & 31));
                    }
                }
            }
        }
        // three rotating tail buffers: the slots read above (rounds r and r-1) are next written
        // in round r+2, after round r+1's barrier
    }
    __syncthreads();
    uint32_t *mrow = mask + (size_t)c * mask_stride_words + (((size_t)b0 * ns) >> 5);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x)
        mrow[i] = s_mask[i];
}

template <int LOGF, int LT>
int launch_one(const float2 *in, size_t in_stride, int channels, int nblocks, int L, int nb,
               const float2 *tw, const float2 *hbr, float thresh, const float2 *tail_in,
               float2 *tail_out, uint32_t *mask, size_t msw, float2 *corr_out, size_t corr_stride,
               cudaStream_t s)
{
    using P = Plan<LOGF>;
    const int ns = P::F - L + 1;
    size_t smem = sizeof(float2) * (size_t)(P::F / 2 + P::F + P::TWP + P::GROUPS * (P::F + P::F / 16) +
                                            3 * P::GROUPS * (L > 1 ? L - 1 : 1)) +
                  sizeof(uint32_t) * (size_t)((nb * ns) >> 5) + 16;
    if (smem > 200 * 1024) {
        set_error("corr_est: %d taps need %zu bytes of shared memory", L, smem);
        return B200AIS_E_INVALID;
    }
    B200_CU(cudaFuncSetAttribute(k_corr_fft<LOGF, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    dim3 grid = channel_grid((nblocks + nb - 1) / nb, channels);
    k_corr_fft<LOGF, LT><<<grid, P::THREADS, smem, s>>>(in, in_stride, nblocks, L, nb, tw, hbr, thresh,
                                                    tail_in, tail_out, mask, msw, corr_out,
                                                    corr_stride, channels);
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
    int nb0 = 32 / g;
    int nb = nb0;
    while (nb < 32 && nb * ns < 8192)
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

int launch_corr_fft(const float2 *in, size_t in_stride, int channels, int n, int L,
                    const float2 *tw, const float2 *hbr, float thresh, const float2 *tail_in,
                    float2 *tail_out, uint8_t *mask, size_t mask_stride, float2 *corr_out,
                    size_t corr_stride, cudaStream_t s)
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
#define B200_CASE(LG)                                                                             \
    case LG:                                                                                      \
        return launch_one<LG, 0>(in, in_stride, channels, nblocks, L, nb, tw, hbr, thresh,       \
                                 tail_in, tail_out, m32, msw, corr_out, corr_stride, s);
    // the three templates of python/ais_demod.py:36-38: tap count folded in
    static int generic = -1; // B200AIS_CORR_GENERIC=1: run-time tap count everywhere (experiment)
    if (generic < 0) {
        const char *e = getenv("B200AIS_CORR_GENERIC");
        generic = (e && atoi(e)) ? 1 : 0;
    }
#define B200_KNOWN(LG, LEN)                                                                       \
    if (L == LEN && !generic)                                                                            \
        return launch_one<LG, LEN>(in, in_stride, channels, nblocks, L, nb, tw, hbr, thresh,      \
                                   tail_in, tail_out, m32, msw, corr_out, corr_stride, s);
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
