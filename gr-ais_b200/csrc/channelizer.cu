// channelizer.cu -- the step before the demod path in the reference's receiver
// (python/radio.py:49-54): firdes.low_pass + freq_xlating_fir_filter_ccf, batched over
// wideband sources, with its C-ABI (b200ais_firdes_low_pass, b200ais_xlat_*).
//
// GNU Radio's block [G] is a decimating FIR with complex band-pass taps
// ctaps[i] = taps[i] * exp(j i fwT0) followed by a rotator at the output rate.  The canonical
// dot product (VOLK's order is machine dependent) visits the taps
// polyphase-major on four real fma chains
//     Pr += cr*xr   Pi += cr*xi   Qr += ci*xr   Qi += ci*xi,     y = (Pr - Qi, Pi + Qr),
// so two filters whose taps are complex conjugates of each other -- the reference's A and B
// channels at -/+25 kHz around 162 MHz (python/radio.py:88-89) -- share one pass:
// y_B = (Pr + Qi, Pi - Qr).
//
// k_xlat_fir: a CTA owns THREADS*8 consecutive outputs of one source.  It stages the input
// tile in shared memory de-interleaved by residue mod D (each polyphase branch is then a
// plain FIR over a contiguous row), and every thread slides an 8-item register window down
// that row: one 8-byte shared load + one broadcast tap load feed 32 fused multiply-adds.
// Rows are padded by one item in eight so the 64-byte lane stride maps to distinct banks.
// FP32-pipe bound: 4*ntaps/D fma per input item (482 at the reference's 250 ksps default).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "device_math.cuh"
#include "internal.h"

using namespace b200ais;

namespace {

constexpr int kMaxFreqs = 16;

struct RotState {
    float pr, pi, ir, ii;
    unsigned counter;
};

// rows are padded by one item per R so that a lane stride of R items maps to distinct banks
template <int R> __host__ __device__ __forceinline__ int padr(int u) { return u + (int)((unsigned)u / (unsigned)R); }

// gr::blocks::rotator [G]: table[j] = phase before item j; phase *= incr (std::complex product,
// separate multiplies and adds); every 512th item phase /= |phase|.  One lane per frequency.
__global__ void k_rot_phase(RotState *__restrict__ st, int nfreqs, int n, float2 *__restrict__ tab,
                            size_t tab_stride)
{
    const int k = threadIdx.x;
    if (k >= nfreqs)
        return;
    float pr = st[k].pr, pi = st[k].pi;
    const float ir = st[k].ir, ii = st[k].ii;
    unsigned counter = st[k].counter;
    float2 *t = tab + (size_t)k * tab_stride;
    // runs that end on a multiple of 512 items (or at n): only a run's last step can normalise,
    // so the others are a bare two-deep multiply-add chain
    for (int j = 0; j < n;) {
        const int run = min(n - j, 512 - (int)(counter & 511u));
#pragma unroll 8
        for (int q = 0; q < run - 1; q++) {
            t[j + q] = make_float2(pr, pi);
            const float nr = pr * ir - pi * ii;
            const float ni = pr * ii + pi * ir;
            pr = nr;
            pi = ni;
        }
        t[j + run - 1] = make_float2(pr, pi);
        float nr = pr * ir - pi * ii;
        float ni = pr * ii + pi * ir;
        counter += (unsigned)run;
        if ((counter & 511u) == 0) {
            const float a = hypot_canon(nr, ni);
            nr = nr / a;
            ni = ni / a;
        }
        pr = nr;
        pi = ni;
        j += run;
    }
    st[k].pr = pr;
    st[k].pi = pi;
    st[k].counter = counter;
}

template <int R> struct Acc {
    float Pr[R], Pi[R], Qr[R], Qi[R];
};

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// R taps of one polyphase branch (the branches are zero-padded to whole groups of R taps:
// fma(0, x, acc) leaves acc as it is for every finite x)
template <int kR>
__device__ __forceinline__ void fir_steps(Acc<kR> &a, float2 (&W)[kR], float2 &nxt, float2 &c,
                                          const float2 *__restrict__ tp, const float2 *__restrict__ xp,
                                          int u_next)
{
#pragma unroll
    for (int s = 0; s < kR; s++) {
        const float2 cc = c;
        c = tp[s + 1]; // the tap array has one slack item behind it
#pragma unroll
        for (int r = 0; r < kR; r++) {
            const float2 x = W[(r - s) & (kR - 1)];
            a.Pr[r] = __fmaf_rn(cc.x, x.x, a.Pr[r]);
            a.Pi[r] = __fmaf_rn(cc.x, x.y, a.Pi[r]);
            a.Qr[r] = __fmaf_rn(cc.y, x.x, a.Qr[r]);
            a.Qi[r] = __fmaf_rn(cc.y, x.y, a.Qi[r]);
        }
        W[(kR - 1 - s) & (kR - 1)] = nxt;
        nxt = xp[padr<kR>(u_next - s)];
    }
}

// one polyphase branch p: a plain FIR over the residue row xp
template <int kR>
__device__ __forceinline__ void fir_branch(Acc<kR> &acc, const float2 *__restrict__ xp,
                                           const float2 *__restrict__ tp, int r0, int b_p, int Qpad)
{
    constexpr int kFront = kR + 2;
    const int u0 = r0 + b_p + kFront; // window at tap 0: items u0 .. u0+R-1
    float2 W[kR];
#pragma unroll
    for (int i = 0; i < kR; i++)
        W[i] = xp[padr<kR>(u0 + i)];
    float2 nxt = xp[padr<kR>(u0 - 1)];
    float2 c = tp[0];
    for (int q = 0; q < Qpad; q += kR)
        fir_steps<kR>(acc, W, nxt, c, tp + q, xp, u0 - q - 2);
}

// G: residue rows resident in shared memory at a time.  G >= D: the whole tile is staged once,
// with coalesced reads (row = residue).  G < D (large decimations, e.g. 25 at 1.2 Msps): the
// branches are walked in groups of G, each group's rows staged just before it (row = branch -
// first branch of the group; a lane reads every D-th item).
template <int THREADS, int kR>
__global__ void __launch_bounds__(THREADS)
k_xlat_fir(const float2 *__restrict__ in, size_t in_stride, int n, int D, int ntaps, int ntp,
           const float2 *__restrict__ taps_pm, const int *__restrict__ pass_a,
           const int *__restrict__ pass_b, int nfreqs, const float2 *__restrict__ rot,
           size_t rot_stride, float2 *__restrict__ out, size_t out_stride, int Mp, int G)
{
    extern __shared__ float2 smem[];
    float2 *tps = smem;            // [ntp + 1] this pass's taps, polyphase-major, zero-padded
    float2 *xs = smem + ntp + 1;   // [min(G, D)][Mp] input tile by residue mod D
    constexpr int J = THREADS * kR;
    constexpr int kFront = kR + 2; // zeroed items in front of every residue row: the window of
                                   // the padded taps and its prefetch end up there
    const int tid = threadIdx.x;
    const int src = blockIdx.y, pass = blockIdx.z;
    const int jb = blockIdx.x * J;
    const int jt = min(J, n - jb);
    const int rows = min(G, D);

    const float2 *tp_g = taps_pm + (size_t)pass * ntp;
    for (int k = tid; k < ntp; k += THREADS)
        cp_async8(tps + k, tp_g + k);
    if (tid == 0)
        tps[ntp] = make_float2(0.f, 0.f);
    for (int k = tid; k < rows * kFront; k += THREADS)
        xs[(k / kFront) * Mp + padr<kR>(k % kFront)] = make_float2(0.f, 0.f);

    // items [jb*D, jb*D + (jt-1)*D + ntaps) of the source row feed this tile:
    // item t -> residue t % D, position t / D + kFront
    const float2 *row = in + (size_t)src * in_stride + (size_t)jb * D;
    const int T = (jt - 1) * D + ntaps;

    Acc<kR> acc;
#pragma unroll
    for (int r = 0; r < kR; r++)
        acc.Pr[r] = acc.Pi[r] = acc.Qr[r] = acc.Qi[r] = 0.f;
    const int r0 = tid * kR;
    const float2 *tp = tps;

    if (G >= D) {
        // asynchronously: every copy is in flight before the first one is waited for
        int a = tid % D, m = tid / D;
        const int da = THREADS % D, dm = THREADS / D;
        for (int t = tid; t < T; t += THREADS) {
            cp_async8(xs + a * Mp + padr<kR>(m + kFront), row + t);
            a += da;
            m += dm;
            if (a >= D) {
                a -= D;
                m++;
            }
        }
        cp_async_wait_all();
        __syncthreads();
        for (int p = 0; p < D; p++) {
            const int lag = ntaps - 1 - p;
            const int b_p = lag / D, a_p = lag - b_p * D;
            const int Qpad = (b_p + kR) / kR * kR; // b_p + 1 taps (p, p+D, ... < ntaps), padded
            fir_branch<kR>(acc, xs + a_p * Mp, tp, r0, b_p, Qpad);
            tp += Qpad;
        }
    } else {
        for (int p0 = 0; p0 < D; p0 += G) {
            const int g = min(G, D - p0);
            __syncthreads(); // the previous group's rows have been read
            for (int r = 0; r < g; r++) {
                const int a_p = (ntaps - 1 - (p0 + r)) % D;
                float2 *xr = xs + r * Mp;
                for (int m = tid; m * D + a_p < T; m += THREADS)
                    cp_async8(xr + padr<kR>(m + kFront), row + (size_t)m * D + a_p);
            }
            cp_async_wait_all();
            __syncthreads();
            for (int r = 0; r < g; r++) {
                const int lag = ntaps - 1 - (p0 + r);
                const int b_p = lag / D;
                const int Qpad = (b_p + kR) / kR * kR;
                fir_branch<kR>(acc, xs + r * Mp, tp, r0, b_p, Qpad);
                tp += Qpad;
            }
        }
    }

    const int ka = pass_a[pass], kb = pass_b[pass];
    float2 *oa = out + ((size_t)src * nfreqs + ka) * out_stride + jb + r0;
    const float2 *ra = rot + (size_t)ka * rot_stride + jb + r0;
#pragma unroll
    for (int r = 0; r < kR; r++) {
        if (r0 + r < jt) {
            const float yr = acc.Pr[r] - acc.Qi[r], yi = acc.Pi[r] + acc.Qr[r];
            const float2 ph = ra[r];
            oa[r] = make_float2(yr * ph.x - yi * ph.y, yr * ph.y + yi * ph.x);
        }
    }
    if (kb >= 0) {
        float2 *ob = out + ((size_t)src * nfreqs + kb) * out_stride + jb + r0;
        const float2 *rb = rot + (size_t)kb * rot_stride + jb + r0;
#pragma unroll
        for (int r = 0; r < kR; r++) {
            if (r0 + r < jt) {
                const float yr = acc.Pr[r] + acc.Qi[r], yi = acc.Pi[r] - acc.Qr[r];
                const float2 ph = rb[r];
                ob[r] = make_float2(yr * ph.x - yi * ph.y, yr * ph.y + yi * ph.x);
            }
        }
    }
}

// taps of one pass, polyphase-major, every branch zero-padded to a multiple of R
int xlat_padded_taps(int R, int D, int ntaps)
{
    int n = 0;
    for (int p = 0; p < D && p < ntaps; p++)
        n += ((ntaps - 1 - p) / D + R) / R * R;
    return n;
}

size_t xlat_smem_bytes(int threads, int R, int D, int ntaps, int *Mp_out, int rows = 0)
{
    const int J = threads * R;
    const int M = J + (ntaps + D - 1) / D + (R + 2) + 1;
    const int Mp = M + M / R + 1;
    if (Mp_out)
        *Mp_out = Mp;
    if (rows <= 0 || rows > D)
        rows = D;
    return sizeof(float2) * ((size_t)xlat_padded_taps(R, D, ntaps) + 1 + (size_t)rows * Mp);
}

} // namespace

// ======================================================== firdes::low_pass

// gr::filter::firdes::low_pass(gain, fs, fc, tw, WIN_HAMMING) [G] (python/radio.py:49):
// ntaps = (int)(53 fs / (22 tw)) made odd; float Hamming window; sinc * window evaluated in
// double, stored as float; normalised to unit DC gain.  Init-time, host (as in the reference).
extern "C" int b200ais_firdes_low_pass(double gain, double fs, double cutoff, double tw,
                                       float *taps, int cap, int *ntaps_out)
{
    if (!(fs > 0) || !(cutoff > 0) || !(cutoff <= fs / 2) || !(tw > 0) || !ntaps_out) {
        set_error("firdes_low_pass: need sampling_freq > 0, 0 < cutoff <= sampling_freq/2, width > 0");
        return B200AIS_E_RANGE; // firdes::sanity_check_1f throws std::out_of_range
    }
    int ntaps = (int)(53.0 * fs / (22.0 * tw));
    if ((ntaps & 1) == 0)
        ntaps++;
    *ntaps_out = ntaps;
    if (!taps)
        return B200AIS_OK;
    if (cap < ntaps) {
        set_error("firdes_low_pass: %d taps do not fit cap %d", ntaps, cap);
        return B200AIS_E_INVALID;
    }
    const int M = (ntaps - 1) / 2;
    const double fwT0 = 2 * M_PI * cutoff / fs;
    for (int n = -M; n <= M; n++) {
        const float w = (float)(0.54 - 0.46 * cos((2 * M_PI * (n + M)) / (ntaps - 1)));
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
    return B200AIS_OK;
}

// ============================================= freq_xlating_fir_filter_ccf

struct b200ais_xlat {
    int D = 0, ntaps = 0, nfreqs = 0, sources = 0, npass = 0;
    double fs = 0;
    std::vector<float> proto;
    std::vector<double> freqs;
    std::vector<std::vector<float2>> ctaps; // [nfreqs][ntaps]
    cudaStream_t stream = nullptr;
    float2 *d_taps[2] = {nullptr, nullptr}; // [npass][ntp] polyphase-major, branches padded to 8 / 16
    int ntp[2] = {0, 0};
    int *d_pass = nullptr;    // [2][kMaxFreqs]
    RotState *d_rot = nullptr;
    DevBuf rottab, in, out;
    bool dirty = true;
};

// build_composite_fir [G]: ctaps and the rotator increment of filter k
static void xlat_compose(b200ais_xlat *h, int k, RotState *rs)
{
    const float fwT0 = (float)(2 * M_PI * h->freqs[k] / h->fs);
    h->ctaps[k].resize(h->ntaps);
    for (int i = 0; i < h->ntaps; i++) {
        const float th = (float)i * fwT0;
        h->ctaps[k][i] = make_float2(h->proto[i] * cosf(th), h->proto[i] * sinf(th));
    }
    const float th = -fwT0 * (float)h->D;
    const float ir = cosf(th), ii = sinf(th);
    const float a = (float)sqrt((double)ir * (double)ir + (double)ii * (double)ii);
    rs->ir = ir / a; // rotator::set_phase_incr normalises
    rs->ii = ii / a;
}

// group filters with conjugate taps into one pass, upload the taps polyphase-major
static int xlat_upload(b200ais_xlat *h)
{
    const int K = h->nfreqs, nt = h->ntaps, D = h->D;
    int pa[kMaxFreqs], pb[kMaxFreqs];
    std::vector<char> used(K, 0);
    int np = 0;
    for (int k = 0; k < K; k++) {
        if (used[k])
            continue;
        used[k] = 1;
        pa[np] = k;
        pb[np] = -1;
        for (int k2 = k + 1; k2 < K && pb[np] < 0; k2++) {
            if (used[k2])
                continue;
            bool conj = true;
            for (int i = 0; i < nt && conj; i++)
                conj = h->ctaps[k][i].x == h->ctaps[k2][i].x && h->ctaps[k][i].y == -h->ctaps[k2][i].y;
            if (conj) {
                pb[np] = k2;
                used[k2] = 1;
            }
        }
        np++;
    }
    h->npass = np;
    for (int v = 0; v < 2; v++) {
        const int R = v ? 16 : 8;
        const int ntp = xlat_padded_taps(R, D, nt);
        std::vector<float2> pm((size_t)np * ntp, make_float2(0.f, 0.f));
        for (int p = 0; p < np; p++) {
            size_t o = (size_t)p * ntp;
            for (int ph = 0; ph < D && ph < nt; ph++) {
                const int Qpad = ((nt - 1 - ph) / D + R) / R * R;
                int q = 0;
                for (int k = ph; k < nt; k += D)
                    pm[o + q++] = h->ctaps[pa[p]][k];
                o += Qpad;
            }
        }
        if (h->d_taps[v])
            cudaFree(h->d_taps[v]);
        h->d_taps[v] = nullptr;
        h->ntp[v] = ntp;
        B200_CU(cudaMalloc(&h->d_taps[v], sizeof(float2) * pm.size()));
        B200_CU(cudaMemcpy(h->d_taps[v], pm.data(), sizeof(float2) * pm.size(), cudaMemcpyHostToDevice));
    }
    int both[2 * kMaxFreqs];
    memcpy(both, pa, sizeof(pa));
    memcpy(both + kMaxFreqs, pb, sizeof(pb));
    B200_CU(cudaMemcpy(h->d_pass, both, sizeof(both), cudaMemcpyHostToDevice));
    h->dirty = false;
    return B200AIS_OK;
}

extern "C" int b200ais_xlat_create(b200ais_xlat **out, int decimation, const float *taps, int ntaps,
                                   const double *center_freqs, int nfreqs, double sampling_freq,
                                   int sources)
{
    if (!out || !taps || !center_freqs || decimation < 1 || ntaps < 1 || nfreqs < 1 ||
        nfreqs > kMaxFreqs || sources < 1 || !(sampling_freq > 0)) {
        set_error("xlat_create: bad arguments (1 <= nfreqs <= %d)", kMaxFreqs);
        return B200AIS_E_INVALID;
    }
    if (xlat_smem_bytes(32, 8, decimation, ntaps, nullptr, 1) > 227 * 1024) {
        set_error("xlat_create: decimation %d x %d taps does not fit shared memory", decimation, ntaps);
        return B200AIS_E_INVALID;
    }
    b200ais_xlat *h = new (std::nothrow) b200ais_xlat;
    if (!h)
        return B200AIS_E_NOMEM;
    h->D = decimation;
    h->ntaps = ntaps;
    h->nfreqs = nfreqs;
    h->sources = sources;
    h->fs = sampling_freq;
    h->proto.assign(taps, taps + ntaps);
    h->freqs.assign(center_freqs, center_freqs + nfreqs);
    h->ctaps.resize(nfreqs);
    std::vector<RotState> rs(nfreqs);
    for (int k = 0; k < nfreqs; k++) {
        xlat_compose(h, k, &rs[k]);
        rs[k].pr = 1.0f;
        rs[k].pi = 0.0f;
        rs[k].counter = 0;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_pass, sizeof(int) * 2 * kMaxFreqs);
    if (e == cudaSuccess)
        e = cudaMalloc(&h->d_rot, sizeof(RotState) * nfreqs);
    if (e == cudaSuccess)
        e = cudaMemcpy(h->d_rot, rs.data(), sizeof(RotState) * nfreqs, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        b200ais_xlat_destroy(h);
        return cuda_fail(e, "xlat_create", __FILE__, __LINE__);
    }
    int rc = xlat_upload(h);
    if (rc) {
        b200ais_xlat_destroy(h);
        return rc;
    }
    *out = h;
    return B200AIS_OK;
}

extern "C" int b200ais_xlat_destroy(b200ais_xlat *h)
{
    if (!h)
        return B200AIS_OK;
    if (h->stream)
        cudaStreamDestroy(h->stream);
    for (int v = 0; v < 2; v++)
        if (h->d_taps[v])
            cudaFree(h->d_taps[v]);
    if (h->d_pass)
        cudaFree(h->d_pass);
    if (h->d_rot)
        cudaFree(h->d_rot);
    h->rottab.release();
    h->in.release();
    h->out.release();
    delete h;
    return B200AIS_OK;
}

extern "C" int b200ais_xlat_history(const b200ais_xlat *h) { return h ? h->ntaps : 0; }
extern "C" int b200ais_xlat_decimation(const b200ais_xlat *h) { return h ? h->D : 0; }

// set_center_freq [G]: the composite taps and the rotator increment are rebuilt before the
// next work(); the rotator keeps its phase and counter.
extern "C" int b200ais_xlat_set_center_freq(b200ais_xlat *h, int k, double center_freq)
{
    if (!h || k < 0 || k >= h->nfreqs) {
        set_error("xlat_set_center_freq: bad filter index");
        return B200AIS_E_INVALID;
    }
    B200_CU(cudaDeviceSynchronize());
    h->freqs[k] = center_freq;
    RotState rs;
    B200_CU(cudaMemcpy(&rs, h->d_rot + k, sizeof(rs), cudaMemcpyDeviceToHost));
    xlat_compose(h, k, &rs);
    B200_CU(cudaMemcpy(h->d_rot + k, &rs, sizeof(rs), cudaMemcpyHostToDevice));
    return xlat_upload(h);
}

extern "C" int b200ais_xlat_set_taps(b200ais_xlat *h, const float *taps, int ntaps)
{
    if (!h || !taps || ntaps < 1 || xlat_smem_bytes(32, 8, h->D, ntaps, nullptr, 1) > 227 * 1024) {
        set_error("xlat_set_taps: bad arguments");
        return B200AIS_E_INVALID;
    }
    B200_CU(cudaDeviceSynchronize());
    h->ntaps = ntaps;
    h->proto.assign(taps, taps + ntaps);
    std::vector<RotState> rs(h->nfreqs);
    B200_CU(cudaMemcpy(rs.data(), h->d_rot, sizeof(RotState) * h->nfreqs, cudaMemcpyDeviceToHost));
    for (int k = 0; k < h->nfreqs; k++)
        xlat_compose(h, k, &rs[k]);
    B200_CU(cudaMemcpy(h->d_rot, rs.data(), sizeof(RotState) * h->nfreqs, cudaMemcpyHostToDevice));
    return xlat_upload(h);
}

extern "C" int b200ais_xlat_reset(b200ais_xlat *h)
{
    if (!h)
        return B200AIS_E_INVALID;
    B200_CU(cudaDeviceSynchronize());
    std::vector<RotState> rs(h->nfreqs);
    B200_CU(cudaMemcpy(rs.data(), h->d_rot, sizeof(RotState) * h->nfreqs, cudaMemcpyDeviceToHost));
    for (auto &r : rs) {
        r.pr = 1.0f;
        r.pi = 0.0f;
        r.counter = 0;
    }
    B200_CU(cudaMemcpy(h->d_rot, rs.data(), sizeof(RotState) * h->nfreqs, cudaMemcpyHostToDevice));
    return B200AIS_OK;
}

template <int THREADS, int R>
static int xlat_launch(b200ais_xlat *h, int n, const float2 *in, size_t in_stride, float2 *out,
                       size_t out_stride, int G, cudaStream_t s)
{
    int Mp = 0;
    const size_t smem = xlat_smem_bytes(THREADS, R, h->D, h->ntaps, &Mp, G);
    B200_CU(cudaFuncSetAttribute(k_xlat_fir<THREADS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
    const int J = THREADS * R;
    dim3 grid((unsigned)((n + J - 1) / J), (unsigned)h->sources, (unsigned)h->npass);
    k_xlat_fir<THREADS, R><<<grid, THREADS, smem, s>>>(in, in_stride, n, h->D, h->ntaps,
                                                       h->ntp[R == 16], h->d_taps[R == 16],
                                                       h->d_pass, h->d_pass + kMaxFreqs, h->nfreqs,
                                                       h->rottab.as<float2>(), (size_t)n, out,
                                                       out_stride, Mp, G);
    B200_LAUNCH_CHECK("k_xlat_fir");
    return B200AIS_OK;
}

// tile shapes, best first: {threads, outputs per thread}
static const int kShapes[][2] = {{64, 16}, {128, 16}, {128, 8}, {64, 8}, {32, 8}};
static int g_xlat_shape = -1; // B200AIS_XLAT_SHAPE=<index> pins one (profiling)

extern "C" int b200ais_xlat_work_dev(b200ais_xlat *h, int noutput_items, const float *in,
                                     size_t in_stride, float *out, size_t out_stride, void *stream)
{
    if (!h || !in || !out || noutput_items < 0 || (size_t)noutput_items > out_stride ||
        (noutput_items > 0 && (size_t)noutput_items * h->D + h->ntaps - 1 > in_stride)) {
        set_error("xlat_work: bad arguments (rows need ntaps-1 + noutput*decimation items)");
        return B200AIS_E_INVALID;
    }
    if (h->sources > 65535) {
        set_error("xlat_work: at most 65535 sources per call");
        return B200AIS_E_INVALID;
    }
    if (noutput_items == 0)
        return B200AIS_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = noutput_items;
    int rc = h->rottab.reserve(sizeof(float2) * (size_t)n * h->nfreqs);
    if (rc)
        return rc;
    k_rot_phase<<<1, 32, 0, s>>>(h->d_rot, h->nfreqs, n, h->rottab.as<float2>(), (size_t)n);
    B200_LAUNCH_CHECK("k_rot_phase");
    const float2 *x = reinterpret_cast<const float2 *>(in);
    float2 *y = reinterpret_cast<float2 *>(out);
    if (g_xlat_shape < 0) {
        const char *e = getenv("B200AIS_XLAT_SHAPE");
        g_xlat_shape = e ? atoi(e) : 99;
    }
    // a 16-outputs-per-thread tile with two CTAs per SM: whole (all D residue rows resident) if
    // it fits, else with as many rows at a time as fit; the small tiles are the last resort
    const size_t half = 113 * 1024;
    int pick = -1, G = h->D;
    if (g_xlat_shape < 5 && xlat_smem_bytes(kShapes[g_xlat_shape][0], kShapes[g_xlat_shape][1], h->D,
                                            h->ntaps, nullptr) <= 227 * 1024)
        pick = g_xlat_shape;
    for (int i = 0; i < 2 && pick < 0; i++)
        if (xlat_smem_bytes(kShapes[i][0], kShapes[i][1], h->D, h->ntaps, nullptr) <= half)
            pick = i;
    for (int i = 0; i < 2 && pick < 0; i++) {
        int Mp = 0;
        const size_t one = xlat_smem_bytes(kShapes[i][0], kShapes[i][1], h->D, h->ntaps, &Mp, 1);
        const size_t rowb = sizeof(float2) * (size_t)Mp;
        if (one + rowb <= half) { // at least two rows at a time
            pick = i;
            G = 1 + (int)((half - one) / rowb);
        }
    }
    for (int i = 2; i < 5 && pick < 0; i++)
        if (xlat_smem_bytes(kShapes[i][0], kShapes[i][1], h->D, h->ntaps, nullptr) <= half)
            pick = i;
    for (int i = 0; i < 5 && pick < 0; i++)
        if (xlat_smem_bytes(kShapes[i][0], kShapes[i][1], h->D, h->ntaps, nullptr) <= 227 * 1024)
            pick = i;
    if (pick < 0) {
        set_error("xlat_work: %d taps at decimation %d do not fit shared memory", h->ntaps, h->D);
        return B200AIS_E_INVALID;
    }
    switch (pick) {
    case 0:
        return xlat_launch<64, 16>(h, n, x, in_stride, y, out_stride, G, s);
    case 1:
        return xlat_launch<128, 16>(h, n, x, in_stride, y, out_stride, G, s);
    case 2:
        return xlat_launch<128, 8>(h, n, x, in_stride, y, out_stride, G, s);
    case 3:
        return xlat_launch<64, 8>(h, n, x, in_stride, y, out_stride, G, s);
    default:
        return xlat_launch<32, 8>(h, n, x, in_stride, y, out_stride, G, s);
    }
}

extern "C" int b200ais_xlat_work(b200ais_xlat *h, int noutput_items, const float *in,
                                 size_t in_stride, float *out, size_t out_stride)
{
    if (!h || !in || !out || noutput_items < 0) {
        set_error("xlat_work: bad arguments");
        return B200AIS_E_INVALID;
    }
    if (noutput_items == 0)
        return B200AIS_OK;
    const size_t nin = (size_t)noutput_items * h->D + h->ntaps - 1;
    if (nin > in_stride || (size_t)noutput_items > out_stride) {
        set_error("xlat_work: rows need ntaps-1 + noutput*decimation input items");
        return B200AIS_E_INVALID;
    }
    const size_t dis = (nin + 1) & ~(size_t)1, dos = ((size_t)noutput_items + 1) & ~(size_t)1;
    const size_t rows_out = (size_t)h->sources * h->nfreqs;
    int rc;
    if ((rc = h->in.reserve(sizeof(float2) * dis * h->sources)) ||
        (rc = h->out.reserve(sizeof(float2) * dos * rows_out)))
        return rc;
    cudaStream_t s = h->stream;
    B200_CU(cudaMemcpy2DAsync(h->in.p, dis * sizeof(float2), in, in_stride * sizeof(float2),
                              nin * sizeof(float2), (size_t)h->sources, cudaMemcpyHostToDevice, s));
    rc = b200ais_xlat_work_dev(h, noutput_items, h->in.as<float>(), dis, h->out.as<float>(), dos, s);
    if (rc)
        return rc;
    B200_CU(cudaMemcpy2DAsync(out, out_stride * sizeof(float2), h->out.p, dos * sizeof(float2),
                              (size_t)noutput_items * sizeof(float2), rows_out,
                              cudaMemcpyDeviceToHost, s));
    B200_CU(cudaStreamSynchronize(s));
    return B200AIS_OK;
}
