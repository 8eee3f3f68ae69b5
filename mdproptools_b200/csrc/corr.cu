// corr.cu -- time-correlation kernels of the Green-Kubo and residence-time paths.
//
// Replaces Conductivity.correlate (dynamical/conductivity.py:97-114), Viscosity.autocorrelate
// (dynamical/viscosity.py:86-120), the cumulative trapezoid of conductivity.py:231 / viscosity.py:151 and
// the per-column acovf loop of ResidenceTime.calc_auto_correlation (dynamical/residence_time.py:112-143).
//
// The reference evaluates the unbiased correlation with a zero-padded FFT; here it is the direct sum
//     out[tau] = (sum_{t < T - tau} a[t + tau] * b[t]) / (T - tau)
// in fp64 FMA with a register-resident sliding window: a thread owns 8 consecutive lags and sweeps a
// contiguous range of t, so each step costs 2 shared-memory loads for 8 FMAs (FP64-pipe bound).
#include "common.cuh"

namespace {

constexpr int XC_LAGS = 128;                 // lags per CTA
constexpr int XC_LPT = 8;                    // lags per thread (register window)
constexpr int XC_LG = XC_LAGS / XC_LPT;      // 16 lag groups
constexpr int XC_TS = 16;                    // t sub-ranges per CTA
constexpr int XC_NT = 2048;                  // t values staged per tile
constexpr int XC_SUB = XC_NT / XC_TS;        // 128 steps per thread per tile
constexpr int XC_THREADS = XC_LG * XC_TS;    // 256

__device__ __forceinline__ int pad8(int i) { return i + (i >> 3); }   // skew: stride-8 doubles hit distinct banks

// grid (nlagblocks, nchan)
__global__ void __launch_bounds__(XC_THREADS) k_xcorr(const double *__restrict__ a, const double *__restrict__ b, int64_t T,
                                                      int64_t nlags, double *__restrict__ out)
{
    constexpr int SA_SIZE = XC_NT + XC_LAGS + (XC_NT + XC_LAGS) / 8 + 8;
    __shared__ double smem[XC_NT + SA_SIZE];   // 36 KB; the final cross-thread reduction reuses it
    double *sb = smem, *sa = smem + XC_NT;
    static_assert(XC_TS * (XC_LAGS + 1) <= XC_NT + SA_SIZE, "reduction scratch must fit");
    double(*red)[XC_LAGS + 1] = reinterpret_cast<double(*)[XC_LAGS + 1]>(smem);
    const int c = blockIdx.y;
    const int64_t L0 = (int64_t)blockIdx.x * XC_LAGS;
    const double *ac = a + (int64_t)c * T, *bc = b + (int64_t)c * T;
    const int tid = threadIdx.x;
    const int lg = tid % XC_LG, ts = tid / XC_LG;
    const int l0 = lg * XC_LPT;   // first lag of this thread relative to L0
    double acc[XC_LPT];
#pragma unroll
    for (int k = 0; k < XC_LPT; ++k) acc[k] = 0.0;

    // only t < T - L0 contributes to any lag of this block
    const int64_t tmax = T - L0;
    for (int64_t t0 = 0; t0 < tmax; t0 += XC_NT) {
        __syncthreads();
        for (int i = tid; i < XC_NT; i += XC_THREADS) {
            const int64_t t = t0 + i;
            sb[i] = t < T ? bc[t] : 0.0;
        }
        for (int i = tid; i < XC_NT + XC_LAGS; i += XC_THREADS) {
            const int64_t t = t0 + L0 + i;
            sa[pad8(i)] = t < T ? ac[t] : 0.0;   // zero padding == the reference's zero-padded FFT
        }
        __syncthreads();
        // s0, l0 and the step s are multiples of 8, so the skewed index is affine inside an octet:
        // pad8(base + j) = pad8(base) + j + (j >> 3) for base % 8 == 0 -- the loads below take immediate offsets
        const double *pa = sa + pad8(ts * XC_SUB + l0);     // window elements of step s: pa[k], k = 0..7; next octet at pa[9..]
        const double2 *pb = reinterpret_cast<const double2 *>(sb + ts * XC_SUB);
        double win[XC_LPT];
#pragma unroll
        for (int k = 0; k < XC_LPT - 1; ++k) win[k] = pa[k];
#pragma unroll 1
        for (int s = 0; s < XC_SUB; s += XC_LPT, pa += XC_LPT + 1, pb += XC_LPT / 2) {
            double bv[XC_LPT];
#pragma unroll
            for (int u = 0; u < XC_LPT; u += 2) {
                const double2 t = pb[u >> 1];
                bv[u] = t.x;
                bv[u + 1] = t.y;
            }
#pragma unroll
            for (int u = 0; u < XC_LPT; ++u) {
                // window element for lag l0+k at step s+u is a[s0+s+u + l0+k]; rotate by renaming
                win[(u + XC_LPT - 1) % XC_LPT] = pa[u + XC_LPT - 1 + ((u + XC_LPT - 1) >> 3)];
#pragma unroll
                for (int k = 0; k < XC_LPT; ++k) acc[k] = fma(win[(u + k) % XC_LPT], bv[u], acc[k]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < XC_LPT; ++k) red[ts][l0 + k] = acc[k];
    __syncthreads();
    if (tid < XC_LAGS) {
        const int64_t lag = L0 + tid;
        if (lag < nlags) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < XC_TS; ++r) s += red[r][tid];
            out[(int64_t)c * nlags + lag] = s / (double)(T - lag);
        }
    }
}

// out[r][k] = scale * cumulative trapezoid; one CTA per row, chunked sequential scan
__global__ void __launch_bounds__(1024) k_cumtrapz(const double *__restrict__ in, int64_t T, double dx, double scale,
                                                   int leading_zero, double *__restrict__ out)
{
    __shared__ double ws[32];
    const double *y = in + (int64_t)blockIdx.x * T;
    const int64_t nout = leading_zero ? T : T - 1;
    double *o = out + (int64_t)blockIdx.x * nout;
    const int64_t m = T - 1;   // number of trapezoids
    const int t = threadIdx.x, nt = blockDim.x;
    const int64_t per = (m + nt - 1) / nt;
    const int64_t b = (int64_t)t * per, e = b + per < m ? b + per : m;
    double s = 0.0;
    for (int64_t k = b; k < e; ++k) s += dx * (y[k + 1] + y[k]) / 2.0;
    const int lane = t & 31, w = t >> 5;
    double inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        const double v = lane < (nt >> 5) ? ws[lane] : 0.0;
        double iv = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, iv, d);
            if (lane >= d) iv += u;
        }
        ws[lane] = iv - v;
    }
    __syncthreads();
    double run = ws[w] + (inc - s);
    if (leading_zero && t == 0) o[0] = 0.0;
    for (int64_t k = b; k < e; ++k) {
        run += dx * (y[k + 1] + y[k]) / 2.0;
        o[k + (leading_zero ? 1 : 0)] = scale * run;
    }
}

// set bit `frame` of the mask row of pair (ia, ib); rows are found by binary search in the sorted keys
__global__ void __launch_bounds__(256) k_bitmask_fill(const int32_t *__restrict__ list, int64_t nentries, int64_t n_b,
                                                      const long long *__restrict__ keys, int64_t npairs, int nwords,
                                                      unsigned long long *__restrict__ masks)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nentries) return;
    const int frame = list[e * 3 + 0];
    const long long key = (long long)list[e * 3 + 1] * n_b + list[e * 3 + 2];
    int64_t lo = 0, hi = npairs - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (keys[lo] == key && (frame >> 6) < nwords) atomicOr(&masks[lo * nwords + (frame >> 6)], 1ull << (frame & 63));
}

// cnt[tau] += sum_p popc(m_p & (m_p >> tau)); grid (ceil(T/256), npair blocks); masks of a pair block in smem
constexpr int BM_PAIRS = 64;
__global__ void __launch_bounds__(256) k_bitmask_autocorr(const unsigned long long *__restrict__ masks, int64_t npairs,
                                                          int nwords, int64_t T, unsigned long long *__restrict__ cnt)
{
    extern __shared__ unsigned long long sm[];   // [BM_PAIRS][nwords + 1]
    const int64_t p0 = (int64_t)blockIdx.y * BM_PAIRS;
    const int np = (int)(npairs - p0 < BM_PAIRS ? npairs - p0 : BM_PAIRS);
    const int W = nwords + 1;
    for (int i = threadIdx.x; i < np * W; i += blockDim.x) {
        const int p = i / W, w = i % W;
        sm[i] = w < nwords ? masks[(p0 + p) * nwords + w] : 0ull;
    }
    __syncthreads();
    const int64_t tau = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= T) return;
    const int ws = (int)(tau >> 6), bs = (int)(tau & 63);
    unsigned long long total = 0;
    for (int p = 0; p < np; ++p) {
        const unsigned long long *m = sm + p * W;
        unsigned int c = 0;
        for (int w = 0; w + ws < nwords; ++w) {
            const unsigned long long lo = m[w + ws], hi = m[w + ws + 1];
            const unsigned long long sh = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
            c += __popcll(m[w] & sh);
        }
        total += c;
    }
    if (total) atomicAdd(&cnt[tau], total);
}

} // namespace

extern "C" {

int mdp_xcorr_unbiased(mdp_ctx *ctx, int nchan, int64_t T, const double *a, const double *b, int64_t nlags, double *out,
                       void *stream)
{
    MDP_REQUIRE(ctx && a && b && out, "mdp_xcorr_unbiased: NULL argument");
    MDP_REQUIRE(nchan > 0 && nchan <= 65535 && T > 0 && nlags > 0 && nlags <= T, "mdp_xcorr_unbiased: bad sizes");
    MDP_CUDA(cudaSetDevice(ctx->device));
    dim3 grid((unsigned)ceil_div<int64_t>(nlags, XC_LAGS), nchan);
    cudaEvent_t tk = ctx->timer_begin(3, (cudaStream_t)stream);
    k_xcorr<<<grid, XC_THREADS, 0, (cudaStream_t)stream>>>(a, b, T, nlags, out);
    ctx->timer_end(tk, (cudaStream_t)stream);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_xcorr");
}

int mdp_cumtrapz(mdp_ctx *ctx, int nrows, int64_t T, const double *in, double dx, double scale, int leading_zero,
                 double *out, void *stream)
{
    MDP_REQUIRE(ctx && in && out, "mdp_cumtrapz: NULL argument");
    MDP_REQUIRE(nrows > 0 && T > 1, "mdp_cumtrapz: bad sizes");
    MDP_CUDA(cudaSetDevice(ctx->device));
    k_cumtrapz<<<nrows, 1024, 0, (cudaStream_t)stream>>>(in, T, dx, scale, leading_zero, out);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_cumtrapz");
}

int mdp_bitmask_fill(mdp_ctx *ctx, int64_t nentries, const int32_t *list, int64_t n_b, const int64_t *pair_keys,
                     int64_t npairs, int nwords, uint64_t *masks, void *stream)
{
    MDP_REQUIRE(ctx && list && pair_keys && masks, "mdp_bitmask_fill: NULL argument");
    MDP_REQUIRE(nentries >= 0 && npairs > 0 && nwords > 0 && n_b > 0, "mdp_bitmask_fill: bad sizes");
    if (nentries == 0) return 0;
    MDP_CUDA(cudaSetDevice(ctx->device));
    k_bitmask_fill<<<(unsigned)ceil_div<int64_t>(nentries, 256), 256, 0, (cudaStream_t)stream>>>(
        list, nentries, n_b, (const long long *)pair_keys, npairs, nwords, (unsigned long long *)masks);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_bitmask_fill");
}

int mdp_bitmask_autocorr(mdp_ctx *ctx, int64_t npairs, int nwords, int64_t T, const uint64_t *masks, uint64_t *cnt_out,
                         void *stream)
{
    MDP_REQUIRE(ctx && masks && cnt_out, "mdp_bitmask_autocorr: NULL argument");
    MDP_REQUIRE(npairs > 0 && nwords > 0 && T > 0 && T <= (int64_t)nwords * 64, "mdp_bitmask_autocorr: bad sizes");
    MDP_CUDA(cudaSetDevice(ctx->device));
    const size_t smem = (size_t)BM_PAIRS * (nwords + 1) * 8;
    MDP_REQUIRE(smem <= ctx->smem_optin, "mdp_bitmask_autocorr: %d mask words per pair exceed shared memory", nwords);
    MDP_CUDA(cudaFuncSetAttribute((const void *)k_bitmask_autocorr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div<int64_t>(T, 256), (unsigned)ceil_div<int64_t>(npairs, BM_PAIRS));
    MDP_REQUIRE(grid.y <= 65535, "mdp_bitmask_autocorr: too many pairs per call (%lld); split the call", (long long)npairs);
    cudaEvent_t tk = ctx->timer_begin(6, (cudaStream_t)stream);
    k_bitmask_autocorr<<<grid, 256, smem, (cudaStream_t)stream>>>((const unsigned long long *)masks, npairs, nwords, T,
                                                                  (unsigned long long *)cnt_out);
    ctx->timer_end(tk, (cudaStream_t)stream);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_bitmask_autocorr");
}

} // extern "C"
