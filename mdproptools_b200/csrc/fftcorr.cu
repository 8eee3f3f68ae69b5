// fftcorr.cu -- unbiased time correlation through a radix-2 Stockham FFT (the default for series of >= 2048 steps,
// ops.XCORR_FFT_MIN_T; shorter ones and MDP_XCORR_FFT=0 take the direct fp64 sum k_xcorr of corr.cu).
//
// Same result as mdp_xcorr_unbiased (Conductivity.correlate conductivity.py:97-114, Viscosity.autocorrelate
// viscosity.py:86-120) to the round-off of an FFT -- which is how the reference itself computes it -- at N log N instead of
// T^2/2: method in fft_corr.h.  Every stage is one launch over all channels (N/2 independent butterflies per channel,
// coalesced 16-byte complex loads and stores); HBM-bound: 2 p stages x 32 B per point.  Channels are processed in
// groups that keep the two work arrays within the scratch arena.
#include <algorithm>

#include "common.cuh"
#include "fft_corr.h"

namespace {

constexpr int FC_THREADS = 256;

__global__ void __launch_bounds__(FC_THREADS) k_fft_twiddle(mdp_c64 *__restrict__ W, long long n)
{
    const long long k = (long long)blockIdx.x * FC_THREADS + threadIdx.x;
    if (k < n / 2) W[k] = mdp_twiddle(k, n);
}

// z[c][i] = a[c][i] + i * b[c][i] for i < T, 0 beyond; grid (ceil(n / FC_THREADS), channels of the group)
__global__ void __launch_bounds__(FC_THREADS) k_fft_load(const double *__restrict__ a, const double *__restrict__ b, long long T,
                                                         long long n, mdp_c64 *__restrict__ z)
{
    const long long i = (long long)blockIdx.x * FC_THREADS + threadIdx.x;
    if (i >= n) return;
    const long long c = blockIdx.y;
    mdp_c64 v;
    v.re = i < T ? a[c * T + i] : 0.0;
    v.im = i < T ? b[c * T + i] : 0.0;
    z[c * n + i] = v;
}

__global__ void __launch_bounds__(FC_THREADS) k_fft_stage(const mdp_c64 *__restrict__ x, mdp_c64 *__restrict__ y,
                                                          const mdp_c64 *__restrict__ W, int t, long long n)
{
    const long long i = (long long)blockIdx.x * FC_THREADS + threadIdx.x;
    if (i >= n / 2) return;
    const long long c = blockIdx.y;
    mdp_fft_butterfly(x + c * n, y + c * n, W, i, t, n);
}

__global__ void __launch_bounds__(FC_THREADS) k_fft_cross(const mdp_c64 *__restrict__ Z, mdp_c64 *__restrict__ P, long long n)
{
    const long long k = (long long)blockIdx.x * FC_THREADS + threadIdx.x;
    if (k >= n) return;
    const long long c = blockIdx.y;
    P[c * n + k] = mdp_cross_spectrum_conj(Z + c * n, k, n);
}

// X = FFT(conj(P)) = N * conj(corr): out[c][tau] = Re X[tau] / N / (T - tau)
__global__ void __launch_bounds__(FC_THREADS) k_fft_store(const mdp_c64 *__restrict__ X, long long T, long long nlags, long long n,
                                                          double *__restrict__ out)
{
    const long long tau = (long long)blockIdx.x * FC_THREADS + threadIdx.x;
    if (tau >= nlags) return;
    const long long c = blockIdx.y;
    out[c * nlags + tau] = X[c * n + tau].re / (double)n / (double)(T - tau);
}

} // namespace

extern "C" {

int mdp_xcorr_fft(mdp_ctx *ctx, int nchan, int64_t T, const double *a, const double *b, int64_t nlags, double *out, void *stream)
{
    MDP_REQUIRE(ctx && a && b && out, "mdp_xcorr_fft: NULL argument");
    MDP_REQUIRE(nchan > 0 && T > 0 && nlags > 0 && nlags <= T, "mdp_xcorr_fft: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const int p = mdp_fft_log2_size(T, nlags);
    MDP_REQUIRE(p <= 30, "mdp_xcorr_fft: %lld steps need a transform of 2^%d points", (long long)T, p);
    const long long n = 1ll << p;
    const size_t per_chan = (size_t)n * sizeof(mdp_c64);
    const size_t wbytes = align256((size_t)(n / 2) * sizeof(mdp_c64));
    // two work arrays per channel of the group; groups sized to 2 GiB of scratch (at least one channel)
    const size_t budget = std::min<size_t>(ctx->slab_limit, (size_t)2 << 30);
    const int group = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::min(nchan, 65535), (budget - std::min(budget, wbytes)) / (2 * per_chan)));
    int rc = ctx->arena_reserve(wbytes + 2 * align256((size_t)group * per_chan) + 4096);
    if (rc) return rc;
    ctx->arena_reset();
    mdp_c64 *W = (mdp_c64 *)ctx->arena_take((size_t)(n / 2) * sizeof(mdp_c64));
    mdp_c64 *x = (mdp_c64 *)ctx->arena_take((size_t)group * per_chan);
    mdp_c64 *y = (mdp_c64 *)ctx->arena_take((size_t)group * per_chan);
    if (!W || !x || !y) {
        mdp_set_error("internal: scratch arena exhausted (fft correlation)");
        return MDP_ERR_OOM;
    }
    cudaEvent_t tk = ctx->timer_begin(3, st);
    k_fft_twiddle<<<(unsigned)ceil_div<long long>(std::max<long long>(n / 2, 1), FC_THREADS), FC_THREADS, 0, st>>>(W, n);
    MDP_LAUNCHED(ctx);
    for (int c0 = 0; c0 < nchan; c0 += group) {
        const int g = std::min(group, nchan - c0);
        const dim3 gn((unsigned)ceil_div<long long>(n, FC_THREADS), (unsigned)g);
        const dim3 gh((unsigned)ceil_div<long long>(std::max<long long>(n / 2, 1), FC_THREADS), (unsigned)g);
        k_fft_load<<<gn, FC_THREADS, 0, st>>>(a + (size_t)c0 * T, b + (size_t)c0 * T, T, n, x);
        MDP_LAUNCHED(ctx);
        mdp_c64 *src = x, *dst = y;
        for (int t = 0; t < p; ++t) {
            k_fft_stage<<<gh, FC_THREADS, 0, st>>>(src, dst, W, t, n);
            MDP_LAUNCHED(ctx);
            std::swap(src, dst);
        }
        k_fft_cross<<<gn, FC_THREADS, 0, st>>>(src, dst, n);
        MDP_LAUNCHED(ctx);
        std::swap(src, dst);
        for (int t = 0; t < p; ++t) {
            k_fft_stage<<<gh, FC_THREADS, 0, st>>>(src, dst, W, t, n);
            MDP_LAUNCHED(ctx);
            std::swap(src, dst);
        }
        const dim3 gl((unsigned)ceil_div<long long>(nlags, FC_THREADS), (unsigned)g);
        k_fft_store<<<gl, FC_THREADS, 0, st>>>(src, T, nlags, n, out + (size_t)c0 * nlags);
        MDP_LAUNCHED(ctx);
    }
    ctx->timer_end(tk, st);
    return mdp_check_launch("k_fft_stage");
}

} // extern "C"
