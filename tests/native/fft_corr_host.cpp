// Host emulation of the FFT correlation of mdproptools_b200/csrc/fftcorr.cu: the same butterflies, cross spectrum and
// normalisation (csrc/fft_corr.h), stage after stage.  TEST INFRASTRUCTURE ONLY.  Built with -ffp-contract=off.
#include <stdint.h>
#include <vector>

#include "../../mdproptools_b200/csrc/fft_corr.h"

extern "C" int emulate_fft_xcorr(const double *a, const double *b, long long T, long long nlags, double *out)
{
    const int p = mdp_fft_log2_size(T, nlags);
    const long long n = 1ll << p;
    std::vector<mdp_c64> W((size_t)n / 2), x((size_t)n), y((size_t)n);
    for (long long k = 0; k < n / 2; ++k) W[(size_t)k] = mdp_twiddle(k, n);
    for (long long i = 0; i < n; ++i) {
        x[(size_t)i].re = i < T ? a[i] : 0.0;
        x[(size_t)i].im = i < T ? b[i] : 0.0;
    }
    mdp_c64 *src = x.data(), *dst = y.data();
    for (int t = 0; t < p; ++t) {
        for (long long i = 0; i < n / 2; ++i) mdp_fft_butterfly(src, dst, W.data(), i, t, n);
        mdp_c64 *tmp = src; src = dst; dst = tmp;
    }
    for (long long k = 0; k < n; ++k) dst[k] = mdp_cross_spectrum_conj(src, k, n);
    { mdp_c64 *tmp = src; src = dst; dst = tmp; }
    for (int t = 0; t < p; ++t) {
        for (long long i = 0; i < n / 2; ++i) mdp_fft_butterfly(src, dst, W.data(), i, t, n);
        mdp_c64 *tmp = src; src = dst; dst = tmp;
    }
    // src = FFT(conj(P)) = N * conj(corr); corr is real
    for (long long tau = 0; tau < nlags; ++tau) out[tau] = src[tau].re / (double)n / (double)(T - tau);
    return p;
}
