// survival_runs.h -- residence-time survival counts from RUNS instead of words (host/device shared body of
// k_survival_runs in corr.cu; exercised on the host by tests/native/survival_runs_host.cpp).
//
// cnt[tau] = sum_pairs sum_t h(t) h(t + tau)   (residence_time.py:112-143; h = the pair's neighbour indicator)
//
// A pair's indicator is a handful of runs [a_i, b_i) -- a neighbour stays for a while, leaves, perhaps returns --, and
// the correlation of two runs is a trapezoid in tau:
//   |[a_i, b_i) n [a_j - tau, b_j - tau)| = r(tau - s0) - r(tau - s1) - r(tau - s2) + r(tau - s3),   r(x) = max(0, x),
//   s0 = a_j - b_i,  s1 = s0 + min(L_i, L_j),  s2 = s0 + max(L_i, L_j),  s3 = b_j - a_i            (L = run length)
// so a run pair is FOUR integer updates of a second-difference array D2 (a break point s < 0 folds into D2[0] and the
// constant V0: r(tau - s) = tau + |s| for tau >= 0; break points >= T are dropped), and
//   cnt[tau] = V0 + sum_{u < tau} slope(u),  slope(u) = sum_{s <= u} D2[s]
// -- two prefix sums at the very end.  Only j >= i contributes for tau >= 0 (runs are disjoint and ordered).  All
// integers: the counts are the ones the AND-shift-popcount kernel produces, bit for bit, for k(k+1)/2 * 4 updates per
// pair instead of T^2 / 128 word operations (C5: ~10^2 against 2*10^5).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MDP_HD __host__ __device__ __forceinline__
#else
#ifndef MDP_HD
#define MDP_HD inline
#endif
#endif

// Runs of set bits of the T-bit mask m[0..W) (bit t of the mask = bit t & 63 of word t >> 6; bits >= T are ignored), in
// ascending order: starts[k] inclusive, ends[k] exclusive.  Returns the number of runs found, which may exceed cap (the
// arrays then hold the first cap runs and the caller must take another route).
MDP_HD int mdp_runs_from_mask(const unsigned long long *m, int W, long long T, int *starts, int *ends, int cap)
{
    int k = 0;
    unsigned long long prev_top = 0;   // bit 63 of the previous word (the bit just below this word's bit 0)
    for (int w = 0; w < W; ++w) {
        unsigned long long x = m[w];
        const long long base = (long long)w * 64;
        if (base >= T) break;
        if (T - base < 64) x &= (1ull << (T - base)) - 1ull;
        // transitions: bit b of tr is set where x's bit b differs from the bit below it
        unsigned long long tr = x ^ ((x << 1) | prev_top);
        while (tr) {
#if defined(__CUDA_ARCH__)
            const int b = __ffsll((long long)tr) - 1;
#else
            const int b = __builtin_ctzll(tr);
#endif
            tr &= tr - 1;
            if ((x >> b) & 1ull) {          // 0 -> 1: a run starts
                if (k < cap) starts[k] = (int)(base + b);
            } else {                         // 1 -> 0: the open run ends
                if (k < cap) ends[k] = (int)(base + b);
                ++k;
            }
        }
        prev_top = x >> 63;
    }
    if (prev_top) {
        // the last run is still open at the end of the trajectory (only possible when T is a multiple of 64: otherwise the
        // masked word has a 0 at bit T and the 1 -> 0 transition was seen above)
        if (k < cap) ends[k] = (int)T;
        ++k;
    }
    return k;
}

// The four second-difference updates of the run pair (i, j), j >= i.  add(index, value) receives 0 <= index < T;
// returns the contribution to V0.
template <class Add>
MDP_HD long long mdp_run_pair_updates(int ai, int bi, int aj, int bj, long long T, const Add add)
{
    const long long Li = bi - ai, Lj = bj - aj;
    const long long s0 = (long long)aj - bi;
    const long long s[4] = {s0, s0 + (Li < Lj ? Li : Lj), s0 + (Li < Lj ? Lj : Li), (long long)bj - ai};
    const long long c[4] = {1, -1, -1, 1};
    long long v0 = 0, at0 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (s[k] < 0) {
            v0 += c[k] * (-s[k]);
            at0 += c[k];
        } else if (s[k] == 0) {
            at0 += c[k];
        } else if (s[k] < T) {
            add(s[k], c[k]);
        }
    }
    if (at0) add(0, at0);
    return v0;
}

// sum_t h(t) h(t + tau) of one pair straight from the words (the arithmetic of k_bitmask_autocorr): the route for a pair
// with more runs than the run buffer holds.  Bits >= T of the mask are zero.
MDP_HD unsigned long long mdp_mask_corr_direct(const unsigned long long *m, int W, long long tau)
{
    const int ws = (int)(tau >> 6), bs = (int)(tau & 63);
    unsigned long long total = 0;
    for (int w = 0; w + ws < W; ++w) {
        const unsigned long long lo = m[w + ws], hi = w + ws + 1 < W ? m[w + ws + 1] : 0ull;
        const unsigned long long sh = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
#if defined(__CUDA_ARCH__)
        total += (unsigned long long)__popcll(m[w] & sh);
#else
        total += (unsigned long long)__builtin_popcountll(m[w] & sh);
#endif
    }
    return total;
}
