// Host emulation of k_survival_runs (mdproptools_b200/csrc/corr.cu): the same run extraction and second-difference
// updates (csrc/survival_runs.h), pair after pair, followed by the two prefix sums.  TEST INFRASTRUCTURE ONLY.
#include <stdint.h>
#include <vector>

#include "../../mdproptools_b200/csrc/survival_runs.h"

struct AddTo {
    long long *d2;
    void operator()(long long i, long long v) const { d2[i] += v; }
};

// masks [P][W] uint64, cnt_out [T]; returns the largest number of runs of any pair (pairs beyond cap take the word route)
extern "C" int emulate_survival_runs(const unsigned long long *masks, long long P, int W, long long T, int cap, long long *cnt_out)
{
    std::vector<long long> d2((size_t)T + 1, 0), direct((size_t)T, 0);
    std::vector<int> st((size_t)cap), en((size_t)cap);
    long long v0 = 0;
    int kmax = 0;
    for (long long p = 0; p < P; ++p) {
        const int k = mdp_runs_from_mask(masks + p * W, W, T, st.data(), en.data(), cap);
        kmax = k > kmax ? k : kmax;
        if (k > cap) {             // more runs than the buffer holds: the word route, straight into cnt (added below)
            for (long long tau = 0; tau < T; ++tau) direct[(size_t)tau] += (long long)mdp_mask_corr_direct(masks + p * W, W, tau);
            continue;
        }
        for (int i = 0; i < k; ++i)
            for (int j = i; j < k; ++j) v0 += mdp_run_pair_updates(st[i], en[i], st[j], en[j], T, AddTo{d2.data()});
    }
    // the two prefix sums exactly as k_survival_finish does them: nt "threads" with contiguous segments, three passes
    const int nt = 1024;
    std::vector<long long> seg_d(nt), seg_s(nt);
    const long long per = (T + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const long long b = (long long)t * per, e = b + per < T ? b + per : T;
        long long sd = 0;
        for (long long k = b; k < e; ++k) sd += d2[k];
        seg_d[t] = sd;
    }
    long long run = 0;
    for (int k = 0; k < nt; ++k) {
        const long long v = seg_d[k];
        seg_d[k] = run;
        run += v;
    }
    for (int t = 0; t < nt; ++t) {
        const long long b = (long long)t * per, e = b + per < T ? b + per : T;
        long long slope = seg_d[t], ss = 0;
        for (long long k = b; k < e; ++k) {
            slope += d2[k];
            ss += slope;
        }
        seg_s[t] = ss;
    }
    run = v0;
    for (int k = 0; k < nt; ++k) {
        const long long v = seg_s[k];
        seg_s[k] = run;
        run += v;
    }
    for (long long t = 0; t < T; ++t) cnt_out[t] = direct[(size_t)t];   // the pairs of the word route
    for (int t = 0; t < nt; ++t) {
        const long long b = (long long)t * per, e = b + per < T ? b + per : T;
        long long slope = seg_d[t], val = seg_s[t];
        for (long long k = b; k < e; ++k) {
            cnt_out[k] += val;
            slope += d2[k];
            val += slope;
        }
    }
    return kmax;
}
