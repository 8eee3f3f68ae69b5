// survival.cu -- residence-time survival counts from the RUNS of the pair bitmasks (the default since its
// first hardware run in round 2; MDP_SURVIVAL_RUNS=0 selects the AND-shift-popcount kernel of corr.cu).
//
// Replaces the same loop as mdp_bitmask_autocorr (residence_time.py:112-143).  Method and proof of equality in
// survival_runs.h: a pair whose indicator consists of k runs costs 4 * k(k+1)/2 integer updates of a second-difference
// array instead of T^2/128 word operations -- for the C5 shape (5 000 frames, neighbours that stay for hundreds of
// frames) two to three orders of magnitude less work, and the result is the same integers.
//
// One warp per pair: lane 0 walks the pair's words and writes the runs to the warp's shared-memory buffer, the lanes
// share the run pairs (i, j >= i) and add the four updates to the CTA's second-difference array in shared memory
// (64-bit shared atomics); a pair with more runs than the buffer holds takes the word route straight into cnt.  CTAs are
// persistent; each flushes its array to the global one at the end, and a last one-block kernel does the two prefix sums.
#include <algorithm>

#include "common.cuh"
#include "survival_runs.h"

namespace {

constexpr int SR_WARPS = 8;
constexpr int SR_CAP = 512;          // runs per pair held in shared memory (2 x 4 B each per warp: 32 KB per CTA)

struct SharedAdd {
    unsigned long long *d2;
    __device__ __forceinline__ void operator()(long long i, long long v) const { atomicAdd(&d2[i], (unsigned long long)v); }
};

__global__ void __launch_bounds__(SR_WARPS * 32) k_survival_runs(const unsigned long long *__restrict__ masks, long long npairs,
                                                                  int W, long long T, unsigned long long *__restrict__ d2g,
                                                                  unsigned long long *__restrict__ v0g,
                                                                  unsigned long long *__restrict__ cnt)
{
    extern __shared__ __align__(16) unsigned char sr_smem[];
    unsigned long long *d2 = reinterpret_cast<unsigned long long *>(sr_smem);                 // [T]
    int *runs = reinterpret_cast<int *>(sr_smem + (size_t)T * 8);                              // [SR_WARPS][2][SR_CAP]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int *st = runs + (size_t)w * 2 * SR_CAP, *en = st + SR_CAP;
    for (long long k = tid; k < T; k += blockDim.x) d2[k] = 0ull;
    __syncthreads();
    long long v0 = 0;
    const long long nwarps = (long long)gridDim.x * SR_WARPS;
    for (long long p = (long long)blockIdx.x * SR_WARPS + w; p < npairs; p += nwarps) {
        const unsigned long long *m = masks + p * W;
        int k = 0;
        if (lane == 0) k = mdp_runs_from_mask(m, W, T, st, en, SR_CAP);
        k = __shfl_sync(0xffffffffu, k, 0);
        __syncwarp();
        if (k > SR_CAP) {
            for (long long tau = lane; tau < T; tau += 32) {
                const unsigned long long c = mdp_mask_corr_direct(m, W, tau);
                if (c) atomicAdd(&cnt[tau], c);
            }
        } else {
            for (int i = 0; i < k; ++i) {
                const int ai = st[i], bi = en[i];
                for (int j = i + lane; j < k; j += 32) v0 += mdp_run_pair_updates(ai, bi, st[j], en[j], T, SharedAdd{d2});
            }
        }
        __syncwarp();   // the buffer is rewritten for the next pair
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v0 += __shfl_xor_sync(0xffffffffu, v0, o);
    if (lane == 0 && v0) atomicAdd(v0g, (unsigned long long)v0);
    __syncthreads();
    for (long long k = tid; k < T; k += blockDim.x) {
        const unsigned long long v = d2[k];
        if (v) atomicAdd(&d2g[k], v);
    }
}

// cnt[t] += V0 + sum_{u < t} slope(u), slope(u) = sum_{s <= u} D2[s]; one block, thread-contiguous segments, two passes
__global__ void __launch_bounds__(1024) k_survival_finish(const unsigned long long *__restrict__ d2g,
                                                          const unsigned long long *__restrict__ v0g, long long T,
                                                          unsigned long long *__restrict__ cnt)
{
    __shared__ long long seg_d[1024], seg_s[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const long long per = (T + nt - 1) / nt;
    const long long b = (long long)t * per, e = b + per < T ? b + per : T;
    // pass 1: sum of D2 over my segment -> slope at the segment's start
    long long sd = 0;
    for (long long k = b; k < e; ++k) sd += (long long)d2g[k];
    seg_d[t] = sd;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int k = 0; k < nt; ++k) {
            const long long v = seg_d[k];
            seg_d[k] = run;          // slope just before the segment
            run += v;
        }
    }
    __syncthreads();
    // pass 2: sum of slope(u) over my segment -> value at the segment's start
    long long slope = seg_d[t], ss = 0;
    for (long long k = b; k < e; ++k) {
        slope += (long long)d2g[k];
        ss += slope;
    }
    seg_s[t] = ss;
    __syncthreads();
    if (t == 0) {
        long long run = (long long)v0g[0];
        for (int k = 0; k < nt; ++k) {
            const long long v = seg_s[k];
            seg_s[k] = run;          // cnt at the segment's first index
            run += v;
        }
    }
    __syncthreads();
    slope = seg_d[t];
    long long val = seg_s[t];
    for (long long k = b; k < e; ++k) {
        cnt[k] += (unsigned long long)val;
        slope += (long long)d2g[k];
        val += slope;
    }
}

} // namespace

extern "C" {

int mdp_survival_runs(mdp_ctx *ctx, int64_t npairs, int nwords, int64_t T, const uint64_t *masks, uint64_t *cnt_out, void *stream)
{
    MDP_REQUIRE(ctx && masks && cnt_out, "mdp_survival_runs: NULL argument");
    MDP_REQUIRE(npairs > 0 && nwords > 0 && T > 0 && T <= (int64_t)nwords * 64, "mdp_survival_runs: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const size_t smem = (size_t)T * 8 + (size_t)SR_WARPS * 2 * SR_CAP * sizeof(int);
    MDP_REQUIRE(smem <= ctx->smem_optin, "mdp_survival_runs: %lld frames exceed the shared-memory second-difference array; "
                                         "use mdp_bitmask_autocorr", (long long)T);
    int rc = ctx->arena_reserve(align256((size_t)(T + 1) * 8) + 4096);
    if (rc) return rc;
    ctx->arena_reset();
    unsigned long long *d2g = (unsigned long long *)ctx->arena_take((size_t)(T + 1) * 8);   // [T] second differences, [T] = V0
    if (!d2g) {
        mdp_set_error("internal: scratch arena exhausted (survival runs)");
        return MDP_ERR_OOM;
    }
    MDP_CUDA(cudaMemsetAsync(d2g, 0, (size_t)(T + 1) * 8, st));
    MDP_CUDA(cudaFuncSetAttribute((const void *)k_survival_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = std::max<int>(1, std::min<int>(4, (int)((size_t)220 * 1024 / std::max<size_t>(smem, 1))));
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(npairs, SR_WARPS), (int64_t)ctx->sm_count * per_sm);
    cudaEvent_t tk = ctx->timer_begin(6, st);
    k_survival_runs<<<grid, SR_WARPS * 32, smem, st>>>((const unsigned long long *)masks, npairs, nwords, T, d2g, d2g + T,
                                                        (unsigned long long *)cnt_out);
    MDP_LAUNCHED(ctx);
    k_survival_finish<<<1, 1024, 0, st>>>(d2g, d2g + T, T, (unsigned long long *)cnt_out);
    ctx->timer_end(tk, st);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_survival_runs");
}

} // extern "C"
