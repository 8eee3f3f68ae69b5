// shell.cu -- neighbour search of a small set A against a large set B through a cell grid over A (EXPERIMENTAL, opt-in:
// MDP_SHELL_GRID=1; the default is the general pair engine's list mode, mdp_pair_list).
//
// Replaces the same search as mdp_pair_list for the residence-time shape (residence_time.py:100-104: ~10^3 central
// atoms, ~10^5 partners, a 3 A shell).  The general engine Hilbert-sorts BOTH sets of every frame; here one CTA per frame
// bins A in shared memory (count, scan, scatter: a few thousand points) and streams B once, coalesced, each B point
// probing the 27 cells around it (shell_grid.h: conservative periodic filter + the reference's own rsq arithmetic, so
// the set of entries is the one mdp_pair_list returns).  24 B of HBM traffic per B point and frame.
#include <algorithm>

#include "common.cuh"
#include "shell_grid.h"

namespace {

constexpr int SG_THREADS = 512;
constexpr int SG_MAX_A = 4096;

struct ListEmit {
    int32_t *list;
    unsigned long long *count;
    long long capacity;
    int frame, ib;
    __device__ __forceinline__ void operator()(int ia) const
    {
        const unsigned long long pos = atomicAdd(count, 1ull);
        if ((long long)pos < capacity) {
            list[pos * 3 + 0] = frame;
            list[pos * 3 + 1] = ia;
            list[pos * 3 + 2] = ib;
        }
    }
};

// grid = frames (strided); dynamic shared memory: sx, sy, sz [na] doubles, sidx [na], cell_of [na], start [ncell_max + 1],
// fill [ncell_max]
__global__ void __launch_bounds__(SG_THREADS) k_shell_grid(const double *__restrict__ xa, long long na, const double *__restrict__ xb,
                                                           long long nb, const double *__restrict__ box, int nframes, double rin2,
                                                           double rout2, int shell_mode, int exclude_same, int32_t *__restrict__ list,
                                                           long long capacity, unsigned long long *__restrict__ count)
{
    extern __shared__ __align__(16) unsigned char sg_smem[];
    constexpr int NCELL_MAX = SG_NC_MAX * SG_NC_MAX * SG_NC_MAX;
    double *sx = reinterpret_cast<double *>(sg_smem);
    double *sy = sx + na;
    double *sz = sy + na;
    int *sidx = reinterpret_cast<int *>(sz + na);
    int *cell_of = sidx + na;
    int *start = cell_of + na;                  // [NCELL_MAX + 1]
    int *fill = start + NCELL_MAX + 1;          // [NCELL_MAX]
    const int tid = threadIdx.x;
    const double r = sqrt(rout2);
    for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
        const double *ax = xa + (long long)f * 3 * na, *ay = ax + na, *az = ay + na;
        const double *bx = xb + (long long)f * 3 * nb, *by = bx + nb, *bz = by + nb;
        ShellGrid g;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            g.len[k] = box[f * 3 + k];
            g.nc[k] = mdp_grid_cells(g.len[k], r);          // >= 3: checked on the host for every frame
            g.inv_w[k] = (double)g.nc[k] / g.len[k];
        }
        g.origin[0] = ax[0];
        g.origin[1] = ay[0];
        g.origin[2] = az[0];
        const int ncell = g.nc[0] * g.nc[1] * g.nc[2];
        __syncthreads();                                     // the previous frame's grid is no longer read
        for (int c = tid; c < ncell; c += SG_THREADS) fill[c] = 0;
        __syncthreads();
        for (int i = tid; i < (int)na; i += SG_THREADS) {
            const int c = mdp_grid_cell(g, ax[i], ay[i], az[i]);
            cell_of[i] = c;
            atomicAdd(&fill[c], 1);
        }
        __syncthreads();
        if (tid < 32) {                                      // exclusive scan of the cell counts by one warp
            const int per = (ncell + 31) / 32;
            const int b = tid * per, e = b + per < ncell ? b + per : ncell;
            int s = 0;
            for (int c = b; c < e; ++c) s += fill[c];
            int inc = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, d);
                if (tid >= d) inc += v;
            }
            int run = inc - s;
            for (int c = b; c < e; ++c) {
                start[c] = run;
                run += fill[c];
                fill[c] = 0;
            }
            if (tid == 31) start[ncell] = inc;
        }
        __syncthreads();
        for (int i = tid; i < (int)na; i += SG_THREADS) {
            const int c = cell_of[i];
            const int p = start[c] + atomicAdd(&fill[c], 1);
            sx[p] = ax[i];
            sy[p] = ay[i];
            sz[p] = az[i];
            sidx[p] = i;
        }
        __syncthreads();
        for (long long j = tid; j < nb; j += SG_THREADS)
            mdp_shell_probe(g, start, sx, sy, sz, sidx, bx[j], by[j], bz[j], (int)j, rin2, rout2, shell_mode, exclude_same,
                            ListEmit{list, count, capacity, f, (int)j});
    }
}

} // namespace

extern "C" {

// returns 0 on success, 1 when the grid search does not apply (the caller then uses mdp_pair_list), < 0 on error
int mdp_shell_search(mdp_ctx *ctx, int nframes, int64_t n_a, const double *xyz_a, int64_t n_b, const double *xyz_b,
                     const double *box, double rin2, double rout2, int shell_mode, int exclude_same_index, int32_t *list_out,
                     int64_t capacity, int64_t *count_out, void *stream)
{
    MDP_REQUIRE(ctx && xyz_a && xyz_b && box && list_out && count_out, "mdp_shell_search: NULL argument");
    MDP_REQUIRE(nframes > 0 && n_a > 0 && n_b > 0 && capacity >= 0 && rout2 > 0.0, "mdp_shell_search: bad sizes");
    if (n_a > SG_MAX_A || n_b > 0x7fffffff) return 1;
    const double r = sqrt(rout2);
    for (int f = 0; f < nframes; ++f)
        for (int k = 0; k < 3; ++k)
            if (mdp_grid_cells(box[(size_t)f * 3 + k], r) < 3) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    constexpr size_t NCELL_MAX = (size_t)SG_NC_MAX * SG_NC_MAX * SG_NC_MAX;
    const size_t smem = (size_t)n_a * (3 * 8 + 4 + 4) + (2 * NCELL_MAX + 1) * 4;
    if (smem > ctx->smem_optin) return 1;
    int rc = ctx->arena_reserve(align256((size_t)nframes * 24) + 4096);
    if (rc) return rc;
    ctx->arena_reset();
    double *d_box = (double *)ctx->arena_take((size_t)nframes * 24);
    if (!d_box) {
        mdp_set_error("internal: scratch arena exhausted (shell search)");
        return MDP_ERR_OOM;
    }
    MDP_CUDA(cudaMemcpyAsync(d_box, box, (size_t)nframes * 24, cudaMemcpyHostToDevice, st));
    MDP_CUDA(cudaMemsetAsync(count_out, 0, 8, st));
    MDP_CUDA(cudaFuncSetAttribute((const void *)k_shell_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = std::max<int>(1, std::min<int>(4, (int)((size_t)220 * 1024 / smem)));
    const unsigned grid = (unsigned)std::min<int64_t>(nframes, (int64_t)ctx->sm_count * per_sm);
    cudaEvent_t tk = ctx->timer_begin(0, st);
    k_shell_grid<<<grid, SG_THREADS, smem, st>>>(xyz_a, n_a, xyz_b, n_b, d_box, nframes, rin2, rout2, shell_mode, exclude_same_index,
                                                 list_out, capacity, (unsigned long long *)count_out);
    ctx->timer_end(tk, st);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_shell_grid");
}

} // extern "C"
