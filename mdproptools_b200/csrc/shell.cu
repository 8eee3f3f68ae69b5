// shell.cu -- neighbour search of a small set A against a large set B through a cell grid over A (the default for
// n_a <= 4096 and n_b >= 8 n_a since round 2; MDP_SHELL_GRID=0 selects the general pair engine's list mode, mdp_pair_list).
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

constexpr int SG_WQ = 384;                      // candidate queue of a warp: (lane << 16 | slot) entries per tile of 32 B points
constexpr int SG_WSTAGE = 128;                  // entries a warp stages in shared memory before one global append

// Hits are staged per WARP in shared memory (one shared atomic each) and appended to the global list with ONE global atomic
// per ~100 entries: a global atomic per hit serialises on the list counter (35 M hits for the C5 residence search -- that,
// not the probing, was most of the 70 ms of the first hardware run), and a per-CTA stage needs a barrier per tile of B.
struct ListEmit {
    int32_t *list;
    unsigned long long *count;
    long long capacity;
    int frame, ib;
    int2 *stage;                 // this warp's SG_WSTAGE entries
    unsigned int *nstage;        // this warp's fill count
    __device__ __forceinline__ void operator()(int ia) const
    {
        const unsigned int p = atomicAdd(nstage, 1u);
        if (p < (unsigned)SG_WSTAGE) {
            stage[p] = make_int2(ia, ib);
            return;
        }
        const unsigned long long pos = atomicAdd(count, 1ull);        // stage full (a tile with an unusual number of hits)
        if ((long long)pos < capacity) {
            list[pos * 3 + 0] = frame;
            list[pos * 3 + 1] = ia;
            list[pos * 3 + 2] = ib;
        }
    }
};

// the warp appends its staged entries to the global list (all lanes call)
__device__ __forceinline__ void warp_flush(int32_t *list, unsigned long long *count, long long capacity, int frame, const int2 *stage,
                                           unsigned int *nstage, int lane)
{
    __syncwarp();
    const unsigned int n = *nstage < (unsigned)SG_WSTAGE ? *nstage : (unsigned)SG_WSTAGE;
    unsigned long long base = 0;
    if (lane == 0 && n) base = atomicAdd(count, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (unsigned int k = lane; k < n; k += 32) {
        const unsigned long long pos = base + k;
        if ((long long)pos < capacity) {
            list[pos * 3 + 0] = frame;
            list[pos * 3 + 1] = stage[k].x;
            list[pos * 3 + 2] = stage[k].y;
        }
    }
    __syncwarp();
    if (lane == 0) *nstage = 0u;
    __syncwarp();
}

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
    // per-warp staging behind the grid (dynamic shared memory, 8-byte aligned: everything before it is a multiple of 4 bytes)
    __shared__ unsigned int nstage[SG_THREADS / 32], ncand[SG_THREADS / 32];
    unsigned char *wbase = reinterpret_cast<unsigned char *>(fill + NCELL_MAX);
    wbase += (8 - (reinterpret_cast<uintptr_t>(wbase) & 7)) & 7;
    double *btile = reinterpret_cast<double *>(wbase);                                   // [warps][96]
    int2 *stage = reinterpret_cast<int2 *>(btile + (SG_THREADS / 32) * 96);              // [warps][SG_WSTAGE]
    unsigned int *candq = reinterpret_cast<unsigned int *>(stage + (SG_THREADS / 32) * SG_WSTAGE);   // [warps][SG_WQ]
    if (tid < SG_THREADS / 32) {
        nstage[tid] = 0u;
        ncand[tid] = 0u;
    }
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
        mdp_grid_set_radius(g, r);
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
        // Every warp streams its own tiles of 32 B points (coalesced), the next tile's coordinates already in flight.  Two
        // phases per tile: (1) every lane walks the cells its point can have partners in and pushes (lane, slot) candidates
        // into the warp's queue -- divergent, but a handful of integer instructions per step; (2) the queued candidates
        // are evaluated 32 at a time, one per lane, with the reference's fp64 arithmetic -- dense.  (A first version
        // evaluated inside the cell walk: 12 of 32 lanes active on average, 500 thread-instructions per B point.)
        {
            const int lane = tid & 31, w = tid >> 5;
            int2 *wst = stage + w * SG_WSTAGE;
            unsigned int *wn = &nstage[w];
            unsigned int *wq = candq + w * SG_WQ;
            unsigned int *wqn = &ncand[w];
            double *wb = btile + w * 96;
            long long j = (long long)w * 32 + lane;
            double cx = 0.0, cy = 0.0, cz = 0.0;
            if (j < nb) {
                cx = bx[j];
                cy = by[j];
                cz = bz[j];
            }
            for (; j - lane < nb; j += SG_THREADS) {
                const long long jn = j + SG_THREADS;
                double nx = 0.0, ny = 0.0, nz = 0.0;
                if (jn < nb) {
                    nx = bx[jn];
                    ny = by[jn];
                    nz = bz[jn];
                }
                wb[lane] = cx;
                wb[32 + lane] = cy;
                wb[64 + lane] = cz;
                const ListEmit emit{list, count, capacity, f, (int)j, wst, wn};
                if (j < nb)
                    mdp_shell_candidates(g, start, cx, cy, cz, [&](int k) {
                        const unsigned int p = atomicAdd(wqn, 1u);
                        if (p < (unsigned)SG_WQ)
                            wq[p] = ((unsigned)lane << 16) | (unsigned)k;
                        else if (mdp_shell_pair_ok(g, sx[k], sy[k], sz[k], sidx[k], cx, cy, cz, (int)j, rin2, rout2, shell_mode, exclude_same))
                            emit(sidx[k]);                            // queue full: settle it here
                    });
                __syncwarp();
                const unsigned int nq = *wqn < (unsigned)SG_WQ ? *wqn : (unsigned)SG_WQ;
                const long long jbase = j - lane;
                for (unsigned int q0 = 0; q0 < nq; q0 += 32) {
                    const unsigned int q = q0 + lane;
                    if (q < nq) {
                        const unsigned int e = wq[q];
                        const int bl = (int)(e >> 16), k = (int)(e & 0xffffu);
                        const int ib = (int)(jbase + bl);
                        if (mdp_shell_pair_ok(g, sx[k], sy[k], sz[k], sidx[k], wb[bl], wb[32 + bl], wb[64 + bl], ib, rin2, rout2, shell_mode,
                                              exclude_same))
                            ListEmit{list, count, capacity, f, ib, wst, wn}(sidx[k]);
                    }
                    __syncwarp();
                    if (*wn >= (unsigned)(SG_WSTAGE / 2)) warp_flush(list, count, capacity, f, wst, wn, lane);
                }
                __syncwarp();
                if (lane == 0) *wqn = 0u;
                __syncwarp();
                cx = nx;
                cy = ny;
                cz = nz;
            }
            warp_flush(list, count, capacity, f, wst, wn, lane);
        }
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
    constexpr size_t SG_WARP_BYTES = (SG_THREADS / 32) * (SG_WSTAGE * 8 + SG_WQ * 4 + 96 * 8);
    const size_t smem = (size_t)n_a * (3 * 8 + 4 + 4) + (2 * NCELL_MAX + 1) * 4 + 8 + SG_WARP_BYTES;
    if (smem + 1024 > ctx->smem_optin) return 1;
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
    const int per_sm = std::max<int>(1, std::min<int>(4, (int)((size_t)220 * 1024 / (smem + 1024))));
    const unsigned grid = (unsigned)std::min<int64_t>(nframes, (int64_t)ctx->sm_count * per_sm);
    cudaEvent_t tk = ctx->timer_begin(0, st);
    k_shell_grid<<<grid, SG_THREADS, smem, st>>>(xyz_a, n_a, xyz_b, n_b, d_box, nframes, rin2, rout2, shell_mode, exclude_same_index,
                                                 list_out, capacity, (unsigned long long *)count_out);
    ctx->timer_end(tk, st);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_shell_grid");
}

} // extern "C"
