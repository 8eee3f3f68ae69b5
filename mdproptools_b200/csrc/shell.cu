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

constexpr int SG_MAX_A = 4096;

constexpr int SG_WQ = 192;                      // candidate queue of a warp: (lane << 16 | slot) entries, drained when it could overflow
constexpr int SG_WSTAGE = 96;                   // entries a warp stages in shared memory before one global append
constexpr unsigned SG_FULL = 0xffffffffu;

// Hits are staged per WARP in shared memory and appended to the global list with ONE global atomic per ~64..96 entries: a
// global atomic per hit serialises on the list counter (35 M hits for the C5 residence search -- that, not the probing,
// was most of the 70 ms of the first hardware run), and a per-CTA stage needs a barrier per tile of B.
// The warp appends its n staged entries to the global list (all lanes call, n uniform).
__device__ __forceinline__ void warp_flush(int32_t *list, unsigned long long *count, long long capacity, int frame, const int2 *stage,
                                           unsigned int n, int lane)
{
    __syncwarp();
    unsigned long long base = 0;
    if (lane == 0 && n) base = atomicAdd(count, (unsigned long long)n);
    base = __shfl_sync(SG_FULL, base, 0);
    for (unsigned int k = lane; k < n; k += 32) {
        const unsigned long long pos = base + k;
        if ((long long)pos < capacity) {
            list[pos * 3 + 0] = frame;
            list[pos * 3 + 1] = stage[k].x;
            list[pos * 3 + 2] = stage[k].y;
        }
    }
    __syncwarp();
}

// The nq queued candidates of the warp's current tile of 32 B points, one per lane and round, through the reference's own
// fp64 test; accepted ones go to the warp's stage by ballot compaction (nst = entries staged, uniform).
__device__ __forceinline__ void shell_drain(const ShellGrid &g, const double *sx, const double *sy, const double *sz, const int *sidx,
                                            const double *wb, const unsigned int *wq, unsigned int nq, long long jbase, double rin2,
                                            double rout2, int shell_mode, int exclude_same, int32_t *list, unsigned long long *count,
                                            long long capacity, int frame, int2 *wst, unsigned int &nst, int lane)
{
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    for (unsigned int q0 = 0; q0 < nq; q0 += 32) {
        const unsigned int q = q0 + lane;
        bool ok = false;
        int ia = 0, ib = 0;
        if (q < nq) {
            const unsigned int e = wq[q];
            const int bl = (int)(e >> 16), k = (int)(e & 0xffffu);
            ib = (int)(jbase + bl);
            ia = sidx[k];
            ok = mdp_shell_pair_ok(g, sx[k], sy[k], sz[k], ia, wb[bl], wb[32 + bl], wb[64 + bl], ib, rin2, rout2, shell_mode, exclude_same);
        }
        const unsigned bal = __ballot_sync(SG_FULL, ok);
        if (ok) wst[nst + __popc(bal & lt)] = make_int2(ia, ib);
        nst += __popc(bal);
        if (nst > (unsigned)(SG_WSTAGE - 32)) {
            warp_flush(list, count, capacity, frame, wst, nst, lane);
            nst = 0;
        }
    }
    __syncwarp();
}

// grid = frames (strided); dynamic shared memory: sx, sy, sz [na] doubles, sidx [na], cell_of [na], start [ncell_max + 1],
// the halo copy of the cell ranges [(nc_max + 2)^3] (its first part doubles as the scatter cursors), then per warp: the B tile (96 doubles), the stage and the candidate queue
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_shell_grid(const double *__restrict__ xa, long long na, const double *__restrict__ xb,
                                                        long long nb, const double *__restrict__ box, int nframes, double rin2,
                                                        double rout2, int shell_mode, int exclude_same, int32_t *__restrict__ list,
                                                        long long capacity, unsigned long long *__restrict__ count)
{
    extern __shared__ __align__(16) unsigned char sg_smem[];
    constexpr int NCELL_MAX = SG_NC_MAX * SG_NC_MAX * SG_NC_MAX;
    constexpr int HALO_MAX = (SG_NC_MAX + 2) * (SG_NC_MAX + 2) * (SG_NC_MAX + 2);
    constexpr int WARPS = THREADS / 32;
    double *sx = reinterpret_cast<double *>(sg_smem);
    double *sy = sx + na;
    double *sz = sy + na;
    int *sidx = reinterpret_cast<int *>(sz + na);
    int *cell_of = sidx + na;
    int *start = cell_of + na;                  // [NCELL_MAX + 1]
    uint32_t *halo = reinterpret_cast<uint32_t *>(start + NCELL_MAX + 1);   // [HALO_MAX]: the ranges the walk reads
    int *fill = reinterpret_cast<int *>(halo);  // [NCELL_MAX] scatter cursors, dead before the halo copy is written
    const int tid = threadIdx.x;
    const double r = sqrt(rout2);
    __shared__ ShellGrid g;
    // per-warp buffers behind the grid (8-byte aligned: everything before them is a multiple of 4 bytes)
    unsigned char *wbase = reinterpret_cast<unsigned char *>(halo + HALO_MAX);
    wbase += (8 - (reinterpret_cast<uintptr_t>(wbase) & 7)) & 7;
    double *btile = reinterpret_cast<double *>(wbase);                                   // [warps][96]
    int2 *stage = reinterpret_cast<int2 *>(btile + WARPS * 96);                          // [warps][SG_WSTAGE]
    unsigned int *candq = reinterpret_cast<unsigned int *>(stage + WARPS * SG_WSTAGE);   // [warps][SG_WQ]
    for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
        const double *ax = xa + (long long)f * 3 * na, *ay = ax + na, *az = ay + na;
        const double *bx = xb + (long long)f * 3 * nb, *by = bx + nb, *bz = by + nb;
        __syncthreads();                                     // the previous frame's grid is no longer read
        if (tid == 0) {                                      // the frame's grid constants live in shared memory: 27 registers less
            ShellGrid t;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                t.len[k] = box[f * 3 + k];
                t.nc[k] = mdp_grid_cells(t.len[k], r);      // >= 3: checked on the host for every frame
                t.inv_w[k] = (double)t.nc[k] / t.len[k];
            }
            mdp_grid_set_radius(t, r);
            t.origin[0] = ax[0];
            t.origin[1] = ay[0];
            t.origin[2] = az[0];
            g = t;
        }
        __syncthreads();
        const int ncell = g.nc[0] * g.nc[1] * g.nc[2];
        for (int c = tid; c < ncell; c += THREADS) fill[c] = 0;
        __syncthreads();
        for (int i = tid; i < (int)na; i += THREADS) {
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
                const int v = __shfl_up_sync(SG_FULL, inc, d);
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
        for (int i = tid; i < (int)na; i += THREADS) {
            const int c = cell_of[i];
            const int p = start[c] + atomicAdd(&fill[c], 1);
            sx[p] = ax[i];
            sy[p] = ay[i];
            sz[p] = az[i];
            sidx[p] = i;
        }
        __syncthreads();
        for (int h = tid, nh = mdp_halo_cells(g); h < nh; h += THREADS) halo[h] = mdp_halo_entry(g, start, h);
        __syncthreads();
        // Every warp streams its own tiles of 32 B points (coalesced), the next tile's coordinates already in flight.  Two
        // phases per tile, both in CONVERGENT control flow:
        //  (1) the cell walk.  Every lane has 1..27 cells to visit (shell_grid.h: 5.4 on average, at most 8 for the residence
        //      shape), consecutive boxes of the halo grid; the warp loops to the LARGEST count among its lanes, lane l
        //      visiting its c-th cell while c is below its own count.  Walked twice: once to count the lane's candidates (one
        //      warp scan then gives every lane its place in the warp's candidate queue), once to write them.  A walk with
        //      per-lane loop bounds serialises lane by lane (round 2a: 43 % of the issue slots used, the rest stalls), and a
        //      ballot compaction per cell costs ~30 instructions per cell and round.
        //  (2) the queued candidates are evaluated 32 at a time, one per lane, with the reference's fp64 arithmetic.
        {
            const int lane = tid & 31, w = tid >> 5;
            const unsigned lt = (1u << lane) - 1u;
            int2 *wst = stage + w * SG_WSTAGE;
            unsigned int *wq = candq + w * SG_WQ;
            double *wb = btile + w * 96;
            unsigned int nst = 0;
            long long j = (long long)w * 32 + lane;
            double cx = 0.0, cy = 0.0, cz = 0.0;
            if (j < nb) {
                cx = bx[j];
                cy = by[j];
                cz = bz[j];
            }
            for (; j - lane < nb; j += THREADS) {
                const long long jn = j + THREADS;
                double nx = 0.0, ny = 0.0, nz = 0.0;
                if (jn < nb) {
                    nx = bx[jn];
                    ny = by[jn];
                    nz = bz[jn];
                }
                wb[lane] = cx;
                wb[32 + lane] = cy;
                wb[64 + lane] = cz;
                const long long jbase = j - lane;
                ShellWalk wk = mdp_shell_walk_begin(g, cx, cy, cz);
                if (j >= nb) wk.tot = 0;
                const int maxtot = __reduce_max_sync(SG_FULL, wk.tot);
                // how many candidates each lane has (first walk), one scan over the warp: every lane knows where its entries go
                int mine = 0;
                {
                    ShellWalk w1 = wk;
                    for (int c = 0; c < maxtot; ++c) {
                        unsigned int e = 0;
                        if (c < w1.tot) e = halo[w1.cell];
                        mdp_shell_walk_next(w1);
                        mine += (int)(e >> 16);
                    }
                }
                int inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(SG_FULL, inc, d);
                    if (lane >= d) inc += v;
                }
                unsigned int nq = (unsigned)__shfl_sync(SG_FULL, inc, 31);
                if (nq <= (unsigned)SG_WQ) {
                    // (second walk) the lane writes its entries behind those of the lanes below it
                    unsigned int off = (unsigned)(inc - mine);
                    const unsigned lanebits = (unsigned)lane << 16;
                    for (int c = 0; c < maxtot; ++c) {
                        unsigned int e = 0;
                        if (c < wk.tot) e = halo[wk.cell];
                        mdp_shell_walk_next(wk);
                        const unsigned val = lanebits | (e & 0xffffu);
                        const int cnt = (int)(e >> 16);
                        if (cnt > 0) wq[off] = val;
                        if (cnt > 1) wq[off + 1] = val + 1;
                        if (cnt > 2) wq[off + 2] = val + 2;
                        if (__builtin_expect(cnt > 3, 0))
                            for (int u = 3; u < cnt; ++u) wq[off + u] = val + u;
                        off += cnt;
                    }
                } else {
                    // more candidates than the queue holds (dense A): cell by cell, ballot-compacted, draining when full
                    nq = 0;
                    for (int c = 0; c < maxtot; ++c) {
                        unsigned int e = 0;
                        if (c < wk.tot) e = halo[wk.cell];
                        mdp_shell_walk_next(wk);
                        const int s = (int)(e & 0xffffu), cnt = (int)(e >> 16);
                        const int maxc = __reduce_max_sync(SG_FULL, cnt);
                        for (int u = 0; u < maxc; ++u) {
                            if (nq + 32u > (unsigned)SG_WQ) {            // uniform: settle what is queued, then go on
                                shell_drain(g, sx, sy, sz, sidx, wb, wq, nq, jbase, rin2, rout2, shell_mode, exclude_same, list, count,
                                            capacity, f, wst, nst, lane);
                                nq = 0;
                            }
                            const bool has = u < cnt;
                            const unsigned bal = __ballot_sync(SG_FULL, has);
                            if (has) wq[nq + __popc(bal & lt)] = ((unsigned)lane << 16) | (unsigned)(s + u);
                            nq += __popc(bal);
                        }
                    }
                }
                shell_drain(g, sx, sy, sz, sidx, wb, wq, nq, jbase, rin2, rout2, shell_mode, exclude_same, list, count, capacity, f, wst,
                            nst, lane);
                cx = nx;
                cy = ny;
                cz = nz;
            }
            warp_flush(list, count, capacity, f, wst, nst, lane);
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
    // 1024 threads per CTA (32 warps hide the shared-memory latencies of the walk) when the grid leaves room, else 512
    constexpr size_t HALO_MAX = (size_t)(SG_NC_MAX + 2) * (SG_NC_MAX + 2) * (SG_NC_MAX + 2);
    const size_t grid_bytes = (size_t)n_a * (3 * 8 + 4 + 4) + (NCELL_MAX + 1 + HALO_MAX) * 4 + 8;
    constexpr size_t SG_WARP_BYTES = SG_WSTAGE * 8 + SG_WQ * 4 + 96 * 8;
    int threads = 1024;
    if (grid_bytes + 32 * SG_WARP_BYTES + 1024 > ctx->smem_optin) threads = 512;
    const size_t smem = grid_bytes + (size_t)(threads / 32) * SG_WARP_BYTES;
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
    const int per_sm = std::max<int>(1, std::min<int>(2048 / threads, (int)((size_t)220 * 1024 / (smem + 1024))));
    const unsigned grid = (unsigned)std::min<int64_t>(nframes, (int64_t)ctx->sm_count * per_sm);
    cudaEvent_t tk = ctx->timer_begin(0, st);
    if (threads == 1024) {
        MDP_CUDA(cudaFuncSetAttribute((const void *)k_shell_grid<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_shell_grid<1024><<<grid, 1024, smem, st>>>(xyz_a, n_a, xyz_b, n_b, d_box, nframes, rin2, rout2, shell_mode, exclude_same_index,
                                                     list_out, capacity, (unsigned long long *)count_out);
    } else {
        MDP_CUDA(cudaFuncSetAttribute((const void *)k_shell_grid<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_shell_grid<512><<<grid, 512, smem, st>>>(xyz_a, n_a, xyz_b, n_b, d_box, nframes, rin2, rout2, shell_mode, exclude_same_index,
                                                   list_out, capacity, (unsigned long long *)count_out);
    }
    ctx->timer_end(tk, st);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_shell_grid");
}

} // extern "C"
