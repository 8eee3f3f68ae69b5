// reduce.cu -- streaming fp64 reductions of the dynamical hot path (HBM-bound kernels).
//
// Replaces: calc_com (common/com_mols.py:5-62), _define_mol_cols (structural/rdf_cn.py:218-241),
// the MSD arithmetic of Diffusion.get_msd_from_dump (dynamical/diffusion.py:207-238), conductivity_loop
// (dynamical/_conductivity.py:7-36) and the OLS sums behind Diffusion.calc_diff (diffusion.py:323-329).
// All arithmetic is unfused fp64 in the reference's operation order; block partial sums are combined in a
// fixed order (two-stage, no floating-point atomics) so results are run-to-run deterministic.
#include <float.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int RB = 256;          // threads per block
constexpr int MSD_APT = 8;       // atoms per thread in the streaming MSD kernel

__device__ __forceinline__ double block_sum(double v, double *sm)
{
    // deterministic block reduction (fixed shuffle tree, fixed warp order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, o));
    }
    return r;   // valid in warp 0
}

__device__ __forceinline__ double2 ld_stream2(const double *p)
{
    double2 v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream1(const double *p)
{
    double v;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// ---- MSD, single time origin ----------------------------------------------------------------
// grid (nchunks, nframes); a chunk is RB*MSD_APT consecutive atoms of [a0, a1).
template <bool VEC, bool PER_ATOM>
__global__ void __launch_bounds__(RB) k_msd_single(const double *__restrict__ traj, const double *__restrict__ ref, int64_t n,
                                                   int64_t a0, int64_t a1, double scale, double *__restrict__ partial,
                                                   int nchunks, double *__restrict__ per_atom)
{
    __shared__ double sm[32];
    const int f = blockIdx.y;
    const double *tx = traj + (int64_t)f * 3 * n;
    const int64_t base = a0 + (int64_t)blockIdx.x * (RB * MSD_APT);
    double sx = 0, sy = 0, sz = 0, st = 0;
    if (VEC) {
#pragma unroll
        for (int k = 0; k < MSD_APT / 2; ++k) {
            const int64_t i = base + (int64_t)k * (RB * 2) + threadIdx.x * 2;
            if (i + 1 < a1) {
                const double2 x = ld_stream2(tx + i), y = ld_stream2(tx + n + i), z = ld_stream2(tx + 2 * n + i);
                const double2 rx = *reinterpret_cast<const double2 *>(ref + i);
                const double2 ry = *reinterpret_cast<const double2 *>(ref + n + i);
                const double2 rz = *reinterpret_cast<const double2 *>(ref + 2 * n + i);
                // SI conversion before differencing (diffusion.py:201-203, 214)
                double dx0 = __dsub_rn(__dmul_rn(x.x, scale), __dmul_rn(rx.x, scale));
                double dy0 = __dsub_rn(__dmul_rn(y.x, scale), __dmul_rn(ry.x, scale));
                double dz0 = __dsub_rn(__dmul_rn(z.x, scale), __dmul_rn(rz.x, scale));
                double dx1 = __dsub_rn(__dmul_rn(x.y, scale), __dmul_rn(rx.y, scale));
                double dy1 = __dsub_rn(__dmul_rn(y.y, scale), __dmul_rn(ry.y, scale));
                double dz1 = __dsub_rn(__dmul_rn(z.y, scale), __dmul_rn(rz.y, scale));
                dx0 = __dmul_rn(dx0, dx0); dy0 = __dmul_rn(dy0, dy0); dz0 = __dmul_rn(dz0, dz0);
                dx1 = __dmul_rn(dx1, dx1); dy1 = __dmul_rn(dy1, dy1); dz1 = __dmul_rn(dz1, dz1);
                const double m0 = __dadd_rn(__dadd_rn(dx0, dy0), dz0), m1 = __dadd_rn(__dadd_rn(dx1, dy1), dz1);
                sx = __dadd_rn(sx, __dadd_rn(dx0, dx1));
                sy = __dadd_rn(sy, __dadd_rn(dy0, dy1));
                sz = __dadd_rn(sz, __dadd_rn(dz0, dz1));
                st = __dadd_rn(st, __dadd_rn(m0, m1));
                if (PER_ATOM) {
                    double *o = per_atom + (int64_t)f * 4 * n;
                    *reinterpret_cast<double2 *>(o + i) = make_double2(dx0, dx1);
                    *reinterpret_cast<double2 *>(o + n + i) = make_double2(dy0, dy1);
                    *reinterpret_cast<double2 *>(o + 2 * n + i) = make_double2(dz0, dz1);
                    *reinterpret_cast<double2 *>(o + 3 * n + i) = make_double2(m0, m1);
                }
            } else if (i < a1) {
                double dx0 = __dsub_rn(__dmul_rn(tx[i], scale), __dmul_rn(ref[i], scale));
                double dy0 = __dsub_rn(__dmul_rn(tx[n + i], scale), __dmul_rn(ref[n + i], scale));
                double dz0 = __dsub_rn(__dmul_rn(tx[2 * n + i], scale), __dmul_rn(ref[2 * n + i], scale));
                dx0 = __dmul_rn(dx0, dx0); dy0 = __dmul_rn(dy0, dy0); dz0 = __dmul_rn(dz0, dz0);
                const double m0 = __dadd_rn(__dadd_rn(dx0, dy0), dz0);
                sx = __dadd_rn(sx, dx0); sy = __dadd_rn(sy, dy0); sz = __dadd_rn(sz, dz0); st = __dadd_rn(st, m0);
                if (PER_ATOM) {
                    double *o = per_atom + (int64_t)f * 4 * n;
                    o[i] = dx0; o[n + i] = dy0; o[2 * n + i] = dz0; o[3 * n + i] = m0;
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < MSD_APT; ++k) {
            const int64_t i = base + (int64_t)k * RB + threadIdx.x;
            if (i < a1) {
                double dx0 = __dsub_rn(__dmul_rn(ld_stream1(tx + i), scale), __dmul_rn(ref[i], scale));
                double dy0 = __dsub_rn(__dmul_rn(ld_stream1(tx + n + i), scale), __dmul_rn(ref[n + i], scale));
                double dz0 = __dsub_rn(__dmul_rn(ld_stream1(tx + 2 * n + i), scale), __dmul_rn(ref[2 * n + i], scale));
                dx0 = __dmul_rn(dx0, dx0); dy0 = __dmul_rn(dy0, dy0); dz0 = __dmul_rn(dz0, dz0);
                const double m0 = __dadd_rn(__dadd_rn(dx0, dy0), dz0);
                sx = __dadd_rn(sx, dx0); sy = __dadd_rn(sy, dy0); sz = __dadd_rn(sz, dz0); st = __dadd_rn(st, m0);
                if (PER_ATOM) {
                    double *o = per_atom + (int64_t)f * 4 * n;
                    o[i] = dx0; o[n + i] = dy0; o[2 * n + i] = dz0; o[3 * n + i] = m0;
                }
            }
        }
    }
    sx = block_sum(sx, sm);
    sy = block_sum(sy, sm);
    sz = block_sum(sz, sm);
    st = block_sum(st, sm);
    if (threadIdx.x == 0) {
        double *o = partial + ((int64_t)f * nchunks + blockIdx.x) * 4;
        o[0] = sx; o[1] = sy; o[2] = sz; o[3] = st;
    }
}

// second stage: out[f][g][c] = sum over chunks (fixed order); one warp per (f, c)
__global__ void __launch_bounds__(128) k_partial_sum(const double *__restrict__ partial, int nchunks, int ncomp,
                                                     double *__restrict__ out, int64_t out_stride_f)
{
    const int f = blockIdx.x;
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= ncomp) return;
    double s = 0.0;
    for (int k = lane; k < nchunks; k += 32) s = __dadd_rn(s, partial[((int64_t)f * nchunks + k) * ncomp + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (lane == 0) out[(int64_t)f * out_stride_f + c] = s;
}

// ---- msd_int (diffusion.py:225-237) ----------------------------------------------------------
__global__ void __launch_bounds__(RB) k_msd_interval(const double *__restrict__ traj, int nframes, int64_t n, double scale,
                                                     int stride, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double px = __dmul_rn(traj[i], scale), py = __dmul_rn(traj[n + i], scale), pz = __dmul_rn(traj[2 * n + i], scale);
    double sx = 0, sy = 0, sz = 0, sm = 0;
    int nint = 1;
    for (int f = stride; f < nframes; f += stride, ++nint) {
        const double *t = traj + (int64_t)f * 3 * n;
        const double x = __dmul_rn(t[i], scale), y = __dmul_rn(t[n + i], scale), z = __dmul_rn(t[2 * n + i], scale);
        double dx = __dsub_rn(x, px), dy = __dsub_rn(y, py), dz = __dsub_rn(z, pz);
        dx = __dmul_rn(dx, dx); dy = __dmul_rn(dy, dy); dz = __dmul_rn(dz, dz);
        sx = __dadd_rn(sx, dx); sy = __dadd_rn(sy, dy); sz = __dadd_rn(sz, dz);
        sm = __dadd_rn(sm, __dadd_rn(__dadd_rn(dx, dy), dz));
        px = x; py = y; pz = z;
    }
    // dx2/dy2/dz2: mean over the nint-1 non-NaN rows; msd: the NaN row became 0 and counts (reference quirk)
    const double d1 = (double)(nint - 1), d0 = (double)nint;
    out[i] = nint > 1 ? sx / d1 : NAN;
    out[n + i] = nint > 1 ? sy / d1 : NAN;
    out[2 * n + i] = nint > 1 ? sz / d1 : NAN;
    out[3 * n + i] = sm / d0;
}

// ---- MSD over all time origins (windowed) ---------------------------------------------------------
// out[lag] = sum_atoms sum_{t0, t0+lag < T} (x(t0+lag) - x(t0))^2 per axis.  FP64-pipe bound for any window longer
// than ~12 lags (2 pipe slots -- DADD + DFMA -- per (atom, axis, origin, lag) against 24 B per atom-frame).
// Tiling: lane = atom (32 consecutive atoms, rows of 256 B are read coalesced), warp = 16 consecutive lags held
// as 16 register accumulators + a 16-deep register window that slides along t0, so one step costs 2 shared-memory
// loads for 32 FP64 instructions.  A CTA (8 warps = 128 lags per pass) stages MW_TT origins and the MW_TT+128 rows
// they pair with in shared memory with cp.async (zero fill beyond the trajectory / the atom range), walks the lag
// blocks of the launch for every time tile (the rows come from L2 after the first touch, HBM is streamed once),
// and keeps the per-lag sums of its atoms in shared memory; CTAs are persistent over (atom block, time chunk)
// items in a fixed order and write one partial row each, which a second kernel adds up in a fixed order
// (deterministic, no floating-point atomics).
constexpr int MW_LB = 16;                    // lags per thread
constexpr int MW_WARPS = 8;
constexpr int MW_LBW = MW_LB * MW_WARPS;     // 128 lags per pass
constexpr int MW_TT = 96;                    // origins per time tile (multiple of MW_LB)
constexpr int MW_BROWS = MW_TT + MW_LBW;     // rows of the partner tile
constexpr int MW_WCAP = 1024;                // lags per launch (per-CTA accumulator 3 x 1024 doubles)

__device__ __forceinline__ void cp_async8_zfill(void *smem_dst, const void *gmem_src, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}

template <bool EDGE>
__device__ __forceinline__ void mw_tile(const double *__restrict__ A, const double *__restrict__ B, int lane, int l0, int steps,
                                        int64_t npair0, double (&acc)[MW_LB])
{
    // B row r holds x(ta + lagbase + r); the pair (origin s, lag l0+k) reads B[s + l0 + k].
    // EDGE: origin s is valid for s < steps, partner valid while s + l0 + k < npair0 (= T - ta - lagbase).
    double win[MW_LB];
#pragma unroll
    for (int k = 0; k < MW_LB - 1; ++k) win[k] = B[(l0 + k) * 32 + lane];
#pragma unroll 1
    for (int s = 0; s < MW_TT; s += MW_LB) {
        if (EDGE && s >= steps) break;
        // Validity is uniform over the warp (lane = atom): a block of MW_LB origins whose every (origin, lag) pair lies
        // inside the chunk and the trajectory takes the unpredicated body even in an edge tile; only the one or two
        // blocks that straddle the end pay for the per-pair test.
        const bool partial = EDGE && (s + MW_LB > steps || npair0 - (s + MW_LB - 1) - l0 < MW_LB);
        if (partial) {
#pragma unroll
            for (int u = 0; u < MW_LB; ++u) {
                win[(u + MW_LB - 1) % MW_LB] = B[(s + u + l0 + MW_LB - 1) * 32 + lane];
                const double a = A[(s + u) * 32 + lane];
                const int64_t kmax = (s + u < steps) ? npair0 - (s + u) - l0 : 0;
#pragma unroll
                for (int k = 0; k < MW_LB; ++k) {
                    const double d = win[(u + k) % MW_LB] - a;
                    if (k < kmax) acc[k] = fma(d, d, acc[k]);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < MW_LB; ++u) {
                win[(u + MW_LB - 1) % MW_LB] = B[(s + u + l0 + MW_LB - 1) * 32 + lane];
                const double a = A[(s + u) * 32 + lane];
#pragma unroll
                for (int k = 0; k < MW_LB; ++k) {
                    const double d = win[(u + k) % MW_LB] - a;
                    acc[k] = fma(d, d, acc[k]);
                }
            }
        }
    }
}

// grid = persistent CTAs; items = nblk atom blocks x ntc time chunks
__global__ void __launch_bounds__(MW_WARPS * 32, 2) k_msd_window(const double *__restrict__ traj, int T, int64_t n, int64_t a0,
                                                                 int64_t a1, int lag0, int nl, int nblk, int ntc, int tc_len,
                                                                 double *__restrict__ partial)
{
    extern __shared__ __align__(16) double mw_smem[];
    double *A = mw_smem;                          // [MW_TT][32]
    double *B = A + MW_TT * 32;                   // [MW_BROWS][32]
    double *accs = B + MW_BROWS * 32;             // [3][nlp]
    const int nlp = (nl + MW_LBW - 1) / MW_LBW * MW_LBW;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int l0 = w * MW_LB;
    for (int k = tid; k < 3 * nlp; k += blockDim.x) accs[k] = 0.0;

    const int nitems = nblk * ntc;
#pragma unroll 1
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int blk = item / ntc, tc = item % ntc;
        const int64_t atom0 = a0 + (int64_t)blk * 32;
        const int tbeg = tc * tc_len;
        const int tend = tbeg + tc_len < T ? tbeg + tc_len : T;
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            const double *xc = traj + (int64_t)c * n + atom0;     // row t of this axis: xc + t*3n
#pragma unroll 1
            for (int lb0 = 0; lb0 < nlp; lb0 += MW_LBW) {
                const int lagbase = lag0 + lb0;
                double acc[MW_LB];
#pragma unroll
                for (int k = 0; k < MW_LB; ++k) acc[k] = 0.0;
#pragma unroll 1
                for (int ta = tbeg; ta < tend; ta += MW_TT) {
                    if (ta + lagbase >= T) break;                 // no partner inside the trajectory
                    __syncthreads();                              // previous tile fully consumed
                    {
                        // warp w copies rows w, w+8, ... (lane = atom): one pointer bump per copy, no index arithmetic
                        const int64_t rs = (int64_t)3 * n;
                        const bool lane_ok = atom0 + lane < a1;
                        const double *src = xc + (int64_t)(ta + w) * rs + lane;
                        double *dst = A + w * 32 + lane;
#pragma unroll 4
                        for (int r = w; r < MW_TT; r += MW_WARPS) {
                            const bool ok = lane_ok && ta + r < T;
                            cp_async8_zfill(dst, ok ? src : xc, ok);
                            src += MW_WARPS * rs;
                            dst += MW_WARPS * 32;
                        }
                        src = xc + (int64_t)(ta + lagbase + w) * rs + lane;
                        dst = B + w * 32 + lane;
#pragma unroll 4
                        for (int r = w; r < MW_BROWS; r += MW_WARPS) {
                            const bool ok = lane_ok && ta + lagbase + r < T;
                            cp_async8_zfill(dst, ok ? src : xc, ok);
                            src += MW_WARPS * rs;
                            dst += MW_WARPS * 32;
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncthreads();
                    const int steps = tend - ta < MW_TT ? tend - ta : MW_TT;
                    const int64_t npair0 = (int64_t)T - ta - lagbase;     // partner row r is inside the trajectory iff r < npair0
                    const bool edge = steps < MW_TT || npair0 < MW_TT + MW_LBW - 1;
                    if (edge)
                        mw_tile<true>(A, B, lane, l0, steps, npair0, acc);
                    else
                        mw_tile<false>(A, B, lane, l0, steps, npair0, acc);
                }
                // sum over the 32 atoms of the block (fixed shuffle tree), then into this CTA's per-lag sums
#pragma unroll
                for (int k = 0; k < MW_LB; ++k) {
                    double v = acc[k];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == k) accs[c * nlp + lb0 + l0 + k] += v;
                }
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < 3 * nl; k += blockDim.x) {
        const int c = k / nl, l = k % nl;
        partial[((int64_t)blockIdx.x * 3 + c) * nl + l] = accs[c * nlp + l];
    }
}

// out[lag0 + l][g][0..3] += scale^2 * sum over CTAs (fixed order); one thread per lag
__global__ void __launch_bounds__(128) k_msd_window_finish(const double *__restrict__ partial, int nctas, int nl, int lag0,
                                                           double scale2, double *__restrict__ out, int64_t out_stride)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    double s[3] = {0.0, 0.0, 0.0};
    for (int b = 0; b < nctas; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) s[c] += partial[((int64_t)b * 3 + c) * nl + l];
    double *o = out + (int64_t)(lag0 + l) * out_stride;
    const double x = s[0] * scale2, y = s[1] * scale2, z = s[2] * scale2;
    o[0] += x;
    o[1] += y;
    o[2] += z;
    o[3] += (x + y) + z;
}

// ---- per-molecule mass-weighted mean (calc_com / _define_mol_cols) ---------------------------------
// thread per (segment, frame); sequential accumulation in atom order, one division at the end.
__global__ void __launch_bounds__(RB) k_segment_com(const double *__restrict__ attr, int ncomp, int64_t n,
                                                    const double *__restrict__ w, int64_t nseg,
                                                    const int32_t *__restrict__ seg_off, double *__restrict__ out,
                                                    double *__restrict__ wsum_out, const double *__restrict__ extra,
                                                    double *__restrict__ extra_out)
{
    const int f = blockIdx.y;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const int b = seg_off[s], e = seg_off[s + 1];
    double ws = 0.0;
    for (int a = b; a < e; ++a) ws = __dadd_rn(ws, w[a]);
    for (int c = 0; c < ncomp; ++c) {
        const double *p = attr + ((int64_t)f * ncomp + c) * n;
        double acc = 0.0;
        for (int a = b; a < e; ++a) acc = __dadd_rn(acc, __dmul_rn(p[a], w[a]));
        out[((int64_t)f * ncomp + c) * nseg + s] = acc / ws;
    }
    if (f == 0) {
        if (wsum_out) wsum_out[s] = ws;
        if (extra && extra_out) {
            double q = 0.0;
            for (int a = b; a < e; ++a) q = __dadd_rn(q, extra[a]);
            extra_out[s] = q;
        }
    }
}

// ---- charge flux (_conductivity.py:7-36) -------------------------------------------------------
// per-molecule mass and charge (sequential sums in atom order); static over the frames of a call
__global__ void __launch_bounds__(RB) k_seg_static(const double *__restrict__ mass, const double *__restrict__ q,
                                                   const int32_t *__restrict__ seg_off, int64_t nseg, double *__restrict__ wsum,
                                                   double *__restrict__ qsum)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    double ws = 0.0, qs = 0.0;
    for (int a = seg_off[s]; a < seg_off[s + 1]; ++a) {
        ws = __dadd_rn(ws, mass[a]);
        qs = __dadd_rn(qs, q[a]);
    }
    wsum[s] = ws;
    qsum[s] = qs;
}

// A chunk = FLUX_B consecutive molecules of one molecule type = one contiguous atom range (molecule membership is
// positional).  Grid (nchunks, frame lanes), sized to the resident CTA slots (2 per SM): a CTA keeps ITS
// chunk -- every per-molecule quantity (atom range, mass, charge) is loaded once -- and walks the frames lane, lane + L, ...
// The three velocity rows of the range (streamed once from HBM) and its masses (L2-resident) arrive in shared memory as
// four TMA bulk copies (cp.async.bulk + mbarrier transaction count) per tile of FLUX_TILE atoms, double buffered: the
// copies of the next frame's tile are in flight while one thread per (component, molecule) adds m*v of its own atoms from
// the current tile in atom order, as the reference does.  History: a thread-per-molecule gather straight from global is
// L1-throughput bound (3.5 TB/s); staging through registers with one tile per CTA serialised every CTA's loads behind
// its barriers (4.7 TB/s); per-thread 8-byte cp.async saturated the LSU queue (3.8 TB/s); persistent CTAs walking
// (chunk, frame) items re-read the per-molecule metadata from global for every item, on the critical path (4.2-4.9 TB/s).
// Bulk copies need 16-byte aligned addresses and sizes: each row is copied from its address rounded down to 16 bytes
// (shift sh = 0 or 1 element, applied when the tile is read) with its length rounded up -- at most one neighbouring
// element on either side, which lies inside the same 16-byte granule of the caller's allocation.
constexpr int FLUX_B = 128;
constexpr int FLUX_TILE = 1408;                 // atoms per tile; 4 rows x (1408 + 2) x 8 B = 45 KB per stage
constexpr int FLUX_ROW = FLUX_TILE + 2;         // + shift + round-up
constexpr int FLUX_STAGES = 2;                  // 90 KB per CTA: two CTAs per SM
constexpr int FLUX_T = 3 * FLUX_B;              // threads: one per (component, molecule), component-major
constexpr size_t FLUX_SMEM = (size_t)FLUX_STAGES * 4 * FLUX_ROW * sizeof(double);

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}

// the four rows of one tile [t0, t0 + cnt) of frame f -> stage buffer; returns the element shifts packed 1 bit per row
__device__ __forceinline__ void flux_issue(const double *vel, const double *mass, int64_t n, int f, int t0, int cnt, unsigned buf,
                                           unsigned bar)
{
    const double *rows[4] = {vel + (int64_t)f * 3 * n + t0, vel + (int64_t)f * 3 * n + n + t0, vel + (int64_t)f * 3 * n + 2 * n + t0,
                             mass + t0};
    unsigned bytes[4], total = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned sh = (unsigned)(((uintptr_t)rows[r] >> 3) & 1u);
        bytes[r] = (((unsigned)cnt + sh + 1u) & ~1u) * 8u;
        total += bytes[r];
    }
    mbar_expect_tx(bar, total);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned sh = (unsigned)(((uintptr_t)rows[r] >> 3) & 1u);
        bulk_g2s(buf + (unsigned)(r * FLUX_ROW * 8), rows[r] - sh, bytes[r], bar);
    }
}

__global__ void __launch_bounds__(FLUX_T) k_charge_flux(const double *__restrict__ vel, int64_t n, const double *__restrict__ mass,
                                                        const double *__restrict__ wsum, const double *__restrict__ qsum,
                                                        const int32_t *__restrict__ seg_off, int64_t s0, int64_t s1, int nframes,
                                                        double vel_scale, double q_scale, double *__restrict__ partial, int nchunks)
{
    extern __shared__ __align__(128) double fs[];   // [FLUX_STAGES][4][FLUX_ROW]: vx, vy, vz, mass
    __shared__ __align__(8) unsigned long long bars[FLUX_STAGES];
    __shared__ double sm[2][3][FLUX_B / 32];
    const unsigned fs_a = (unsigned)__cvta_generic_to_shared(fs);
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bars);
    const int tid = threadIdx.x, lane = tid & 31;
    const int comp = tid / FLUX_B, mol = tid - comp * FLUX_B, wc = mol >> 5;   // component (warp uniform), molecule, warp of the component
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < FLUX_STAGES; ++k) mbar_init(bar_a + 8u * k, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // my chunk and my molecule: loaded once
    const int chunk = blockIdx.x;
    const int64_t sfirst = s0 + (int64_t)chunk * FLUX_B;
    const int64_t slast = sfirst + FLUX_B < s1 ? sfirst + FLUX_B : s1;
    const int64_t s = sfirst + mol;
    const int A0 = seg_off[sfirst], A1 = seg_off[slast];
    const bool mine = s < s1;
    const int b = mine ? seg_off[s] : A1, e = mine ? seg_off[s + 1] : A1;
    const double ws = mine ? wsum[s] : 1.0;
    const double qsi = mine ? __dmul_rn(qsum[s], q_scale) : 0.0;
    const int shm_par = (int)(((uintptr_t)mass >> 3) & 1);

    // producer cursor (the tile whose copies are issued next), advanced identically by every thread
    int pf = blockIdx.y, pt0 = A0;
    unsigned issued = 0, consumed = 0;
    auto produce = [&]() {
        if (A1 <= A0 || pf >= nframes) return;
        const int cnt = A1 - pt0 < FLUX_TILE ? A1 - pt0 : FLUX_TILE;
        if (tid == 0) {
            const unsigned st = issued % FLUX_STAGES;
            flux_issue(vel, mass, n, pf, pt0, cnt, fs_a + st * (unsigned)(4 * FLUX_ROW * 8), bar_a + 8u * st);
        }
        ++issued;
        pt0 += cnt;
        if (pt0 >= A1) {
            pt0 = A0;
            pf += gridDim.y;
        }
    };
#pragma unroll 1
    for (int k = 0; k < FLUX_STAGES - 1; ++k) produce();

    int item_k = 0;
    for (int f = blockIdx.y; f < nframes; f += gridDim.y, ++item_k) {
        double acc = 0.0;
        const int64_t rowoff = (int64_t)f * 3 * n + (int64_t)comp * n;
        for (int t0 = A0; t0 < A1; t0 += FLUX_TILE) {
            const int cnt = A1 - t0 < FLUX_TILE ? A1 - t0 : FLUX_TILE;
            produce();   // the stage it overwrites was released by the barrier that ended the previous tile
            const unsigned st = consumed % FLUX_STAGES;
            mbar_wait(bar_a + 8u * st, (consumed / FLUX_STAGES) & 1u);
            ++consumed;
            const double *tv = fs + (size_t)st * 4 * FLUX_ROW + comp * FLUX_ROW + (((uintptr_t)(vel + rowoff + t0) >> 3) & 1);
            const double *tm = fs + (size_t)st * 4 * FLUX_ROW + 3 * FLUX_ROW + ((shm_par + t0) & 1);
            const int lo = (b > t0 ? b : t0) - t0, hi = (e < t0 + cnt ? e : t0 + cnt) - t0;
#pragma unroll 4
            for (int i = lo; i < hi; ++i) acc = __dadd_rn(acc, __dmul_rn(tv[i], tm[i]));
            if (t0 + cnt < A1) __syncthreads();   // tile consumed (the last tile's barrier is the one of the CTA sum)
        }
        double j = mine ? __dmul_rn(__dmul_rn(acc / ws, vel_scale), qsi) : 0.0;
        // deterministic CTA sum per component with a single barrier (fixed shuffle tree, fixed warp order)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) j = __dadd_rn(j, __shfl_xor_sync(0xffffffffu, j, o));
        double(*smk)[FLUX_B / 32] = sm[item_k & 1];
        if (lane == 0) smk[comp][wc] = j;
        __syncthreads();
        if (tid < 3) {
            double r = 0.0;
#pragma unroll
            for (int k = 0; k < FLUX_B / 32; ++k) r = __dadd_rn(r, smk[tid][k]);
            partial[((int64_t)f * nchunks + chunk) * 3 + tid] = r;
        }
    }
}

// out[c][g][frame0+f] = sum of chunk partials (fixed order)
__global__ void __launch_bounds__(96) k_flux_finish(const double *__restrict__ partial, int nchunks, int g, int ngroups,
                                                    double *__restrict__ out, int64_t out_stride, int64_t frame0)
{
    const int f = blockIdx.x;
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int k = lane; k < nchunks; k += 32) s = __dadd_rn(s, partial[((int64_t)f * nchunks + k) * 3 + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (lane == 0) out[((int64_t)c * ngroups + g) * out_stride + frame0 + f] = s;
}

// ---- OLS sums --------------------------------------------------------------------------------
__global__ void __launch_bounds__(RB) k_ols_sums(const double *__restrict__ t, const double *__restrict__ y, int64_t T,
                                                 int64_t i0, int64_t i1, double *__restrict__ out)
{
    __shared__ double sm[32];
    const int c = blockIdx.x;
    const double *yc = y + (int64_t)c * T;
    double a = 0, b = 0, d = 0;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const double tv = t[i], yv = yc[i];
        a = __dadd_rn(a, __dmul_rn(tv, tv));
        b = __dadd_rn(b, __dmul_rn(tv, yv));
        d = __dadd_rn(d, __dmul_rn(yv, yv));
    }
    a = block_sum(a, sm);
    b = block_sum(b, sm);
    d = block_sum(d, sm);
    if (threadIdx.x == 0) {
        out[c * 3 + 0] = a;
        out[c * 3 + 1] = b;
        out[c * 3 + 2] = d;
    }
}

// ---- number density along one axis (structural/number_density.py:30-154) -------------------------------
// Pass 1: min and max of the coordinate over the atoms whose key (type or altered id) equals the surface key, per frame
// (number_density.py:77-83).  Doubles are ordered through the usual monotone map to uint64 so that atomicMin/Max work.
__device__ __forceinline__ unsigned long long ord_enc(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_dec(unsigned long long u)
{
    const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}
__global__ void __launch_bounds__(RB) k_axis_minmax(const double *__restrict__ coord, const double *__restrict__ key, int64_t n,
                                                    double surface_key, unsigned long long *__restrict__ mm)
{
    const int f = blockIdx.y;
    const double *c = coord + (int64_t)f * n, *k = key + (int64_t)f * n;
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (k[i] == surface_key) {
            const unsigned long long e = ord_enc(c[i]);
            lo = e < lo ? e : lo;
            hi = e > hi ? e : hi;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo != ~0ull) atomicMin(&mm[f * 2 + 0], lo);
        if (hi != 0ull) atomicMax(&mm[f * 2 + 1], hi);
    }
}

// Pass 2: per target key, the 1-D histogram of ((x - min) - range) / bin_size truncated towards zero (positive
// dist_from_interface, atoms with x - min < dist) or of (x - min) / bin_size (negative, atoms with x - min > dist), in the
// reference's operation order (:84-110).  A negative index k counts in bin nbins + k, as numpy's negative indexing does
// (:99); indices outside [-nbins, nbins) -- an IndexError in the reference -- are dropped.
constexpr int AXD_MAX_TARGETS = 32;
struct AxisTargets {
    double key[AXD_MAX_TARGETS];
    int n;
};
__global__ void __launch_bounds__(RB) k_axis_hist(const double *__restrict__ coord, const double *__restrict__ key, int64_t n,
                                                  const unsigned long long *__restrict__ mm, AxisTargets tg, double dist,
                                                  double bin_size, int nbins, unsigned long long *__restrict__ out,
                                                  double *__restrict__ mm_out)
{
    const int f = blockIdx.y;
    const double *c = coord + (int64_t)f * n, *k = key + (int64_t)f * n;
    const bool have = mm[f * 2 + 0] != ~0ull;
    const double mn = have ? ord_dec(mm[f * 2 + 0]) : __longlong_as_double(0x7ff8000000000000ll);
    const double mx = have ? ord_dec(mm[f * 2 + 1]) : __longlong_as_double(0x7ff8000000000000ll);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        mm_out[f * 2 + 0] = mn;
        mm_out[f * 2 + 1] = mx;
    }
    if (!have) return;   // no surface atom in this frame: the reference's min() is NaN and every comparison fails
    const double range = __dsub_rn(mx, mn);
    unsigned long long *o = out + (int64_t)f * tg.n * nbins;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = __dsub_rn(c[i], mn);
        const bool sel = dist > 0.0 ? x < dist : x > dist;
        if (!sel) continue;
        const double b = dist > 0.0 ? __dsub_rn(x, range) : x;
        const double q = __ddiv_rn(b, bin_size);
        if (!(q > -2147483648.0 && q < 2147483648.0)) continue;
        int kb = (int)q;                      // truncation towards zero, as ndarray.astype(int)
        if (kb < 0) kb += nbins;
        if (kb < 0 || kb >= nbins) continue;
        const double ki = k[i];
        for (int t = 0; t < tg.n; ++t)
            if (ki == tg.key[t]) atomicAdd(&o[(int64_t)t * nbins + kb], 1ull);
    }
}

} // namespace

extern "C" {

int mdp_msd_single_origin(mdp_ctx *ctx, int nframes, int64_t n, const double *traj, const double *ref, double scale,
                          const int64_t *group_off, int ngroups, double *sums_out, double *per_atom_out, void *stream)
{
    MDP_REQUIRE(ctx && traj && ref && sums_out, "mdp_msd_single_origin: NULL argument");
    MDP_REQUIRE(nframes > 0 && n > 0 && ngroups >= 1, "mdp_msd_single_origin: sizes must be positive");
    MDP_REQUIRE(nframes <= 65535, "mdp_msd_single_origin: at most 65535 frames per call");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const int64_t per_chunk = RB * MSD_APT;
    const int max_chunks = (int)ceil_div<int64_t>(n, per_chunk) + 1;
    int rc = ctx->arena_reserve((size_t)nframes * max_chunks * 32 + 4096);
    if (rc) return rc;
    ctx->arena_reset();
    double *partial = (double *)ctx->arena_take((size_t)nframes * max_chunks * 32);
    const bool vec = (n % 2 == 0) && (((uintptr_t)traj | (uintptr_t)ref | (uintptr_t)per_atom_out) % 16 == 0);
    for (int g = 0; g < ngroups; ++g) {
        const int64_t a0 = group_off ? group_off[g] : 0, a1 = group_off ? group_off[g + 1] : n;
        MDP_REQUIRE(a0 >= 0 && a1 >= a0 && a1 <= n, "mdp_msd_single_origin: bad group range");
        if (a1 == a0) {
            MDP_CUDA(cudaMemset2DAsync(sums_out + g * 4, (size_t)ngroups * 32, 0, 32, nframes, st));
            continue;
        }
        const bool v = vec && (a0 % 2 == 0);
        const int nchunks = (int)ceil_div<int64_t>(a1 - a0, per_chunk);
        dim3 grid(nchunks, nframes);
        cudaEvent_t tk = ctx->timer_begin(2, st);
        if (v && per_atom_out)
            k_msd_single<true, true><<<grid, RB, 0, st>>>(traj, ref, n, a0, a1, scale, partial, nchunks, per_atom_out);
        else if (v)
            k_msd_single<true, false><<<grid, RB, 0, st>>>(traj, ref, n, a0, a1, scale, partial, nchunks, nullptr);
        else if (per_atom_out)
            k_msd_single<false, true><<<grid, RB, 0, st>>>(traj, ref, n, a0, a1, scale, partial, nchunks, per_atom_out);
        else
            k_msd_single<false, false><<<grid, RB, 0, st>>>(traj, ref, n, a0, a1, scale, partial, nchunks, nullptr);
        ctx->timer_end(tk, st);
        MDP_LAUNCHED(ctx);
        k_partial_sum<<<nframes, 128, 0, st>>>(partial, nchunks, 4, sums_out + g * 4, (int64_t)ngroups * 4);
        MDP_LAUNCHED(ctx);
    }
    return mdp_check_launch("k_msd_single");
}

int mdp_msd_interval(mdp_ctx *ctx, int nframes, int64_t n, const double *traj, double scale, int stride, double *out,
                     void *stream)
{
    MDP_REQUIRE(ctx && traj && out, "mdp_msd_interval: NULL argument");
    MDP_REQUIRE(nframes > 0 && n > 0 && stride > 0, "mdp_msd_interval: sizes must be positive");
    MDP_CUDA(cudaSetDevice(ctx->device));
    k_msd_interval<<<(unsigned)ceil_div<int64_t>(n, RB), RB, 0, (cudaStream_t)stream>>>(traj, nframes, n, scale, stride, out);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_msd_interval");
}

int mdp_msd_all_origins(mdp_ctx *ctx, int nframes, int64_t n, const double *traj, double scale, const int64_t *group_off,
                        int ngroups, int max_lag, double *sums_out, void *stream)
{
    MDP_REQUIRE(ctx && traj && sums_out, "mdp_msd_all_origins: NULL argument");
    MDP_REQUIRE(nframes > 0 && n > 0 && ngroups >= 1 && max_lag > 0 && max_lag <= nframes,
                "mdp_msd_all_origins: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const int nctas_max = ctx->sm_count * 2;
    int rc = ctx->arena_reserve((size_t)nctas_max * 3 * MW_WCAP * 8 + 4096);
    if (rc) return rc;
    ctx->arena_reset();
    double *partial = (double *)ctx->arena_take((size_t)nctas_max * 3 * MW_WCAP * 8);
    MDP_REQUIRE(partial != nullptr, "mdp_msd_all_origins: scratch arena exhausted");
    const size_t smem = (size_t)(MW_TT + MW_BROWS) * 32 * 8 + (size_t)3 * MW_WCAP * 8;
    MDP_CUDA(cudaFuncSetAttribute((const void *)k_msd_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int g = 0; g < ngroups; ++g) {
        const int64_t a0 = group_off ? group_off[g] : 0, a1 = group_off ? group_off[g + 1] : n;
        MDP_REQUIRE(a0 >= 0 && a1 >= a0 && a1 <= n, "mdp_msd_all_origins: bad group range");
        if (a1 == a0) continue;
        const int64_t nblk64 = ceil_div<int64_t>(a1 - a0, 32);
        MDP_REQUIRE(nblk64 < (1 << 26), "mdp_msd_all_origins: too many atoms in one group");
        const int nblk = (int)nblk64;
        // few atom blocks (molecule centres of mass): split the origins into chunks so that every SM has work
        int ntc = 1;
        const int tiles = (int)ceil_div<int64_t>(nframes, MW_TT);
        while ((int64_t)nblk * ntc < 4LL * nctas_max && ntc < tiles) ntc *= 2;
        if (ntc > tiles) ntc = tiles;
        const int tc_len = (int)ceil_div<int64_t>(tiles, ntc) * MW_TT;
        ntc = (int)ceil_div<int64_t>(nframes, tc_len);
        const int nctas = (int)std::min<int64_t>(nctas_max, (int64_t)nblk * ntc);
        for (int lag0 = 0; lag0 < max_lag; lag0 += MW_WCAP) {
            const int nl = std::min(MW_WCAP, max_lag - lag0);
            cudaEvent_t tk = ctx->timer_begin(5, st);
            k_msd_window<<<nctas, MW_WARPS * 32, smem, st>>>(traj, nframes, n, a0, a1, lag0, nl, nblk, ntc, tc_len, partial);
            ctx->timer_end(tk, st);
            MDP_LAUNCHED(ctx);
            k_msd_window_finish<<<(unsigned)ceil_div<int>(nl, 128), 128, 0, st>>>(partial, nctas, nl, lag0, scale * scale,
                                                                                sums_out + g * 4, (int64_t)ngroups * 4);
            MDP_LAUNCHED(ctx);
        }
    }
    return mdp_check_launch("k_msd_window");
}

int mdp_segment_com(mdp_ctx *ctx, int nframes, int ncomp, int64_t n, const double *attr, const double *w, int64_t nseg,
                    const int32_t *seg_off, double *out, double *wsum_out, const double *extra, double *extra_out,
                    void *stream)
{
    MDP_REQUIRE(ctx && attr && w && seg_off && out, "mdp_segment_com: NULL argument");
    MDP_REQUIRE(nframes > 0 && nframes <= 65535 && ncomp > 0 && n > 0 && nseg > 0, "mdp_segment_com: bad sizes");
    MDP_CUDA(cudaSetDevice(ctx->device));
    dim3 grid((unsigned)ceil_div<int64_t>(nseg, RB), nframes);
    k_segment_com<<<grid, RB, 0, (cudaStream_t)stream>>>(attr, ncomp, n, w, nseg, seg_off, out, wsum_out, extra, extra_out);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_segment_com");
}

int mdp_charge_flux(mdp_ctx *ctx, int nframes, int64_t n, const double *vel, const double *mass, const double *q,
                    int64_t nseg, const int32_t *seg_off, const int64_t *group_seg_off, int ngroups, double vel_scale,
                    double q_scale, double *out, int64_t out_stride, int64_t frame0, void *stream)
{
    MDP_REQUIRE(ctx && vel && mass && q && seg_off && group_seg_off && out, "mdp_charge_flux: NULL argument");
    MDP_REQUIRE(nframes > 0 && nframes <= 65535 && n > 0 && nseg > 0 && ngroups > 0, "mdp_charge_flux: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const int max_chunks = (int)ceil_div<int64_t>(nseg, FLUX_B) + ngroups;
    int rc = ctx->arena_reserve((size_t)nframes * max_chunks * 24 + (size_t)nseg * 16 + 8192);
    if (rc) return rc;
    ctx->arena_reset();
    double *partial = (double *)ctx->arena_take((size_t)nframes * max_chunks * 24);
    double *wsum = (double *)ctx->arena_take((size_t)nseg * 8);
    double *qsum = (double *)ctx->arena_take((size_t)nseg * 8);
    if (!partial || !wsum || !qsum) {
        mdp_set_error("internal: scratch arena exhausted (charge flux)");
        return MDP_ERR_OOM;
    }
    k_seg_static<<<(unsigned)ceil_div<int64_t>(nseg, RB), RB, 0, st>>>(mass, q, seg_off, nseg, wsum, qsum);
    MDP_LAUNCHED(ctx);
    for (int g = 0; g < ngroups; ++g) {
        const int64_t s0 = group_seg_off[g], s1 = group_seg_off[g + 1];
        MDP_REQUIRE(s0 >= 0 && s1 > s0 && s1 <= nseg, "mdp_charge_flux: bad molecule-type range");
        const int nchunks = (int)ceil_div<int64_t>(s1 - s0, FLUX_B);
        // frame lanes: fill the resident CTA slots (2 per SM) without exceeding them (a partial second wave would double the time)
        const int slots = ctx->sm_count * 2;
        const int lanes = std::max(1, std::min(nframes, slots / std::max(1, std::min(nchunks, slots))));
        dim3 grid((unsigned)nchunks, (unsigned)lanes);
        MDP_CUDA(cudaFuncSetAttribute((const void *)k_charge_flux, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLUX_SMEM));
        cudaEvent_t tk = ctx->timer_begin(4, st);
        k_charge_flux<<<grid, FLUX_T, FLUX_SMEM, st>>>(vel, n, mass, wsum, qsum, seg_off, s0, s1, nframes, vel_scale, q_scale, partial,
                                                       nchunks);
        ctx->timer_end(tk, st);
        MDP_LAUNCHED(ctx);
        k_flux_finish<<<nframes, 96, 0, st>>>(partial, nchunks, g, ngroups, out, out_stride, frame0);
        MDP_LAUNCHED(ctx);
    }
    return mdp_check_launch("k_charge_flux");
}

int mdp_ols_sums(mdp_ctx *ctx, int ncol, int64_t T, const double *t, const double *y, int64_t i0, int64_t i1, double *out,
                 void *stream)
{
    MDP_REQUIRE(ctx && t && y && out, "mdp_ols_sums: NULL argument");
    MDP_REQUIRE(ncol > 0 && T > 0 && i0 >= 0 && i1 <= T && i1 > i0, "mdp_ols_sums: bad sizes");
    MDP_CUDA(cudaSetDevice(ctx->device));
    k_ols_sums<<<ncol, RB, 0, (cudaStream_t)stream>>>(t, y, T, i0, i1, out);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_ols_sums");
}

int mdp_axis_density(mdp_ctx *ctx, int nframes, int64_t n, const double *coord, const double *key, double surface_key,
                     int ntargets, const double *target_keys, double dist_from_interface, double bin_size, int nbins,
                     uint64_t *counts_out, double *minmax_out, void *stream)
{
    MDP_REQUIRE(ctx && coord && key && target_keys && counts_out && minmax_out, "mdp_axis_density: NULL argument");
    MDP_REQUIRE(nframes > 0 && nframes <= 65535 && n > 0 && nbins > 0 && bin_size > 0.0 && dist_from_interface != 0.0,
                "mdp_axis_density: bad sizes");
    MDP_REQUIRE(ntargets > 0 && ntargets <= AXD_MAX_TARGETS, "mdp_axis_density: 1..%d target types per call", AXD_MAX_TARGETS);
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    int rc = ctx->arena_reserve((size_t)nframes * 16 + 4096);
    if (rc) return rc;
    ctx->arena_reset();
    unsigned long long *mm = (unsigned long long *)ctx->arena_take((size_t)nframes * 16);
    if (!mm) {
        mdp_set_error("internal: scratch arena exhausted (axis density)");
        return MDP_ERR_OOM;
    }
    // {min, max} in the ordered encoding: min starts at all ones, max at zero
    std::vector<unsigned long long> init((size_t)nframes * 2);
    for (int f = 0; f < nframes; ++f) {
        init[2 * f] = ~0ull;
        init[2 * f + 1] = 0ull;
    }
    MDP_CUDA(cudaMemcpyAsync(mm, init.data(), init.size() * 8, cudaMemcpyHostToDevice, st));
    MDP_CUDA(cudaMemsetAsync(counts_out, 0, (size_t)nframes * ntargets * nbins * 8, st));
    AxisTargets tg;
    tg.n = ntargets;
    for (int t = 0; t < ntargets; ++t) tg.key[t] = target_keys[t];
    dim3 grid((unsigned)std::min<int64_t>(ceil_div<int64_t>(n, RB), 4 * ctx->sm_count), nframes);
    k_axis_minmax<<<grid, RB, 0, st>>>(coord, key, n, surface_key, mm);
    MDP_LAUNCHED(ctx);
    k_axis_hist<<<grid, RB, 0, st>>>(coord, key, n, mm, tg, dist_from_interface, bin_size, nbins,
                                     (unsigned long long *)counts_out, minmax_out);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_axis_hist");
}

} // extern "C"
