// epilogue.cu -- device-side epilogues of the cutoff searches: what the reference does in pandas per central atom after
// _calc_rsq has found the neighbours (hydration_number.py:13-32, 60-75; cluster_analysis.py:150-182).
//
// Input is always the neighbour list of mdp_pair_list / mdp_shell_search: int32 (frame, ia, ib) entries in no particular
// order.  Three building blocks, all on the caller's stream:
//
//   mdp_list_group       entries -> segments: one segment per (frame, ia), in (frame, ia) order, the entries of a segment
//                        sorted by a caller-supplied 32-bit key (default: ib).  Counting sort by segment (count, scan,
//                        place) + a per-segment shell sort of the handful of entries a cutoff sphere holds.  Returns the
//                        segment offsets and, per output slot, the index of the entry that landed there (a permutation), so
//                        any per-entry payload can be gathered in the reference's row order.
//   mdp_hydration_count  per grouped entry the cosine between the minimum-image cation->O displacement and the water
//                        bisector (H1 + H2) - 2 O, in the reference's fp64 expression order (numpy: products, (a+b)+c sums,
//                        sqrt, one division), and per (frame, cation) the two counters the hydration factor needs:
//                        waters in range, waters with cos < threshold (-0.72, hydration_number.py:32).
//   mdp_cluster_members  per (frame, central atom) the sorted, duplicate-free list of molecules that own a neighbour atom
//                        and pass the force filter min(sum fx, sum fy, sum fz) * c < max_force (cluster_analysis.py:169-182;
//                        per-molecule sums in atom order).
#include <float.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int EB = 256;

__global__ void __launch_bounds__(EB) k_lg_count(const int32_t *__restrict__ list, int64_t m, int64_t n_a, uint32_t *__restrict__ cnt)
{
    const int64_t e = (int64_t)blockIdx.x * EB + threadIdx.x;
    if (e >= m) return;
    atomicAdd(&cnt[(int64_t)list[e * 3] * n_a + list[e * 3 + 1]], 1u);
}

// exclusive scan of nseg counters in three coalesced passes: (1) sum of every tile of SCAN_TILE counters, (2) scan of the
// tile sums by one CTA, (3) scan inside every tile + its base; off[nseg] = total; the counters are zeroed for the placement pass
constexpr int SCAN_T = 256, SCAN_PER = 8, SCAN_TILE = SCAN_T * SCAN_PER;

__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long *ws, unsigned long long &total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        const unsigned long long x = lane < (int)(blockDim.x >> 5) ? ws[lane] : 0ull;
        unsigned long long ix = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, ix, o);
            if (lane >= o) ix += u;
        }
        ws[lane] = ix - x;
        if (lane == 31) ws[32] = ix;
    }
    __syncthreads();
    total = ws[32];
    return ws[w] + inc - v;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tile_sums(const uint32_t *__restrict__ cnt, int64_t nseg, unsigned long long *__restrict__ tsum)
{
    __shared__ unsigned long long ws[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_T + threadIdx.x;
        if (i < nseg) s += cnt[i];
    }
    unsigned long long total;
    block_excl_scan(s, ws, total);
    if (threadIdx.x == 0) tsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_tiles(unsigned long long *__restrict__ tsum, int64_t ntiles, int64_t *__restrict__ off_total)
{
    __shared__ unsigned long long ws[33];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    for (int64_t b = 0; b < ntiles; b += 1024) {
        const int64_t i = b + threadIdx.x;
        const unsigned long long v = i < ntiles ? tsum[i] : 0ull;
        unsigned long long total;
        const unsigned long long ex = block_excl_scan(v, ws, total);
        if (i < ntiles) tsum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *off_total = (int64_t)carry;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_apply(uint32_t *__restrict__ cnt, int64_t nseg, const unsigned long long *__restrict__ tsum,
                                                       int64_t *__restrict__ off)
{
    __shared__ unsigned long long ws[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER;   // a thread owns SCAN_PER consecutive counters
    uint32_t v[SCAN_PER];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        v[k] = base + k < nseg ? cnt[base + k] : 0u;
        s += v[k];
    }
    unsigned long long total;
    unsigned long long run = tsum[blockIdx.x] + block_excl_scan(s, ws, total);
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        if (base + k < nseg) {
            off[base + k] = (int64_t)run;
            cnt[base + k] = 0u;
        }
        run += v[k];
    }
}

__global__ void __launch_bounds__(EB) k_lg_place(const int32_t *__restrict__ list, const uint32_t *__restrict__ key, int64_t m,
                                                 int64_t n_a, const int64_t *__restrict__ off, uint32_t *__restrict__ fill,
                                                 uint32_t *__restrict__ okey, int64_t *__restrict__ perm)
{
    const int64_t e = (int64_t)blockIdx.x * EB + threadIdx.x;
    if (e >= m) return;
    const int64_t seg = (int64_t)list[e * 3] * n_a + list[e * 3 + 1];
    const int64_t slot = off[seg] + atomicAdd(&fill[seg], 1u);
    okey[slot] = key ? key[e] : (uint32_t)list[e * 3 + 2];
    perm[slot] = e;
}

// one thread per segment: shell sort of (key, perm) by key, ties by entry index (deterministic whatever the placement order)
__global__ void __launch_bounds__(EB) k_lg_sort(const int64_t *__restrict__ off, int64_t nseg, uint32_t *__restrict__ okey,
                                                int64_t *__restrict__ perm)
{
    const int64_t s = (int64_t)blockIdx.x * EB + threadIdx.x;
    if (s >= nseg) return;
    const int64_t b = off[s];
    const int64_t n = off[s + 1] - b;
    uint32_t *k = okey + b;
    int64_t *p = perm + b;
    int64_t gap = 1;
    while (gap < n / 3) gap = 3 * gap + 1;
    for (; gap >= 1; gap /= 3) {
        for (int64_t i = gap; i < n; ++i) {
            const uint32_t kv = k[i];
            const int64_t pv = p[i];
            int64_t j = i;
            while (j >= gap && (k[j - gap] > kv || (k[j - gap] == kv && p[j - gap] > pv))) {
                k[j] = k[j - gap];
                p[j] = p[j - gap];
                j -= gap;
            }
            k[j] = kv;
            p[j] = pv;
        }
    }
}

// cnt: scratch of nseg counters followed (256-byte aligned) by ceil(nseg / SCAN_TILE) tile sums: list_group_scratch(nseg) bytes
static size_t list_group_scratch(int64_t nseg) { return align256((size_t)nseg * 4) + align256((size_t)ceil_div<int64_t>(nseg, SCAN_TILE) * 8) + 256; }

static int list_group(mdp_ctx *ctx, int nframes, int64_t n_a, int64_t m, const int32_t *list, const uint32_t *key, int64_t *seg_off,
                      uint32_t *okey, int64_t *perm, uint32_t *cnt, cudaStream_t st)
{
    const int64_t nseg = (int64_t)nframes * n_a;
    unsigned long long *tsum = (unsigned long long *)((char *)cnt + align256((size_t)nseg * 4));
    MDP_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nseg * 4, st));
    if (m > 0) {
        k_lg_count<<<(unsigned)ceil_div<int64_t>(m, EB), EB, 0, st>>>(list, m, n_a, cnt);
        MDP_LAUNCHED(ctx);
    }
    const int64_t ntiles = ceil_div<int64_t>(nseg, SCAN_TILE);
    k_scan_tile_sums<<<(unsigned)ntiles, SCAN_T, 0, st>>>(cnt, nseg, tsum);
    MDP_LAUNCHED(ctx);
    k_scan_tiles<<<1, 1024, 0, st>>>(tsum, ntiles, seg_off + nseg);
    MDP_LAUNCHED(ctx);
    k_scan_apply<<<(unsigned)ntiles, SCAN_T, 0, st>>>(cnt, nseg, tsum, seg_off);
    MDP_LAUNCHED(ctx);
    if (m > 0) {
        k_lg_place<<<(unsigned)ceil_div<int64_t>(m, EB), EB, 0, st>>>(list, key, m, n_a, seg_off, cnt, okey, perm);
        MDP_LAUNCHED(ctx);
        k_lg_sort<<<(unsigned)ceil_div<int64_t>(nseg, EB), EB, 0, st>>>(seg_off, nseg, okey, perm);
        MDP_LAUNCHED(ctx);
    }
    return mdp_check_launch("list_group");
}

// ---- hydration ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mic_ref(double d, double l)   // rdf_cn.py:49-54: d - sign(d) * l when |d| > l/2
{
    const double h = l / 2;
    if (d > h) return __dsub_rn(d, l);
    if (d < -h) return __dadd_rn(d, l);
    return d;
}

// one thread per grouped entry: cosine of (cation - O, minimum image) and (H1 + H2) - 2 O (raw coordinates)
__global__ void __launch_bounds__(EB) k_hyd_cos(const int32_t *__restrict__ list, const int64_t *__restrict__ perm, int64_t m,
                                                const double *__restrict__ cat, int64_t n_a, const double *__restrict__ ox,
                                                const double *__restrict__ h1, const double *__restrict__ h2, int64_t n_w,
                                                const double *__restrict__ box, double *__restrict__ cos_out)
{
    const int64_t k = (int64_t)blockIdx.x * EB + threadIdx.x;
    if (k >= m) return;
    const int64_t e = perm[k];
    const int f = list[e * 3];
    const int64_t ia = list[e * 3 + 1], ib = list[e * 3 + 2];
    const double *c = cat + (int64_t)f * 3 * n_a, *o = ox + (int64_t)f * 3 * n_w, *p1 = h1 + (int64_t)f * 3 * n_w,
                 *p2 = h2 + (int64_t)f * 3 * n_w;
    double d[3], v[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double oa = o[a * n_w + ib];
        d[a] = mic_ref(__dsub_rn(c[a * n_a + ia], oa), box[f * 3 + a]);
        v[a] = __dsub_rn(__dadd_rn(p1[a * n_w + ib], p2[a * n_w + ib]), __dmul_rn(2.0, oa));
    }
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(d[0], v[0]), __dmul_rn(d[1], v[1])), __dmul_rn(d[2], v[2]));
    const double n1 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2])));
    const double n2 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2])));
    cos_out[k] = dot / __dmul_rn(n1, n2);
}

// one thread per (frame, cation): waters in range, waters with cos < threshold
__global__ void __launch_bounds__(EB) k_hyd_counts(const int64_t *__restrict__ off, int64_t nseg, const double *__restrict__ cosv,
                                                   double thr, int32_t *__restrict__ counts)
{
    const int64_t s = (int64_t)blockIdx.x * EB + threadIdx.x;
    if (s >= nseg) return;
    int n = 0;
    for (int64_t k = off[s]; k < off[s + 1]; ++k) n += cosv[k] < thr ? 1 : 0;
    counts[s * 2] = (int32_t)(off[s + 1] - off[s]);
    counts[s * 2 + 1] = n;
}

// ---- clusters ----------------------------------------------------------------------------------------------------------
// flag[f][mol] = min(sum fx, sum fy, sum fz) * c < max_force  (sums over the molecule's atoms in atom order)
__global__ void __launch_bounds__(EB) k_mol_force_flag(const double *__restrict__ force, int64_t n, const int32_t *__restrict__ seg_off,
                                                       int64_t nmol, double c, double max_force, uint8_t *__restrict__ flag)
{
    const int64_t mol = (int64_t)blockIdx.x * EB + threadIdx.x;
    const int f = blockIdx.y;
    if (mol >= nmol) return;
    const double *fr = force + (int64_t)f * 3 * n;
    double mn = DBL_MAX;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double s = 0.0;
        for (int i = seg_off[mol]; i < seg_off[mol + 1]; ++i) s = __dadd_rn(s, fr[a * n + i]);
        mn = fmin(mn, s);
    }
    flag[(int64_t)f * nmol + mol] = __dmul_rn(mn, c) < max_force ? 1 : 0;
}

// key of an entry = molecule of the neighbour atom, or the sentinel when the molecule fails the force filter
__global__ void __launch_bounds__(EB) k_cl_key(const int32_t *__restrict__ list, int64_t m, const int32_t *__restrict__ mol_of_atom,
                                               const uint8_t *__restrict__ flag, int64_t nmol, uint32_t *__restrict__ key)
{
    const int64_t e = (int64_t)blockIdx.x * EB + threadIdx.x;
    if (e >= m) return;
    const int mol = mol_of_atom[list[e * 3 + 2]];
    key[e] = flag[(int64_t)list[e * 3] * nmol + mol] ? (uint32_t)mol : 0xffffffffu;
}

// one thread per segment: the sorted keys -> unique keys at the front of the segment (sentinels dropped), their number
__global__ void __launch_bounds__(EB) k_seg_unique(const int64_t *__restrict__ off, int64_t nseg, uint32_t *__restrict__ okey,
                                                   int32_t *__restrict__ ucount)
{
    const int64_t s = (int64_t)blockIdx.x * EB + threadIdx.x;
    if (s >= nseg) return;
    const int64_t b = off[s], e = off[s + 1];
    int64_t w = b;
    for (int64_t k = b; k < e; ++k) {
        const uint32_t v = okey[k];
        if (v == 0xffffffffu) break;                    // sorted: sentinels are last
        if (w == b || okey[w - 1] != v) okey[w++] = v;
    }
    ucount[s] = (int32_t)(w - b);
}

// ---- unique (ia, ib) pairs of a neighbour list, sorted by key = ia * n_b + ib ------------------------------------------------
// (what the residence-time correlation needs before it can build one time bitmask per ever-neighbour pair,
// residence_time.py:100-111).  One bit per possible pair (n_a * n_b / 8 bytes: 16 MB for 2 000 x 62 666), set by the entries;
// the set bits are then enumerated in order by a popcount scan -- no sort.
__global__ void __launch_bounds__(EB) k_pk_set(const int32_t *__restrict__ list, int64_t m, int64_t n_b, uint32_t *__restrict__ bits)
{
    const int64_t e = (int64_t)blockIdx.x * EB + threadIdx.x;
    if (e >= m) return;
    const int64_t key = (int64_t)list[e * 3 + 1] * n_b + list[e * 3 + 2];
    atomicOr(&bits[key >> 5], 1u << (key & 31));
}

__global__ void __launch_bounds__(SCAN_T) k_pk_tile_sums(const uint32_t *__restrict__ bits, int64_t nwords, unsigned long long *__restrict__ tsum)
{
    __shared__ unsigned long long ws[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_T + threadIdx.x;
        if (i < nwords) s += __popc(bits[i]);
    }
    unsigned long long total;
    block_excl_scan(s, ws, total);
    if (threadIdx.x == 0) tsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_T) k_pk_emit(const uint32_t *__restrict__ bits, int64_t nwords, const unsigned long long *__restrict__ tsum,
                                                    int64_t *__restrict__ keys, int64_t cap)
{
    __shared__ unsigned long long ws[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER;
    uint32_t v[SCAN_PER];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        v[k] = base + k < nwords ? bits[base + k] : 0u;
        s += __popc(v[k]);
    }
    unsigned long long total;
    unsigned long long run = tsum[blockIdx.x] + block_excl_scan(s, ws, total);
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        uint32_t w = v[k];
        while (w) {
            const int b = __ffs(w) - 1;
            w &= w - 1;
            if ((int64_t)run < cap) keys[run] = (base + k) * 32 + b;
            ++run;
        }
    }
}

} // namespace

extern "C" {

int mdp_unique_pair_keys(mdp_ctx *ctx, int64_t m, const int32_t *list, int64_t n_a, int64_t n_b, int64_t *keys_out, int64_t capacity,
                         int64_t *count_out, void *stream)
{
    MDP_REQUIRE(ctx && count_out && (m == 0 || (list && keys_out)), "mdp_unique_pair_keys: NULL argument");
    MDP_REQUIRE(m >= 0 && n_a > 0 && n_b > 0 && n_a * n_b <= ((int64_t)1 << 36), "mdp_unique_pair_keys: bad sizes (n_a * n_b <= 2^36)");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const int64_t nwords = ceil_div<int64_t>(n_a * n_b, 32);
    const int64_t ntiles = ceil_div<int64_t>(nwords, SCAN_TILE);
    int rc = ctx->arena_reserve(align256((size_t)nwords * 4) + align256((size_t)ntiles * 8) + 4096);
    if (rc) return rc;
    ctx->arena_reset();
    uint32_t *bits = (uint32_t *)ctx->arena_take((size_t)nwords * 4);
    unsigned long long *tsum = (unsigned long long *)ctx->arena_take((size_t)ntiles * 8);
    if (!bits || !tsum) {
        mdp_set_error("internal: scratch arena exhausted (unique pair keys)");
        return MDP_ERR_OOM;
    }
    MDP_CUDA(cudaMemsetAsync(bits, 0, (size_t)nwords * 4, st));
    if (m > 0) {
        k_pk_set<<<(unsigned)ceil_div<int64_t>(m, EB), EB, 0, st>>>(list, m, n_b, bits);
        MDP_LAUNCHED(ctx);
    }
    k_pk_tile_sums<<<(unsigned)ntiles, SCAN_T, 0, st>>>(bits, nwords, tsum);
    MDP_LAUNCHED(ctx);
    k_scan_tiles<<<1, 1024, 0, st>>>(tsum, ntiles, count_out);
    MDP_LAUNCHED(ctx);
    k_pk_emit<<<(unsigned)ntiles, SCAN_T, 0, st>>>(bits, nwords, tsum, keys_out, capacity);
    MDP_LAUNCHED(ctx);
    return mdp_check_launch("k_pk_emit");
}

int mdp_list_group(mdp_ctx *ctx, int nframes, int64_t n_a, int64_t m, const int32_t *list, const uint32_t *key, int64_t *seg_off,
                   uint32_t *key_out, int64_t *perm_out, void *stream)
{
    MDP_REQUIRE(ctx && seg_off && (m == 0 || (list && key_out && perm_out)), "mdp_list_group: NULL argument");
    MDP_REQUIRE(nframes > 0 && n_a > 0 && m >= 0, "mdp_list_group: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const int64_t nseg = (int64_t)nframes * n_a;
    int rc = ctx->arena_reserve(list_group_scratch(nseg) + 4096);
    if (rc) return rc;
    ctx->arena_reset();
    uint32_t *cnt = (uint32_t *)ctx->arena_take(list_group_scratch(nseg));
    if (!cnt) {
        mdp_set_error("internal: scratch arena exhausted (list group)");
        return MDP_ERR_OOM;
    }
    return list_group(ctx, nframes, n_a, m, list, key, seg_off, key_out, perm_out, cnt, st);
}

int mdp_hydration_count(mdp_ctx *ctx, int nframes, int64_t n_cat, const double *xyz_cat, int64_t n_wat, const double *xyz_o,
                        const double *xyz_h1, const double *xyz_h2, const double *box, int64_t m, const int32_t *list, double threshold,
                        double *cos_out, int64_t *seg_off, int32_t *counts, void *stream)
{
    MDP_REQUIRE(ctx && xyz_cat && xyz_o && xyz_h1 && xyz_h2 && box && seg_off && counts && (m == 0 || (list && cos_out)),
                "mdp_hydration_count: NULL argument");
    MDP_REQUIRE(nframes > 0 && n_cat > 0 && n_wat > 0 && m >= 0, "mdp_hydration_count: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const int64_t nseg = (int64_t)nframes * n_cat;
    int rc = ctx->arena_reserve(list_group_scratch(nseg) + align256((size_t)m * 4) + align256((size_t)m * 8) +
                                align256((size_t)nframes * 24) + 8192);
    if (rc) return rc;
    ctx->arena_reset();
    uint32_t *cnt = (uint32_t *)ctx->arena_take(list_group_scratch(nseg));
    uint32_t *okey = (uint32_t *)ctx->arena_take((size_t)std::max<int64_t>(m, 1) * 4);
    int64_t *perm = (int64_t *)ctx->arena_take((size_t)std::max<int64_t>(m, 1) * 8);
    double *d_box = (double *)ctx->arena_take((size_t)nframes * 24);
    if (!cnt || !okey || !perm || !d_box) {
        mdp_set_error("internal: scratch arena exhausted (hydration)");
        return MDP_ERR_OOM;
    }
    MDP_CUDA(cudaMemcpyAsync(d_box, box, (size_t)nframes * 24, cudaMemcpyHostToDevice, st));
    cudaEvent_t tk = ctx->timer_begin(8, st);
    rc = list_group(ctx, nframes, n_cat, m, list, nullptr, seg_off, okey, perm, cnt, st);
    if (rc) return rc;
    if (m > 0) {
        k_hyd_cos<<<(unsigned)ceil_div<int64_t>(m, EB), EB, 0, st>>>(list, perm, m, xyz_cat, n_cat, xyz_o, xyz_h1, xyz_h2, n_wat, d_box,
                                                                     cos_out);
        MDP_LAUNCHED(ctx);
    }
    k_hyd_counts<<<(unsigned)ceil_div<int64_t>(nseg, EB), EB, 0, st>>>(seg_off, nseg, cos_out, threshold, counts);
    MDP_LAUNCHED(ctx);
    ctx->timer_end(tk, st);
    return mdp_check_launch("k_hyd_counts");
}

int mdp_cluster_members(mdp_ctx *ctx, int nframes, int64_t n_central, int64_t n_atoms, const double *force, int64_t n_mol,
                        const int32_t *mol_seg_off, const int32_t *mol_of_atom, double force_constant, double max_force, int64_t m,
                        const int32_t *list, int64_t *seg_off, uint32_t *mol_out, int32_t *mol_count, void *stream)
{
    MDP_REQUIRE(ctx && force && mol_seg_off && mol_of_atom && seg_off && mol_count && (m == 0 || (list && mol_out)),
                "mdp_cluster_members: NULL argument");
    MDP_REQUIRE(nframes > 0 && nframes <= 65535 && n_central > 0 && n_atoms > 0 && n_mol > 0 && m >= 0, "mdp_cluster_members: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    const int64_t nseg = (int64_t)nframes * n_central;
    int rc = ctx->arena_reserve(list_group_scratch(nseg) + align256((size_t)m * 4) + align256((size_t)m * 8) +
                                align256((size_t)nframes * n_mol) + 8192);
    if (rc) return rc;
    ctx->arena_reset();
    uint32_t *cnt = (uint32_t *)ctx->arena_take(list_group_scratch(nseg));
    uint32_t *key = (uint32_t *)ctx->arena_take((size_t)std::max<int64_t>(m, 1) * 4);
    int64_t *perm = (int64_t *)ctx->arena_take((size_t)std::max<int64_t>(m, 1) * 8);
    uint8_t *flag = (uint8_t *)ctx->arena_take((size_t)nframes * n_mol);
    if (!cnt || !key || !perm || !flag) {
        mdp_set_error("internal: scratch arena exhausted (cluster members)");
        return MDP_ERR_OOM;
    }
    cudaEvent_t tk = ctx->timer_begin(9, st);
    dim3 gf((unsigned)ceil_div<int64_t>(n_mol, EB), (unsigned)nframes);
    k_mol_force_flag<<<gf, EB, 0, st>>>(force, n_atoms, mol_seg_off, n_mol, force_constant, max_force, flag);
    MDP_LAUNCHED(ctx);
    if (m > 0) {
        k_cl_key<<<(unsigned)ceil_div<int64_t>(m, EB), EB, 0, st>>>(list, m, mol_of_atom, flag, n_mol, key);
        MDP_LAUNCHED(ctx);
    }
    rc = list_group(ctx, nframes, n_central, m, list, key, seg_off, mol_out, perm, cnt, st);
    if (rc) return rc;
    k_seg_unique<<<(unsigned)ceil_div<int64_t>(nseg, EB), EB, 0, st>>>(seg_off, nseg, mol_out, mol_count);
    MDP_LAUNCHED(ctx);
    ctx->timer_end(tk, st);
    return mdp_check_launch("k_seg_unique");
}

} // extern "C"
