// shell_grid.h -- neighbour search of a SMALL set A against a large set B through a cell grid over A (host/device shared
// body of k_shell_grid in shell.cu; exercised on the host by tests/native/shell_grid_host.cpp).
//
// The residence-time search (residence_time.py:100-104) asks, per frame, which of ~10^3 central atoms have which of ~10^5
// partners inside a 3 A shell.  The general pair engine sorts BOTH sets of every frame; here only A is binned (in shared
// memory) and B is streamed once, each B point probing the cells around it that can hold a partner (5.4 on average).
//
// Exactness: the grid is only a candidate filter; a candidate pair is accepted by the reference's own arithmetic
// (_calc_rsq, rdf_cn.py:46-56: d = a - b, one shift by -sign(d)*l when |d| > l/2, rsq = (dx*dx + dy*dy) + dz*dz, unfused).
// The filter is conservative for ANY coordinates: the reference's single shift makes |d'| <= r only if d is within r of
// 0, +l or -l, i.e. close modulo l, and the grid is periodic with period l; the cell width l/nc exceeds r by a factor
// (1 + 1e-9), far more than the rounding of the cell coordinate, so two points that close are always in the same or in
// adjacent cells (nc >= 3 keeps the three neighbours per axis distinct).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MDP_HD __host__ __device__ __forceinline__
#else
#ifndef MDP_HD
#define MDP_HD inline
#endif
#endif

constexpr int SG_NC_MAX = 16;                     // cells per axis at most: 4096 cells, 16 KB of shared-memory offsets

struct ShellGrid {
    double origin[3], inv_w[3], len[3];           // cell coordinate of x on axis k: (x - origin[k]) * inv_w[k]
    double rw[3];                                 // r / cell width * (1 + 1e-6) per axis, set by mdp_grid_set_radius
    int nc[3];
};

MDP_HD void mdp_grid_set_radius(ShellGrid &g, double r)
{
    for (int k = 0; k < 3; ++k) g.rw[k] = r * (1.0 + 1e-6) * g.inv_w[k];
}

// cells per axis for box length l and outer radius r: as many as fit with width >= r * (1 + 1e-9), at most SG_NC_MAX;
// < 3 means "this search is not for the grid" (the caller takes the general engine)
MDP_HD int mdp_grid_cells(double l, double r)
{
    if (!(l > 0.0) || !(r > 0.0)) return 0;
    const double q = l / (r * (1.0 + 1e-9));
    int n = q >= (double)SG_NC_MAX ? SG_NC_MAX : (int)q;
    return n;
}

// cell of x on one axis and the position inside it: t = (x - origin) * inv_w, cell = floor(t) mod nc, frac = t - floor(t).
// Wrapped coordinates (the usual case) take no modulo at all; anything else one integer remainder.  NaN / inf / absurd
// magnitudes fall into cell 0 with frac 0.5 (they never pass the exact test).
MDP_HD int mdp_grid_cell1f(double x, double origin, double inv_w, int nc, double &frac)
{
    const double t = (x - origin) * inv_w;
    const double fl = floor(t);
    frac = t - fl;
    if (!(fl > -2.0e9 && fl < 2.0e9)) {
        frac = 0.5;
        return 0;
    }
    int c = (int)fl;
    if ((unsigned)c >= (unsigned)nc) {
        c += c < 0 ? nc : -nc;                    // one period off (coordinates wrapped relative to another origin): no division
        if ((unsigned)c >= (unsigned)nc) {
            c %= nc;
            if (c < 0) c += nc;
        }
    }
    return c;
}

MDP_HD int mdp_grid_cell1(double x, double origin, double inv_w, int nc)
{
    double frac;
    return mdp_grid_cell1f(x, origin, inv_w, nc, frac);
}

MDP_HD int mdp_grid_cell(const ShellGrid &g, double x, double y, double z)
{
    const int cx = mdp_grid_cell1(x, g.origin[0], g.inv_w[0], g.nc[0]);
    const int cy = mdp_grid_cell1(y, g.origin[1], g.inv_w[1], g.nc[1]);
    const int cz = mdp_grid_cell1(z, g.origin[2], g.inv_w[2], g.nc[2]);
    return (cz * g.nc[1] + cy) * g.nc[0] + cx;
}

MDP_HD double mdp_mic1(double d, double l)       // rdf_cn.py:49-54
{
    const double h = l / 2;
    if (d > h) return d - l;
    if (d < -h) return d + l;
    return d;
}

// rsq of head a against b in the reference's operation order (unfused: compile without FMA contraction)
MDP_HD double mdp_rsq_ref(double ax, double ay, double az, double bx, double by, double bz, const double *len)
{
    const double dx = mdp_mic1(ax - bx, len[0]), dy = mdp_mic1(ay - by, len[1]), dz = mdp_mic1(az - bz, len[2]);
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
#else
    return (dx * dx + dy * dy) + dz * dz;
#endif
}

MDP_HD bool mdp_shell_accepts(double rsq, double rin2, double rout2, int shell_mode)
{
    return shell_mode ? (rsq > rin2 && rsq <= rout2) : (rsq < rout2);   // residence_time.py:102 / cluster_analysis.py:160
}

// Probe the grid with one B point: emit(ia) for every A point accepted.  The halo copy of the cell ranges (below) and the
// A points sorted by cell (sx, sy, sz, sidx = original index) describe the grid; ib is B's index (for exclude_same).
//
// Which neighbour cells can hold a partner is decided per axis from B's position INSIDE its cell: an A point in the cell
// below is at least frac * w away (w = cell width), one in the cell above at least (1 - frac) * w, so the lower neighbour
// is needed only when frac <= rw and the upper one only when frac >= 1 - rw, rw = r / w * (1 + 1e-6) (the margin is six
// orders of magnitude above the rounding of t and covers an A point that sits on a cell boundary and was rounded into the
// neighbour).  With cells two to three radii wide (2 000 ions in a 126 A box: w = 7.9 A, r = 3 A) that is 5.4 cells per
// point on average instead of 27.
// the cells of axis k that can hold a partner of coordinate b: first (may be -1 or nc: wrap it with mdp_grid_wrap1) and
// how many (1..3)
MDP_HD void mdp_shell_axis_range(const ShellGrid &g, int k, double b, int &first, int &count)
{
    double frac;
    const int c0 = mdp_grid_cell1f(b, g.origin[k], g.inv_w[k], g.nc[k], frac);
    const int lo = frac <= g.rw[k] ? -1 : 0, hi = frac >= 1.0 - g.rw[k] ? 1 : 0;
    first = c0 + lo;
    count = hi - lo + 1;
}

MDP_HD int mdp_grid_wrap1(int c, int nc)          // c in [-1, nc + 1] -> [0, nc)
{
    return c < 0 ? c + nc : (c >= nc ? c - nc : c);
}

// The grid is read through a HALO copy: (nc + 2) cells per axis, entry (x', y', z') = the slot range of cell
// (x' - 1, y' - 1, z' - 1) modulo nc, packed (count << 16 | first slot) (n_a <= 4096).  A point's up-to-27 cells are then
// a box of consecutive halo indices -- no per-cell wrap, one load per cell.
MDP_HD int mdp_halo_cells(const ShellGrid &g)
{
    return (g.nc[0] + 2) * (g.nc[1] + 2) * (g.nc[2] + 2);
}

MDP_HD uint32_t mdp_halo_entry(const ShellGrid &g, const int *cell_start, int idx)
{
    const int hx = g.nc[0] + 2, hy = g.nc[1] + 2;
    const int x = idx % hx, y = (idx / hx) % hy, z = idx / (hx * hy);
    const int cx = mdp_grid_wrap1(x - 1, g.nc[0]), cy = mdp_grid_wrap1(y - 1, g.nc[1]), cz = mdp_grid_wrap1(z - 1, g.nc[2]);
    const int c = (cz * g.nc[1] + cy) * g.nc[0] + cx;
    return ((uint32_t)(cell_start[c + 1] - cell_start[c]) << 16) | (uint32_t)cell_start[c];
}

// the walk of one B point over its cells: tot halo cells, visited in x-fastest order
struct ShellWalk {
    int cell;          // halo index of the cell to visit now
    int n0, n1, tot;   // cells along x, along y, in total
    int dx, dy;        // index step on an x carry (after the +1) and on a y carry
    int ix, iy;
};

MDP_HD ShellWalk mdp_shell_walk_begin(const ShellGrid &g, double bx, double by, double bz)
{
    int fx, fy, fz, n2;
    ShellWalk w;
    mdp_shell_axis_range(g, 0, bx, fx, w.n0);
    mdp_shell_axis_range(g, 1, by, fy, w.n1);
    mdp_shell_axis_range(g, 2, bz, fz, n2);
    const int hx = g.nc[0] + 2, hy = g.nc[1] + 2;
    w.cell = ((fz + 1) * hy + (fy + 1)) * hx + (fx + 1);
    w.tot = w.n0 * w.n1 * n2;
    w.dx = hx - w.n0;
    w.dy = hx * hy - w.n1 * hx;
    w.ix = w.iy = 0;
    return w;
}

MDP_HD void mdp_shell_walk_next(ShellWalk &w)
{
    w.cell += 1;
    if (++w.ix == w.n0) {
        w.ix = 0;
        w.cell += w.dx;
        if (++w.iy == w.n1) {
            w.iy = 0;
            w.cell += w.dy;
        }
    }
}

// cand(k) for every slot k of the sorted A arrays that lies in a cell B's point can have a partner in
template <class Cand>
MDP_HD void mdp_shell_candidates(const ShellGrid &g, const uint32_t *halo, double bx, double by, double bz, const Cand cand)
{
    ShellWalk w = mdp_shell_walk_begin(g, bx, by, bz);
    for (int c = 0; c < w.tot; ++c) {
        const uint32_t e = halo[w.cell];
        const int s = (int)(e & 0xffffu), cnt = (int)(e >> 16);
        for (int k = s; k < s + cnt; ++k) cand(k);
        mdp_shell_walk_next(w);
    }
}

// the exact test of one candidate: the reference's own arithmetic decides
MDP_HD bool mdp_shell_pair_ok(const ShellGrid &g, double ax, double ay, double az, int ia, double bx, double by, double bz, int ib,
                              double rin2, double rout2, int shell_mode, int exclude_same)
{
    const double rsq = mdp_rsq_ref(ax, ay, az, bx, by, bz, g.len);
    return mdp_shell_accepts(rsq, rin2, rout2, shell_mode) && !(exclude_same && ia == ib);
}

template <class Emit>
struct ShellProbeCand {
    const ShellGrid &g;
    const double *sx, *sy, *sz;
    const int *sidx;
    double bx, by, bz, rin2, rout2;
    int ib, shell_mode, exclude_same;
    const Emit &emit;
    MDP_HD void operator()(int k) const
    {
        if (mdp_shell_pair_ok(g, sx[k], sy[k], sz[k], sidx[k], bx, by, bz, ib, rin2, rout2, shell_mode, exclude_same)) emit(sidx[k]);
    }
};

template <class Emit>
MDP_HD void mdp_shell_probe(const ShellGrid &g, const uint32_t *halo, const double *sx, const double *sy, const double *sz,
                            const int *sidx, double bx, double by, double bz, int ib, double rin2, double rout2, int shell_mode,
                            int exclude_same, const Emit emit)
{
    mdp_shell_candidates(g, halo, bx, by, bz,
                         ShellProbeCand<Emit>{g, sx, sy, sz, sidx, bx, by, bz, rin2, rout2, ib, shell_mode, exclude_same, emit});
}
