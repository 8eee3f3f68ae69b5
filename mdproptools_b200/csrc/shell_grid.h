// shell_grid.h -- neighbour search of a SMALL set A against a large set B through a cell grid over A (host/device shared
// body of k_shell_grid in shell.cu; exercised on the host by tests/native/shell_grid_host.cpp).
//
// The residence-time search (residence_time.py:100-104) asks, per frame, which of ~10^3 central atoms have which of ~10^5
// partners inside a 3 A shell.  The general pair engine sorts BOTH sets of every frame; here only A is binned (in shared
// memory) and B is streamed once, each B point probing the 27 cells around it.
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
    int nc[3];
};

// cells per axis for box length l and outer radius r: as many as fit with width >= r * (1 + 1e-9), at most SG_NC_MAX;
// < 3 means "this search is not for the grid" (the caller takes the general engine)
MDP_HD int mdp_grid_cells(double l, double r)
{
    if (!(l > 0.0) || !(r > 0.0)) return 0;
    const double q = l / (r * (1.0 + 1e-9));
    int n = q >= (double)SG_NC_MAX ? SG_NC_MAX : (int)q;
    return n;
}

MDP_HD int mdp_grid_cell1(double x, double origin, double inv_w, int nc)
{
    const double t = floor((x - origin) * inv_w);
    // true modulo; |t| stays far below 2^53 for any sane coordinate, NaN / inf fall into cell 0 (they never pass the
    // exact test)
    double m = t - floor(t / (double)nc) * (double)nc;
    int c = (m >= 0.0 && m < (double)nc) ? (int)m : 0;
    return c;
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

// Probe the grid with one B point: emit(ia) for every A point accepted.  cell_start[ncell + 1] and the A points sorted by
// cell (sx, sy, sz, sidx = original index) describe the grid; ib is B's index (for exclude_same).
template <class Emit>
MDP_HD void mdp_shell_probe(const ShellGrid &g, const int *cell_start, const double *sx, const double *sy, const double *sz,
                            const int *sidx, double bx, double by, double bz, int ib, double rin2, double rout2, int shell_mode,
                            int exclude_same, const Emit emit)
{
    const int cx = mdp_grid_cell1(bx, g.origin[0], g.inv_w[0], g.nc[0]);
    const int cy = mdp_grid_cell1(by, g.origin[1], g.inv_w[1], g.nc[1]);
    const int cz = mdp_grid_cell1(bz, g.origin[2], g.inv_w[2], g.nc[2]);
    for (int oz = -1; oz <= 1; ++oz) {
        int z = cz + oz;
        z = z < 0 ? z + g.nc[2] : (z >= g.nc[2] ? z - g.nc[2] : z);
        for (int oy = -1; oy <= 1; ++oy) {
            int y = cy + oy;
            y = y < 0 ? y + g.nc[1] : (y >= g.nc[1] ? y - g.nc[1] : y);
            for (int ox = -1; ox <= 1; ++ox) {
                int x = cx + ox;
                x = x < 0 ? x + g.nc[0] : (x >= g.nc[0] ? x - g.nc[0] : x);
                const int c = (z * g.nc[1] + y) * g.nc[0] + x;
                for (int k = cell_start[c]; k < cell_start[c + 1]; ++k) {
                    const double rsq = mdp_rsq_ref(sx[k], sy[k], sz[k], bx, by, bz, g.len);
                    if (mdp_shell_accepts(rsq, rin2, rout2, shell_mode) && !(exclude_same && sidx[k] == ib)) emit(sidx[k]);
                }
            }
        }
    }
}
