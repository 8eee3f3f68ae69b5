// Host emulation of k_shell_grid (mdproptools_b200/csrc/shell.cu): the grid over A is built with the same cell function,
// the halo copy with the same mdp_halo_entry, and every B point probes it with the same walk (mdp_shell_probe).  TEST INFRASTRUCTURE ONLY.  Built with -ffp-contract=off.
#include <stdint.h>
#include <vector>

#include "../../mdproptools_b200/csrc/shell_grid.h"

struct Mark {
    unsigned char *row;   // out[ia * nb + ib]
    long long nb;
    int ib;
    void operator()(int ia) const { row[(long long)ia * nb + ib] += 1; }
};

// out [na][nb] bytes = how many times the pair was emitted (must be 0 or 1); returns 0, or -1 when the grid does not apply
extern "C" int emulate_shell_grid(const double *xa, const double *ya, const double *za, int na, const double *xb, const double *yb,
                                  const double *zb, int nb, const double *len, double rin2, double rout2, int shell_mode,
                                  int exclude_same, unsigned char *out)
{
    ShellGrid g;
    const double r = sqrt(rout2);
    for (int k = 0; k < 3; ++k) {
        g.nc[k] = mdp_grid_cells(len[k], r);
        if (g.nc[k] < 3) return -1;
        g.len[k] = len[k];
        g.inv_w[k] = (double)g.nc[k] / len[k];
    }
    mdp_grid_set_radius(g, r);
    g.origin[0] = g.origin[1] = g.origin[2] = 0.0;
    if (na > 0) {
        g.origin[0] = xa[0];
        g.origin[1] = ya[0];
        g.origin[2] = za[0];
    }
    const int ncell = g.nc[0] * g.nc[1] * g.nc[2];
    std::vector<int> start(ncell + 1, 0), fill(ncell, 0), sidx(na);
    std::vector<double> sx(na), sy(na), sz(na);
    for (int i = 0; i < na; ++i) start[mdp_grid_cell(g, xa[i], ya[i], za[i]) + 1]++;
    for (int c = 0; c < ncell; ++c) start[c + 1] += start[c];
    for (int i = 0; i < na; ++i) {
        const int c = mdp_grid_cell(g, xa[i], ya[i], za[i]);
        const int p = start[c] + fill[c]++;
        sx[p] = xa[i];
        sy[p] = ya[i];
        sz[p] = za[i];
        sidx[p] = i;
    }
    std::vector<uint32_t> halo((size_t)mdp_halo_cells(g));
    for (int h = 0; h < (int)halo.size(); ++h) halo[h] = mdp_halo_entry(g, start.data(), h);
    for (int j = 0; j < nb; ++j)
        mdp_shell_probe(g, halo.data(), sx.data(), sy.data(), sz.data(), sidx.data(), xb[j], yb[j], zb[j], j, rin2, rout2, shell_mode,
                        exclude_same, Mark{out, nb, j});
    return 0;
}
