#!/usr/bin/env python
"""Randomised stress of the pair engine against the ORACLE (the CPU restatement of the reference): every mode the public
API uses -- symmetric and rectangular histograms with 1..3 classes (both pair kernels, orthogonal and triclinic image),
coordination numbers through the table-bin mode, neighbour lists in both shell modes (orthogonal and triclinic).
Sizes are kept small enough for the oracle's O(N^2) loops.

    python tests/fuzz/fuzz_oracle.py [seconds per component] [seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main(budget=None, seed=None, max_cases=None):
    import torch
    from mdproptools_b200 import ops
    from mdproptools_b200._lib import PAIR_F64, PAIR_NO_SORT, PAIR_TRICLINIC, bin_edges
    from oracle import oracle as O
    if budget is None:
        budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4242
    rng = np.random.default_rng(seed)
    torch.cuda.set_device(0)
    os.environ["MDP_SHELL_GRID"] = "0"                      # the general engine's list mode is the subject here
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    report = {}

    def cell_and_points(n, tri, kind):
        L = rng.uniform(9.0, 45.0, 3)
        cell = tuple(L) + ((float(rng.uniform(-0.45, 0.45) * L[0]), float(rng.uniform(-0.45, 0.45) * L[0]),
                            float(rng.uniform(-0.45, 0.45) * L[1])) if tri else ())
        s = rng.uniform(0, 1, (n, 3))
        if kind == "unwrapped":
            s = rng.uniform(-1.6, 2.6, (n, 3))
        if kind == "clustered":
            s = (rng.uniform(0, 1, (12, 3))[rng.integers(0, 12, n)] + np.round(rng.normal(0, 0.03, (n, 3)), 2)) % 1.0
        if tri:
            lx, ly, lz, xy, xz, yz = cell
            pos = s[:, 0:1] * np.array([lx, 0, 0]) + s[:, 1:2] * np.array([xy, ly, 0]) + s[:, 2:3] * np.array([xz, yz, lz])
        else:
            pos = s * L[None, :]
        if kind == "lattice":
            a = 0.05 * int(rng.integers(10, 60))
            m = int(np.ceil(n ** (1 / 3)))
            pos = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij")).reshape(3, -1)[:, :n].T * a
        return cell, np.ascontiguousarray(pos.T)

    def common():
        tri = rng.uniform() < 0.4
        kind = str(rng.choice(["uniform", "unwrapped", "clustered", "lattice"]))
        n = int(rng.choice([1, 2, 33, 257, 1000, 3000]))
        cell, pos = cell_and_points(n, tri, kind)
        lim = (0.49 if tri else 0.95) * min(cell[:3])
        rc = float(rng.uniform(1.0, min(12.0, lim)))
        flags = (PAIR_TRICLINIC if tri else 0) | (PAIR_F64 if rng.uniform() < 0.4 else 0) | (PAIR_NO_SORT if rng.uniform() < 0.1 else 0)
        return tri, kind, n, cell, pos, rc, flags

    def hist_sym():
        tri, kind, n, cell, pos, rc, flags = common()
        ncls = int(rng.integers(1, 4))
        ddr = float(rng.choice([0.01, 0.05, 0.1, 0.37]))
        nb = max(1, int(rc / ddr))
        typ = rng.integers(1, ncls + 1, n).astype(np.float64)
        rel = np.array([[a, b] for a in range(1, ncls + 1) for b in range(a, ncls + 1)])
        fn = O.rdf_loop_tri if tri else O.rdf_loop
        full, part = fn(typ, pos[0], pos[1], pos[2], rel, cell, rc, ddr, nb, nthreads=0)
        hist = ops.pair_hist(dev(pos[None]), dev((typ - 1).astype(np.int32)), ncls, [cell], O.rcut_sq(rc), bin_edges(ddr, nb), ddr, flags=flags)
        w = [np.full(ops.sym_rows(ncls), 2)]
        for a, b in rel:
            r = np.zeros(ops.sym_rows(ncls), dtype=np.int64)
            r[ops.sym_row(a - 1, b - 1, ncls)] = 2 if a == b else 1
            w.append(r)
        red = ops.hist_reduce(hist, np.stack(w)).cpu().numpy()[0]
        if not (np.array_equal(red[0], full) and np.array_equal(red[1:], part)):
            return dict(tri=tri, kind=kind, n=n, cell=cell, rc=rc, ddr=ddr, ncls=ncls, flags=flags)

    def hist_rect():
        tri, kind, n, cell, pos, rc, flags = common()
        m = int(rng.choice([1, 40, 700]))
        _, posb = cell_and_points(m, tri, "uniform")
        if tri:                                              # the second set must live in the same cell
            s = rng.uniform(0, 1, (m, 3))
            lx, ly, lz, xy, xz, yz = cell
            posb = np.ascontiguousarray((s[:, 0:1] * np.array([lx, 0, 0]) + s[:, 1:2] * np.array([xy, ly, 0]) + s[:, 2:3] * np.array([xz, yz, lz])).T)
        else:
            posb = np.ascontiguousarray((rng.uniform(0, 1, (m, 3)) * np.array(cell[:3])[None, :]).T)
        na_, nb_ = int(rng.integers(1, 4)), int(rng.integers(1, 3))
        ddr = float(rng.choice([0.05, 0.1, 0.37]))
        nb = max(1, int(rc / ddr))
        ta, tb = rng.integers(1, na_ + 1, n).astype(np.float64), rng.integers(1, nb_ + 1, m).astype(np.float64)
        rel = np.array([[a, b] for a in range(1, na_ + 1) for b in range(1, nb_ + 1)])
        fn = O.rdf_rect_tri if tri else O.rdf_rect
        part = fn(ta, pos[0], pos[1], pos[2], tb, posb[0], posb[1], posb[2], rel, cell, rc, ddr, nb, nthreads=0)
        hist = ops.pair_hist(dev(pos[None]), dev((ta - 1).astype(np.int32)), na_, [cell], O.rcut_sq(rc), bin_edges(ddr, nb), ddr,
                             xyz_b=dev(posb[None]), cls_b=dev((tb - 1).astype(np.int32)), ncls_b=nb_, flags=flags)
        w = np.zeros((len(rel), na_ * nb_), dtype=np.int64)
        for k, (a, b) in enumerate(rel):
            w[k, (a - 1) * nb_ + (b - 1)] = 1
        red = ops.hist_reduce(hist, w).cpu().numpy()[0]
        if not np.array_equal(red, part):
            return dict(tri=tri, kind=kind, n=n, m=m, cell=cell, rc=rc, ddr=ddr, na=na_, nb=nb_, flags=flags)

    def cn_table():
        tri, kind, n, cell, pos, rc, flags = common()
        if tri:
            return None                                      # (the oracle's cn loops are orthogonal, as the reference's)
        ncls = int(rng.integers(1, 4))
        typ = rng.integers(1, ncls + 1, n).astype(np.float64)
        rel = np.array([[a, b] for a in range(1, ncls + 1) for b in range(a, ncls + 1)])
        rcs = [float(rng.uniform(0.5, rc)) for _ in rel]
        cn = O.cn_loop(typ, pos[0], pos[1], pos[2], rel, cell, rcs, nthreads=0)
        rc2 = np.array([r * r for r in rcs])
        thr = np.unique(rc2)
        edges = np.concatenate(([0.0], thr))
        hist = ops.pair_hist(dev(pos[None]), dev((typ - 1).astype(np.int32)), ncls, [cell], float(thr[-1]), edges, 0.0,
                             flags=flags & ~PAIR_F64)
        w = np.zeros((len(rel), ops.sym_rows(ncls)), dtype=np.int64)
        for k, (a, b) in enumerate(rel):
            w[k, ops.sym_row(a - 1, b - 1, ncls)] = 2 if a == b else 1
        red = ops.hist_reduce(hist, w, cumulative=True).cpu().numpy()[0]
        upto = np.searchsorted(thr, rc2)
        got = np.array([red[k, upto[k]] for k in range(len(rel))])
        if not np.array_equal(got, cn):
            return dict(kind=kind, n=n, cell=cell, rcs=rcs, ncls=ncls, flags=flags)

    def lists():
        tri, kind, n, cell, pos, rc, flags = common()
        n = min(n, 1000)
        pos = pos[:, :n]
        m = int(rng.choice([1, 50, 900]))
        same = rng.uniform() < 0.3
        if same:
            posb = pos
        elif tri:
            s = rng.uniform(0, 1, (m, 3))
            lx, ly, lz, xy, xz, yz = cell
            posb = np.ascontiguousarray((s[:, 0:1] * np.array([lx, 0, 0]) + s[:, 1:2] * np.array([xy, ly, 0]) + s[:, 2:3] * np.array([xz, yz, lz])).T)
        else:
            posb = np.ascontiguousarray((rng.uniform(0, 1, (m, 3)) * np.array(cell[:3])[None, :]).T)
        r_in = float(rng.choice([0.0, 0.4 * rc]))
        fn = O.shell_mask_tri if tri else O.shell_mask
        h = fn(pos[0], pos[1], pos[2], posb[0], posb[1], posb[2], cell, r_in, rc, same)
        lst, _ = ops.pair_list(dev(pos[None]), dev(posb[None]), [cell], r_in * r_in, rc * rc, 1, exclude_same_index=same,
                               flags=flags & PAIR_TRICLINIC)
        l = lst.cpu().numpy()
        got = np.zeros_like(h)
        got[l[:, 1], l[:, 2]] = 1
        if not (len(l) == int(h.sum()) and np.array_equal(got, h)):
            return dict(tri=tri, kind=kind, n=n, m=posb.shape[1], cell=cell, r_in=r_in, rc=rc, same=same)

    fails = 0
    for name, fn in (("hist_sym", hist_sym), ("hist_rect", hist_rect), ("cn_table", cn_table), ("lists", lists)):
        t0, k, bad = time.time(), 0, 0
        while time.time() - t0 < budget and (max_cases is None or k < max_cases):
            desc = fn()
            k += 1
            if desc is not None:
                bad += 1
                print("MISMATCH", name, desc, flush=True)
        report[name] = (k, bad)
        fails += bad
    os.environ.pop("MDP_SHELL_GRID", None)
    print("cases (mismatches):", ", ".join(f"{k} {v[0]} ({v[1]})" for k, v in report.items()))
    return report if max_cases is not None else fails


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
