#!/usr/bin/env python
"""Randomised stress of the pair engine: k_pair_fast (fp32 filter + fp64 exact path) must give the integers of the
all-fp64 kernel k_pair for EVERY geometry -- random atom counts, boxes (orthogonal and triclinic), cutoffs, bin widths,
class counts, point distributions (uniform, clustered, simple-cubic lattice = thousands of pairs exactly on bin edges,
unwrapped, far from the origin), symmetric and rectangular sets.  Every 10th case is also checked against the oracle.

    python tests/fuzz/fuzz_pair.py [seconds] [seed]          # prints one line per failure and a summary; exit 1 on failure
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
    from mdproptools_b200._lib import PAIR_F64, PAIR_NO_SORT, PAIR_TRICLINIC, Context, bin_edges
    from oracle import oracle as O
    if budget is None:
        budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    rng = np.random.default_rng(seed)
    torch.cuda.set_device(0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t0, cases, bad, exact, evald, kinds, rejected = time.time(), 0, 0, 0, 0, {}, {}
    while time.time() - t0 < budget and (max_cases is None or cases < max_cases):
        n = int(rng.choice([50, 300, 1000, 3000, 8000, 20000]))
        F = int(rng.choice([1, 1, 2, 3]))
        L = rng.uniform(12.0, 60.0, 3)
        tri = rng.uniform() < 0.3
        rc = float(rng.uniform(2.0, min(14.0, 0.49 * L.min() if tri else 0.9 * L.min())))
        ddr = float(rng.choice([0.0025, 0.01, 0.05, 0.05, 0.1, 0.37, 1.5]))
        nb = max(1, int(rc / ddr))
        if nb > 16000:
            continue
        ncls = int(rng.choice([1, 1, 2, 3]))
        dist = rng.choice(["uniform", "clustered", "lattice", "unwrapped", "far"])
        Ls = np.stack([L * (1.0 + 0.01 * f) for f in range(F)])
        pos = rng.uniform(0, 1, (F, 3, n)) * Ls[:, :, None]
        if dist == "clustered":
            c = rng.uniform(0, 1, (F, 3, 20)) * Ls[:, :, None]
            pos = c[:, :, rng.integers(0, 20, n)] + np.round(rng.normal(0, 0.7, (F, 3, n)), 1)
        elif dist == "lattice":
            a = ddr * int(rng.integers(3, 40))                       # lattice constant = a whole number of bins
            m = int(np.ceil(n ** (1 / 3)))
            g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij")).reshape(3, -1)[:, :n] * a
            pos = np.broadcast_to(g[None], (F, 3, n)).copy()
        elif dist == "unwrapped":
            pos += rng.integers(-4, 5, (F, 3, n)) * Ls[:, :, None]
        elif dist == "far":
            pos += 1.0e4
        flags = PAIR_NO_SORT if rng.uniform() < 0.1 else 0
        boxes = Ls
        if tri:
            tilt = np.stack([rng.uniform(-0.45, 0.45, F) * Ls[:, 0], rng.uniform(-0.45, 0.45, F) * Ls[:, 0], rng.uniform(-0.45, 0.45, F) * Ls[:, 1]], axis=1)
            boxes = np.concatenate([Ls, tilt], axis=1)
            flags |= PAIR_TRICLINIC
        typ = rng.integers(1, ncls + 1, n).astype(np.float64)
        cls = dev((typ - 1).astype(np.int32))
        kw = {}
        rect = rng.uniform() < 0.2
        if rect:
            mB = int(rng.choice([40, 700, 5000]))
            posb = rng.uniform(0, 1, (F, 3, mB)) * Ls[:, :, None]
            typb = rng.integers(1, ncls + 1, mB).astype(np.float64)
            kw = dict(xyz_b=dev(posb), cls_b=dev((typb - 1).astype(np.int32)), ncls_b=ncls)
        edges = bin_edges(ddr, nb)
        x = dev(pos)
        only = os.environ.get("FUZZ_ONLY")
        if only is not None and cases != int(only):
            cases += 1
            if cases > int(only):
                break
            continue
        case = dict(n=n, F=F, L=L.tolist(), tri=bool(tri), rc=rc, ddr=ddr, nb=nb, ncls=ncls, dist=str(dist), flags=flags, rect=bool(rect),
                    mB=(kw['xyz_b'].shape[2] if rect else None))
        if os.environ.get("FUZZ_VERBOSE"):
            print("case", cases, case, flush=True)
        try:
            fast = ops.pair_hist(x, cls, ncls, boxes, rc * rc, edges, ddr, flags=flags, **kw)
            if os.environ.get("FUZZ_VERBOSE"):
                torch.cuda.synchronize()
                print("  fast ok", flush=True)
            st = Context.get(0).pair_stats()
            f64 = ops.pair_hist(x, cls, ncls, boxes, rc * rc, edges, ddr, flags=flags | PAIR_F64, **kw)
        except Exception as exc:                                   # a shape the ABI refuses (it says so): not a parity case
            rejected[str(exc)[:80]] = rejected.get(str(exc)[:80], 0) + 1
            continue
        ok = bool(torch.equal(fast, f64))
        if ok and cases % 10 == 0 and not rect and not tri and n <= 8000:
            full, _ = O.rdf_loop(typ, pos[0, 0], pos[0, 1], pos[0, 2], np.array([[1, 1]]), tuple(Ls[0]), rc, ddr, nb, nthreads=0)
            ok = bool(np.array_equal(fast[0].sum(dim=0).cpu().numpy() * 2, full))
        cases += 1
        exact += int(st.get("exact_path_pairs", 0))
        evald += int(st.get("pair_evals", 0))
        key = f"{dist}{'/tri' if tri else ''}{'/rect' if rect else ''}"
        kinds[key] = kinds.get(key, 0) + 1
        if not ok:
            bad += 1
            print("MISMATCH", case, flush=True)
    print(f"{cases} random cases in {time.time() - t0:.0f} s: {bad} mismatches between k_pair_fast and k_pair (fp64); "
          f"{evald:.3e} pairs evaluated by the fast kernel, {exact:.3e} of them settled on the exact fp64 path; kinds: {kinds}; refused by the ABI: {rejected}")
    return (cases, bad) if max_cases is not None else (1 if bad else 0)


if __name__ == "__main__":
    sys.exit(main())
