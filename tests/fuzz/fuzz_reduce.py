#!/usr/bin/env python
"""Randomised stress of the reductions and list epilogues against numpy / the oracle:

  list_group   counting sort + per-segment sort of a neighbour list          vs  np.lexsort (permutation, offsets: exact)
  pair_keys    distinct (central, partner) keys                              vs  np.unique (exact)
  hydration    cosines and the two counters per (frame, cation)              vs  the numpy expressions of the reference (bit for bit)
  clusters     molecule completion + force filter                            vs  numpy (exact)
  segment_com  per-molecule weighted means                                   vs  sequential fp64 in atom order (bit for bit)
  msd_window   MSD over all time origins                                     vs  the oracle (1e-10)
  msd_interval interval MSD                                                  vs  the oracle (1e-13)
  cumtrapz     cumulative trapezoid                                          vs  the oracle (1e-12 of max)

    python tests/fuzz/fuzz_reduce.py [seconds per component] [seed]
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
    from oracle import oracle as O
    if budget is None:
        budget = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 31337
    rng = np.random.default_rng(seed)
    torch.cuda.set_device(0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def rand_list(F, na, nb, m):
        if m == 0:
            return np.zeros((0, 3), dtype=np.int32)
        lst = np.stack([rng.integers(0, F, m), rng.integers(0, na, m), rng.integers(0, nb, m)], axis=1).astype(np.int32)
        lst = np.unique(lst, axis=0)
        return lst[rng.permutation(len(lst))]

    def dims():
        return (int(rng.integers(1, 6)), int(rng.choice([1, 2, 9, 50, 300])), int(rng.choice([1, 5, 40, 1000, 20000])),
                int(rng.choice([0, 1, 30, 2000, 60000])))

    def list_group():
        F, na, nb, m = dims()
        lst = rand_list(F, na, nb, m)
        seg_off, key, perm = (t.cpu().numpy() for t in ops.list_group(dev(lst), na, F))
        order = np.lexsort((lst[:, 2], lst[:, 1], lst[:, 0])) if len(lst) else np.zeros(0, dtype=np.int64)
        counts = np.bincount(lst[:, 0].astype(np.int64) * na + lst[:, 1], minlength=F * na) if len(lst) else np.zeros(F * na, dtype=np.int64)
        if not (np.array_equal(perm, order) and np.array_equal(key, lst[order, 2]) and
                np.array_equal(seg_off, np.concatenate(([0], np.cumsum(counts))))):
            return dict(F=F, na=na, nb=nb, m=len(lst))

    def pair_keys():
        F, na, nb, m = dims()
        lst = rand_list(F, na, nb, m)
        got = ops.unique_pair_keys(dev(lst), na, nb).cpu().numpy()
        want = np.unique(lst[:, 1].astype(np.int64) * nb + lst[:, 2])
        if not np.array_equal(got, want):
            return dict(F=F, na=na, nb=nb, m=len(lst))

    def hydration():
        F, ncat, nw, m = dims()
        nw, m = min(nw, 2000), min(m, 5000)
        L = rng.uniform(8.0, 40.0, (F, 3))
        cat = rng.uniform(-0.5, 1.5, (F, 3, ncat)) * L[:, :, None]
        o = rng.uniform(-0.5, 1.5, (F, 3, nw)) * L[:, :, None]
        h1, h2 = o + rng.normal(0, 0.6, o.shape), o + rng.normal(0, 0.6, o.shape)
        lst = rand_list(F, ncat, nw, m)
        if len(lst) == 0:
            return None
        thr = float(rng.uniform(-0.9, 0.2))
        cos, seg_off, counts = (t.cpu().numpy() for t in ops.hydration_count(dev(lst), dev(cat), dev(o), dev(h1), dev(h2), L, thr))
        order = np.lexsort((lst[:, 2], lst[:, 1], lst[:, 0]))
        f, ia, ib = lst[order].T
        d = np.stack([cat[f, a, ia] - o[f, a, ib] for a in range(3)], axis=1)
        for a in range(3):
            l = L[f, a]
            cond = (d[:, a] > l / 2) | (d[:, a] < -l / 2)
            d[cond, a] = d[cond, a] - np.sign(d[cond, a]) * l[cond]
        v = np.stack([(h1[f, a, ib] + h2[f, a, ib]) - 2 * o[f, a, ib] for a in range(3)], axis=1)
        want = np.sum(d * v, axis=1) / (np.linalg.norm(d, axis=1) * np.linalg.norm(v, axis=1))
        seg = f.astype(np.int64) * ncat + ia
        if not (np.array_equal(cos, want) and np.array_equal(counts[:, :, 0].ravel(), np.bincount(seg, minlength=F * ncat)) and
                np.array_equal(counts[:, :, 1].ravel(), np.bincount(seg[want < thr], minlength=F * ncat))):
            return dict(F=F, ncat=ncat, nw=nw, m=len(lst), thr=thr)

    def clusters():
        F, ncen = int(rng.integers(1, 4)), int(rng.choice([1, 3, 11, 40]))
        sizes = rng.integers(1, 17, int(rng.choice([1, 5, 60, 400])))
        seg_off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
        n, nmol = int(seg_off[-1]), len(sizes)
        mol_of_atom = np.repeat(np.arange(nmol), sizes).astype(np.int32)
        force = rng.normal(0, 40.0, (F, 3, n))
        const, max_force = 0.043363 / 16, float(rng.choice([0.05, 0.75, -0.2]))
        lst = rand_list(F, ncen, n, int(rng.choice([0, 10, 1500, 20000])))
        so, mols, cnt = (t.cpu().numpy() for t in ops.cluster_members(dev(lst), ncen, dev(force), dev(seg_off), dev(mol_of_atom), const, max_force))
        fsum = np.stack([np.stack([np.add.reduceat(force[f, a], seg_off[:-1]) for a in range(3)]) for f in range(F)])
        # (reduceat adds pairwise inside long segments; segments here are <= 16 atoms: sequential, as the kernel)
        ok = fsum.min(axis=1) * const < max_force
        for f in range(F):
            for c in range(ncen):
                s = f * ncen + c
                atoms = lst[(lst[:, 0] == f) & (lst[:, 1] == c)][:, 2]
                want = np.unique(mol_of_atom[atoms])
                want = want[ok[f, want]]
                if not np.array_equal(mols[so[s]: so[s] + cnt[s]], want):
                    return dict(F=F, ncen=ncen, nmol=nmol, m=len(lst), max_force=max_force, f=f, c=c)

    def segment_com():
        S = int(rng.choice([1, 2, 31, 500, 5000]))
        sizes = rng.integers(1, int(rng.choice([2, 17, 70])), S)
        off = np.concatenate(([0], np.cumsum(sizes)))
        n, F, C = int(off[-1]), int(rng.integers(1, 4)), int(rng.integers(1, 4))
        attr, w, q = rng.normal(0, 10, (F, C, n)), rng.uniform(1, 30, n), rng.normal(0, 1, n)
        out, wsum, qsum = ops.segment_com(dev(attr), dev(w), dev(off.astype(np.int32)), extra=dev(q))
        # sequential accumulation in atom order, vectorised over segments: step k adds the k-th atom of every segment that has one
        ws, acc = np.zeros(S), np.zeros((F, C, S))
        for k in range(int(sizes.max())):
            sel = np.flatnonzero(sizes > k)
            idx = off[sel] + k
            ws[sel] = ws[sel] + w[idx]
            acc[:, :, sel] = acc[:, :, sel] + attr[:, :, idx] * w[idx]
        if not (np.array_equal(out.cpu().numpy(), acc / ws) and np.allclose(wsum.cpu().numpy(), ws, rtol=1e-14)):
            return dict(S=S, n=n, F=F, C=C)

    def msd_window():
        T, n = int(rng.choice([2, 33, 97, 300, 700])), int(rng.choice([1, 33, 70, 700]))
        lag = int(rng.integers(1, T + 1))
        traj = 50.0 + np.cumsum(rng.normal(0, 0.1, (T, 3, n)), axis=0)
        sums = ops.msd_all_origins(dev(traj), lag).cpu().numpy()[:, 0]
        norm = (T - np.arange(lag))[:, None] * n
        ref = O.msd_all_origins(traj, lag)
        got = sums / norm
        if not (np.all(sums[0] == 0) and np.allclose(got[1:], ref[1:], rtol=1e-10, atol=0)):
            return dict(T=T, n=n, lag=lag)

    def msd_interval():
        T, n, st = int(rng.choice([2, 9, 33, 200])), int(rng.choice([1, 31, 700, 5000])), int(rng.integers(1, 5))
        traj = np.cumsum(rng.normal(0, 0.1, (T, 3, n)), axis=0)
        if len(traj[::st]) < 2:
            return None
        got = ops.msd_interval(dev(traj[::st]), 1e-10, 1).cpu().numpy()
        if not np.allclose(got, O.msd_interval(traj, 1e-10, st), rtol=1e-13, atol=0):
            return dict(T=T, n=n, stride=st)

    def cumtrapz():
        R, T = int(rng.integers(1, 5)), int(rng.choice([2, 3, 255, 256, 257, 5001, 100000]))
        y = rng.normal(0, 1, (R, T))
        lead = bool(rng.integers(0, 2))
        got = ops.cumtrapz(dev(y), 0.37, 2.5, leading_zero=lead).cpu().numpy()
        ref = np.stack([2.5 * O.cumtrapz(r, 0.37, lead) for r in y])
        if not (got.shape == ref.shape and np.allclose(got, ref, rtol=0, atol=1e-12 * (np.abs(ref).max() + 1e-300))):
            return dict(R=R, T=T, lead=lead)

    report, fails = {}, 0
    for name, fn in (("list_group", list_group), ("pair_keys", pair_keys), ("hydration", hydration), ("clusters", clusters),
                     ("segment_com", segment_com), ("msd_window", msd_window), ("msd_interval", msd_interval), ("cumtrapz", cumtrapz)):
        t0, k, bad = time.time(), 0, 0
        while time.time() - t0 < budget and (max_cases is None or k < max_cases):
            desc = fn()
            k += 1
            if desc is not None:
                bad += 1
                print("MISMATCH", name, desc, flush=True)
        report[name] = (k, bad)
        fails += bad
    print("cases (mismatches):", ", ".join(f"{k} {v[0]} ({v[1]})" for k, v in report.items()))
    return report if max_cases is not None else fails


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
