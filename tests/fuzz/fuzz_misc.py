#!/usr/bin/env python
"""Randomised stress of the kernels beside the pair engine, each against an independent implementation:

  shell     k_shell_grid (small-set cell-grid search)        vs  the general pair engine's list mode (MDP_SHELL_GRID=0)
  survival  key build + bitmask fill + run-based correlation  vs  the AND-shift-popcount kernel and the oracle
  fft       Stockham FFT correlation (three stages per pass)  vs  the direct fp64 sum (1e-10 of max|C|)
  parser    text -> HBM -> k_dump_rows (FrameBatches)         vs  the host parser, bit for bit (random spellings, CRLF, blanks)
  msd       k_msd_single per-atom values                      vs  the oracle, bit for bit

    python tests/fuzz/fuzz_misc.py [seconds per component] [seed]
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main(budget=None, seed=None, only=None, max_cases=None):
    import torch
    from mdproptools_b200 import ops
    from mdproptools_b200.io import dump as D
    from mdproptools_b200.io.pipeline import FrameBatches
    from oracle import oracle as O
    if budget is None:
        budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 777
    rng = np.random.default_rng(seed)
    torch.cuda.set_device(0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    report, failures = {}, 0

    def loop(name, fn):
        nonlocal failures
        if only and name not in only:
            return
        t0, n, bad = time.time(), 0, 0
        while time.time() - t0 < budget and (max_cases is None or n < max_cases):
            desc = fn()
            n += 1
            if desc is not None:
                bad += 1
                print("MISMATCH", name, desc, flush=True)
        report[name] = (n, bad)
        failures += bad

    def packed(lst, na, nb):
        l = lst.cpu().numpy().astype(np.int64)
        return np.sort((l[:, 0] * na + l[:, 1]) * nb + l[:, 2])

    def shell():
        F = int(rng.integers(1, 5))
        na = int(rng.choice([1, 7, 60, 400, 2000, 4096]))
        nb = int(max(8 * na, rng.choice([100, 3000, 40000])))
        L = rng.uniform(10.0, 70.0, 3)
        Ls = np.stack([L * (1 + 0.02 * f) for f in range(F)])
        r_out = float(rng.uniform(1.0, L.min() / 3.05))
        r_in = float(rng.choice([0.0, 0.5 * r_out]))
        mode = int(rng.integers(0, 2))
        a = rng.uniform(0, 1, (F, 3, na)) * Ls[:, :, None]
        b = rng.uniform(0, 1, (F, 3, nb)) * Ls[:, :, None]
        kind = rng.choice(["wrapped", "unwrapped", "same", "blob"])
        if kind == "unwrapped":
            b += rng.integers(-3, 4, (F, 3, nb)) * Ls[:, :, None]
            a += rng.integers(-1, 2, (F, 3, na)) * Ls[:, :, None]
        if kind == "blob":                                   # all central atoms in one corner: dense cells, long candidate lists
            a = rng.uniform(0, 0.15, (F, 3, na)) * Ls[:, :, None]
        excl = False
        if kind == "same":
            b[:, :, :na] = a                                 # the central atoms are among the partners: exclude_same_index
            excl = True
        res = []
        for flag in ("0", "1"):
            os.environ["MDP_SHELL_GRID"] = flag
            lst, _ = ops.pair_list(dev(a), dev(b), Ls, r_in ** 2, r_out ** 2, mode, exclude_same_index=excl)
            res.append(packed(lst, na, nb))
        os.environ.pop("MDP_SHELL_GRID", None)
        if not np.array_equal(res[0], res[1]):
            return dict(F=F, na=na, nb=nb, L=L.tolist(), r_in=r_in, r_out=r_out, mode=mode, kind=str(kind), n0=len(res[0]), n1=len(res[1]))

    def survival():
        T = int(rng.choice([1, 2, 63, 64, 65, 130, 1000, 2600, 5000]))
        na, nb = int(rng.integers(1, 12)), int(rng.integers(1, 60))
        kind = rng.choice(["dense", "sparse", "walk", "full"])
        if kind == "dense":
            h = rng.uniform(size=(T, na, nb)) < 0.5
        elif kind == "sparse":
            h = rng.uniform(size=(T, na, nb)) < 0.02
        elif kind == "full":
            h = np.ones((T, na, nb), dtype=bool)
        else:
            h = np.abs(np.cumsum(rng.normal(0, 0.3, (T, na, nb)), axis=0) + rng.normal(0, 1.5, (na, nb))) < 1.0
        f, a, b = np.nonzero(h)
        if len(f) == 0:
            return None
        lst = torch.from_numpy(np.stack([f, a, b], axis=1).astype(np.int32)).cuda()
        lst = lst[torch.randperm(lst.shape[0], device="cuda")]
        os.environ["MDP_SURVIVAL_RUNS"] = "0"
        ref, P = ops.bitmask_autocorr_from_list(lst, nb, T, n_a=na)
        os.environ["MDP_SURVIVAL_RUNS"] = "1"
        got, P2 = ops.bitmask_autocorr_from_list(lst, nb, T, n_a=na)
        os.environ.pop("MDP_SURVIVAL_RUNS", None)
        ok = P == P2 and torch.equal(ref, got)
        if ok and T <= 1000:
            ok = np.array_equal(got.cpu().numpy(), O.survival_counts(h.reshape(T, -1).astype(np.uint8)))
        if not ok:
            return dict(T=T, na=na, nb=nb, kind=str(kind))

    def fft():
        T = int(rng.choice([2048, 2049, 3000, 4095, 4096, 4097, 10000, 65536, 70001]))
        C = int(rng.integers(1, 5))
        nl = int(rng.choice([T, max(1, T // 3), 1]))
        a = np.cumsum(rng.normal(0, 1, (C, T)), axis=1) * 0.05 + rng.normal(0, 1, (C, T))
        b = a if rng.uniform() < 0.5 else rng.normal(0.3, 2, (C, T))
        os.environ["MDP_XCORR_FFT"] = "1"
        got = ops.xcorr_unbiased(dev(a), dev(b), nl).cpu().numpy()
        os.environ["MDP_XCORR_FFT"] = "0"
        ref = ops.xcorr_unbiased(dev(a), dev(b), nl).cpu().numpy()
        os.environ.pop("MDP_XCORR_FFT", None)
        err = np.max(np.abs(got - ref)) / (np.abs(ref).max() + 1e-300)
        if not err < 1e-10:
            return dict(T=T, C=C, nlags=nl, err=float(err))

    spell = [lambda v: repr(float(v)), lambda v: "%g" % v, lambda v: "%.17g" % v, lambda v: "%+.3e" % v, lambda v: "%.12f" % v,
             lambda v: "%.6f" % v, lambda v: "%.25f" % v, lambda v: "%dE0" % int(v), lambda v: "%.10e" % v, lambda v: "%d" % int(v)]
    tmp = tempfile.mkdtemp(prefix="mdp_fuzz_")

    def parser():
        nfiles, per = int(rng.integers(1, 4)), int(rng.integers(1, 3))
        n = int(rng.choice([1, 2, 31, 64, 65, 500, 5000]))
        cols = ["id", "type", "x", "y", "z"] + (["vx", "q"] if rng.uniform() < 0.5 else [])
        order = list(rng.permutation(len(cols)))
        fcols = [cols[k] for k in order]
        want = [c for c in cols if rng.uniform() < 0.8 or c == "id"]
        eol = "\r\n" if rng.uniform() < 0.2 else "\n"
        scale = float(rng.choice([1e-5, 1.0, 50.0, 1e6]))
        for old in os.listdir(tmp):
            os.remove(os.path.join(tmp, old))
        for k in range(nfiles):
            with open(os.path.join(tmp, f"dump.f.{k}.dump"), "w", newline="") as fh:
                for j in range(per):
                    ids = rng.permutation(n) + 1
                    vals = rng.normal(0, scale, (n, len(cols)))
                    fh.write(f"ITEM: TIMESTEP{eol}{k * 10 + j}{eol}ITEM: NUMBER OF ATOMS{eol}{n}{eol}ITEM: BOX BOUNDS pp pp pp{eol}0 9{eol}0 9{eol}0 9{eol}")
                    fh.write("ITEM: ATOMS " + " ".join(fcols) + eol)
                    for i in range(n):
                        row = []
                        for c in fcols:
                            if c == "id":
                                row.append(str(int(ids[i])))
                            elif c == "type":
                                row.append(str(1 + int(ids[i]) % 3))
                            else:
                                row.append(spell[int(rng.integers(0, len(spell)))](vals[i, cols.index(c)]))
                        fh.write((" " * int(rng.integers(0, 3))) + " ".join(row) + (" " if rng.uniform() < 0.3 else "") + eol)
                        if rng.uniform() < 0.01:
                            fh.write("  " + eol)
        pat = os.path.join(tmp, "dump.f.*.dump")
        ref = list(FrameBatches(pat, want, device_parse=False))
        fb = FrameBatches(pat, want, device_parse=True)
        got = list(fb)
        ok = len(ref) == len(got) and fb.device_parsed_frames + fb.host_reparsed_frames == nfiles * per
        for a_, b_ in zip(got, ref):
            ok = ok and [m.timestep for m in a_.metas] == [m.timestep for m in b_.metas]
            ok = ok and torch.equal(a_.wait().view(torch.int64), b_.wait().view(torch.int64)) and torch.equal(a_.host.view(torch.int64), b_.host.view(torch.int64))
        if not ok:
            return dict(nfiles=nfiles, per=per, n=n, fcols=fcols, want=want, eol=repr(eol), scale=scale)

    def msd():
        T, n = int(rng.integers(1, 12)), int(rng.choice([1, 31, 32, 33, 2049, 4096, 50000]))
        traj = np.cumsum(rng.normal(0, 0.1, (T, 3, n)), axis=0) + rng.uniform(-50, 50, (1, 3, n))
        t0 = int(rng.integers(0, T))
        per_atom, mean = O.msd_single_origin(traj, t0, 1e-10)
        t = dev(traj)
        sums, pa = ops.msd_single_origin(t, t[t0].contiguous(), 1e-10, per_atom=True)
        if not (np.array_equal(pa.cpu().numpy(), per_atom) and np.allclose(sums[:, 0].cpu().numpy() / n, mean, rtol=1e-12, atol=0)):
            return dict(T=T, n=n, t0=t0)

    for name, fn in (("shell", shell), ("survival", survival), ("fft", fft), ("parser", parser), ("msd", msd)):
        loop(name, fn)
    print("cases (mismatches):", ", ".join(f"{k} {v[0]} ({v[1]})" for k, v in report.items()))
    return report if max_cases is not None else failures


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
