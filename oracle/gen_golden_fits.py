#!/usr/bin/env python
"""Golden vectors for the HOST-SIDE fits from the UNMODIFIED reference:

    Conductivity.detect_time_range / fit_curve / green_kubo      conductivity.py:116-165, 234-274
    Viscosity.fit_avg_visc / bootstrapping                       viscosity.py:239-434
    ResidenceTime.fit_auto_correlation                           residence_time.py:150-208
    Diffusion.get_msd_from_log                                   diffusion.py:241-265

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Inputs are synthetic and seeded; they are stored
together with the reference's outputs in tests/golden/ref_fits.npz (+ tests/golden/msd_log.tar.gz), so the tests run
where the reference is absent.  The discrete results -- the plateau window indices, idx_start/idx_cut of the viscosity
fit, the truncation length of the residence fit -- are stored explicitly: the tests report index agreement first, fitted
scalars second ("hard parts").

    python oracle/gen_golden_fits.py
"""
from __future__ import annotations

import os
import sys
import tarfile
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def synth_flux_corr(rng, T, ntypes=2):
    """Charge-flux correlation functions: a damped oscillation that dies out, then noise whose amplitude shrinks -- the
    shape detect_time_range looks for a quiet window in."""
    t = np.arange(T, dtype=np.float64)
    rows = []
    for i in range(ntypes + 1):
        sig = (1.0 + 0.3 * i) * np.exp(-t / (T / 40.0)) * np.cos(t / (T / 300.0))
        noise = rng.normal(0, 1, T) * (0.05 * np.exp(-t / (T / 6.0)) + 1e-4)
        burst = np.where((t > 0.7 * T) & (t < 0.75 * T), rng.normal(0, 0.05, T), 0.0)
        rows.append(sig + noise + burst)
    return np.stack(rows).astype(np.float32).astype(np.float64)     # stored as float32: the reference sees exactly these values


def synth_visc(rng, T, nrep, dt_fs=1.0):
    """Running viscosity integrals of nrep replicates: eta (1 - a e^{-t/t1} - (1-a) e^{-t/t2}) plus a random walk whose
    spread grows with time (so that std >= 0.4 * mean is reached at a finite time)."""
    t = np.arange(1, T + 1, dtype=np.float64) * dt_fs
    eta, a, t1, t2 = 8.0e-4, 0.7, 3.0e3, 2.0e4
    base = eta * (1 - a * np.exp(-t / t1) - (1 - a) * np.exp(-t / t2))
    # increments grow with time: the replicates agree at early times (std << mean, as for real running integrals that all
    # start at 0) and spread late, so that std >= 0.4 * mean is first met far beyond the 2 ps where the fit starts
    walks = np.cumsum(rng.normal(0, 1, (nrep, T)) * (5e-5 * (1.0 + (t / 2500.0) ** 2)), axis=1) * eta
    return t, (base[None, :] + walks).astype(np.float32).astype(np.float64)


def main():
    from oracle import ref_harness as H
    H.install()
    from mdproptools.dynamical.conductivity import Conductivity
    from mdproptools.dynamical.diffusion import Diffusion
    from mdproptools.dynamical.residence_time import ResidenceTime
    from mdproptools.dynamical.viscosity import Viscosity

    rng = np.random.default_rng(20261017)
    out = {}
    work = tempfile.mkdtemp(prefix="mdp_fits_")

    # ---- conductivity: plateau window, average, sigma -------------------------------------------------
    c = Conductivity.__new__(Conductivity)            # the constructor parses dump files; the fits need only these fields
    c.temp, c.volume, c.working_dir = 298.15, (40.0e-10) ** 3, work
    for tag, T, nt in (("a", 3000, 2), ("b", 60000, 1)):
        tot = synth_flux_corr(rng, T, nt)
        c.time = (np.arange(T) * 2.0e-15).tolist()
        integ = c.integrate_charge_flux_correlation(tot)
        out[f"cond_{tag}_tot_flux"] = tot.astype(np.float32)
        out[f"cond_{tag}_integral_last"] = integ[:, -1]            # (the integral itself is recomputed by the test)
        for tol in (1e-4, 1e-2, 0.3):
            try:
                win = np.array([Conductivity.detect_time_range(tot[i], tol) for i in range(len(tot))], dtype=np.int64)
                ave, tr = c.fit_curve(tot, integ, tol)
                out[f"cond_{tag}_win_{tol}"] = win
                out[f"cond_{tag}_ave_{tol}"] = ave
                out[f"cond_{tag}_range_{tol}"] = np.array([list(x) for x in tr], dtype=np.float64)
                out[f"cond_{tag}_sigma_{tol}"] = c.green_kubo(ave)
            except TypeError:
                out[f"cond_{tag}_win_{tol}"] = np.zeros((0, 2), dtype=np.int64)      # no quiet window: the reference raises

    # ---- viscosity: weighted double-exponential fit, bootstrap ---------------------------------------
    v = Viscosity("log.none_*", cutoff_time=500, volume=40.0 ** 3, temp=298.15, timestep=1, acf_method="wkt", units="real",
                  working_dir=work)
    t, visc_avg = synth_visc(rng, 40000, 4)
    v.time = t
    out["visc_fit_avg"] = visc_avg.astype(np.float32)               # time = 1, 2, ..., T fs
    mean, std = np.average(visc_avg, axis=0), np.std(visc_avg, axis=0)
    out["visc_fit_idx"] = np.array([np.where(t > 2000)[0][0], np.where(std >= 0.4 * mean)[0][0]], dtype=np.int64)
    out["visc_fit_eta"] = np.array(v.fit_avg_visc(visc_avg, plot=False))
    import random
    random.seed(7)
    eta_b, std_b = v.bootstrapping(visc_avg, 3, 4, plot=False)
    out["visc_boot"] = np.array([eta_b, std_b])

    # ---- residence time: stretched-exponential fit --------------------------------------------------
    r = ResidenceTime.__new__(ResidenceTime)
    r.working_dir = work
    tt = np.arange(400) * 0.5
    cols = {"Time (ps)": tt}
    for name, (a, tr, ts, beta) in {"9-1": (0.8, 60.0, 2.0, 0.7), "9-4": (0.55, 15.0, 0.8, 0.9)}.items():
        cols[name] = (ResidenceTime._stretched_exp_function(tt, a, tr, ts, beta) + rng.normal(0, 2e-3, len(tt))).astype(np.float32).astype(np.float64)
    r.corr_df = pd.DataFrame(cols)
    res = r.fit_auto_correlation(cut_percent=0.9, plot=False)
    out["res_fit_time"] = tt
    out["res_fit_cols"] = np.array(list(cols)[1:])
    out["res_fit_corr"] = np.stack([cols[k] for k in list(cols)[1:]]).astype(np.float32)
    out["res_fit_len"] = np.array([int(len(tt) * 0.9)], dtype=np.int64)
    out["res_fit_params"] = np.array([res[k] for k in list(cols)[1:]])

    # ---- MSD from thermo logs -----------------------------------------------------------------------
    logd = os.path.join(work, "msdlog")
    os.makedirs(logd)
    step0 = 0
    for k in range(2):
        n = 40
        steps = step0 + np.arange(n + 1) * 1000
        msd1 = 0.6 * steps / 1000.0 + rng.normal(0, 0.05, n + 1)
        msd2 = 0.1 * steps / 1000.0 + rng.normal(0, 0.02, n + 1)
        with open(os.path.join(logd, f"log.msd_{k + 1}"), "w") as f:
            f.write("LAMMPS (synthetic)\nPer MPI rank memory allocation (min/avg/max) = 1 | 1 | 1 Mbytes\n")
            f.write("Step Temp c_msd1[4] c_msd2[4]\n")
            for s_, a_, b_ in zip(steps, msd1, msd2):
                f.write(f"{int(s_)} 298.0 {a_:.8f} {b_:.8f}\n")
            f.write("Loop time of 1.0 on 1 procs for 40000 steps with 10 atoms\n")
        step0 = int(steps[-1])
    d = Diffusion(timestep=1, units="real", outputs_dir=logd, diff_dir=work)
    msd = d.get_msd_from_log("log.msd_*")
    out["msdlog_cols"] = np.array(list(msd.columns))
    out["msdlog_values"] = msd.values.astype(np.float64)
    with tarfile.open(os.path.join(GOLD, "msd_log.tar.gz"), "w:gz") as tar:
        for name in sorted(os.listdir(logd)):
            tar.add(os.path.join(logd, name), arcname=name)

    np.savez_compressed(os.path.join(GOLD, "ref_fits.npz"), **out)
    print("wrote", os.path.join(GOLD, "ref_fits.npz"), {k: np.shape(v_) for k, v_ in out.items() if "win" in k or "idx" in k})
    for k in ("visc_fit_idx", "visc_fit_eta", "visc_boot", "res_fit_params"):
        print(k, out[k])


if __name__ == "__main__":
    main()
