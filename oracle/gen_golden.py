"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  ``python -m oracle.gen_golden``

Inputs that travel with the repo (so the GPU box, which has no /root/reference, can run the parity tests):
  tests/golden/sample_frames.tar.gz  two real frames of data/mg_tfsi_dme (timesteps 0 and 2500000)
  tests/golden/mini_traj.tar.gz      26-frame, few-molecule sub-trajectory cut from the same data
  tests/golden/visc_log.tar.gz       two synthetic LAMMPS thermo logs (Step Pxy Pxz Pyz), seeded
  tests/golden/water_box.tar.gz      synthetic cation + 3-site water box (3 frames), seeded
Outputs of the reference on those inputs:
  tests/golden/ref_structural.npz, ref_clusters.json, ref_unique_conf.json, ref_dynamical.npz, ref_hydration.npz
"""
from __future__ import annotations

import glob
import io
import json
import os
import shutil
import sys
import tarfile
import tempfile
import time

import numpy as np
import pandas as pd
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MASS = [16.0, 12.01, 1.008, 14.01, 32.06, 16.0, 12.01, 19.0, 24.305]
NUM_MOLS = [591, 66, 33]
NUM_ATOMS = [16, 15, 1]
ELEMENTS = ["O", "C", "H", "N", "S", "O", "C", "F", "Mg"]


def _tar_dir(src_dir, out_path):
    with tarfile.open(out_path, "w:gz") as tf:
        for name in sorted(os.listdir(src_dir)):
            tf.add(os.path.join(src_dir, name), arcname=name)


# ------------------------------------------------------------------------------------------
def make_sample_frames(work):
    d = os.path.join(work, "sample")
    os.makedirs(d)
    for ts in (0, 2500000):
        shutil.copy(os.path.join(H.sample_dir(), f"dump.nvt.{ts}.dump"), d)
    _tar_dir(d, os.path.join(GOLD, "sample_frames.tar.gz"))
    return d


def make_mini_traj(work):
    """Cut 3 Mg + their first-shell molecules (+ fillers) out of every 4th sample frame; renumber ids."""
    from oracle import oracle as O

    d = os.path.join(work, "mini")
    os.makedirs(d)
    f0 = next(O.read_dumps(os.path.join(H.sample_dir(), "dump.nvt.0.dump")))
    c = f0.sorted_by_id()
    mt, mi, off = O.mol_membership(NUM_MOLS, NUM_ATOMS)
    mol_of_atom = np.repeat(np.arange(len(mt)), np.diff(off))
    L = f0.lattice_lengths
    mg_mols = [591 + 66 + k for k in range(3)]
    chosen = set()
    for m in mg_mols:
        a = off[m]
        rsq = O.calc_rsq([c["x"][a], c["y"][a], c["z"][a]], c["x"], c["y"], c["z"], L)
        chosen.update(mol_of_atom[rsq < 3.0 ** 2].tolist())
    dme = sorted(m for m in chosen if mt[m] == 1)
    tfsi = sorted(m for m in chosen if mt[m] == 2)
    k = 0
    while len(dme) < 24:
        if k not in dme:
            dme.append(k)
        k += 1
    k = 591
    while len(tfsi) < 6:
        if k not in tfsi:
            tfsi.append(k)
        k += 1
    mols = sorted(dme) + sorted(tfsi) + mg_mols
    keep_ids = np.concatenate([np.arange(off[m], off[m + 1]) + 1 for m in mols])  # original ids, new order
    new_id = {int(o): n + 1 for n, o in enumerate(keep_ids)}
    num_mols = [len(dme), len(tfsi), len(mg_mols)]
    files = O.dump_files(os.path.join(H.sample_dir(), "dump.nvt.*.dump"))[::4]
    for fn in files:
        with open(fn) as f:
            lines = f.read().split("\n")
        head = lines[:9]
        rows = []
        for ln in lines[9:]:
            if not ln.strip():
                continue
            tok = ln.split()
            i = int(tok[0])
            if i in new_id:
                tok[0] = str(new_id[i])
                rows.append(" ".join(tok) + " ")
        head[3] = str(len(rows))
        ts = int(head[1])
        with open(os.path.join(d, f"dump.mini.{ts}.dump"), "w") as f:
            f.write("\n".join(head + rows) + "\n")
    _tar_dir(d, os.path.join(GOLD, "mini_traj.tar.gz"))
    return d, num_mols


def make_visc_logs(work, T=4001, nrep=2):
    d = os.path.join(work, "visc")
    os.makedirs(d)
    rng = np.random.default_rng(20261019)
    for r in range(nrep):
        p = np.zeros((T, 3))
        x = rng.normal(0, 500.0, 3)
        a = np.exp(-1.0 / 50.0)
        for t in range(T):
            x = a * x + np.sqrt(1 - a * a) * rng.normal(0, 500.0, 3)
            p[t] = x
        with open(os.path.join(d, f"log.visc_{r + 1}"), "w") as f:
            f.write("LAMMPS (synthetic)\nunits real\n")
            f.write("Per MPI rank memory allocation (min/avg/max) = 1 | 1 | 1 Mbytes\n")
            f.write("Step Temp Pxy Pxz Pyz \n")
            for t in range(T):
                f.write(f"{t * 5:d} 298.15 {p[t, 0]:.6f} {p[t, 1]:.6f} {p[t, 2]:.6f} \n")
            f.write("Loop time of 1.0 on 1 procs for 20000 steps with 100 atoms\n")
    _tar_dir(d, os.path.join(GOLD, "visc_log.tar.gz"))
    return d


def make_water_box(work, ncat=6, nwat=150, nframes=3, L=18.0):
    """cations (1 atom, mol type 1) then 3-site waters O,H,H (mol type 2); ids contiguous; rows shuffled."""
    d = os.path.join(work, "water")
    os.makedirs(d)
    rng = np.random.default_rng(20261020)
    cat = rng.uniform(0, L, (ncat, 3))
    wo = rng.uniform(0, L, (nwat, 3))
    for fr in range(nframes):
        cat = cat + rng.normal(0, 0.15, cat.shape)
        wo = wo + rng.normal(0, 0.15, wo.shape)
        rows = []
        aid = 1
        for k in range(ncat):
            p = np.mod(cat[k], L)
            rows.append((aid, 1, *p)); aid += 1
        for k in range(nwat):
            o = np.mod(wo[k], L)
            u = rng.normal(size=3); u /= np.linalg.norm(u)
            v = np.cross(u, rng.normal(size=3)); v /= np.linalg.norm(v)
            h1 = o + 0.9572 * (np.cos(0.9122) * u + np.sin(0.9122) * v)
            h2 = o + 0.9572 * (np.cos(0.9122) * u - np.sin(0.9122) * v)
            rows.append((aid, 2, *o)); aid += 1
            rows.append((aid, 3, *h1)); aid += 1
            rows.append((aid, 3, *h2)); aid += 1
        order = rng.permutation(len(rows))
        with open(os.path.join(d, f"dump.water.{fr * 1000}.dump"), "w") as f:
            f.write(f"ITEM: TIMESTEP\n{fr * 1000}\nITEM: NUMBER OF ATOMS\n{len(rows)}\n")
            f.write("ITEM: BOX BOUNDS pp pp pp\n" + f"0.0000000000000000e+00 {L:.16e}\n" * 3)
            f.write("ITEM: ATOMS id type x y z \n")
            for k in order:
                a, t, x, y, z = rows[k]
                f.write(f"{a} {t} {x:.6g} {y:.6g} {z:.6g} \n")
    _tar_dir(d, os.path.join(GOLD, "water_box.tar.gz"))
    return d, [ncat, nwat], [1, 3]


def make_slab_box(work, nframes=2, L=(10.0, 11.0, 14.0)):
    """A 2 A surface slab (type 1) at z in [3, 5] with a two-species fluid (molecules of 2 atoms: types 2, 3) below and
    above it; ids contiguous (slab first), rows shuffled.  Fixture for calc_number_density."""
    d = os.path.join(work, "slab")
    os.makedirs(d)
    rng = np.random.default_rng(20261021)
    nsurf, nmol = 40, 120
    surf = np.column_stack([rng.uniform(0, L[0], nsurf), rng.uniform(0, L[1], nsurf), rng.uniform(3.0, 5.0, nsurf)])
    zf = np.where(rng.uniform(size=nmol) < 0.25, rng.uniform(0.0, 3.0, nmol), rng.uniform(5.0, 13.0, nmol))
    mol = np.column_stack([rng.uniform(0, L[0], nmol), rng.uniform(0, L[1], nmol), zf])
    for fr in range(nframes):
        surf = surf + rng.normal(0, 0.02, surf.shape)
        mol = mol + rng.normal(0, 0.1, mol.shape)
        rows, aid = [], 1
        for p in surf:
            rows.append((aid, 1, *p)); aid += 1
        for p in mol:
            rows.append((aid, 2, *p)); aid += 1
            rows.append((aid, 3, *(p + rng.normal(0, 0.5, 3)))); aid += 1
        order = rng.permutation(len(rows))
        with open(os.path.join(d, f"dump.slab.{fr * 500}.dump"), "w") as f:
            f.write(f"ITEM: TIMESTEP\n{fr * 500}\nITEM: NUMBER OF ATOMS\n{len(rows)}\n")
            f.write("ITEM: BOX BOUNDS pp pp pp\n" + "".join(f"0.0000000000000000e+00 {l:.16e}\n" for l in L))
            f.write("ITEM: ATOMS id type x y z \n")
            for k in order:
                a, t, x, y, z = rows[k]
                f.write(f"{a} {t} {x:.6g} {y:.6g} {z:.6g} \n")
    _tar_dir(d, os.path.join(GOLD, "slab_box.tar.gz"))
    return d, [nsurf, nmol], [1, 2]


def run_number_density(slab_dir, num_mols, num_atoms, out):
    """The reference function needs numpy < 1.24 (np.int, :50; np.product, :118); both aliases are restored for the call --
    the reference source itself is untouched."""
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "product"):
        np.product = np.prod
    from mdproptools.structural.number_density import calc_number_density as _cnd

    def calc_number_density(*a, **k):
        # pandas >= 3 (copy-on-write) hands out read-only ``Series.values``; the reference does ``b -= dist_range`` on the
        # values of a boolean-filtered (hence already copied) Series (:89-96), so a writable copy is what old pandas gave it
        orig = pd.Series.values
        with mock.patch.object(pd.Series, "values", property(lambda self: np.array(orig.fget(self)))):
            return _cnd(*a, **k)

    tmp = tempfile.mkdtemp()
    for f in glob.glob(os.path.join(slab_dir, "*.dump")):
        shutil.copy(f, tmp)
    df = calc_number_density("dump.slab.*.dump", 1, [2, 3, 1], 0.5, 8.0, "z", working_dir=tmp, save_mode=False)
    out["nd_pos"] = df.values; out["nd_pos_cols"] = np.array(list(df.columns))
    df = calc_number_density("dump.slab.*.dump", 1, [2, 3], 0.5, -20.0, "z", working_dir=tmp, save_mode=False)
    out["nd_neg"] = df.values
    # altered ids: slab atoms -> 1, molecule atoms -> 2 and 3 by position in the molecule (rdf_cn.py:197-215)
    df = calc_number_density("dump.slab.*.dump", 1, [3, 2], 0.25, 6.0, "z", num_mols=num_mols, num_atoms_per_mol=num_atoms,
                             working_dir=tmp, save_mode=False)
    out["nd_alt"] = df.values
    df = calc_number_density("dump.slab.*.dump", 1, [2], 0.5, 12.0, "x", working_dir=tmp, save_mode=False)
    out["nd_x"] = df.values
    out["nd_num_mols"] = np.array(num_mols); out["nd_num_atoms"] = np.array(num_atoms)


# ------------------------------------------------------------------------------------------
def run_structural(sample_dir, out):
    import pandas as pd
    from mdproptools.structural import rdf_cn

    f0 = os.path.join(sample_dir, "dump.nvt.0.dump")
    both = os.path.join(sample_dir, "dump.nvt.*.dump")
    rel = [[9, 9, 9, 9], [1, 4, 6, 9]]

    # raw integer histograms straight from the numba kernel, frame 0 (rdf_cn.py:72-97)
    dump = next(H.parse_lammps_dumps(f0))
    ref_df = dump.data[["id", "type", "x", "y", "z"]].sort_values("id").drop("id", axis=1)
    lengths = dump.box.to_lattice().lengths
    full = np.zeros(400); part = np.zeros((4, 400))
    t = time.time()
    rdf_cn._rdf_loop(ref_df.values, np.asarray(rel).transpose(), 4, lengths, 20, 0.05, full, part)
    print("  _rdf_loop frame 0:", time.time() - t, "s (incl. JIT)")
    out["rdf_raw_full_f0"] = full.astype(np.int64)
    out["rdf_raw_part_f0"] = part.astype(np.int64)
    out["box_lengths_f0"] = np.array(lengths)

    # smaller cutoff / odd bin so that edge handling is exercised too
    full = np.zeros(int(7.3 / 0.07)); part = np.zeros((4, int(7.3 / 0.07)))
    rdf_cn._rdf_loop(ref_df.values, np.asarray(rel).transpose(), 4, lengths, 7.3, 0.07, full, part)
    out["rdf_raw_full_f0_rc7p3"] = full.astype(np.int64)
    out["rdf_raw_part_f0_rc7p3"] = part.astype(np.int64)

    df = rdf_cn.calc_atomic_rdf(20, 0.05, 9, MASS, rel, f0, save_mode=False)
    out["atomic_rdf_f0"] = df.values
    out["atomic_rdf_columns"] = np.array(list(df.columns))
    df = rdf_cn.calc_atomic_rdf(20, 0.05, 9, MASS, rel, both, save_mode=False)
    out["atomic_rdf_2frames"] = df.values
    df = rdf_cn.calc_atomic_rdf(12, 0.05, 9, MASS, [[32, 32], [17, 32]], f0, num_mols=NUM_MOLS,
                                num_atoms_per_mol=NUM_ATOMS, save_mode=False)
    out["atomic_rdf_altered_f0"] = df.values
    out["atomic_rdf_altered_columns"] = np.array(list(df.columns))

    df = rdf_cn.calc_atomic_cn([2.325, 4.375, 2.375, 13.0], 0.05, 9, MASS, rel, f0, save_mode=False)
    out["atomic_cn_f0"] = df.values
    out["atomic_cn_columns"] = np.array(list(df.columns))
    df = rdf_cn.calc_atomic_cn([2.325, 4.375, 2.375, 13.0], 0.05, 9, MASS, rel, both, save_mode=False)
    out["atomic_cn_2frames"] = df.values
    df = rdf_cn.calc_atomic_cn([4.375, 13.0], 0.05, 9, MASS, [[32, 32], [17, 32]], f0, num_mols=NUM_MOLS,
                               num_atoms_per_mol=NUM_ATOMS, save_mode=False)
    out["atomic_cn_altered_f0"] = df.values

    mrel = [[9, 9, 4], [1, 2, 3]]
    df = rdf_cn.calc_molecular_rdf(20, 0.05, 9, MASS, mrel, f0, NUM_MOLS, NUM_ATOMS, save_mode=False)
    out["molecular_rdf_f0"] = df.values
    out["molecular_rdf_columns"] = np.array(list(df.columns))
    df = rdf_cn.calc_molecular_cn([2.325, 3.775, 4.375], 0.05, 9, MASS, mrel, f0, NUM_MOLS, NUM_ATOMS,
                                  save_mode=False)
    out["molecular_cn_f0"] = df.values
    # (calc_intermolecular_rdf is exercised on the mini trajectory instead: its _calc_props call uses
    #  the molecule table as 'atoms' and needs num_types == number of molecule types.)

    # molecule COMs of frame 0 as the reference computes them (BLAS dot order) -- rdf_cn.py:218-241
    d0 = dump.data[["id", "type", "x", "y", "z"]].sort_values("id").copy()
    mol_df = rdf_cn._define_mol_cols(d0, NUM_MOLS, NUM_ATOMS, MASS)
    out["mol_com_f0"] = mol_df.values


def run_clusters(sample_dir, work):
    from mdproptools.structural.cluster_analysis import get_clusters

    wd = os.path.join(work, "clusters")
    os.makedirs(wd)
    n = get_clusters(filename=os.path.join(sample_dir, "dump.nvt.*.dump"), atom_type=9, r_cut=2.3,
                     num_mols=NUM_MOLS, num_atoms_per_mol=NUM_ATOMS, full_trajectory=False, frame=1,
                     elements=ELEMENTS, alter_atom_types=False, max_force=0.75, working_dir=wd)
    files = {os.path.basename(p): open(p).read() for p in sorted(glob.glob(os.path.join(wd, "Cluster_*.xyz")))}
    # cross-check against the reference's own golden files (tests/structural/test_files)
    ref_dir = os.path.join(H.REFERENCE_ROOT, "tests", "structural", "test_files")
    same = sum(open(os.path.join(ref_dir, k)).read() == v for k, v in files.items())
    print(f"  get_clusters: {n} clusters, {same}/{len(files)} byte-identical to the reference's own goldens")
    wd2 = os.path.join(work, "clusters_alt")
    os.makedirs(wd2)
    n2 = get_clusters(filename=os.path.join(sample_dir, "dump.nvt.*.dump"), atom_type=32, r_cut=2.3,
                      num_mols=NUM_MOLS, num_atoms_per_mol=NUM_ATOMS, full_trajectory=True, frame=None,
                      elements=ELEMENTS, alter_atom_types=True, max_force=0.75, working_dir=wd2)
    files2 = {os.path.basename(p): open(p).read() for p in sorted(glob.glob(os.path.join(wd2, "Cluster_*.xyz")))}
    with open(os.path.join(GOLD, "ref_clusters.json"), "w") as f:
        json.dump({"frame1_type9": {"count": n, "files": files, "identical_to_reference_goldens": same},
                   "full_altered32": {"count": n2, "files": files2}}, f)


def run_dynamical(mini_dir, num_mols, visc_dir, out):
    import pandas as pd
    from mdproptools.dynamical.diffusion import Diffusion
    from mdproptools.dynamical.conductivity import Conductivity
    from mdproptools.dynamical.residence_time import ResidenceTime
    from mdproptools.dynamical.viscosity import Viscosity
    from mdproptools.structural import rdf_cn

    out["mini_num_mols"] = np.array(num_mols)
    tmp = tempfile.mkdtemp()
    d = Diffusion(timestep=1, units="real", outputs_dir=mini_dir, diff_dir=tmp)
    msd, msd_all, msd_int = d.get_msd_from_dump("dump.mini.*.dump", msd_type="com", num_mols=num_mols,
                                                num_atoms_per_mol=NUM_ATOMS, mass=MASS, com_drift=True,
                                                avg_interval=True, tao_coeff=4)
    out["msd_com_cols"] = np.array(list(msd.columns)); out["msd_com"] = msd.values
    out["msd_all_com_cols"] = np.array(list(msd_all.columns)); out["msd_all_com"] = msd_all.values
    out["msd_int_com_cols"] = np.array(list(msd_int.columns)); out["msd_int_com"] = msd_int.values
    diff = d.calc_diff(msd)
    out["diff_com"] = diff.values
    msd, msd_all = d.get_msd_from_dump("dump.mini.*.dump", msd_type="com", num_mols=num_mols,
                                       num_atoms_per_mol=NUM_ATOMS, mass=MASS, com_drift=False)
    out["msd_com_nodrift"] = msd.values
    msd, msd_all, msd_int = d.get_msd_from_dump("dump.mini.*.dump", msd_type="allatom", avg_interval=True,
                                                tao_coeff=4)
    out["msd_allatom_cols"] = np.array(list(msd.columns)); out["msd_allatom"] = msd.values
    out["msd_all_allatom_cols"] = np.array(list(msd_all.columns)); out["msd_all_allatom"] = msd_all.values
    out["msd_int_allatom_cols"] = np.array(list(msd_int.columns)); out["msd_int_allatom"] = msd_int.values
    diff = d.calc_diff(msd, initial_time={0: 1e-9}, final_time={0: 4e-9})
    out["diff_allatom_window"] = diff.values

    c = Conductivity("dump.mini.*.dump", num_mols, NUM_ATOMS, volume=49.182348836183905 ** 3, mass=MASS,
                     temp=298.15, timestep=1, units="real", working_dir=mini_dir)
    j = c.get_charge_flux()
    tot = c.correlate_charge_flux(j)
    integ = c.integrate_charge_flux_correlation(tot)
    out["cond_flux"] = j; out["cond_time"] = np.array(c.time); out["cond_tot_flux"] = tot; out["cond_integral"] = integ
    out["cond_green_kubo_of_last"] = c.green_kubo(integ[:, -1])

    # shells chosen so that membership actually toggles in the 26 frames: Mg-O(DME) distances fluctuate
    # around 1.85-2.05 A, Mg-O(TFSI, altered type 27) around 2.1 A; the third relation is a like pair (k == l,
    # self exclusion at residence_time.py:103-104)
    rt = ResidenceTime([[0, 2.0], [1.9, 2.2], [0, 3.2]], [[32, 32, 1], [1, 27, 1]],
                       os.path.join(mini_dir, "dump.mini.*.dump"),
                       dt=1, num_mols=num_mols, num_atoms_per_mol=NUM_ATOMS, working_dir=tmp)
    rt.calc_auto_correlation()
    out["residence_cols"] = np.array(list(rt.corr_df.columns)); out["residence_corr"] = rt.corr_df.values

    df = rdf_cn.calc_intermolecular_rdf(20, 0.05, 3, MASS, [[3, 3, 2], [1, 2, 2]],
                                        os.path.join(mini_dir, "dump.mini.0.dump"), num_mols, NUM_ATOMS,
                                        save_mode=False)
    out["intermolecular_rdf_mini_f0"] = df.values

    v = Viscosity("log.visc_*", cutoff_time=500, volume=40.0 ** 3, temp=298.15, timestep=1, acf_method="wkt",
                  units="real", working_dir=visc_dir)
    # glob order is filesystem dependent in the reference (viscosity.py:209); pin it for the fixture
    import mdproptools.dynamical.viscosity as vmod
    real_glob = vmod.glob.glob
    vmod.glob.glob = lambda p: sorted(real_glob(p))
    visc_avg, visc_data, acf_data, tvec = v.calc_avg_visc(output_all_data=True)
    vmod.glob.glob = real_glob
    out["visc_avg"] = np.array(visc_avg); out["visc_data"] = np.array(visc_data)
    out["visc_acf"] = np.array(acf_data); out["visc_time"] = np.array(tvec)
    series = np.array(acf_data)[0, 0, :200]
    out["visc_acf_bruteforce_first200_in"] = series
    out["visc_acf_bruteforce_first200"] = Viscosity.autocorrelate(series, "brute_force")


def run_hydration(water_dir, num_mols, num_atoms, out):
    from hydration_number import get_hydration_number  # bare-import module (hydration_number.py:8)

    tmp = tempfile.mkdtemp()
    for f in glob.glob(os.path.join(water_dir, "*.dump")):
        shutil.copy(f, tmp)
    df = get_hydration_number("dump.water.*.dump", cation_type=1, water_type=2, r_cut=5.0, alter_atom_ids=False,
                              num_mols=num_mols, num_atoms_per_mol=num_atoms, working_dir=tmp)
    out["hyd_angles"] = df["angles_distribution"].values
    out["hyd_factor"] = df["hydration_factor"].values[:1]
    out["hyd_num_mols"] = np.array(num_mols); out["hyd_num_atoms"] = np.array(num_atoms)


def run_unique_configurations(work):
    """get_unique_configurations (cluster_analysis.py:238-457) exactly as the reference's own test drives it
    (tests/structural/test_cluster_analysis.py:62-100): clusters of frame 50 around the altered atom type 32, then the
    configuration table.  The reference's five conf_*.xyz goldens are checked here; its csv goldens are git-lfs stubs,
    so the DataFrames of this run are what gets stored (together with the cluster files that are the function's input)."""
    import pandas as pd
    from mdproptools.structural.cluster_analysis import get_clusters, get_unique_configurations
    data_dir = os.path.join(H.REFERENCE_ROOT, "data", "mg_tfsi_dme")
    wd = os.path.join(work, "uniq")
    os.makedirs(wd)
    get_clusters(filename=os.path.join(data_dir, "dump.nvt.*.dump"), atom_type=32, r_cut=2.3, num_mols=NUM_MOLS,
                 num_atoms_per_mol=NUM_ATOMS, full_trajectory=False, frame=50, elements=ELEMENTS, alter_atom_types=True,
                 max_force=0.75, working_dir=wd)
    inputs = {os.path.basename(p): open(p).read() for p in sorted(glob.glob(os.path.join(wd, "Cluster_*.xyz")))}
    mols = [H.ShimMolecule.from_file(os.path.join(data_dir, f)) for f in ("dme.pdb", "tfsi.pdb", "mg.pdb")]
    species = [[str(x) for x in m.species] for m in mols]
    df, df1 = get_unique_configurations(cluster_pattern="Cluster_*.xyz", r_cut=2.3, molecules=mols, mol_num=2,
                                        type_coord_atoms=["O", "N", "Mg"], working_dir=wd, find_top=True, perc=None,
                                        cum_perc=100, mol_names=["dme", "tfsi", "mg"], zip=False)
    ref_dir = os.path.join(H.REFERENCE_ROOT, "tests", "structural", "test_files")
    conf = {os.path.basename(p): open(p).read() for p in sorted(glob.glob(os.path.join(wd, "conf_*.xyz")))}
    same = sum(open(os.path.join(ref_dir, k)).read() == v for k, v in conf.items())
    print(f"  get_unique_configurations: {len(inputs)} clusters, {len(conf)} conf files, {same} byte-identical to the "
          f"reference's own goldens; counts {df1['count'].tolist()}")
    with open(os.path.join(GOLD, "ref_unique_conf.json"), "w") as f:
        json.dump({"cluster_files": inputs, "species": species, "conf_files": conf, "conf_identical_to_reference_goldens": same,
                   "clusters_csv": open(os.path.join(wd, "clusters.csv")).read(),
                   "configurations_csv": open(os.path.join(wd, "configurations.csv")).read(),
                   "top_conf_csv": open(os.path.join(wd, "top_conf.csv")).read(),
                   "counts": df1["count"].tolist(), "percent": df1["%"].tolist()}, f)


def main():
    H.install()
    os.makedirs(GOLD, exist_ok=True)
    work = tempfile.mkdtemp(prefix="golden_work_")
    print("work dir", work)
    sample_dir = make_sample_frames(work)
    mini_dir, mini_num_mols = make_mini_traj(work)
    visc_dir = make_visc_logs(work)
    water_dir, w_mols, w_atoms = make_water_box(work)
    slab_dir, s_mols, s_atoms = make_slab_box(work)
    which = sys.argv[1:] or ["structural", "clusters", "unique", "dynamical", "hydration", "density"]
    if "structural" in which:
        out = {}
        run_structural(sample_dir, out)
        np.savez_compressed(os.path.join(GOLD, "ref_structural.npz"), **out)
    if "clusters" in which:
        run_clusters(sample_dir, work)
    if "unique" in which:
        run_unique_configurations(work)
    if "dynamical" in which:
        out = {}
        run_dynamical(mini_dir, mini_num_mols, visc_dir, out)
        np.savez_compressed(os.path.join(GOLD, "ref_dynamical.npz"), **out)
    if "hydration" in which:
        out = {}
        run_hydration(water_dir, w_mols, w_atoms, out)
        np.savez_compressed(os.path.join(GOLD, "ref_hydration.npz"), **out)
    if "density" in which:
        out = {}
        run_number_density(slab_dir, s_mols, s_atoms, out)
        np.savez_compressed(os.path.join(GOLD, "ref_number_density.npz"), **out)
    print("done")


if __name__ == "__main__":
    main()
