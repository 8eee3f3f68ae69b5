"""Pins the CPU oracle (oracle/oracle.c + oracle/oracle.py) to outputs of the UNMODIFIED reference.

The fixtures under tests/golden/ were produced by oracle/gen_golden.py, which imports the reference from
/root/reference and runs its own functions (numba kernels included) on two real frames of the bundled
Mg-TFSI/DME trajectory, on a 26-frame sub-trajectory cut from it, and on seeded synthetic inputs.
Integer histograms must match bit for bit; floats that only pass through the host normalisation must be
identical; quantities whose summation order is library dependent (BLAS dot, pandas Kahan sums, FFT) carry an
explicit tolerance.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import MASS, NUM_ATOMS, NUM_MOLS

REL = [[9, 9, 9, 9], [1, 4, 6, 9]]


@pytest.fixture(scope="module")
def frames(sample_dir):
    return list(O.read_dumps(os.path.join(sample_dir, "dump.nvt.*.dump")))


def test_frame_order_and_box(frames, gold_structural):
    assert [f.timestep for f in frames] == [0, 2500000]
    assert frames[0].natoms == 10479
    assert np.array_equal(np.array(frames[0].lattice_lengths), gold_structural["box_lengths_f0"])


def test_rdf_loop_counts_bit_exact(frames, gold_structural):
    c = frames[0].sorted_by_id()
    rel = np.asarray(REL).T
    full, part = O.rdf_loop(c["type"], c["x"], c["y"], c["z"], rel, frames[0].lattice_lengths, 20, 0.05, 400, nthreads=0)
    assert np.array_equal(full, gold_structural["rdf_raw_full_f0"])
    assert np.array_equal(part, gold_structural["rdf_raw_part_f0"])
    # known answers recorded in SURVEY.md 8(c) from the unmodified reference
    assert full.sum() == 30926986
    assert part.sum(axis=1).tolist() == [10670, 782, 3127, 358]
    assert full[40] == 3170 and full[399] == 230684 and np.nonzero(full)[0][0] == 19
    sha = hashlib.sha256(np.concatenate([full, part.ravel()]).astype(np.int64).tobytes()).hexdigest()
    assert sha == "fff0ecea439519c02628f45eec96844689a9a21891beab179eaf7bb781ea04ef"


def test_rdf_loop_odd_bins(frames, gold_structural):
    c = frames[0].sorted_by_id()
    nb = int(7.3 / 0.07)
    full, part = O.rdf_loop(c["type"], c["x"], c["y"], c["z"], np.asarray(REL).T, frames[0].lattice_lengths, 7.3, 0.07, nb,
                            nthreads=0)
    assert np.array_equal(full, gold_structural["rdf_raw_full_f0_rc7p3"])
    # int(7.3/0.07) = 104 but pairs with 7.28 <= r < 7.3 have bin index 104: the reference writes them out of
    # bounds, i.e. rdf_part[k][104] lands in rdf_part[k+1][0] (rdf_full[104] falls off the array).  The oracle
    # and the CUDA path drop such pairs (documented divergence); the golden file shows exactly that corruption.
    _, part_wide = O.rdf_loop(c["type"], c["x"], c["y"], c["z"], np.asarray(REL).T, frames[0].lattice_lengths, 7.3, 0.07,
                              nb + 1, nthreads=0)
    assert np.array_equal(part_wide[:, :nb], part)
    expect = part.copy()
    expect[1:, 0] += part_wide[:-1, nb]
    assert part_wide[:, nb].sum() > 0
    assert np.array_equal(expect, gold_structural["rdf_raw_part_f0_rc7p3"])


def test_atomic_rdf_floats_identical(frames, gold_structural):
    out = O.atomic_rdf(frames[:1], 20, 0.05, REL)
    assert np.array_equal(out, gold_structural["atomic_rdf_f0"])
    out2 = O.atomic_rdf(frames, 20, 0.05, REL)
    assert np.array_equal(out2, gold_structural["atomic_rdf_2frames"])


def test_atomic_rdf_altered_ids(frames, gold_structural):
    out = O.atomic_rdf(frames[:1], 12, 0.05, [[32, 32], [17, 32]], num_mols=NUM_MOLS, num_atoms_per_mol=NUM_ATOMS)
    assert np.array_equal(out, gold_structural["atomic_rdf_altered_f0"])


def test_atomic_cn(frames, gold_structural):
    cn = O.atomic_cn(frames[:1], [2.325, 4.375, 2.375, 13.0], REL)
    assert np.array_equal(cn, gold_structural["atomic_cn_f0"][0])
    assert np.allclose(cn * 33, [141, 41, 57, 120], rtol=0, atol=1e-12)          # SURVEY 8(c) known answer
    cn2 = O.atomic_cn(frames, [2.325, 4.375, 2.375, 13.0], REL)
    assert np.array_equal(cn2, gold_structural["atomic_cn_2frames"][0])
    cna = O.atomic_cn(frames[:1], [4.375, 13.0], [[32, 32], [17, 32]], num_mols=NUM_MOLS, num_atoms_per_mol=NUM_ATOMS)
    assert np.array_equal(cna, gold_structural["atomic_cn_altered_f0"][0])


def test_molecule_com_and_molecular_rdf(frames, gold_structural):
    c = frames[0].sorted_by_id()
    mt, mx, my, mz = O.mol_com_wrapped(c["type"], c["x"], c["y"], c["z"], NUM_MOLS, NUM_ATOMS, MASS)
    g = gold_structural["mol_com_f0"]
    assert np.array_equal(mt, g[:, 0])
    # the reference uses a BLAS dot (order library dependent): a few ulp of the coordinate (|x| < 50)
    assert np.max(np.abs(np.stack([mx, my, mz], 1) - g[:, 1:])) < 1e-13
    out = O.molecular_rdf(frames[:1], 20, 0.05, [[9, 9, 4], [1, 2, 3]], NUM_MOLS, NUM_ATOMS, MASS)
    ref = gold_structural["molecular_rdf_f0"]
    # identical unless a COM sits within an ulp of a bin edge; allow at most a couple of moved counts
    assert np.array_equal(out[:, 0], ref[:, 0])
    assert np.count_nonzero(out != ref) <= 4
    cn = O.molecular_cn(frames[:1], [2.325, 3.775, 4.375], [[9, 9, 4], [1, 2, 3]], NUM_MOLS, NUM_ATOMS, MASS)
    assert np.allclose(cn, gold_structural["molecular_cn_f0"][0], rtol=0, atol=1e-12)
    assert np.allclose(cn, [58 / 33, 2 / 33, 41 / 66], rtol=1e-15)             # SURVEY 8(c) known answer


def test_calc_atom_type_restatement():
    ids = np.arange(1, sum(n * a for n, a in zip(NUM_MOLS, NUM_ATOMS)) + 1, dtype=np.float64)
    out = O.calc_atom_type(ids, NUM_MOLS, NUM_ATOMS)
    # literal transcription of the reference's double loop (rdf_cn.py:197-215) on a subsample
    cut = np.cumsum(np.multiply(NUM_MOLS, NUM_ATOMS))
    for n in list(range(0, len(ids), 97)) + [len(ids) - 1]:
        v = ids[n]
        for i, c in enumerate(cut):
            if v <= c:
                v = (v - c) % NUM_ATOMS[i]
                if v == 0:
                    v = NUM_ATOMS[i]
                if i > 0:
                    v += np.sum(NUM_ATOMS[:i])
                break
        assert out[n] == v
    assert len(np.unique(out)) == sum(NUM_ATOMS)


# ---- dynamical ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mini(mini_dir):
    return list(O.read_dumps(os.path.join(mini_dir, "dump.mini.*.dump")))


def test_msd_allatom(mini, gold_dynamical):
    traj = np.stack([np.stack([f.sorted_by_id()[c] for c in ("xu", "yu", "zu")]) for f in mini])
    per_atom, mean = O.msd_single_origin(traj, 0, 1e-10)
    g = gold_dynamical["msd_all_allatom"]            # Time, id, dx2, dy2, dz2, msd sorted by (Time, id)
    T, _, N = traj.shape
    assert g.shape[0] == T * N
    assert np.array_equal(per_atom.transpose(0, 2, 1).reshape(T * N, 4), g[:, 2:])
    assert np.allclose(mean, gold_dynamical["msd_allatom"][:, 1:], rtol=1e-13, atol=0)   # pandas mean = Kahan sum
    mi = O.msd_interval(traj, 1e-10, 4)
    assert np.allclose(mi.T, gold_dynamical["msd_int_allatom"][:, 1:], rtol=1e-13, atol=0)


def test_ols_matches_statsmodels_formulas(gold_dynamical):
    msd = gold_dynamical["msd_com"]
    cols = list(gold_dynamical["msd_com_cols"])
    for k, name in enumerate(["msd1", "msd2", "msd3"]):
        slope, bse, r2 = O.ols_origin(msd[:, 0], msd[:, cols.index(name)])
        assert np.allclose([slope / 6, bse / 6, r2], gold_dynamical["diff_com"][k], rtol=1e-12)


def test_charge_flux_and_correlation(mini, gold_dynamical):
    num_mols = gold_dynamical["mini_num_mols"].tolist()
    mt, mi, off = O.mol_membership(num_mols, NUM_ATOMS)
    c0 = mini[0].sorted_by_id()
    m_atom = np.array([MASS[int(t) - 1] for t in c0["type"]])
    vel = np.stack([np.stack([f.sorted_by_id()[c] for c in ("vx", "vy", "vz")]) for f in mini])
    type_off = np.concatenate(([0], np.cumsum(num_mols)))
    J = O.charge_flux(vel, m_atom, c0["q"], off, type_off, 1e-10 / 1e-15, 1.602176634e-19)
    # neutral molecule types give J ~ rounding noise, so the tolerance is relative to max|J|
    assert np.max(np.abs(J - gold_dynamical["cond_flux"])) < 1e-12 * np.abs(gold_dynamical["cond_flux"]).max()
    tot = np.zeros((4, J.shape[2]))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                corr = O.correlate_fft(J[k, i], J[k, j])
                tot[i] += corr
                tot[-1] += corr
    scale = np.abs(gold_dynamical["cond_tot_flux"]).max()
    assert np.max(np.abs(tot - gold_dynamical["cond_tot_flux"])) < 1e-12 * scale
    # the direct long-double sum agrees with the FFT form to FFT round-off
    d = O.xcorr_direct(J[0, 0], J[0, 1])
    f = O.correlate_fft(J[0, 0], J[0, 1])
    assert np.max(np.abs(d - f)) < 1e-13 * np.abs(f).max()
    t = gold_dynamical["cond_time"]
    integ = np.stack([O.cumtrapz(r, t[1] - t[0], True) for r in gold_dynamical["cond_tot_flux"]])
    assert np.allclose(integ, gold_dynamical["cond_integral"], rtol=1e-13, atol=0)


def test_residence_correlation(mini, gold_dynamical):
    num_mols = gold_dynamical["mini_num_mols"].tolist()
    cols = list(gold_dynamical["residence_cols"])
    ref = gold_dynamical["residence_corr"]
    assert cols[1:] == ["32-1", "32-27", "1-1"]
    for name, (k, l), (r_in, r_out) in zip(cols[1:], [(32, 1), (32, 27), (1, 1)], [(0, 2.0), (1.9, 2.2), (0, 3.2)]):
        xa, xb, L = [], [], []
        for f in mini:
            c = f.sorted_by_id()
            typ = O.calc_atom_type(c["id"], num_mols, NUM_ATOMS)
            a, b = typ == k, typ == l
            xa.append((c["x"][a], c["y"][a], c["z"][a]))
            xb.append((c["x"][b], c["y"][b], c["z"][b]))
            L.append(f.lattice_lengths)
        corr, cnt = O.residence_correlation(xa, xb, L, r_in, r_out, k == l)
        assert cnt[0] > 0 and len(np.unique(corr)) > 5                             # the fixture is not degenerate
        assert np.allclose(corr, ref[:, cols.index(name)], rtol=0, atol=1e-12)     # reference = FFT round-off
    assert np.allclose(ref[:, 0], [f.timestep * 1e-3 for f in mini])


def test_viscosity_acf_and_integral(visc_dir, gold_dynamical):
    acf = gold_dynamical["visc_acf"]                # [rep, 3, T'] already times PRESSURE_CONVERSION**2
    series = gold_dynamical["visc_acf_bruteforce_first200_in"]
    bf = gold_dynamical["visc_acf_bruteforce_first200"]
    d = O.xcorr_direct(series, series)
    assert np.allclose(d, bf, rtol=1e-12, atol=1e-12 * np.abs(bf).max())
    # running integral of the stored ACF reproduces the stored viscosity
    V = 40.0 ** 3 * (1e-10) ** 3
    kB = 1.380649 * 10 ** -23
    dt = 5 * 1e-15
    visc = np.stack([np.multiply(V / (kB * 298.15), O.cumtrapz(a, dt, False)) for a in acf[0]])
    ref = gold_dynamical["visc_data"][0]
    assert np.max(np.abs(visc - ref)) < 1e-11 * np.abs(ref).max()


def test_msd_all_origins_oracle_consistency():
    rng = np.random.default_rng(5)
    traj = np.cumsum(rng.normal(0, 0.1, (40, 3, 17)), axis=0)
    out = O.msd_all_origins(traj, 40)
    # lag 0 is zero; the single-origin MSD is the t0 = 0 term of the all-origins average
    assert np.all(out[0] == 0)
    lag = 7
    d = traj[lag:] - traj[:-lag]
    assert np.allclose(out[lag, :3], (d ** 2).mean(axis=(0, 2)), rtol=1e-13)
    assert np.allclose(out[lag, 3], (d ** 2).sum(axis=1).mean(), rtol=1e-13)


def test_triclinic_image_is_the_nearest_image_within_half_the_cell_width():
    """mic="triclinic" is an extension with no reference implementation: the oracle's sequential z,y,x shift
    (oracle.c pair_rsq_tri) is checked against a 27-image brute force wherever the nearest-image distance is
    below half the smallest cell width, for wrapped points and LAMMPS-legal tilts."""
    rng = np.random.default_rng(5)
    for cell in [(30.0, 28.0, 26.0, 6.0, 3.0, -4.2), (20.0, 20.0, 20.0, 10.0, -10.0, 10.0), (167.19, 167.19, 167.19, 33.438, 16.719, -25.0785)]:
        lx, ly, lz, xy, xz, yz = cell
        n = 3000
        s = rng.uniform(0, 1, (n, 3))
        pos = s[:, 0:1] * np.array([lx, 0, 0]) + s[:, 1:2] * np.array([xy, ly, 0]) + s[:, 2:3] * np.array([xz, yz, lz])
        x, y, z = np.ascontiguousarray(pos.T)
        a, b, c = np.array([lx, 0, 0]), np.array([xy, ly, 0]), np.array([xz, yz, lz])
        vol = lx * ly * lz
        widths = [vol / np.linalg.norm(np.cross(b, c)), vol / np.linalg.norm(np.cross(c, a)), vol / np.linalg.norm(np.cross(a, b))]
        rmax2 = (0.5 * min(widths)) ** 2
        for h in range(0, n, 250):
            got = O.calc_rsq_tri(pos[h], x, y, z, cell)
            ref = O.nearest_image_rsq_bruteforce(pos[h], x, y, z, cell)
            m = ref < rmax2
            assert m.sum() > 10
            assert np.allclose(got[m], ref[m], rtol=1e-12, atol=1e-12)
            assert (got >= ref - 1e-9).all()      # never closer than the true nearest image
    # zero tilt: identical bits to the reference's orthogonal single shift
    x, y, z = rng.uniform(0, 25, (3, 2000))
    assert np.array_equal(O.calc_rsq_tri((3.0, 4.0, 5.0), x, y, z, (25.0, 25.0, 25.0, 0, 0, 0)),
                          O.calc_rsq((3.0, 4.0, 5.0), x, y, z, (25.0, 25.0, 25.0)))


def test_number_density_oracle_vs_reference(slab_dir, gold_density):
    """calc_number_density (number_density.py:30-154): positive and negative distance, altered ids, another axis; the
    reference output is reproduced to the last bit (integer counts, same division order)."""
    frames = list(O.read_dumps(os.path.join(slab_dir, "dump.slab.*.dump")))
    assert len(frames) == 2
    g = gold_density
    assert np.array_equal(O.number_density(frames, 1, [2, 3, 1], 0.5, 8.0, "z"), g["nd_pos"])
    assert np.array_equal(O.number_density(frames, 1, [2, 3], 0.5, -20.0, "z"), g["nd_neg"])
    assert np.array_equal(O.number_density(frames, 1, [3, 2], 0.25, 6.0, "z", num_mols=g["nd_num_mols"].tolist(),
                                           num_atoms_per_mol=g["nd_num_atoms"].tolist()), g["nd_alt"])
    assert np.array_equal(O.number_density(frames, 1, [2], 0.5, 12.0, "x"), g["nd_x"])
    assert g["nd_pos"][:, 1:].sum() > 0 and g["nd_neg"][:, 1:].sum() > 0
