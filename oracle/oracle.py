"""CPU oracle for the mdproptools hot path (numpy + the C restatement in ``oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never by ``mdproptools_b200``.

Parity status: PINNED against fixtures produced by the unmodified reference
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; checked in ``tests/test_oracle_golden.py``).
Two north-star extensions have no reference implementation (true triclinic minimum image, MSD over
all time origins); for those this file *is* the definition and their parity is unpinned.

All ``file:line`` citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import ctypes
import glob as _glob
import gzip
import os
import re
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int64)
c_bp = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_max_threads.restype = ctypes.c_int
    return _LIB


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_ip)


# ----------------------------------------------------------------------------------------------
# pair kernels (C)
# ----------------------------------------------------------------------------------------------
def rcut_sq(r_cut) -> float:
    """``r_cut ** 2`` as numba evaluates it (rdf_cn.py:66): exact for ints, x*x for floats."""
    if isinstance(r_cut, (int, np.integer)):
        return float(int(r_cut) ** 2)
    r = float(r_cut)
    return r * r


def rdf_loop(typ, x, y, z, rel, lengths, r_cut, ddr, nb, nthreads=1):
    """_rdf_loop (rdf_cn.py:72-97) -> (full int64[nb], part int64[R, nb])."""
    typ, tp = _d(typ); x, xp = _d(x); y, yp = _d(y); z, zp = _d(z)
    rel, rp = _i(np.asarray(rel).reshape(-1, 2))
    R = rel.shape[0]
    full = np.zeros(nb); part = np.zeros((R, nb))
    lib().orc_rdf_loop(tp, xp, yp, zp, ctypes.c_int64(len(x)), rp, ctypes.c_int64(R),
                       ctypes.c_double(lengths[0]), ctypes.c_double(lengths[1]), ctypes.c_double(lengths[2]),
                       ctypes.c_double(rcut_sq(r_cut)), ctypes.c_double(ddr), ctypes.c_int64(nb),
                       full.ctypes.data_as(c_dp), part.ctypes.data_as(c_dp), ctypes.c_int(nthreads))
    return full.astype(np.int64), part.astype(np.int64)


def cn_loop(typ, x, y, z, rel, lengths, r_cuts, nthreads=1):
    """_cn_loop (rdf_cn.py:100-119) -> int64[R]."""
    typ, tp = _d(typ); x, xp = _d(x); y, yp = _d(y); z, zp = _d(z)
    rel, rp = _i(np.asarray(rel).reshape(-1, 2))
    R = rel.shape[0]
    rc2, rcp = _d([rcut_sq(r) for r in r_cuts])
    cn = np.zeros(R)
    lib().orc_cn_loop(tp, xp, yp, zp, ctypes.c_int64(len(x)), rp, ctypes.c_int64(R),
                      ctypes.c_double(lengths[0]), ctypes.c_double(lengths[1]), ctypes.c_double(lengths[2]),
                      rcp, cn.ctypes.data_as(c_dp), ctypes.c_int(nthreads))
    return cn.astype(np.int64)


def rdf_rect(ta, xa, ya, za, tb, xb, yb, zb, rel, lengths, r_cut, ddr, nb, nthreads=1):
    """_rdf_mol_loop (rdf_cn.py:122-141) -> int64[R, nb]."""
    ta, tap = _d(ta); xa, xap = _d(xa); ya, yap = _d(ya); za, zap = _d(za)
    tb, tbp = _d(tb); xb, xbp = _d(xb); yb, ybp = _d(yb); zb, zbp = _d(zb)
    rel, rp = _i(np.asarray(rel).reshape(-1, 2))
    R = rel.shape[0]
    part = np.zeros((R, nb))
    lib().orc_rdf_rect(tap, xap, yap, zap, ctypes.c_int64(len(xa)), tbp, xbp, ybp, zbp, ctypes.c_int64(len(xb)),
                       rp, ctypes.c_int64(R), ctypes.c_double(lengths[0]), ctypes.c_double(lengths[1]),
                       ctypes.c_double(lengths[2]), ctypes.c_double(rcut_sq(r_cut)), ctypes.c_double(ddr),
                       ctypes.c_int64(nb), part.ctypes.data_as(c_dp), ctypes.c_int(nthreads))
    return part.astype(np.int64)


def cn_rect(ta, xa, ya, za, tb, xb, yb, zb, rel, lengths, r_cuts, nthreads=1):
    """_cn_mol_loop (rdf_cn.py:144-162) -> int64[R]."""
    ta, tap = _d(ta); xa, xap = _d(xa); ya, yap = _d(ya); za, zap = _d(za)
    tb, tbp = _d(tb); xb, xbp = _d(xb); yb, ybp = _d(yb); zb, zbp = _d(zb)
    rel, rp = _i(np.asarray(rel).reshape(-1, 2))
    R = rel.shape[0]
    rc2, rcp = _d([rcut_sq(r) for r in r_cuts])
    cn = np.zeros(R)
    lib().orc_cn_rect(tap, xap, yap, zap, ctypes.c_int64(len(xa)), tbp, xbp, ybp, zbp, ctypes.c_int64(len(xb)),
                      rp, ctypes.c_int64(R), ctypes.c_double(lengths[0]), ctypes.c_double(lengths[1]),
                      ctypes.c_double(lengths[2]), rcp, cn.ctypes.data_as(c_dp), ctypes.c_int(nthreads))
    return cn.astype(np.int64)


def calc_rsq(head, x, y, z, lengths):
    """_calc_rsq (rdf_cn.py:35-58): rsq of one head point against M others."""
    head, hp = _d(head); x, xp = _d(x); y, yp = _d(y); z, zp = _d(z)
    out = np.empty(len(x))
    lib().orc_calc_rsq(hp, xp, yp, zp, ctypes.c_int64(len(x)), ctypes.c_double(lengths[0]),
                       ctypes.c_double(lengths[1]), ctypes.c_double(lengths[2]), out.ctypes.data_as(c_dp))
    return out


def shell_mask(xa, ya, za, xb, yb, zb, lengths, r_in, r_out, same_set):
    """h = (rsq > r_in**2) & (rsq <= r_out**2), self cleared (residence_time.py:100-104)."""
    xa, xap = _d(xa); ya, yap = _d(ya); za, zap = _d(za)
    xb, xbp = _d(xb); yb, ybp = _d(yb); zb, zbp = _d(zb)
    out = np.zeros((len(xa), len(xb)), dtype=np.uint8)
    lib().orc_shell_mask(xap, yap, zap, ctypes.c_int64(len(xa)), xbp, ybp, zbp, ctypes.c_int64(len(xb)),
                         ctypes.c_double(lengths[0]), ctypes.c_double(lengths[1]), ctypes.c_double(lengths[2]),
                         ctypes.c_double(rcut_sq(r_in)), ctypes.c_double(rcut_sq(r_out)),
                         ctypes.c_int(1 if same_set else 0), out.ctypes.data_as(c_bp))
    return out


# ---- general triclinic minimum image (extension, parity unpinned by the reference; see oracle.c) ----
def _cell(cell):
    """cell = (lx, ly, lz, xy, xz, yz)"""
    c = np.ascontiguousarray(cell, dtype=np.float64).reshape(6)
    return c, c.ctypes.data_as(c_dp)


def calc_rsq_tri(head, x, y, z, cell):
    head, hp = _d(head); x, xp = _d(x); y, yp = _d(y); z, zp = _d(z)
    cell, cp = _cell(cell)
    out = np.empty(len(x))
    lib().orc_calc_rsq_tri(hp, xp, yp, zp, ctypes.c_int64(len(x)), cp, out.ctypes.data_as(c_dp))
    return out


def rdf_loop_tri(typ, x, y, z, rel, cell, r_cut, ddr, nb, nthreads=1):
    """_rdf_loop counting rules with the LAMMPS-style triclinic image -> (full int64[nb], part int64[R, nb])."""
    typ, tp = _d(typ); x, xp = _d(x); y, yp = _d(y); z, zp = _d(z)
    rel, rp = _i(np.asarray(rel).reshape(-1, 2))
    cell, cp = _cell(cell)
    R = rel.shape[0]
    full = np.zeros(nb); part = np.zeros((R, nb))
    lib().orc_rdf_loop_tri(tp, xp, yp, zp, ctypes.c_int64(len(x)), rp, ctypes.c_int64(R), cp,
                           ctypes.c_double(rcut_sq(r_cut)), ctypes.c_double(ddr), ctypes.c_int64(nb),
                           full.ctypes.data_as(c_dp), part.ctypes.data_as(c_dp), ctypes.c_int(nthreads))
    return full.astype(np.int64), part.astype(np.int64)


def rdf_rect_tri(ta, xa, ya, za, tb, xb, yb, zb, rel, cell, r_cut, ddr, nb, nthreads=1):
    ta, tap = _d(ta); xa, xap = _d(xa); ya, yap = _d(ya); za, zap = _d(za)
    tb, tbp = _d(tb); xb, xbp = _d(xb); yb, ybp = _d(yb); zb, zbp = _d(zb)
    rel, rp = _i(np.asarray(rel).reshape(-1, 2))
    cell, cp = _cell(cell)
    R = rel.shape[0]
    part = np.zeros((R, nb))
    lib().orc_rdf_rect_tri(tap, xap, yap, zap, ctypes.c_int64(len(xa)), tbp, xbp, ybp, zbp, ctypes.c_int64(len(xb)),
                           rp, ctypes.c_int64(R), cp, ctypes.c_double(rcut_sq(r_cut)), ctypes.c_double(ddr),
                           ctypes.c_int64(nb), part.ctypes.data_as(c_dp), ctypes.c_int(nthreads))
    return part.astype(np.int64)


def shell_mask_tri(xa, ya, za, xb, yb, zb, cell, r_in, r_out, same_set):
    xa, xap = _d(xa); ya, yap = _d(ya); za, zap = _d(za)
    xb, xbp = _d(xb); yb, ybp = _d(yb); zb, zbp = _d(zb)
    cell, cp = _cell(cell)
    out = np.zeros((len(xa), len(xb)), dtype=np.uint8)
    lib().orc_shell_mask_tri(xap, yap, zap, ctypes.c_int64(len(xa)), xbp, ybp, zbp, ctypes.c_int64(len(xb)), cp,
                             ctypes.c_double(rcut_sq(r_in)), ctypes.c_double(rcut_sq(r_out)),
                             ctypes.c_int(1 if same_set else 0), out.ctypes.data_as(c_bp))
    return out


def nearest_image_rsq_bruteforce(head, x, y, z, cell):
    """min over the 27 neighbouring images of |head - other - (i a + j b + k c)|^2 (numpy; definition check only)."""
    lx, ly, lz, xy, xz, yz = [float(v) for v in cell]
    a = np.array([lx, 0.0, 0.0]); b = np.array([xy, ly, 0.0]); c = np.array([xz, yz, lz])
    d = np.stack([head[0] - np.asarray(x), head[1] - np.asarray(y), head[2] - np.asarray(z)], axis=1)
    best = np.full(len(d), np.inf)
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                v = d - (i * a + j * b + k * c)
                best = np.minimum(best, np.sum(v * v, axis=1))
    return best


def survival_counts(h):
    """cnt[tau] = sum_pairs sum_t h(t) h(t+tau) for h uint8[T, npairs] (residence_time.py:112-143)."""
    h = np.ascontiguousarray(h, dtype=np.uint8)
    T, P = h.shape
    cnt = np.zeros(T, dtype=np.int64)
    lib().orc_survival_counts(h.ctypes.data_as(c_bp), ctypes.c_int64(T), ctypes.c_int64(P),
                              cnt.ctypes.data_as(c_ip))
    return cnt


def xcorr_direct(a, b):
    """long-double direct form of the unbiased correlation (conductivity.py:97-114)."""
    a, ap = _d(a); b, bp = _d(b)
    out = np.empty(len(a))
    lib().orc_xcorr_direct(ap, bp, ctypes.c_int64(len(a)), out.ctypes.data_as(c_dp))
    return out


def msd_all_origins(traj, max_lag):
    """brute-force windowed MSD over all time origins; traj float64[T,3,N] -> [max_lag,4]."""
    traj, tp = _d(traj)
    T, _, N = traj.shape
    out = np.empty((max_lag, 4))
    lib().orc_msd_all_origins(tp, ctypes.c_int64(T), ctypes.c_int64(N), ctypes.c_int64(max_lag),
                              out.ctypes.data_as(c_dp))
    return out


# ----------------------------------------------------------------------------------------------
# LAMMPS dump / log reading (restating the pymatgen behaviour the reference relies on)
# ----------------------------------------------------------------------------------------------
@dataclass
class Frame:
    timestep: int
    natoms: int
    bounds: np.ndarray           # (3,2) after the tilt correction pymatgen applies
    tilt: np.ndarray | None      # (xy, xz, yz) or None
    cols: dict = field(default_factory=dict)   # column name -> float64[N] in FILE order

    @property
    def lattice_lengths(self):
        """dump.box.to_lattice().lengths (used at rdf_cn.py:260, residence_time.py:79)."""
        m = np.diag(self.bounds[:, 1] - self.bounds[:, 0])
        if self.tilt is not None:
            m[1, 0], m[2, 0], m[2, 1] = self.tilt
        return tuple(np.sqrt(np.sum(m ** 2, axis=1)).tolist())

    @property
    def bound_lengths(self):
        """bounds[hi]-bounds[lo] (cluster_analysis.py:110-112, hydration_number.py:38-40, diffusion.py:75-77)."""
        return tuple(float(self.bounds[k, 1] - self.bounds[k, 0]) for k in range(3))

    def sorted_by_id(self):
        order = np.argsort(self.cols["id"], kind="stable")
        return {k: v[order] for k, v in self.cols.items()}


def _open(fname):
    return gzip.open(fname, "rt") if fname.endswith(".gz") else open(fname, "rt")


def _frame_from_lines(lines):
    timestep = int(lines[1]); natoms = int(lines[3])
    box = np.array([[float(v) for v in lines[k].split()] for k in (5, 6, 7)])
    bounds = box[:, :2].copy()
    tilt = None
    if "xy xz yz" in lines[4]:
        tilt = box[:, 2].copy()
        xs = (0.0, tilt[0], tilt[1], tilt[0] + tilt[1]); ys = (0.0, tilt[2])
        bounds -= np.array([[min(xs), max(xs)], [min(ys), max(ys)], [0.0, 0.0]])
    names = lines[8].replace("ITEM: ATOMS", "").split()
    body = [l.split() for l in lines[9:] if l.strip()]
    arr = np.array(body, dtype=np.float64) if body else np.zeros((0, len(names)))
    return Frame(timestep, natoms, bounds, tilt, {n: np.ascontiguousarray(arr[:, k]) for k, n in enumerate(names)})


def dump_files(pattern):
    files = _glob.glob(pattern)
    if len(files) > 1:
        pat = pattern.replace("*", "([0-9]+)").replace("\\", "\\\\")
        files = sorted(files, key=lambda f: int(re.match(pat, f).group(1)))
    return files


def read_dumps(pattern):
    """Generator of Frames, ordering and splitting as pymatgen.parse_lammps_dumps does."""
    for fname in dump_files(pattern):
        with _open(fname) as f:
            cache = []
            for line in f:
                if line.startswith("ITEM: TIMESTEP"):
                    if cache:
                        yield _frame_from_lines(cache)
                    cache = [line.rstrip("\n")]
                else:
                    cache.append(line.rstrip("\n"))
            if cache:
                yield _frame_from_lines(cache)


# ----------------------------------------------------------------------------------------------
# host-side algebra of the structural API (numpy), in the reference's operation order
# ----------------------------------------------------------------------------------------------
CON_CONSTANT = 1.660538921  # rdf_cn.py:30


def calc_atom_type(ids, num_mols, num_atoms):
    """Vectorised restatement of _calc_atom_type (rdf_cn.py:197-215): id -> position-in-molecule type."""
    ids = np.asarray(ids, dtype=np.float64)
    totals = np.multiply(num_mols, num_atoms)
    cut = np.cumsum(totals)
    out = ids.copy()
    done = np.zeros(len(ids), dtype=bool)
    for i, c in enumerate(cut):
        sel = (~done) & (ids <= c)
        v = (ids[sel] - c) % num_atoms[i]
        v = np.where(v == 0, num_atoms[i], v)
        if i > 0:
            v = v + np.sum(num_atoms[:i])
        out[sel] = v
        done |= sel
    return out


def mol_membership(num_mols, num_atoms_per_mol):
    """(mol_type, mol_id, segment offsets) implied by id order (rdf_cn.py:222-230, com_mols.py:31-42)."""
    mt, mi, off = [], [], [0]
    for t, nm in enumerate(num_mols):
        for m in range(nm):
            mt.append(t + 1); mi.append(m + 1)
            off.append(off[-1] + num_atoms_per_mol[t])
    return np.array(mt), np.array(mi), np.array(off)


def mol_com_wrapped(typ, x, y, z, num_mols, num_atoms_per_mol, mass):
    """_define_mol_cols (rdf_cn.py:218-241): per-molecule mass-weighted mean of wrapped coordinates.

    The reference evaluates ``mass_vec @ coords / mass.sum()`` (BLAS dot, order library dependent);
    here: sequential fp64 accumulation in atom-id order, then one division (SURVEY 7 'hard parts').
    """
    mt, mi, off = mol_membership(num_mols, num_atoms_per_mol)
    m = np.array([mass[int(t) - 1] for t in typ], dtype=np.float64)
    M = len(mt)
    out = np.zeros((M, 3))
    for k in range(M):
        s, e = off[k], off[k + 1]
        msum = 0.0; ax = ay = az = 0.0
        for a in range(s, e):
            ax += m[a] * x[a]; ay += m[a] * y[a]; az += m[a] * z[a]; msum += m[a]
        out[k] = (ax / msum, ay / msum, az / msum)
    return mt.astype(np.float64), out[:, 0].copy(), out[:, 1].copy(), out[:, 2].copy()


def shell_volume(bin_size, num_bins):
    """rdf_cn.py:312-318."""
    return 4 / 3 * np.pi * bin_size ** 3 * (np.arange(1, num_bins + 1) ** 3 - np.arange(num_bins) ** 3)


def type_counts(typ):
    u, c = np.unique(np.asarray(typ).astype(np.int64), return_counts=True)
    return dict(zip(u.tolist(), c.tolist()))


def normalize_rdf(bin_size, rho_pairs, atom_types, partial_relations, num_bins, rdf_part, rdf_full=None,
                  num_atoms=None, rho=None):
    """_normalize_rdf (rdf_cn.py:297-329)."""
    sv = shell_volume(bin_size, num_bins)
    R = len(partial_relations[0])
    if rdf_full is not None:
        rdf_full = rdf_full / (num_atoms * rho * sv)
    nref = np.array([atom_types[a] for a in partial_relations[0]], dtype=np.int64).reshape(R, 1)
    num_atoms_matrix = np.tile(nref, num_bins)
    rho_pairs_matrix = np.tile(np.asarray(rho_pairs).reshape((R, 1)), num_bins)
    sv_matrix = np.tile(sv, (R, 1))
    rdf_part = rdf_part / (num_atoms_matrix * rho_pairs_matrix * sv_matrix)
    return rdf_full, rdf_part


def atomic_rdf(frames, r_cut, bin_size, partial_relations, num_mols=None, num_atoms_per_mol=None, nthreads=0,
               return_counts=False, mic="reference"):
    """calc_atomic_rdf (rdf_cn.py:385-530) on a list of Frames -> float64[nb, 2+R] like df.values."""
    nb = int(r_cut / bin_size)
    radii = (np.arange(nb) + 0.5) * bin_size
    rel = np.asarray(partial_relations).transpose()
    R = rel.shape[0]
    full_sum = np.zeros(nb); part_sum = np.zeros((R, nb))
    counts = []
    for fr in frames:
        c = fr.sorted_by_id()
        typ = c["type"]
        if num_mols and num_atoms_per_mol:
            typ = calc_atom_type(c["id"], num_mols, num_atoms_per_mol)
        if mic == "triclinic":      # extension (no reference implementation): true cell volume, general triclinic image
            lengths = fr.bound_lengths
            tilt = fr.tilt if fr.tilt is not None else (0.0, 0.0, 0.0)
        else:
            lengths = fr.lattice_lengths
        volume = np.prod(lengths)
        at = type_counts(typ)
        n = len(typ)
        rho = n / volume
        rho_pairs = np.array([at[b] / volume for b in partial_relations[1]])
        if mic == "triclinic":
            full, part = rdf_loop_tri(typ, c["x"], c["y"], c["z"], rel, tuple(lengths) + tuple(tilt), r_cut, bin_size, nb,
                                      nthreads)
        else:
            full, part = rdf_loop(typ, c["x"], c["y"], c["z"], rel, lengths, r_cut, bin_size, nb, nthreads)
        counts.append((full, part))
        f, p = normalize_rdf(bin_size, rho_pairs, at, partial_relations, nb, part.astype(np.float64),
                             full.astype(np.float64), n, rho)
        full_sum += f; part_sum += p
    full_sum = full_sum / len(frames); part_sum = part_sum / len(frames)
    out = np.vstack((radii, full_sum, part_sum)).transpose()
    return (out, counts) if return_counts else out


def number_density(frames, surface_atom, atom_types, bin_size, dist_from_interface, axis, num_mols=None,
                   num_atoms_per_mol=None, return_counts=False):
    """calc_number_density (structural/number_density.py:30-154) on a list of Frames -> float64[nb, 1+R] like df.values.
    Same operation order: shift by the surface minimum (:84), side selection (:88-110), ``b -= dist_range`` for a positive
    distance (:96), ``(b / bin_size).astype(int)`` (:97), ``rho_part[i][k] += 1`` with numpy's negative-index wrap (:99);
    indices outside [-nb, nb) (IndexError in the reference) are dropped."""
    nb = int(abs(dist_from_interface) / bin_size)
    radii = (np.arange(nb) + 0.5) * bin_size
    R = len(atom_types)
    total = np.zeros((R, nb))
    counts = []
    for fr in frames:
        c = fr.sorted_by_id()
        key = calc_atom_type(c["id"], num_mols, num_atoms_per_mol) if (num_mols and num_atoms_per_mol) else c["type"]
        x = np.asarray(c[axis], dtype=np.float64)
        surf = x[key == surface_atom]
        cnt = np.zeros((R, nb), dtype=np.int64)
        if len(surf):
            mn, mx = surf.min(), surf.max()
            rng_ = mx - mn
            xs = x - mn
            for i, j in enumerate(atom_types):
                if dist_from_interface > 0:
                    b = xs[(key == j) & (xs < dist_from_interface)] - rng_
                else:
                    b = xs[(key == j) & (xs > dist_from_interface)]
                k = (b / bin_size).astype(int)
                k = np.where(k < 0, k + nb, k)
                k = k[(k >= 0) & (k < nb)]
                np.add.at(cnt[i], k, 1)
        counts.append(cnt)
        lengths = dict(zip("xyz", fr.lattice_lengths))
        area = np.prod([lengths[a] for a in lengths if a != axis])
        total += cnt.astype(np.float64) / (area * bin_size)
    total = total / len(frames)
    out = np.vstack((radii, total)).transpose()
    return (out, counts) if return_counts else out


def atomic_cn(frames, r_cuts, partial_relations, num_mols=None, num_atoms_per_mol=None, nthreads=0):
    """calc_atomic_cn (rdf_cn.py:533-651) -> float64[R]."""
    rel = np.asarray(partial_relations).transpose()
    cn_sum = np.zeros(rel.shape[0])
    for fr in frames:
        c = fr.sorted_by_id()
        typ = c["type"]
        if num_mols and num_atoms_per_mol:
            typ = calc_atom_type(c["id"], num_mols, num_atoms_per_mol)
        at = type_counts(typ)
        cn = cn_loop(typ, c["x"], c["y"], c["z"], rel, fr.lattice_lengths, r_cuts, nthreads).astype(np.float64)
        cn = cn / [at[a] for a in partial_relations[0]]       # _normalize_cn rdf_cn.py:332-338
        cn_sum += cn
    return cn_sum / len(frames)


def molecular_rdf(frames, r_cut, bin_size, partial_relations, num_mols, num_atoms_per_mol, mass, nthreads=0,
                  inter=False):
    """calc_molecular_rdf (rdf_cn.py:654-756); inter=True -> calc_intermolecular_rdf (:857-902)."""
    nb = int(r_cut / bin_size)
    radii = (np.arange(nb) + 0.5) * bin_size
    rel = np.asarray(partial_relations).transpose()
    R = rel.shape[0]
    part_sum = np.zeros((R, nb))
    for fr in frames:
        c = fr.sorted_by_id()
        mt, mx, my, mz = mol_com_wrapped(c["type"], c["x"], c["y"], c["z"], num_mols, num_atoms_per_mol, mass)
        lengths = fr.lattice_lengths
        volume = np.prod(lengths)
        ot = type_counts(mt)
        if inter:
            at = ot
            part = rdf_rect(mt, mx, my, mz, mt, mx, my, mz, rel, lengths, r_cut, bin_size, nb, nthreads)
        else:
            at = type_counts(c["type"])
            part = rdf_rect(c["type"], c["x"], c["y"], c["z"], mt, mx, my, mz, rel, lengths, r_cut, bin_size,
                            nb, nthreads)
        rho_pairs = np.array([ot[b] / volume for b in partial_relations[1]])
        _, p = normalize_rdf(bin_size, rho_pairs, at, partial_relations, nb, part.astype(np.float64))
        part_sum += p
    part_sum = part_sum / len(frames)
    return np.vstack((radii, part_sum)).transpose()


def molecular_cn(frames, r_cuts, partial_relations, num_mols, num_atoms_per_mol, mass, nthreads=0):
    """calc_molecular_cn (rdf_cn.py:759-854)."""
    rel = np.asarray(partial_relations).transpose()
    cn_sum = np.zeros(rel.shape[0])
    for fr in frames:
        c = fr.sorted_by_id()
        mt, mx, my, mz = mol_com_wrapped(c["type"], c["x"], c["y"], c["z"], num_mols, num_atoms_per_mol, mass)
        at = type_counts(c["type"])
        cn = cn_rect(c["type"], c["x"], c["y"], c["z"], mt, mx, my, mz, rel, fr.lattice_lengths, r_cuts,
                     nthreads).astype(np.float64)
        cn = cn / [at[a] for a in partial_relations[0]]
        cn_sum += cn
    return cn_sum / len(frames)


# ----------------------------------------------------------------------------------------------
# dynamical API (numpy)
# ----------------------------------------------------------------------------------------------
def correlate_fft(a, b):
    """Conductivity.correlate (conductivity.py:97-114): zero-pad to 2T, ifft(fft(a) conj fft(b)), / (T..1)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    al = np.concatenate((a, np.zeros(len(a)))); bl = np.concatenate((b, np.zeros(len(b))))
    c = np.fft.ifft(np.fft.fft(al) * np.conjugate(np.fft.fft(bl))).real
    d = c[: len(c) // 2]
    return d / (np.arange(len(d)) + 1)[::-1]


def cumtrapz(y, dx, initial_zero):
    """scipy cumulative_trapezoid(dx=..); leading 0 only for conductivity (conductivity.py:231 vs viscosity.py:151)."""
    y = np.asarray(y, dtype=np.float64)
    res = np.cumsum(dx * (y[1:] + y[:-1]) / 2.0)
    return np.concatenate(([0.0], res)) if initial_zero else res


def ols_origin(t, y):
    """sm.OLS(y, t).fit() without intercept (diffusion.py:323-329): slope, bse, uncentred R^2."""
    t = np.asarray(t, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    sxx = float(np.dot(t, t)); beta = float(np.dot(t, y)) / sxx
    r = y - beta * t
    ssr = float(np.dot(r, r))
    return beta, float(np.sqrt(ssr / (len(y) - 1) / sxx)), 1.0 - ssr / float(np.dot(y, y))


def com_unwrapped(cols, num_mols, num_atoms_per_mol, mass, attrs=("xu", "yu", "zu"), charge=False):
    """calc_com (com_mols.py:5-62): sum(m*a)/sum(m) per molecule (pandas groupby.sum order = id order)."""
    mt, mi, off = mol_membership(num_mols, num_atoms_per_mol)
    if mass:
        m = np.array([mass[int(t) - 1] for t in cols["type"]], dtype=np.float64)
    else:
        m = cols["mass"]
    seg = off[:-1]
    msum = np.add.reduceat(m, seg)
    out = {"type": mt, "mol_id": mi, "mass": msum}
    for a in attrs:
        out[a] = np.add.reduceat(cols[a] * m, seg) / msum
    if charge:
        out["q"] = np.add.reduceat(cols["q"], seg)
    return out


def msd_single_origin(traj, t0, scale):
    """Diffusion.get_msd_from_dump arithmetic (diffusion.py:201-218) for one group:
    traj float64[T,3,N] -> (per_atom [T,4,N], mean [T,4]); SI conversion happens before differencing."""
    x = np.asarray(traj, dtype=np.float64) * scale
    d2 = (x - x[t0][None]) ** 2
    msd = (d2[:, 0] + d2[:, 1]) + d2[:, 2]
    per_atom = np.concatenate([d2, msd[:, None, :]], axis=1)
    return per_atom, per_atom.mean(axis=2)


def msd_interval(traj, scale, stride):
    """msd_int (diffusion.py:225-237): NaN first row -> dx2.. averaged over n-1 rows, msd over n rows."""
    x = np.asarray(traj, dtype=np.float64)[::stride] * scale
    d2 = (x[1:] - x[:-1]) ** 2
    n = x.shape[0]
    out = np.empty((4, x.shape[2]))
    out[:3] = d2.sum(axis=0) / (n - 1)
    out[3] = ((d2[:, 0] + d2[:, 1]) + d2[:, 2]).sum(axis=0) / n
    return out


def charge_flux(vel, mass_atom, q_atom, seg_off, type_off, vel_scale, q_scale):
    """conductivity_loop (_conductivity.py:7-36): vel [T,3,N] -> J [3, ntypes, T]."""
    vel = np.asarray(vel, dtype=np.float64)
    seg = np.asarray(seg_off[:-1])
    msum = np.add.reduceat(mass_atom, seg)
    qmol = np.add.reduceat(q_atom, seg) * q_scale
    T = vel.shape[0]
    G = len(type_off) - 1
    J = np.zeros((3, G, T))
    for t in range(T):
        for c in range(3):
            vcom = np.add.reduceat(vel[t, c] * mass_atom, seg) / msum * vel_scale
            for g in range(G):
                s, e = type_off[g], type_off[g + 1]
                J[c, g, t] = np.dot(vcom[s:e], qmol[s:e])
    return J


def residence_correlation(frames_xyz_a, frames_xyz_b, lengths, r_in, r_out, same_set):
    """ResidenceTime.calc_auto_correlation (residence_time.py:70-148) for one relation, exact-integer form:
    frames_xyz_* = list over frames of (x, y, z) arrays; returns C(tau) normalised by C(0)."""
    T = len(frames_xyz_a)
    h = np.stack([shell_mask(*frames_xyz_a[t], *frames_xyz_b[t], lengths[t], r_in, r_out, same_set).reshape(-1)
                  for t in range(T)])
    cnt = survival_counts(h).astype(np.float64)
    ncols = h.shape[1]
    c = (cnt / (T - np.arange(T))) / ncols
    return c / c[0], cnt
