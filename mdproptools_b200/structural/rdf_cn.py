"""Radial distribution functions and coordination numbers from LAMMPS dump files -- drop-in for
``mdproptools.structural.rdf_cn`` (reference file mdproptools/structural/rdf_cn.py; citations below are
lines of that file).

Public functions keep the reference's names, positional order, defaults, return types and side-effect
files: calc_atomic_rdf (:385), calc_atomic_cn (:533), calc_molecular_rdf (:654), calc_molecular_cn (:759),
calc_intermolecular_rdf (:857).

Division of labour
  device  all pair arithmetic (_calc_rsq/_remove_outliers/_rdf_loop/_cn_loop/_rdf_mol_loop/_cn_mol_loop,
          :35-162) -> per-frame *integer* histograms via libmdprop_b200 (csrc/pair.cu); molecule centres of
          mass (_define_mol_cols, :218-241) via the segmented reduction kernel;
  host    everything that only defines how integers become the final floats (_calc_props :244-294,
          _normalize_rdf :297-329, _normalize_cn :332-338, _save_* :341-382), restated in numpy in the
          reference's operation order so that equal counts give bit-identical DataFrames.
Frames are sharded round-robin over the ranks of an initialised torch.distributed group; the per-frame
integer histograms are merged with one int64 all-reduce.

Extension (keyword ``mic``, default "reference"): the reference has no triclinic minimum image -- it wraps tilted
cells as if they were orthogonal with the lattice-vector lengths and takes volume = prod(lengths) (:260-261).
``mic="reference"`` reproduces exactly that; ``mic="triclinic"`` uses the general triclinic image of
MDP_PAIR_TRICLINIC (include/mdprop_b200.h) with the cell (lx, ly, lz, xy, xz, yz) from the dump header and the true
cell volume lx*ly*lz.  For orthogonal boxes the two are identical bit for bit.

Documented divergences from the reference:
  * a bin index >= num_bins (possible when r_cut/bin_size is not an integer) is dropped instead of
    written out of bounds (the reference corrupts memory);
  * molecule centres of mass use sequential fp64 accumulation in atom-id order (the reference's BLAS dot
    has a library-dependent order; differences are <= 1 ulp of the coordinate).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from .. import dist, ops
from .._lib import PAIR_TRICLINIC, bin_edges
from ..io.pipeline import ArrayBatches, FrameBatches

CON_CONSTANT = 1.660538921  # :30

VERBOSE = False   # the reference prints per-frame progress; off by default here


def _log(*a):
    if VERBOSE:
        print(*a)


# ------------------------------------------------------------------------------------------------
# host-side restatements
# ------------------------------------------------------------------------------------------------
def _num_bins(r_cut, bin_size):
    """_initialize (:165-180)."""
    if isinstance(r_cut, list):
        nb = [int(i / bin_size) for i in r_cut]
        radii = [(np.arange(i) + 0.5) * bin_size for i in nb]
    else:
        nb = int(r_cut / bin_size)
        radii = (np.arange(nb) + 0.5) * bin_size
    return nb, radii


def _rcut_sq(r_cut) -> float:
    """``r_cut ** 2`` as the numba kernels evaluate it (:66): exact for ints, x*x for floats."""
    if isinstance(r_cut, (int, np.integer)):
        return float(int(r_cut) ** 2)
    r = float(r_cut)
    return r * r


def calc_atom_type_ids(ids, num_mols, num_atoms):
    """_calc_atom_type (:197-215) vectorised: atom id -> 1-based position inside its molecule type block,
    offset by the atoms-per-molecule of the earlier molecule types.  ids beyond the last block are left as is."""
    ids = np.asarray(ids, dtype=np.float64)
    cut = np.cumsum(np.multiply(num_mols, num_atoms))
    out = ids.copy()
    which = np.searchsorted(cut, ids, side="left")          # first block with id <= cutoff
    for i in range(len(cut)):
        sel = which == i
        if not sel.any():
            continue
        v = np.mod(ids[sel] - cut[i], num_atoms[i])
        v[v == 0] = num_atoms[i]
        if i > 0:
            v = v + np.sum(num_atoms[:i])
        out[sel] = v
    return out


def _same_values(a, b):
    """a == b element for element.  Contiguous arrays of one dtype and shape are compared with one memcmp (ctypes releases
    the GIL; 10^5 doubles take ~50 us, where ``np.array_equal`` builds a temporary and holds the GIL next to the pipeline's
    reader threads)."""
    if a is b:
        return True
    if a.shape != b.shape:
        return False
    if a.dtype == b.dtype and a.flags.c_contiguous and b.flags.c_contiguous:
        import ctypes

        return _memcmp()(ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data), ctypes.c_size_t(a.nbytes)) == 0
    return bool(np.array_equal(a, b))


def _memcmp():
    import ctypes

    fn = getattr(_memcmp, "fn", None)
    if fn is None:
        fn = ctypes.CDLL(None).memcmp
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        _memcmp.fn = fn
    return fn


class _TypeCache:
    """Per-frame host work that depends only on the frame's id/type column: the altered types, the {type: count} table and
    the class of every atom.  Trajectories almost always repeat the same column frame after frame; one comparison (tens
    of microseconds for 10^5 atoms) then replaces ~1 ms of recomputation per frame, which is what the consumer thread of
    the file pipeline spent next to a pair kernel that needs 0.1 ms per frame."""

    def __init__(self):
        self.key = None
        self.val = None
        self.src = None            # the object last asked about (an identical object needs no comparison at all)

    def get(self, column, make):
        if column is self.src:
            return self.val
        if self.key is None or not _same_values(self.key, column):
            self.key = np.array(column, copy=True)
            self.val = make(column)
        self.src = column if column.base is None else None      # (a view into a staging buffer is not an identity)
        return self.val


class _DeviceCache:
    """The class vector on the device: uploaded again only when the host vector is another object than last time (a
    pageable H2D copy waits for the stream -- the consumer would stall behind its own kernels once per batch)."""

    def __init__(self):
        self.src, self.dev = None, None

    def get(self, arr, device):
        if arr is not self.src or self.dev is None or self.dev.device != device:
            self.dev = torch.from_numpy(arr).to(device)
            self.src = arr
        return self.dev


def _value_counts(typ):
    """{type: count}, keys ascending.  Types are small non-negative integers in practice: one bincount pass instead of
    np.unique's sort (this runs once per frame on the host, next to a pair kernel that takes 0.16 ms per frame)."""
    t = np.asarray(typ).astype(np.int64)
    if t.size and t.min() >= 0 and t.max() < (1 << 20):
        c = np.bincount(t)
        u = np.flatnonzero(c)
        return dict(zip(u.tolist(), c[u].tolist()))
    u, c = np.unique(t, return_counts=True)
    return dict(zip(u.tolist(), c.tolist()))


def _calc_props(box_lengths, n_objects, atom_types, object_types, num_types, mass, partial_relations, atom_types_col,
                num_atoms_per_mol=None):
    """_calc_props (:244-294): densities and the consistency checks (same exceptions)."""
    num_relations = len(partial_relations[0])
    volume = np.prod(box_lengths)
    set_id = set(atom_types.keys())
    if atom_types_col == "type":
        if num_types != len(set_id):
            raise ValueError(
                f"""Consistency check failed: Number of specified
                            atomic types is different from the calculated value
                            specified= {num_types}, calculated= {len(set_id)}"""
            )
    elif atom_types_col == "id":
        if np.sum(num_atoms_per_mol) != len(set_id):
            raise ValueError(
                f"""Consistency check failed: Number of specified
                            atomic types is different from the calculated value
                            specified= {num_atoms_per_mol},
                            calculated= {len(set_id)}"""
            )
    total_mass = np.sum([float(mass[i]) * float(atom_types[i + 1]) for i in range(num_types)])
    total_density = float((total_mass / volume) * CON_CONSTANT)
    _log("{0:s}{1:10.8f}".format("Average density=", float(total_density)))
    rho = n_objects / volume
    rho_pairs = np.zeros(num_relations)
    for index, object_type in enumerate(partial_relations[1]):
        rho_pairs[index] = object_types[object_type] / volume
        if rho_pairs[index] < 1.0e-22:
            raise ValueError("Error: Density is zero for mol type: " + str(object_type))
    return rho, rho_pairs


def _rdf_denominators(bin_size, rho_pairs, atom_types, partial_relations, num_relations, num_bins, num_atoms=None, rho=None):
    """The divisors of _normalize_rdf (:297-329), built with the reference's expressions in the reference's order, so that
    ``counts / divisor`` is bit-identical to the reference whether it is evaluated per frame or for a batch of frames."""
    shell_volume = (
        4 / 3 * np.pi * bin_size ** 3 * (np.arange(1, num_bins + 1) ** 3 - np.arange(num_bins) ** 3)
    )
    den_full = None if num_atoms is None else num_atoms * rho * shell_volume
    ref_atoms = np.asarray(partial_relations[0]).reshape((num_relations, 1))
    ref_atoms_matrix = np.tile(ref_atoms, num_bins)
    num_atoms_matrix = np.vectorize(atom_types.get)(ref_atoms_matrix)
    rho_pairs_matrix = np.tile(rho_pairs.reshape((num_relations, 1)), num_bins)
    shell_volume_matrix = np.tile(shell_volume, (num_relations, 1))
    return den_full, num_atoms_matrix * rho_pairs_matrix * shell_volume_matrix


def _normalize_rdf(bin_size, rho_pairs, atom_types, partial_relations, num_relations, num_bins, rdf_part, rdf_full=None,
                   num_atoms=None, rho=None):
    """_normalize_rdf (:297-329)."""
    den_full, den_part = _rdf_denominators(bin_size, rho_pairs, atom_types, partial_relations, num_relations, num_bins,
                                           num_atoms if rdf_full is not None else None, rho)
    if rdf_full is not None:
        rdf_full = rdf_full / den_full
    rdf_part = rdf_part / den_part
    return rdf_full, rdf_part


def _normalize_cn(atom_types, partial_relations, cn):
    """_normalize_cn (:332-338)."""
    num_ref_atoms = [atom_types[i] for i in partial_relations[0]]
    return cn / num_ref_atoms


def _save_rdf(radii, relation_matrix, path_or_buf, save_mode, rdf_part_sum, rdf_full_sum=None):
    """_save_rdf (:341-365)."""
    if rdf_full_sum is not None:
        result_tuple = (radii, rdf_full_sum, rdf_part_sum)
        final_label = ["r ($\\AA$)", "g_full(r)"]
    else:
        result_tuple = (radii, rdf_part_sum)
        final_label = ["r ($\\AA$)"]
    final_array = np.vstack(result_tuple).transpose()
    final_labels = final_label + [f"g_{str(pair[0])}-{str(pair[1])}" for pair in relation_matrix]
    final_df = pd.DataFrame(final_array, columns=final_labels)
    if save_mode:
        final_df.to_csv(path_or_buf, index=False)
        _log("Results are written to pd.DataFrame and csv file")
    else:
        _log(final_df)
    return final_df


def _save_cn(relation_matrix, path_or_buff, cn_sum, save_mode):
    """_save_cn (:368-382)."""
    final_array = np.vstack(cn_sum).transpose()
    final_labels = [f"cn_{str(pair[0])}-{str(pair[1])}" for pair in relation_matrix]
    final_df = pd.DataFrame(final_array, columns=final_labels)
    if save_mode:
        final_df.to_csv(path_or_buff, index=False)
        _log("CN results are written to pd.DataFrame and csv file")
    else:
        _log(final_df)
    return final_df


# ------------------------------------------------------------------------------------------------
# class maps: relation types -> small dense class ids for the device histogram
# ------------------------------------------------------------------------------------------------
class _ClassMap:
    """Types named by the relations get their own class; everything else shares the class 'other'."""

    def __init__(self, named_types, present_types=None):
        self.named = sorted(set(int(t) for t in named_types))
        self.index = {t: k for k, t in enumerate(self.named)}
        # the class 'other' is only materialised when some point actually has an unnamed type (one class fewer
        # keeps single-species runs on the kernel variant without per-pair class bookkeeping)
        self.has_other = present_types is None or bool(set(int(t) for t in present_types) - set(self.named))
        self.ncls = len(self.named) + (1 if self.has_other else 0)
        self.other = self.ncls - 1 if self.has_other else -1

    def classes_of(self, typ: np.ndarray) -> np.ndarray:
        t = np.asarray(typ).astype(np.int64)
        if t.size and t.min() >= 0 and t.max() < (1 << 20):      # one table lookup instead of a mask per named type
            lut = np.full(int(t.max()) + 1, self.other, dtype=np.int32)
            for ty, k in self.index.items():
                if 0 <= ty < len(lut):
                    lut[ty] = k
            cls = lut[t]
        else:
            cls = np.full(t.shape, self.other, dtype=np.int32)
            for ty, k in self.index.items():
                cls[t == ty] = k
        if not self.has_other and (cls < 0).any():
            raise ValueError("a frame contains atom types that were absent from the first frame")
        return cls

    def cls(self, t):
        return self.index.get(int(t), self.other)


def _sym_weights(cmap: _ClassMap, relation_matrix, with_full: bool):
    """hist_reduce weights for the symmetric kernel: g_full counts every pair twice (:85-86); relation (a,b)
    counts a pair once per matching role order (:89-96) -> 2 on the like row, 1 on the unlike row."""
    rows = ops.sym_rows(cmap.ncls)
    w = []
    if with_full:
        w.append(np.full(rows, 2, dtype=np.int32))
    for a, b in relation_matrix:
        r = np.zeros(rows, dtype=np.int32)
        r[ops.sym_row(cmap.cls(a), cmap.cls(b), cmap.ncls)] = 2 if int(a) == int(b) else 1
        w.append(r)
    return np.stack(w)


def _rect_weights(cmap_a: _ClassMap, cmap_b: _ClassMap, relation_matrix):
    rows = cmap_a.ncls * cmap_b.ncls
    w = []
    for a, b in relation_matrix:
        r = np.zeros(rows, dtype=np.int32)
        r[cmap_a.cls(a) * cmap_b.ncls + cmap_b.cls(b)] = 1
        w.append(r)
    return np.stack(w)


def _cn_edges(r_cut_list):
    """Sorted distinct squared cutoffs as a histogram edge table; bin k = #{thresholds <= rsq}."""
    rc2 = np.array([_rcut_sq(r) for r in r_cut_list], dtype=np.float64)
    thr = np.unique(rc2)                       # ascending
    edges = np.concatenate(([0.0], thr))       # edges[k>=1] = thr[k-1]
    # a pair with rsq < rc2[kl] lies in bins 0 .. (index of rc2[kl] in thr); cumulative bin index per relation
    upto = np.searchsorted(thr, rc2)           # bins 0..upto inclusive
    return edges, float(thr[-1]), upto


# ------------------------------------------------------------------------------------------------
# minimum-image convention
# ------------------------------------------------------------------------------------------------
def _mic_flags(mic) -> int:
    if mic == "reference":
        return 0
    if mic == "triclinic":
        return PAIR_TRICLINIC
    raise ValueError(f"mic must be 'reference' or 'triclinic', got {mic!r}")


def _frame_box(box, flags):
    """(lengths whose product is the normalisation volume, the row handed to the pair kernel)."""
    if flags & PAIR_TRICLINIC:
        lengths = box.bound_lengths()
        tilt = box.tilt if box.tilt is not None else (0.0, 0.0, 0.0)
        return lengths, tuple(lengths) + tuple(tilt)
    lengths = box.lattice_lengths()          # :260
    return lengths, lengths


class _PropsCache:
    """_frame_box + _calc_props of a frame, remembered for the next frames with the same box, atom count and type table
    (every frame of an NVT trajectory): ~0.1 ms of small numpy calls per frame otherwise -- as much as the pair kernel
    itself needs for a 10^5-atom frame.  The checks of _calc_props run (and raise) on the first frame of each kind."""

    def __init__(self, flags, num_types, mass, partial_relations, col, num_atoms_per_mol):
        self.args = (num_types, mass, partial_relations, col, num_atoms_per_mol)
        self.flags, self.key, self.val = flags, None, None

    def get(self, meta, at):
        b = meta.box
        key = (meta.natoms, id(at), tuple(map(tuple, b.bounds)), None if b.tilt is None else tuple(b.tilt))
        if key != self.key:
            lengths, boxrow = _frame_box(b, self.flags)
            num_types, mass, rel, col, napm = self.args
            rho, rho_pairs = _calc_props(lengths, meta.natoms, at, at, num_types, mass, rel, col, napm)
            self.key, self.val = key, (at, boxrow, rho, rho_pairs)           # (holding `at` keeps its id unique)
        return self.val[1:]


# ------------------------------------------------------------------------------------------------
# frame iteration shared by all entry points
# ------------------------------------------------------------------------------------------------
class _RetryUnsharded(Exception):
    """Some rank met a dump file with more than one frame while reading only its share of the files."""


_SHARD_FILES = [True]


def _frame_batches(filename, columns):
    """This rank's frames.  With several ranks every rank READS only every world-th file (frame index = file index for the
    usual one-frame-per-file dumps); ``_agree_sharding`` then checks collectively that no rank met a multi-frame file --
    otherwise the call is repeated with every rank reading everything and selecting its frames (``file_sharded``)."""
    w, r = dist.world_size(), dist.rank()
    if w > 1 and _SHARD_FILES[0]:
        return FrameBatches(filename, columns, file_shard=(r, w))
    sel = (lambda i: i % w == r) if w > 1 else None
    return FrameBatches(filename, columns, frame_select=sel)


def _agree_sharding(batches, device):
    """Call after the batch loop and before the merge collective: all ranks learn whether the sharded read held."""
    if getattr(batches, "file_shard", None) is None:
        return
    flag = torch.tensor([1 if batches.multi_frame_seen else 0], dtype=torch.int64, device=device)
    dist.all_reduce_max_(flag)
    if int(flag.item()):
        raise _RetryUnsharded()


def file_sharded(fn):
    """Decorator of the file-based entry points: run with sharded file reads, fall back (on every rank at once) when a
    multi-frame file was met."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        try:
            return fn(*args, **kwargs)
        except _RetryUnsharded:
            _SHARD_FILES[0] = False
            try:
                return fn(*args, **kwargs)
            finally:
                _SHARD_FILES[0] = True

    return wrapper


def _merge_frames(per_frame: dict, total_frames: int, shape, device):
    """Place this rank's per-frame integer results into [T, *shape] and all-reduce (int64, exact)."""
    out = torch.zeros((total_frames,) + tuple(shape), dtype=torch.int64, device=device)
    for idx, t in per_frame.items():
        out[idx] = t
    dist.all_reduce_sum_(out)
    return out


def _mol_segments(num_mols, num_atoms_per_mol):
    """Molecule membership implied by id order (:222-230): (mol_type[M], seg_off[M+1])."""
    mt = np.repeat(np.arange(1, len(num_mols) + 1), num_mols)
    sizes = np.repeat(np.asarray(num_atoms_per_mol), num_mols)
    off = np.concatenate(([0], np.cumsum(sizes)))
    return mt, off


# ------------------------------------------------------------------------------------------------
# public API
# ------------------------------------------------------------------------------------------------
@file_sharded
def calc_atomic_rdf(r_cut, bin_size, num_types, mass, partial_relations, filename, num_mols=None, num_atoms_per_mol=None,
                    path_or_buff="rdf.csv", save_mode=True, mic="reference"):
    """Full and partial atom-atom RDF; see the reference docstring (:397-452) for the arguments."""
    flags = _mic_flags(mic)
    num_bins, radii = _num_bins(r_cut, bin_size)
    num_relations = len(partial_relations[0])
    relation_matrix = np.asarray(partial_relations).transpose()
    altered = bool(num_mols and num_atoms_per_mol)
    named = list(partial_relations[0]) + list(partial_relations[1])
    cmap = weights = None          # fixed by the types present in this rank's first frame
    edges = bin_edges(bin_size, num_bins)
    rcut2 = _rcut_sq(r_cut)

    batches = _frame_batches(filename, ["id", "type", "x", "y", "z"])
    counts, props = {}, {}
    device = torch.device("cuda", torch.cuda.current_device())   # a rank that receives no frames still joins the merge
    tcache, ccache, dcache = _TypeCache(), _TypeCache(), _DeviceCache()
    pcache = _PropsCache(flags, num_types, mass, partial_relations, "id" if altered else "type", num_atoms_per_mol)

    def _typ_and_counts(col):
        typ_ = calc_atom_type_ids(col, num_mols, num_atoms_per_mol) if altered else np.array(col, copy=True)
        return typ_, _value_counts(typ_)

    for batch in batches:
        dev = batch.wait()
        device = dev.device
        F = len(batch.metas)
        host = batch.host.numpy()
        cls_rows = []
        boxes = np.empty((F, 6 if flags else 3))
        same_cls, cls_first = True, None
        for k, meta in enumerate(batch.metas):
            _log("The timestep of the current file is: " + str(meta.timestep))
            typ, at = tcache.get(host[k, 0] if altered else host[k, 1], _typ_and_counts)
            boxrow, rho, rho_pairs = pcache.get(meta, at)
            props[meta.index] = (at, rho, rho_pairs, meta.natoms)
            if cmap is None:
                cmap = _ClassMap(named, present_types=at.keys())
                weights = _sym_weights(cmap, relation_matrix, with_full=True)
            cls_k = ccache.get(typ, cmap.classes_of)
            same_cls = same_cls and (k == 0 or cls_k is cls_first)
            cls_first = cls_k if k == 0 else cls_first
            cls_rows.append(cls_k)
            boxes[k] = boxrow
        xyz = dev[:, 2:5, :].contiguous()
        # one [N] vector when the batch shares its classes (and no upload at all when it is the previous batch's)
        cls_d = dcache.get(cls_first, device) if same_cls else torch.from_numpy(np.stack(cls_rows)).to(device)
        hist = ops.pair_hist(xyz, cls_d, cmap.ncls, boxes, rcut2, edges, bin_size, flags=flags)
        red = ops.hist_reduce(hist, weights)                      # [F, 1+R, nb]
        for k, meta in enumerate(batch.metas):
            counts[meta.index] = red[k]
    _agree_sharding(batches, device)
    T = batches.total_frames or 0
    if T == 0:
        raise ValueError(f"no dump frames found for {filename!r}")
    allc = _merge_frames(counts, T, (1 + num_relations, num_bins), device).cpu().numpy()
    if dist.world_size() > 1:
        props = _gather_props(props, T)

    rdf_full_sum = np.zeros(num_bins)
    rdf_part_sum = np.zeros((num_relations, num_bins))
    den_cache = {}   # frames of one composition and volume share the divisors of _normalize_rdf (the NVT case: one entry)
    for idx in range(T):
        at, rho, rho_pairs, natoms = props[idx]
        key = (natoms, rho, np.asarray(rho_pairs).tobytes(), tuple(sorted(at.items())))
        den = den_cache.get(key)
        if den is None:
            den = den_cache[key] = _rdf_denominators(bin_size, rho_pairs, at, partial_relations, num_relations, num_bins,
                                                     natoms, rho)
        rdf_full_sum += allc[idx, 0].astype(np.float64) / den[0]
        rdf_part_sum += allc[idx, 1:].astype(np.float64) / den[1]
    rdf_full_sum = rdf_full_sum / T
    rdf_part_sum = rdf_part_sum / T
    return _save_rdf(radii, relation_matrix, path_or_buff, save_mode, rdf_part_sum, rdf_full_sum=rdf_full_sum)


def _gather_props(props: dict, total_frames: int) -> dict:
    """Every rank needs every frame's host-side normalisation inputs (tiny python objects)."""
    import torch.distributed as d

    gathered = [None] * d.get_world_size()
    d.all_gather_object(gathered, props)
    out = {}
    for g in gathered:
        out.update(g)
    assert len(out) == total_frames
    return out


@file_sharded
def calc_atomic_cn(r_cut, bin_size, num_types, mass, partial_relations, filename, num_mols=None, num_atoms_per_mol=None,
                   path_or_buff="cn.csv", save_mode=True, mic="reference"):
    """Atom-atom coordination numbers, one cutoff per relation (:533-651)."""
    flags = _mic_flags(mic)
    _num_bins(r_cut, bin_size)
    num_relations = len(partial_relations[0])
    relation_matrix = np.asarray(partial_relations).transpose()
    altered = bool(num_mols and num_atoms_per_mol)
    named = list(partial_relations[0]) + list(partial_relations[1])
    cmap = weights = None
    edges, rcut2_max, upto = _cn_edges(r_cut)
    nthr = len(edges) - 1

    batches = _frame_batches(filename, ["id", "type", "x", "y", "z"])
    counts, props = {}, {}
    device = torch.device("cuda", torch.cuda.current_device())   # a rank that receives no frames still joins the merge
    tcache, ccache, dcache = _TypeCache(), _TypeCache(), _DeviceCache()
    pcache = _PropsCache(flags, num_types, mass, partial_relations, "id" if altered else "type", num_atoms_per_mol)

    def _typ_and_counts(col):
        typ_ = calc_atom_type_ids(col, num_mols, num_atoms_per_mol) if altered else np.array(col, copy=True)
        return typ_, _value_counts(typ_)

    for batch in batches:
        dev = batch.wait()
        device = dev.device
        F = len(batch.metas)
        host = batch.host.numpy()
        cls_rows = []
        boxes = np.empty((F, 6 if flags else 3))
        same_cls, cls_first = True, None
        for k, meta in enumerate(batch.metas):
            typ, at = tcache.get(host[k, 0] if altered else host[k, 1], _typ_and_counts)
            boxrow = pcache.get(meta, at)[0]
            props[meta.index] = at
            if cmap is None:
                cmap = _ClassMap(named, present_types=at.keys())
                weights = _sym_weights(cmap, relation_matrix, with_full=False)
            cls_k = ccache.get(typ, cmap.classes_of)
            same_cls = same_cls and (k == 0 or cls_k is cls_first)
            cls_first = cls_k if k == 0 else cls_first
            cls_rows.append(cls_k)
            boxes[k] = boxrow
        xyz = dev[:, 2:5, :].contiguous()
        cls_d = dcache.get(cls_first, device) if same_cls else torch.from_numpy(np.stack(cls_rows)).to(device)
        hist = ops.pair_hist(xyz, cls_d, cmap.ncls, boxes, rcut2_max, edges, 0.0, flags=flags)
        red = ops.hist_reduce(hist, weights, cumulative=True)     # [F, R, nthr] cumulative over thresholds
        for k, meta in enumerate(batch.metas):
            counts[meta.index] = red[k]
    _agree_sharding(batches, device)
    T = batches.total_frames or 0
    if T == 0:
        raise ValueError(f"no dump frames found for {filename!r}")
    allc = _merge_frames(counts, T, (num_relations, nthr), device).cpu().numpy()
    if dist.world_size() > 1:
        props = _gather_props(props, T)
    cn_sum = np.zeros(num_relations)
    for idx in range(T):
        cn = np.array([allc[idx, kl, upto[kl]] for kl in range(num_relations)], dtype=np.float64)
        cn_sum += _normalize_cn(props[idx], partial_relations, cn)
    cn_sum = cn_sum / T
    return _save_cn(relation_matrix, path_or_buff, cn_sum, save_mode)


def _molecular_common(filename, num_types, mass, partial_relations, num_mols, num_atoms_per_mol, inter, run_pairs, flags=0):
    """Shared frame loop of the atom/molecule-COM entry points (:654-902).  ``run_pairs`` gets the device
    tensors of one batch and returns per-frame integer results [F, R, *]."""
    num_relations = len(partial_relations[0])
    relation_matrix = np.asarray(partial_relations).transpose()
    mol_type, seg_off = _mol_segments(num_mols, num_atoms_per_mol)
    cmap_a = _ClassMap(partial_relations[0])
    cmap_b = _ClassMap(partial_relations[1])
    weights = _rect_weights(cmap_a, cmap_b, relation_matrix)
    mol_types_count = _value_counts(mol_type)
    batches = _frame_batches(filename, ["id", "type", "x", "y", "z"])
    counts, props = {}, {}
    device = torch.device("cuda", torch.cuda.current_device())   # a rank that receives no frames still joins the merge
    for batch in batches:
        dev = batch.wait()
        device = dev.device
        F = len(batch.metas)
        host = batch.host.numpy()
        n = host.shape[2]
        if seg_off[-1] != n:
            raise ValueError(f"Length of values ({seg_off[-1]}) does not match length of index ({n})")
        boxes = np.empty((F, 6 if flags else 3))
        cls_a = np.empty((F, n), dtype=np.int32)
        for k, meta in enumerate(batch.metas):
            lengths, boxes[k] = _frame_box(meta.box, flags)
            if inter:
                at = mol_types_count
                n_objects = len(mol_type)
                atc = "type"
            else:
                at = _value_counts(host[k, 1])
                n_objects = len(mol_type)
                atc = "type"
                cls_a[k] = cmap_a.classes_of(host[k, 1])
            rho, rho_pairs = _calc_props(lengths, n_objects, at, mol_types_count, num_types, mass, partial_relations, atc,
                                         num_atoms_per_mol)
            props[meta.index] = (at, rho_pairs)
        xyz = dev[:, 2:5, :].contiguous()
        # masses per atom from the type column of the first frame of the batch (:231); types are static per id
        m_atom = torch.from_numpy(np.array([mass[int(t - 1)] for t in host[0, 1]], dtype=np.float64)).to(device)
        com, _, _ = ops.segment_com(xyz, m_atom, torch.from_numpy(seg_off.astype(np.int32)).to(device))   # [F,3,M]
        cls_m = torch.from_numpy(cmap_b.classes_of(mol_type)).to(device)
        if inter:
            cm_a = _ClassMap(partial_relations[0])
            res = run_pairs(com, torch.from_numpy(cm_a.classes_of(mol_type)).to(device), cm_a.ncls, com, cls_m,
                            cmap_b.ncls, boxes, _rect_weights(cm_a, cmap_b, relation_matrix))
        else:
            res = run_pairs(xyz, torch.from_numpy(cls_a).to(device), cmap_a.ncls, com, cls_m, cmap_b.ncls, boxes, weights)
        for k, meta in enumerate(batch.metas):
            counts[meta.index] = res[k]
    _agree_sharding(batches, device)
    T = batches.total_frames or 0
    if T == 0:
        raise ValueError(f"no dump frames found for {filename!r}")
    return counts, props, T, device, relation_matrix, num_relations


@file_sharded
def calc_molecular_rdf(r_cut, bin_size, num_types, mass, partial_relations, filename, num_mols, num_atoms_per_mol,
                       path_or_buff="rdf_mol.csv", save_mode=True, _inter=False, mic="reference"):
    """Partial RDF between atoms and molecule centres of mass (:654-756)."""
    flags = _mic_flags(mic)
    num_bins, radii = _num_bins(r_cut, bin_size)
    edges = bin_edges(bin_size, num_bins)
    rcut2 = _rcut_sq(r_cut)

    def run(xa, ca, na, xb, cb, nb_, boxes, w):
        hist = ops.pair_hist(xa, ca, na, boxes, rcut2, edges, bin_size, xyz_b=xb, cls_b=cb, ncls_b=nb_, flags=flags)
        return ops.hist_reduce(hist, w)

    counts, props, T, device, relation_matrix, R = _molecular_common(filename, num_types, mass, partial_relations, num_mols,
                                                                      num_atoms_per_mol, _inter, run, flags)
    allc = _merge_frames(counts, T, (R, num_bins), device).cpu().numpy()
    if dist.world_size() > 1:
        props = _gather_props(props, T)
    rdf_part_sum = np.zeros((R, num_bins))
    for idx in range(T):
        at, rho_pairs = props[idx]
        _, rdf_part = _normalize_rdf(bin_size, rho_pairs, at, partial_relations, R, num_bins, allc[idx].astype(np.float64))
        rdf_part_sum += rdf_part
    rdf_part_sum = rdf_part_sum / T
    return _save_rdf(radii, relation_matrix, path_or_buff, save_mode, rdf_part_sum)


@file_sharded
def calc_molecular_cn(r_cut, bin_size, num_types, mass, partial_relations, filename, num_mols, num_atoms_per_mol,
                      path_or_buff="cn_mol.csv", save_mode=True, mic="reference"):
    """Coordination numbers between atoms and molecule centres of mass (:759-854)."""
    flags = _mic_flags(mic)
    _num_bins(r_cut, bin_size)
    edges, rcut2_max, upto = _cn_edges(r_cut)
    nthr = len(edges) - 1

    def run(xa, ca, na, xb, cb, nb_, boxes, w):
        hist = ops.pair_hist(xa, ca, na, boxes, rcut2_max, edges, 0.0, xyz_b=xb, cls_b=cb, ncls_b=nb_, flags=flags)
        return ops.hist_reduce(hist, w, cumulative=True)

    counts, props, T, device, relation_matrix, R = _molecular_common(filename, num_types, mass, partial_relations, num_mols,
                                                                      num_atoms_per_mol, False, run, flags)
    allc = _merge_frames(counts, T, (R, nthr), device).cpu().numpy()
    if dist.world_size() > 1:
        props = _gather_props(props, T)
    cn_sum = np.zeros(R)
    for idx in range(T):
        at, _ = props[idx]
        cn = np.array([allc[idx, kl, upto[kl]] for kl in range(R)], dtype=np.float64)
        cn_sum += _normalize_cn(at, partial_relations, cn)
    cn_sum = cn_sum / T
    return _save_cn(relation_matrix, path_or_buff, cn_sum, save_mode)


def calc_atomic_rdf_from_arrays(positions, types, box_lengths, r_cut, bin_size, partial_relations, batch_frames=64,
                                return_counts=False, mic="reference", frame_range=None):
    """Array front end of :func:`calc_atomic_rdf` for trajectories that are already in memory.

    positions   float64 [T, 3, N] host array (numpy, or a pinned torch tensor for async copies), rows in id order
    types       [N] or [T, N] atom types (ints or floats, as the dump's ``type`` column)
    box_lengths [3] or [T, 3] box lengths as ``dump.box.to_lattice().lengths`` gives them; with ``mic="triclinic"``
                [6] or [T, 6] = (lx, ly, lz, xy, xz, yz) and the normalisation volume is lx*ly*lz
    Everything else as in calc_atomic_rdf; same per-frame normalisation (rdf_cn.py:297-329, 502-521), same DataFrame.
    Frames stream through pinned staging buffers: the copy of batch k+1 overlaps the kernels of batch k.
    With ``return_counts`` the raw integer histograms [T, 1+R, nbins] (g_full row first) are returned as well.

    Under an initialised process group the T frames are split in contiguous blocks over the ranks (``dist.shard_range``);
    ``positions`` then holds either all T frames on every rank, or only this rank's block with ``frame_range=(lo, hi, T)``
    saying which block of a T-frame trajectory it is.  The per-frame integer histograms are all-gathered (exact) and every
    rank returns the same DataFrame.
    """
    pos = positions if isinstance(positions, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(positions, dtype=np.float64))
    T, _, N = pos.shape
    if frame_range is not None:
        lo, hi, T = (int(v) for v in frame_range)
        if (lo, hi) != dist.shard_range(T) or pos.shape[0] != hi - lo:
            raise ValueError(f"frame_range {frame_range} is not this rank's block {dist.shard_range(T)} of {T} frames")
    else:
        lo, hi = dist.shard_range(T)
        pos = pos[lo:hi]
    num_bins, radii = _num_bins(r_cut, bin_size)
    num_relations = len(partial_relations[0])
    relation_matrix = np.asarray(partial_relations).transpose()
    dev = torch.device("cuda", torch.cuda.current_device())
    batches = iter(ArrayBatches(pos, batch_frames, dev)) if hi > lo else iter(())   # (a rank beyond the last frame has no block)
    first = next(batches, None)                                   # the copy of the first batch starts now, under the host prep
    types = np.asarray(types)
    at_all = _value_counts(types)                                 # one pass over the types serves the class map too
    cmap = _ClassMap(list(partial_relations[0]) + list(partial_relations[1]), present_types=list(at_all))
    weights = _sym_weights(cmap, relation_matrix, with_full=True)
    edges = bin_edges(bin_size, num_bins)
    rcut2 = _rcut_sq(r_cut)
    flags = _mic_flags(mic)
    boxes = np.broadcast_to(np.asarray(box_lengths, dtype=np.float64), (T, 6 if flags else 3))
    static_types = types.ndim == 1
    if static_types:
        cls_static = torch.from_numpy(cmap.classes_of(types)).to(dev)
        at_static = at_all
    out = torch.empty((hi - lo, 1 + num_relations, num_bins), dtype=torch.int64, device=dev)

    def _all_batches():
        if first is not None:
            yield first
            yield from batches

    for f0, f1, x in _all_batches():                             # copy of batch k+1 overlaps the kernels of batch k
        cls = cls_static if static_types else torch.from_numpy(np.stack([cmap.classes_of(t) for t in types[lo + f0:lo + f1]])).to(dev)
        hist = ops.pair_hist(x, cls, cmap.ncls, boxes[lo + f0:lo + f1], rcut2, edges, bin_size, flags=flags)
        out[f0:f1] = ops.hist_reduce(hist, weights)
    counts = dist.all_gather_blocks(out, T).cpu().numpy()
    rdf_full_sum = np.zeros(num_bins)
    rdf_part_sum = np.zeros((num_relations, num_bins))
    den_cache = {}   # frames with the same composition and volume share their divisors (the usual NVT case: one entry)
    volumes = np.prod(boxes[:, :3], axis=1)
    if static_types and T > 1 and np.all(volumes == volumes[0]):
        # one composition, one volume: the per-frame loop below divides every frame by the same arrays and adds the frames in
        # order -- which is what an axis-0 reduction of the divided [T, ...] array does (numpy accumulates rows sequentially
        # along a non-contiguous axis: same additions in the same order, bit-identical; 1000 frames: 10 ms -> 1 ms)
        rho = N / volumes[0]
        rho_pairs = np.array([at_static[b] / volumes[0] for b in partial_relations[1]])
        den = _rdf_denominators(bin_size, rho_pairs, at_static, partial_relations, num_relations, num_bins, N, rho)
        rdf_full_sum = np.add.reduce(counts[:, 0].astype(np.float64) / den[0], axis=0)
        rdf_part_sum = np.add.reduce(counts[:, 1:].astype(np.float64) / den[1], axis=0)
        df = _save_rdf(radii, relation_matrix, None, False, rdf_part_sum / T, rdf_full_sum=rdf_full_sum / T)
        return (df, counts) if return_counts else df
    for t in range(T):
        at = at_static if static_types else _value_counts(types[t])
        volume = np.prod(boxes[t, :3])
        key = (volume,) if static_types else (volume, tuple(sorted(at.items())))
        den = den_cache.get(key)
        if den is None:
            rho = N / volume
            rho_pairs = np.array([at[b] / volume for b in partial_relations[1]])
            den = den_cache[key] = _rdf_denominators(bin_size, rho_pairs, at, partial_relations, num_relations, num_bins, N, rho)
        rdf_full_sum += counts[t, 0].astype(np.float64) / den[0]
        rdf_part_sum += counts[t, 1:].astype(np.float64) / den[1]
    df = _save_rdf(radii, relation_matrix, None, False, rdf_part_sum / T, rdf_full_sum=rdf_full_sum / T)
    return (df, counts) if return_counts else df


def calc_intermolecular_rdf(r_cut, bin_size, num_types, mass, partial_relations, filename, num_mols, num_atoms_per_mol,
                            path_or_buff="rdf_mol.csv", save_mode=True, mic="reference"):
    """Molecule-COM / molecule-COM partial RDF, self pairs included as in the reference (:857-902)."""
    return calc_molecular_rdf(r_cut, bin_size, num_types, mass, partial_relations, filename, num_mols, num_atoms_per_mol,
                              path_or_buff=path_or_buff, save_mode=save_mode, _inter=True, mic=mic)
