"""Hydration factor of cations -- drop-in for ``mdproptools/structural/hydration_number.py`` (citations are
lines of that file; note the reference module is not importable as part of its package because of the bare
``from rdf_cn import ...`` at :8).

Per frame the reference loops over cation atoms, finds the waters (first atom of each water molecule = O)
within ``r_cut`` with ``_calc_rsq`` (:16-19) and evaluates the cosine between the minimum-image cation->O
displacement and the water bisector (H1 + H2 - 2 O, raw coordinates, :60-63); a water counts as oriented
when cos < -0.72 (:32).  Here the cation x water cutoff search of all frames of a batch is one
``mdp_pair_list`` call on the device, and so is the epilogue (``mdp_hydration_count``, csrc/epilogue.cu): the
entries are grouped in the reference's row order (frame, cation, water), the cosines are formed in numpy's fp64
expression order (bit-identical to the host expressions they replace), and the two counters per cation -- waters
in range, waters with cos < -0.72 -- come back as integers.  The host only averages them.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch

from .. import dist, ops
from ..io.pipeline import FrameBatches
from .rdf_cn import _mol_segments


def get_hydration_number(dump_pattern, cation_type, water_type, r_cut, alter_atom_ids=False, num_mols=None,
                         num_atoms_per_mol=None, working_dir=None):
    """Returns a DataFrame with column ``angles_distribution`` (all cosines) and the constant column
    ``hydration_factor`` (mean over frames of the mean over cations of the oriented fraction); also written to
    ``angles_df.csv`` in ``working_dir`` (:78-101).  ``alter_atom_ids`` only changes the unused ``type`` column
    in the reference (:41-43) and is accepted for signature compatibility."""
    if not working_dir:
        working_dir = os.getcwd()
    mol_type_m, seg_off = _mol_segments(num_mols, num_atoms_per_mol)
    mol_of_atom = np.repeat(np.arange(len(mol_type_m)), np.diff(seg_off))
    cat_rows = np.nonzero(mol_type_m[mol_of_atom] == cation_type)[0]           # every atom of a cation molecule
    wat_mols = np.nonzero(mol_type_m == water_type)[0]
    o_rows = seg_off[wat_mols]                                                  # groupby(mol_id).first()  (:52)
    sizes = np.diff(seg_off)[wat_mols]
    if np.any(sizes < 3):
        raise ValueError("water molecules need at least 3 atoms (O, H, H)")
    h1_rows, h2_rows = o_rows + 1, o_rows + 2                                   # nth([1, 2])           (:54)
    rc2 = r_cut ** 2

    w, r = dist.world_size(), dist.rank()
    sel = (lambda i: i % w == r) if w > 1 else None
    batches = FrameBatches(os.path.join(working_dir, dump_pattern), ["id", "x", "y", "z"], frame_select=sel)
    per_frame = {}
    for batch in batches:
        dev = batch.wait()
        host = batch.host.numpy()
        F, _, n = host.shape
        if seg_off[-1] != n:
            raise ValueError(f"Length of values ({seg_off[-1]}) does not match length of index ({n})")
        xyz = dev[:, 1:4, :]
        xa = xyz.index_select(2, torch.from_numpy(cat_rows).to(dev.device)).contiguous()
        xb = xyz.index_select(2, torch.from_numpy(o_rows).to(dev.device)).contiguous()
        boxes = np.array([m.box.bound_lengths() for m in batch.metas])
        lst, _ = ops.pair_list(xa, xb, boxes, 0.0, rc2, shell_mode=0)
        dv = dev.device
        xh1 = xyz.index_select(2, torch.from_numpy(h1_rows).to(dv)).contiguous()
        xh2 = xyz.index_select(2, torch.from_numpy(h2_rows).to(dv)).contiguous()
        cos_d, seg_off, counts = ops.hydration_count(lst, xa, xb, xh1, xh2, boxes, threshold=-0.72)
        cos_h = cos_d.cpu().numpy()
        off = seg_off.cpu().numpy()
        cnt = counts.cpu().numpy().astype(np.int64)                        # [F, ncat, 2]
        ncat = len(cat_rows)
        if np.any(cnt[:, :, 0] == 0):
            raise ZeroDivisionError("division by zero")                   # a cation without water in range, as the reference (:32)
        frac = cnt[:, :, 1] / cnt[:, :, 0]
        for k, meta in enumerate(batch.metas):
            factor = 0.0
            for a in range(ncat):                                          # the reference's running sum over cations (:72-73)
                factor += float(frac[k, a])                               # (Python floats: the final sum() below is the reference's)
            per_frame[meta.index] = (cos_h[off[k * ncat]: off[(k + 1) * ncat]].tolist(), factor / ncat)
    T = batches.total_frames or 0
    if w > 1:
        import torch.distributed as d_

        gathered = [None] * w
        d_.all_gather_object(gathered, per_frame)
        per_frame = {}
        for g in gathered:
            per_frame.update(g)
    res = [per_frame[i] for i in range(T)]
    angles_df = pd.DataFrame([item for sub in res for item in sub[0]], columns=["angles_distribution"])
    angles_df["hydration_factor"] = sum([i[1] for i in res]) / len(res)
    angles_df.to_csv(os.path.join(working_dir, "angles_df.csv"))
    return angles_df
