"""Per-molecule mass-weighted reduction -- device counterpart of ``mdproptools.common.com_mols.calc_com``
(reference mdproptools/common/com_mols.py:5-62).

The reference takes a pymatgen ``LammpsDump`` and returns a pandas frame per snapshot; the hot-path version
works on a whole batch of frames resident on the device: ``attr [F, C, N]`` -> ``[F, C, M]`` with one thread
per molecule accumulating ``sum(m * a)`` sequentially in atom-id order and dividing by ``sum(m)`` once, which is
the reference's ``(a * m).groupby(mol).sum() / m.groupby(mol).sum()`` (:57-60) in a fixed order.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def mol_membership(num_mols, num_atoms_per_mol):
    """(mol_type[M], mol_id[M], seg_off[M+1]) implied by id order (com_mols.py:31-42)."""
    mol_type = np.repeat(np.arange(1, len(num_mols) + 1), num_mols)
    mol_id = np.concatenate([np.arange(1, k + 1) for k in num_mols]) if len(num_mols) else np.zeros(0, dtype=int)
    sizes = np.repeat(np.asarray(num_atoms_per_mol), num_mols)
    seg_off = np.concatenate(([0], np.cumsum(sizes)))
    return mol_type, mol_id, seg_off


def atom_masses(types: np.ndarray, mass) -> np.ndarray:
    """``mass[int(type - 1)]`` per atom (com_mols.py:54)."""
    return np.asarray(mass, dtype=np.float64)[np.asarray(types).astype(np.int64) - 1]


def calc_com(attr: torch.Tensor, masses: torch.Tensor, seg_off: np.ndarray, charges: torch.Tensor | None = None):
    """attr [F,C,N] (device) -> (com [F,C,M], mol_mass [M], mol_charge [M] or None)."""
    so = torch.from_numpy(np.ascontiguousarray(seg_off, dtype=np.int32)).to(attr.device)
    return ops.segment_com(attr, masses, so, extra=charges)
