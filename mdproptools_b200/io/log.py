"""LAMMPS thermo-log reading (replaces ``pymatgen.io.lammps.outputs.parse_lammps_log`` used at
mdproptools/dynamical/viscosity.py:211 and mdproptools/utilities/log.py:21).

A log holds one thermo block per ``run``: it starts after the line beginning with
``Per MPI rank memory allocation`` (or the older ``Memory usage per processor =``) and ends before
``Loop time of``.  One-line style blocks are a whitespace table with a header row; multi-line style blocks
repeat ``---- Step N ----`` followed by ``key = value`` pairs.  Host-side only (the data volume is tiny).
"""
from __future__ import annotations

import glob
import re
from io import StringIO

import numpy as np
import pandas as pd

_BEGIN = ("Memory usage per processor =", "Per MPI rank memory allocation (min/avg/max) =")
_END = "Loop time of"
_MULTI = re.compile(r"-+\s+Step\s+([0-9]+)\s+-+")
_KV = re.compile(r"([0-9A-Za-z_\[\]]+)\s+=\s+([0-9eE\.+-]+)")


def _parse_block(lines):
    if _MULTI.match(lines[0]):
        rows, keys = [], None
        cur = None
        for ln in lines:
            m = _MULTI.match(ln)
            if m:
                if cur is not None:
                    rows.append(cur)
                cur = {"Step": int(m.group(1))}
            elif cur is not None:
                for k, v in _KV.findall(ln):
                    cur[k] = float(v)
        if cur is not None:
            rows.append(cur)
        keys = list(rows[0].keys())
        return pd.DataFrame(rows)[keys]
    return pd.read_csv(StringIO("".join(lines)), sep=r"\s+")


def parse_lammps_log(filename="log.lammps"):
    """List of DataFrames, one per thermo block."""
    with open(filename, "rt") as f:
        lines = f.readlines()
    begins = [i for i, ln in enumerate(lines) if ln.startswith(_BEGIN)]
    ends = [i for i, ln in enumerate(lines) if ln.startswith(_END)]
    return [_parse_block(lines[b + 1:e]) for b, e in zip(begins, ends)]


def concat_log(log_pattern, step=None, working_dir=None):
    """mdproptools/utilities/log.py:10-28: first thermo block of every matching log, files ordered by the
    integer that ``*`` matches, the last row of every log but the final one dropped (it repeats as the first
    row of the next run), optionally thinned to rows 1, 50001, ... when ``step`` is truthy."""
    import os

    working_dir = working_dir or os.getcwd()
    files = glob.glob(f"{working_dir}/{log_pattern}")
    if len(files) > 1:
        pattern = r"%s" % log_pattern.replace("*", "([0-9]+)")
        pattern = ".*" + pattern.replace("\\", "\\\\")
        files = sorted(files, key=lambda f: int(re.match(pattern, f).group(1)))
    logs = [parse_lammps_log(file)[0] for file in files]
    for p, l in enumerate(logs[:-1]):
        logs[p] = l[:-1]
    full_log = pd.concat(logs, ignore_index=True)
    if step:
        full_log = full_log.loc[range(1, full_log.shape[0], 50000)]
    return full_log
