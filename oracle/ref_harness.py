"""Harness that lets the UNMODIFIED reference (``/root/reference/mdproptools``) be imported in the
build container so that golden vectors can be generated from it.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  It only works where
``/root/reference`` is mounted (the build container); it never travels to the GPU box -- the outputs
it produces are committed as small fixtures under ``tests/golden/`` by ``oracle/gen_golden.py``.

What it does (all of it is glue written for this repo; no reference source is copied):

* registers an empty ``mdproptools`` package object whose ``__path__`` points at the read-only
  reference tree, so sub-modules import without executing ``mdproptools/__init__.py`` (which pulls
  seaborn/statsmodels/matplotlib, absent here);
* provides a stand-in for the un-vendored third-party dependency **pymatgen**
  (``requirements.txt:1`` -> fork ``molmd/pymatgen@molmd_fix_3-9``, branch-pinned, not in
  ``/root/reference``): ``pymatgen.io.lammps.outputs.{parse_lammps_dumps, parse_lammps_log,
  LammpsDump, LammpsBox}`` and ``pymatgen.core.structure.Molecule`` (the few methods get_unique_configurations calls), restating the
  published upstream behaviour (glob + integer sort of ``*``, frame split at ``ITEM: TIMESTEP``,
  box bounds/tilt handling, ``pandas.read_csv`` of the atom block, thermo-block log parsing);
* mocks matplotlib / seaborn / statsmodels (plotting only) and supplies the two numerical
  statsmodels entry points the reference calls: ``statsmodels.tsa.stattools.acovf`` (FFT, unbiased,
  no demean -- the call at ``residence_time.py:135``) and ``statsmodels.api.OLS`` (no-intercept
  least squares: params, bse, uncentred R^2 -- the call at ``diffusion.py:323``);
* aliases ``scipy.integrate.cumtrapz`` (removed in scipy>=1.14; used at ``viscosity.py:151``);
* points ``NUMBA_CACHE_DIR`` at a fresh writable dir (the reference kernels are ``cache=True`` and
  the tree is read-only).
"""
from __future__ import annotations

import glob
import os
import re
import sys
import tempfile
import types
from io import StringIO
from unittest import mock

import numpy as np
import pandas as pd

REFERENCE_ROOT = os.environ.get("MDPROP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mdproptools"))


# ----------------------------------------------------------------------------------------------
# pymatgen stand-in
# ----------------------------------------------------------------------------------------------

class _Site:
    """pymatgen.core.sites.Site as far as get_unique_configurations (cluster_analysis.py:337-372) uses it: species_string,
    coords, equality = same species and coordinates within Site.position_atol (1e-5)."""

    def __init__(self, sym, xyz):
        self.species_string = sym
        self.coords = np.asarray(xyz, dtype=np.float64)

    def __eq__(self, other):
        return (isinstance(other, _Site) and self.species_string == other.species_string
                and bool(np.allclose(self.coords, other.coords, atol=1e-5)))

    def __hash__(self):
        return hash(self.species_string)

    def __str__(self):
        return self.species_string


class _Species(str):
    """str(mol.species[i]) is the element symbol"""


class ShimMolecule:
    """Stand-in for pymatgen.core.structure.Molecule restricted to what the reference calls:
    ``Molecule.from_file`` (.xyz: count line, comment line, ``sym x y z`` rows; .pdb: HETATM/ATOM records, element in
    columns 77-78), ``.species``, ``mol[i]`` / ``mol[a:b]`` (site / list of sites), ``site in mol``, and
    ``get_neighbors(site, r)`` = the sites within r of site.coords (``<= r``), the site itself excluded, in site order
    (published upstream behaviour: IMolecule.get_sites_in_sphere / get_neighbors)."""

    def __init__(self, sites):
        self.sites = list(sites)

    @classmethod
    def from_file(cls, filename):
        filename = str(filename)
        sites = []
        with open(filename) as f:
            lines = f.read().splitlines()
        if filename.lower().endswith(".pdb"):
            for ln in lines:
                if ln.startswith(("HETATM", "ATOM")):
                    sym = ln[76:78].strip().capitalize()
                    sites.append(_Site(sym, [float(ln[30:38]), float(ln[38:46]), float(ln[46:54])]))
        else:
            n = int(lines[0].split()[0])
            for ln in lines[2:2 + n]:
                p = ln.split()
                sites.append(_Site(p[0], [float(p[1]), float(p[2]), float(p[3])]))
        return cls(sites)

    @property
    def species(self):
        return [_Species(s.species_string) for s in self.sites]

    def __len__(self):
        return len(self.sites)

    def __iter__(self):
        return iter(self.sites)

    def __getitem__(self, i):
        return self.sites[i]

    def get_neighbors(self, site, r):
        out = []
        for s in self.sites:
            d = float(np.linalg.norm(s.coords - site.coords))
            if d <= r and not (s is site or s == site):
                out.append(s)
        return out

class _Lattice:
    def __init__(self, matrix):
        self._matrix = np.array(matrix, dtype=np.float64).reshape((3, 3))

    @property
    def lengths(self):
        return tuple(np.sqrt(np.sum(self._matrix ** 2, axis=1)).tolist())

    @property
    def volume(self):
        m = self._matrix
        return float(abs(np.dot(np.cross(m[0], m[1]), m[2])))


class LammpsBox:
    def __init__(self, bounds, tilt=None):
        bounds_arr = np.array(bounds)
        assert bounds_arr.shape == (3, 2)
        self.bounds = bounds_arr.tolist()
        matrix = np.diag(bounds_arr[:, 1] - bounds_arr[:, 0])
        self.tilt = None
        if tilt is not None:
            tilt_arr = np.array(tilt)
            assert tilt_arr.shape == (3,)
            matrix[1, 0] = tilt_arr[0]
            matrix[2, 0] = tilt_arr[1]
            matrix[2, 1] = tilt_arr[2]
            self.tilt = tilt_arr.tolist()
        self._matrix = matrix

    @property
    def volume(self):
        m = self._matrix
        return np.dot(np.cross(m[0], m[1]), m[2])

    def to_lattice(self):
        return _Lattice(self._matrix)


class LammpsDump:
    def __init__(self, timestep, natoms, box, data):
        self.timestep = timestep
        self.natoms = natoms
        self.box = box
        self.data = data

    @classmethod
    def from_string(cls, string):
        lines = string.split("\n")
        timestep = int(lines[1])
        natoms = int(lines[3])
        box_arr = np.loadtxt(StringIO("\n".join(lines[5:8])))
        bounds = box_arr[:, :2]
        tilt = None
        if "xy xz yz" in lines[4]:
            tilt = box_arr[:, 2]
            x = (0, tilt[0], tilt[1], tilt[0] + tilt[1])
            y = (0, tilt[2])
            bounds -= np.array([[min(x), max(x)], [min(y), max(y)], [0, 0]])
        box = LammpsBox(bounds, tilt)
        data_head = lines[8].replace("ITEM: ATOMS", "").split()
        data = pd.read_csv(StringIO("\n".join(lines[9:])), names=data_head, sep=r"\s+")
        return cls(timestep, natoms, box, data)


def parse_lammps_dumps(file_pattern):
    files = glob.glob(file_pattern)
    if len(files) > 1:
        pattern = file_pattern.replace("*", "([0-9]+)").replace("\\", "\\\\")
        files = sorted(files, key=lambda f: int(re.match(pattern, f).group(1)))
    for fname in files:
        with open(fname, "rt") as f:
            dump_cache = []
            for line in f:
                if line.startswith("ITEM: TIMESTEP"):
                    if len(dump_cache) > 0:
                        yield LammpsDump.from_string("".join(dump_cache))
                    dump_cache = [line]
                else:
                    dump_cache.append(line)
            yield LammpsDump.from_string("".join(dump_cache))


def parse_lammps_log(filename="log.lammps"):
    with open(filename, "rt") as f:
        lines = f.readlines()
    begin_flag = ("Memory usage per processor =", "Per MPI rank memory allocation (min/avg/max) =")
    end_flag = "Loop time of"
    begins, ends = [], []
    for i, line in enumerate(lines):
        if line.startswith(begin_flag):
            begins.append(i)
        elif line.startswith(end_flag):
            ends.append(i)

    def _parse_thermo(thermo_lines):
        multi_pattern = r"-+\s+Step\s+([0-9]+)\s+-+"
        if re.match(multi_pattern, thermo_lines[0]):
            timestep_marks = [i for i, l in enumerate(thermo_lines) if re.match(multi_pattern, l)]
            timesteps = np.split(thermo_lines, timestep_marks)[1:]
            dicts = []
            kv_pattern = r"([0-9A-Za-z_\[\]]+)\s+=\s+([0-9eE\.+-]+)"
            for ts in timesteps:
                data = {}
                data["Step"] = int(re.match(multi_pattern, ts[0]).group(1))
                data.update({k: float(v) for k, v in re.findall(kv_pattern, "".join(ts[1:]))})
                dicts.append(data)
            df = pd.DataFrame(dicts)
            columns = ["Step"] + [k for k, v in re.findall(kv_pattern, "".join(timesteps[0][1:]))]
            df = df[columns]
        else:
            df = pd.read_csv(StringIO("".join(thermo_lines)), sep=r"\s+")
        return df

    runs = []
    for b, e in zip(begins, ends):
        runs.append(_parse_thermo(lines[b + 1 : e]))
    return runs


# ----------------------------------------------------------------------------------------------
# statsmodels stand-ins (numerical entry points only)
# ----------------------------------------------------------------------------------------------
def _next_regular(target):
    """smallest 5-smooth number >= target (statsmodels.compat.scipy._next_regular)."""
    if target <= 6:
        return target
    if not (target & (target - 1)):
        return target
    match = float("inf")
    p5 = 1
    while p5 < target:
        p35 = p5
        while p35 < target:
            quotient = -(-target // p35)
            p2 = 2 ** ((quotient - 1).bit_length())
            N = p2 * p35
            if N == target:
                return N
            elif N < match:
                match = N
            p35 *= 3
            if p35 == target:
                return p35
        if p35 < match:
            match = p35
        p5 *= 5
        if p5 == target:
            return p5
    if p5 < match:
        match = p5
    return match


def acovf(x, adjusted=False, demean=True, fft=True, missing="none", nlag=None, unbiased=None):
    if unbiased is not None:
        adjusted = unbiased
    x = np.asarray(x, dtype=np.float64).squeeze()
    xo = x - x.mean() if demean else x
    n = len(x)
    d = n - np.arange(n) if adjusted else n * np.ones(n)
    if fft:
        nobs = len(xo)
        nfft = _next_regular(2 * nobs + 1)
        Frf = np.fft.fft(xo, n=nfft)
        acov = np.fft.ifft(Frf * np.conjugate(Frf))[:nobs] / d
        return acov.real
    return (np.correlate(xo, xo, "full")[n - 1 :]) / d


class _OLSResult:
    def __init__(self, y, x):
        y = np.asarray(y, dtype=np.float64)
        x = np.asarray(x, dtype=np.float64)
        n = len(y)
        sxx = float(np.dot(x, x))
        beta = float(np.dot(x, y)) / sxx
        resid = y - beta * x
        ssr = float(np.dot(resid, resid))
        self.params = [beta]
        self.bse = [float(np.sqrt(ssr / (n - 1) / sxx))]
        self.rsquared = 1.0 - ssr / float(np.dot(y, y))
        self._pred = beta * x

    def predict(self):
        return self._pred

    def summary(self):
        return f"OLS(no intercept): slope={self.params[0]!r} bse={self.bse[0]!r} R2={self.rsquared!r}"


class _OLS:
    def __init__(self, endog, exog):
        self._y, self._x = endog, exog

    def fit(self):
        return _OLSResult(self._y, self._x)


_INSTALLED = False


def install():
    """Make ``import mdproptools.<sub>.<mod>`` resolve to the unmodified reference."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_ref_cache_"))

    pkg = types.ModuleType("mdproptools")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "mdproptools")]
    sys.modules["mdproptools"] = pkg
    # hydration_number.py:8 does a bare ``from rdf_cn import ...``
    sys.path.append(os.path.join(REFERENCE_ROOT, "mdproptools", "structural"))

    # pymatgen
    this = sys.modules[__name__]
    for name in ("pymatgen", "pymatgen.io", "pymatgen.io.lammps", "pymatgen.core"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    outputs = types.ModuleType("pymatgen.io.lammps.outputs")
    for sym in ("parse_lammps_dumps", "parse_lammps_log", "LammpsDump", "LammpsBox"):
        setattr(outputs, sym, getattr(this, sym))
    sys.modules["pymatgen.io.lammps.outputs"] = outputs
    structure = types.ModuleType("pymatgen.core.structure")
    structure.Molecule = ShimMolecule
    sys.modules["pymatgen.core.structure"] = structure

    # plotting mocks
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "seaborn", "tqdm"):
        if name == "tqdm":
            try:
                import tqdm  # noqa: F401
                continue
            except Exception:
                pass
        sys.modules[name] = mock.MagicMock(name=name)

    # statsmodels
    sm = types.ModuleType("statsmodels")
    sm.__path__ = []
    sm_api = types.ModuleType("statsmodels.api")
    sm_api.OLS = _OLS
    sm_tsa = types.ModuleType("statsmodels.tsa")
    sm_tsa.__path__ = []
    sm_st = types.ModuleType("statsmodels.tsa.stattools")
    sm_st.acovf = acovf
    sm.api, sm.tsa, sm_tsa.stattools = sm_api, sm_tsa, sm_st
    sys.modules.update(
        {
            "statsmodels": sm,
            "statsmodels.api": sm_api,
            "statsmodels.tsa": sm_tsa,
            "statsmodels.tsa.stattools": sm_st,
        }
    )

    import scipy.integrate as _si

    if not hasattr(_si, "cumtrapz"):
        _si.cumtrapz = _si.cumulative_trapezoid
    _INSTALLED = True


def sample_dir() -> str:
    return os.path.join(REFERENCE_ROOT, "data", "mg_tfsi_dme")
