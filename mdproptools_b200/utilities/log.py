"""``mdproptools.utilities.log`` under its reference import path (mdproptools/utilities/log.py:10): the reader lives in
``mdproptools_b200.io.log`` (pymatgen-free thermo-log parser)."""
from ..io.log import concat_log, parse_lammps_log  # noqa: F401
