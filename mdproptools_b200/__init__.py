"""mdproptools_b200 -- B200-native (sm_100a) implementation of the mdproptools trajectory post-processing
hot path behind mdproptools' own Python API.

    from mdproptools_b200.structural.rdf_cn import calc_atomic_rdf, calc_atomic_cn, ...
    from mdproptools_b200.structural.cluster_analysis import get_clusters
    from mdproptools_b200.structural.hydration_number import get_hydration_number
    from mdproptools_b200.dynamical.diffusion import Diffusion
    from mdproptools_b200.dynamical.conductivity import Conductivity
    from mdproptools_b200.dynamical.viscosity import Viscosity
    from mdproptools_b200.dynamical.residence_time import ResidenceTime

Host code is Python; all arithmetic of the hot path runs in hand-written CUDA kernels reached through the C
ABI of ``libmdprop_b200.so`` (include/mdprop_b200.h).  There is no CPU fallback.
"""
from . import dynamical, structural  # noqa: F401  (mirrors mdproptools/__init__.py:1)

__version__ = "0.1.0"
