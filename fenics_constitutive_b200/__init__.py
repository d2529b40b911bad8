"""fenics_constitutive_b200 -- B200-native (sm_100a, fp64) per-quadrature-point
constitutive updates behind fenics-constitutive's ``IncrSmallStrainModel``
interface.

Layout
    models/      host-side mirror of the reference's ``fenics_constitutive.models``
                 (same class names, constructor dicts, history keys, exceptions);
                 every ``evaluate`` forwards to the C ABI below
    gather.py    companion operator: grad_del_u at the quadrature points
    partition.py QP-axis sharding helpers for one-rank-per-GPU runs
    _lib.py      ctypes binding of libfcx.so (include/fcx.h); fails loudly if the
                 CUDA library is missing -- there is no CPU fallback
    csrc/        hand-written CUDA kernels + the C ABI
"""
from __future__ import annotations

from . import models  # noqa: F401
from .models import (  # noqa: F401
    IncrSmallStrainModel,
    LinearElasticityModel,
    SpringKelvinModel,
    SpringMaxwellModel,
    StressStrainConstraint,
    VonMises3D,
)

__version__ = "0.1.0"
