"""Drop-in mirror of ``fenics_constitutive.models`` (reference
src/fenics_constitutive/models/__init__.py:1-21): same public names, every
``evaluate`` runs on the B200 through libfcx.so."""
from __future__ import annotations

from .interfaces import *  # noqa: F401,F403
from .interfaces import IncrSmallStrainModel, StressStrainConstraint
from .linear_elasticity_model import LinearElasticityModel
from .mises_plasticity_isotropic_hardening import VonMises3D
from .rust_models import (DruckerPrager3D, DruckerPragerHyperbolic3D, LinearElasticity3D,
                          MisesPlasticityLinearHardening3D)
from .spring_kelvin_model import SpringKelvinModel
from .spring_maxwell_model import SpringMaxwellModel
from .utils import *  # noqa: F401,F403

__all__ = [
    "DruckerPrager3D",
    "DruckerPragerHyperbolic3D",
    "IncrSmallStrainModel",
    "LinearElasticity3D",
    "LinearElasticityModel",
    "MisesPlasticityLinearHardening3D",
    "SpringKelvinModel",
    "SpringMaxwellModel",
    "StressStrainConstraint",
    "VonMises3D",
]
