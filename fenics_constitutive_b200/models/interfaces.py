"""The model contract, mirrored from the reference so user code is drop-in.

Reference: src/fenics_constitutive/models/interfaces.py
  StressStrainConstraint  :14-73   (values 1..5; dims 1/1/4/4/6; gdim 1/1/2/2/3)
  IncrSmallStrainModel    :76-143  (evaluate / constraint / history_dim, derived dims)
The integer values double as the `constraint` codes of the C ABI (include/fcx.h).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum

import numpy as np

__all__ = ["IncrSmallStrainModel", "StressStrainConstraint"]

# value -> (Mandel dimension s, geometric dimension g)
_DIMS = {1: (1, 1), 2: (1, 1), 3: (4, 2), 4: (4, 2), 5: (6, 3)}


class StressStrainConstraint(Enum):
    """Constraint on the stress or strain state (reference interfaces.py:14-27)."""

    UNIAXIAL_STRAIN = 1
    UNIAXIAL_STRESS = 2
    PLANE_STRAIN = 3
    PLANE_STRESS = 4
    FULL = 5

    @property
    def stress_strain_dim(self) -> int:
        """Length of the Mandel stress/strain vector (reference :29-50)."""
        return _DIMS[self.value][0]

    @property
    def geometric_dim(self) -> int:
        """Spatial dimension of grad_del_u (reference :52-73)."""
        return _DIMS[self.value][1]


class IncrSmallStrainModel(ABC):
    """Interface for incremental small-strain models (reference :76-143)."""

    @abstractmethod
    def evaluate(
        self,
        t: float,
        del_t: float,
        grad_del_u: np.ndarray,
        stress: np.ndarray,
        tangent: np.ndarray,
        history: dict[str, np.ndarray] | None,
    ) -> None:
        """Update ``stress`` (sigma_n -> sigma_n+1), overwrite ``tangent`` and
        advance ``history`` in place for all quadrature points.

        Args:
            t: time at the start of the increment.
            del_t: time increment.
            grad_del_u: flat [n*g*g] gradient of the displacement increment
                (``ufl.nabla_grad`` convention).
            stress: flat [n*s] Mandel stress.
            tangent: flat [n*s*s] Mandel tangent, row-major per point.
            history: dict of flat history arrays or None.
        """

    @property
    @abstractmethod
    def constraint(self) -> StressStrainConstraint:
        """The constraint the model instance is built for."""

    @property
    def stress_strain_dim(self) -> int:
        return self.constraint.stress_strain_dim

    @property
    def geometric_dim(self) -> int:
        return self.constraint.geometric_dim

    @property
    @abstractmethod
    def history_dim(self) -> dict[str, int | tuple[int, int]] | None:
        """Name -> per-point dimension of each history array, or None."""
