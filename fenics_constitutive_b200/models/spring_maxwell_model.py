"""SpringMaxwellModel on the GPU.

Reference: src/fenics_constitutive/models/spring_maxwell_model.py:8-99 -- spring
parallel to a Maxwell arm, deviatoric viscous strain, backward Euler.
Kernel: MaxwellModel<S,G> in csrc/fcx_models.cuh.
"""
from __future__ import annotations

import numpy as np

from .. import _buffers as B
from .._lib import check, lib
from ._base import CudaModel
from .interfaces import StressStrainConstraint
from .utils import get_elastic_tangent, lame_parameters


class SpringMaxwellModel(CudaModel):
    """Args:
        parameters: ``{"E0", "E1", "tau", "nu"}`` (``nu`` := 0 for UNIAXIAL_STRESS,
            reference :31-34).
        constraint: the stress-strain constraint.
    History: ``{"strain_visco": s, "strain": s}`` (reference :94-99).
    """

    def __init__(self, parameters: dict[str, float], constraint: StressStrainConstraint):
        self._constraint = constraint
        self.E0 = parameters["E0"]
        self.E1 = parameters["E1"]
        self.tau = parameters["tau"]
        if constraint == StressStrainConstraint.UNIAXIAL_STRESS:
            self.nu = 0.0
        else:
            self.nu = parameters["nu"]
        self.D_0 = get_elastic_tangent(self.E0, self.nu, constraint)
        self.D_1 = get_elastic_tangent(self.E1, self.nu, constraint)
        self.mu1, _ = lame_parameters(self.E1, self.nu)

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history) -> None:
        _ = t
        s = self.stress_strain_dim
        if history is None:
            self._collect(grad_del_u, stress, tangent, [])
            msg = "history must not be None"
            raise ValueError(msg)
        n, kind, bufs, dev = self._collect(
            grad_del_u, stress, tangent,
            [("strain_visco", history["strain_visco"], s), ("strain", history["strain"], s)],
        )
        assert del_t > 0, "Time step must be defined and positive."
        bg, bs, bt, bev, bet = bufs
        D0 = np.ascontiguousarray(self.D_0, dtype=np.float64)
        D1 = np.ascontiguousarray(self.D_1, dtype=np.float64)
        L = lib()
        common = (self._constraint.value, D0.ctypes.data, D1.ctypes.data, self.mu1, self.tau,
                  del_t, n, bg.ptr, bs.ptr, bt.ptr, bev.ptr, bet.ptr)
        if kind == B.HOST:
            rc = L.fcx_maxwell_evaluate_host(*common)
        else:
            rc = L.fcx_maxwell_evaluate(*common, self._bind(dev))
        check(rc, "SpringMaxwellModel.evaluate")

    @property
    def constraint(self) -> StressStrainConstraint:
        return self._constraint

    @property
    def history_dim(self) -> dict[str, int]:
        return {"strain_visco": self.stress_strain_dim, "strain": self.stress_strain_dim}
