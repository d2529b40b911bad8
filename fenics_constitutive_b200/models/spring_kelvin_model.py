"""SpringKelvinModel on the GPU.

Reference: src/fenics_constitutive/models/spring_kelvin_model.py:9-99 -- spring in
series with a Kelvin body (three-parameter solid), deviatoric viscous strain,
backward Euler.  Kernel: KelvinModel<S,G> in csrc/fcx_models.cuh.
"""
from __future__ import annotations

import numpy as np

from .. import _buffers as B
from .._lib import check, lib
from ._base import CudaModel
from .interfaces import StressStrainConstraint
from .utils import get_elastic_tangent, get_identity, lame_parameters


class SpringKelvinModel(CudaModel):
    """Args:
        parameters: ``{"E0", "E1", "tau", "nu"}`` (``nu`` is forced to 0 for
            UNIAXIAL_STRESS, reference :33-36).
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
        self.I2 = get_identity(self.stress_strain_dim, constraint)
        self.mu0, self.lam0 = lame_parameters(self.E0, self.nu)
        self.mu1, _ = lame_parameters(self.E1, self.nu)

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history) -> None:
        _ = t
        s = self.stress_strain_dim
        if history is None:
            # reference :63-65 (after its size assertion, which needs no history)
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
        I2 = np.ascontiguousarray(self.I2, dtype=np.float64)
        L = lib()
        common = (self._constraint.value, D0.ctypes.data, I2.ctypes.data, self.mu0, self.lam0,
                  self.mu1, self.tau, del_t, n, bg.ptr, bs.ptr, bt.ptr, bev.ptr, bet.ptr)
        if kind == B.HOST:
            rc = L.fcx_kelvin_evaluate_host(*common)
        else:
            rc = L.fcx_kelvin_evaluate(*common, self._bind(dev))
        check(rc, "SpringKelvinModel.evaluate")

    @property
    def constraint(self) -> StressStrainConstraint:
        return self._constraint

    @property
    def history_dim(self) -> dict[str, int]:
        return {"strain_visco": self.stress_strain_dim, "strain": self.stress_strain_dim}
