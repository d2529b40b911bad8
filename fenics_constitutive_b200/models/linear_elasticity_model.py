"""LinearElasticityModel on the GPU.

Reference: src/fenics_constitutive/models/linear_elasticity_model.py:10-53
  stress += strain_from_grad_u(grad_del_u) @ D ; tangent[:] = tile(D.flatten(), n)
Kernel: ElasticModel<S,G> in csrc/fcx_models.cuh via fcx_elastic_evaluate[_host].
"""
from __future__ import annotations

import numpy as np

from .. import _buffers as B
from .._lib import check, lib
from ._base import CudaModel
from .interfaces import StressStrainConstraint
from .utils import get_elastic_tangent


class LinearElasticityModel(CudaModel):
    """Linear elasticity for all five constraints.

    Args:
        parameters: ``{"E": Young's modulus, "nu": Poisson ratio}``.
        constraint: the stress-strain constraint.
    Public attribute ``D`` (s x s) as in the reference (:24); it is re-read on
    every ``evaluate``, so user modifications take effect.
    """

    def __init__(self, parameters: dict[str, float], constraint: StressStrainConstraint):
        self._constraint = constraint
        self.D = get_elastic_tangent(parameters["E"], parameters["nu"], constraint)

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history=None) -> None:
        # unused, as in the reference: t, del_t, history
        n, kind, (bg, bs, bt), dev = self._collect(grad_del_u, stress, tangent, [])
        D = np.ascontiguousarray(self.D, dtype=np.float64)
        L = lib()
        if kind == B.HOST:
            rc = L.fcx_elastic_evaluate_host(
                self._constraint.value, D.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr
            )
        else:
            stream = self._bind(dev)
            rc = L.fcx_elastic_evaluate(
                self._constraint.value, D.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr, stream
            )
        check(rc, "LinearElasticityModel.evaluate")

    @property
    def constraint(self) -> StressStrainConstraint:
        return self._constraint

    @property
    def history_dim(self) -> None:
        return None
