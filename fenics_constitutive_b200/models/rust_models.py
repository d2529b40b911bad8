"""CUDA-backed mirrors of the reference's Rust models
(src/fenics_constitutive/models/rust_models.py:84-161, compiled crate comfe-rs):

    LinearElasticity3D({"mu": array, "kappa": array})
    MisesPlasticityLinearHardening3D({"mu", "kappa", "y_0", "h"})

Same constructor shape as the pyo3 classes (parameter values are length-1 numpy
arrays, bindings/src/lib.rs:61-74), same ``history_dim`` convention -- ONE history
array under the key ``"history"`` (bindings/src/lib.rs:90-100,131-136) -- and FULL
constraint only.  The Drucker-Prager models of the crate are not exported by the
reference's ``models.__all__`` and are out of scope.
"""
from __future__ import annotations

import numpy as np

from .. import _buffers as B
from .._lib import check, lib
from ._base import CudaModel
from .interfaces import StressStrainConstraint

__all__ = ["LinearElasticity3D", "MisesPlasticityLinearHardening3D"]


def _scalar(parameters, key: str) -> float:
    return float(np.asarray(parameters[key], dtype=np.float64).reshape(-1)[0])


class LinearElasticity3D(CudaModel):
    """comfe-rs ``LinearElasticity3D`` (comfe-rs/src/linear_elasticity.rs:49-74):
    C = 2 mu P_dev + 3 kappa P_vol (mandel.rs:126-128), stress += C de, tangent = C.
    Runs on the elastic FULL kernel with that C as the tangent matrix."""

    def __init__(self, parameters: dict[str, np.ndarray]) -> None:
        self.mu = _scalar(parameters, "mu")
        self.kappa = _scalar(parameters, "kappa")
        oo = np.zeros((6, 6))
        oo[:3, :3] = 1.0
        p_vol = oo * (1.0 / 3.0)              # consts.rs:106-108
        p_dev = np.eye(6) + p_vol * -1.0      # consts.rs:113-115
        self.D = (2.0 * self.mu) * p_dev + (3.0 * self.kappa) * p_vol

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history=None) -> None:
        n, kind, (bg, bs, bt), dev = self._collect(grad_del_u, stress, tangent, [])
        D = np.ascontiguousarray(self.D, dtype=np.float64)
        L = lib()
        if kind == B.HOST:
            rc = L.fcx_elastic_evaluate_host(self.constraint.value, D.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr)
        else:
            rc = L.fcx_elastic_evaluate(self.constraint.value, D.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr,
                                        self._bind(dev))
        check(rc, "LinearElasticity3D.evaluate")

    @property
    def constraint(self) -> StressStrainConstraint:
        return StressStrainConstraint.FULL

    @property
    def history_dim(self) -> None:
        return None


class MisesPlasticityLinearHardening3D(CudaModel):
    """comfe-rs ``MisesPlasticity3D`` (comfe-rs/src/mises_plasticity.rs:58-126): J2 plasticity,
    linear isotropic hardening sigma_y = y_0 + h alpha, closed-form radial return.
    History ``{"history": 7}`` = [alpha, plastic_strain[6]] per quadrature point.

    ``record_plastic_flag`` (extra, default False): keep a uint8 array with 1 where
    ``s_tr_eq >= sigma_y`` in ``self.plastic_flag``."""

    def __init__(self, parameters: dict[str, np.ndarray]) -> None:
        self.parameters = np.array([_scalar(parameters, k) for k in ("mu", "kappa", "y_0", "h")],
                                   dtype=np.float64)
        self.record_plastic_flag = False
        self.plastic_flag = None

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history) -> None:
        if history is None or "history" not in history:
            raise ValueError("'history' entry not found in input")  # bindings/src/lib.rs:93-95
        n, kind, bufs, dev = self._collect(grad_del_u, stress, tangent, [("history", history["history"], 7)])
        bg, bs, bt, bh = bufs
        P = self.parameters
        L = lib()
        if kind == B.HOST:
            flag = np.zeros(n, dtype=np.uint8) if self.record_plastic_flag else None
            rc = L.fcx_mises_linear_hardening_evaluate_host(
                P.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr, bh.ptr, flag.ctypes.data if flag is not None else None)
        else:
            import torch

            flag = torch.zeros(n, dtype=torch.uint8, device=f"cuda:{dev}") if self.record_plastic_flag else None
            rc = L.fcx_mises_linear_hardening_evaluate(
                P.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr, bh.ptr,
                flag.data_ptr() if flag is not None else None, self._bind(dev))
        self.plastic_flag = flag
        check(rc, "MisesPlasticityLinearHardening3D.evaluate")

    @property
    def constraint(self) -> StressStrainConstraint:
        return StressStrainConstraint.FULL

    @property
    def history_dim(self) -> dict[str, int]:
        return {"history": 7}
