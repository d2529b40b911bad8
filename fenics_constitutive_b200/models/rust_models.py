"""CUDA-backed mirrors of the reference's Rust models
(src/fenics_constitutive/models/rust_models.py:84-161, compiled crate comfe-rs):

    LinearElasticity3D({"mu": array, "kappa": array})
    MisesPlasticityLinearHardening3D({"mu", "kappa", "y_0", "h"})
    DruckerPrager3D({"mu", "kappa", "a", "b", "b_flow"})
    DruckerPragerHyperbolic3D({"mu", "kappa", "a", "b", "d", "b_flow"})

Same constructor shape as the pyo3 classes (parameter values are length-1 numpy
arrays, bindings/src/lib.rs:61-74), same ``history_dim`` convention -- ONE history
array under the key ``"history"`` (bindings/src/lib.rs:90-100,131-136) -- and FULL
constraint only.
"""
from __future__ import annotations

import numpy as np

from .. import _buffers as B
from .._lib import check, lib
from ._base import CudaModel
from .interfaces import StressStrainConstraint

__all__ = ["LinearElasticity3D", "MisesPlasticityLinearHardening3D", "DruckerPrager3D",
           "DruckerPragerHyperbolic3D"]

_INT_MAX = 2**31 - 1


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


class _DruckerPragerBase(CudaModel):
    """comfe-rs ``IsotropicPlasticityModel3D<..., MODEL>`` (comfe-rs/src/plasticity/general.rs:105-266):
    implicit return mapping on [sigma, del_lambda, kappa] with the reference's Newton stop rule
    (atol = rtol = 1e-8, maxit 25) and the consistent tangent ``inverse(dres)[0:6,0:6] C``.
    History ``{"history": 7}`` = [alpha, plastic_strain[6]] per quadrature point.

    Where the Rust code panics (Newton not converged, or the classic model's
    ``assert!(i_1 < a/b)`` at the apex of the cone) ``evaluate`` raises RuntimeError; the
    offending points keep their input stress and history.

    ``record_plastic_flag`` (extra, default False): keep a uint8 array with 1 where the
    trial state violated the yield condition (``f > 0``) in ``self.plastic_flag``."""

    _keys: tuple = ()
    _hyperbolic = 0
    _MSG = "Plasticity3D: Newton-Raphson did not converge"

    def __init__(self, parameters: dict[str, np.ndarray]) -> None:
        self.parameters = np.array([_scalar(parameters, k) for k in self._keys], dtype=np.float64)
        self.record_plastic_flag = False
        self.plastic_flag = None
        self._status = {}

    def _status_tensor(self, dev):
        import torch

        if dev not in self._status:
            self._status[dev] = torch.tensor([0, _INT_MAX], dtype=torch.int32, device=f"cuda:{dev}")
        return self._status[dev]

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history) -> None:
        if history is None or "history" not in history:
            raise ValueError("'history' entry not found in input")  # bindings/src/lib.rs:93-95
        n, kind, bufs, dev = self._collect(grad_del_u, stress, tangent, [("history", history["history"], 7)])
        bg, bs, bt, bh = bufs
        P = self.parameters
        L = lib()
        name = type(self).__name__
        if kind == B.HOST:
            flag = np.zeros(n, dtype=np.uint8) if self.record_plastic_flag else None
            rc = L.fcx_drucker_prager_evaluate_host(
                self._hyperbolic, P.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr, bh.ptr,
                flag.ctypes.data if flag is not None else None)
            self.plastic_flag = flag
            if check(rc, f"{name}.evaluate") > 0:
                raise RuntimeError(f"{self._MSG} ({rc} point(s))")
            return
        import torch

        flag = torch.zeros(n, dtype=torch.uint8, device=f"cuda:{dev}") if self.record_plastic_flag else None
        status = self._status_tensor(dev)
        stream = self._bind(dev)
        rc = L.fcx_drucker_prager_evaluate(
            self._hyperbolic, P.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr, bh.ptr,
            flag.data_ptr() if flag is not None else None, status.data_ptr(), stream)
        self.plastic_flag = flag
        check(rc, f"{name}.evaluate")
        count, first = (int(x) for x in status.cpu())
        if count > 0:
            status.copy_(torch.tensor([0, _INT_MAX], dtype=torch.int32))
            raise RuntimeError(f"{self._MSG} ({count} point(s), first index {first})")

    @property
    def constraint(self) -> StressStrainConstraint:
        return StressStrainConstraint.FULL

    @property
    def history_dim(self) -> dict[str, int]:
        return {"history": 7}

    @property
    def symmetric_tangent(self) -> bool:
        """Non-associated flow (b_flow != b) makes the consistent tangent non-symmetric."""
        b = self.parameters[self._keys.index("b")]
        return bool(b == self.parameters[self._keys.index("b_flow")])


class DruckerPrager3D(_DruckerPragerBase):
    """Classic Drucker-Prager, f = sqrt(J2) + b I1 - a, flow potential slope ``b_flow``
    (comfe-rs/src/plasticity/drucker_prager_classic.rs:46-166; reference wrapper
    models/rust_models.py:96-116).  Parameters ``mu, kappa, a, b, b_flow`` (length-1 arrays)."""

    _keys = ("mu", "kappa", "a", "b", "b_flow")
    _hyperbolic = 0


class DruckerPragerHyperbolic3D(_DruckerPragerBase):
    """Hyperbolically smoothed Drucker-Prager, f = sqrt(J2 + d^2) + b I1 - a
    (comfe-rs/src/plasticity/drucker_prager_hyperbolic.rs:48-164; reference wrapper
    models/rust_models.py:119-141).  Parameters ``mu, kappa, a, b, d, b_flow``."""

    _keys = ("mu", "kappa", "a", "b", "d", "b_flow")
    _hyperbolic = 1
