"""Host-side constants and the Mandel conversion.

Reference: src/fenics_constitutive/models/utils.py
  lame_parameters      :18-22
  get_elastic_tangent  :25-93
  get_identity         :96-129
  strain_from_grad_u   :132-208   (here: CUDA kernel via fcx_strain_from_grad_u)
The small matrices are built on the host with the same floating-point
expressions as the reference, so the `D` handed to the kernels is bit-identical
to the reference's `self.D`.
"""
from __future__ import annotations

import numpy as np

from .. import _buffers as B
from .._lib import check, lib
from .interfaces import StressStrainConstraint

__all__ = [
    "lame_parameters",
    "get_elastic_tangent",
    "get_identity",
    "strain_from_grad_u",
]

_C = StressStrainConstraint


def lame_parameters(E: float, nu: float) -> tuple[float, float]:
    """(mu, lam) from Young's modulus and Poisson ratio (reference :18-22)."""
    mu = E / (2.0 * (1.0 + nu))
    lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    return mu, lam


def _isotropic_block(mu: float, lam: float, s: int) -> np.ndarray:
    """Normal 3x3 block [2mu+lam on the diagonal, lam off it] plus 2mu on the
    shear diagonal, for s = 4 (plane strain) or 6 (full)."""
    D = np.zeros((s, s))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] = 2.0 * mu + lam
    D[np.arange(3, s), np.arange(3, s)] = 2.0 * mu
    return D


def get_elastic_tangent(E: float, nu: float, constraint: StressStrainConstraint) -> np.ndarray:
    """Linear-elastic Mandel tangent for a constraint (reference :25-93)."""
    mu, lam = lame_parameters(E, nu)
    if constraint is _C.FULL:
        return _isotropic_block(mu, lam, 6)
    if constraint is _C.PLANE_STRAIN:
        return _isotropic_block(mu, lam, 4)
    if constraint is _C.PLANE_STRESS:
        pattern = np.zeros((4, 4))
        pattern[0, 0] = pattern[1, 1] = 1.0
        pattern[0, 1] = pattern[1, 0] = nu
        pattern[3, 3] = 1.0 - nu
        return E / (1 - nu**2.0) * pattern
    if constraint is _C.UNIAXIAL_STRAIN:
        return np.array([[E * (1.0 - nu) / ((1.0 + nu) * (1.0 - 2.0 * nu))]])
    if constraint is _C.UNIAXIAL_STRESS:
        return np.array([[E]])
    raise NotImplementedError("Constraint not implemented")


def get_identity(stress_strain_dim: int, constraint: StressStrainConstraint) -> np.ndarray:
    """Second-order identity in Mandel notation (reference :96-129)."""
    ones = {_C.FULL: 3, _C.PLANE_STRAIN: 3, _C.PLANE_STRESS: 2, _C.UNIAXIAL_STRAIN: 1,
            _C.UNIAXIAL_STRESS: 1}
    if constraint not in ones:
        raise NotImplementedError("Constraint not implemented")
    I2 = np.zeros(stress_strain_dim, dtype=np.float64)
    I2[: ones[constraint]] = 1.0
    return I2


def strain_from_grad_u(grad_u, constraint: StressStrainConstraint):
    """Mandel strain [n*s] from a displacement gradient [n*g*g] (reference :132-208).

    Runs the CUDA kernel.  A numpy input is uploaded, converted on the GPU and
    returned as numpy; a CUDA tensor input returns a CUDA tensor.
    """
    import torch

    s, g = constraint.stress_strain_dim, constraint.geometric_dim
    host = isinstance(grad_u, np.ndarray)
    if host:
        src = torch.from_numpy(np.ascontiguousarray(grad_u, dtype=np.float64).reshape(-1)).cuda()
    else:
        src = B.as_buf(grad_u, "grad_u").owner.reshape(-1)
    n = int(src.numel() / (g * g))
    out = torch.zeros(n * s, dtype=torch.float64, device=src.device)
    L = lib()
    check(L.fcx_set_device(src.device.index))
    check(
        L.fcx_strain_from_grad_u(
            constraint.value, n, src.data_ptr(), out.data_ptr(), B.current_stream_ptr(src.device.index)
        ),
        "strain_from_grad_u",
    )
    return out.cpu().numpy() if host else out
