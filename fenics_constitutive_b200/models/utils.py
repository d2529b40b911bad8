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
from .interfaces import IncrSmallStrainModel, StressStrainConstraint

__all__ = [
    "lame_parameters",
    "get_elastic_tangent",
    "get_identity",
    "strain_from_grad_u",
    "UniaxialStrainFrom3D",
    "PlaneStrainFrom3D",
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


class _From3D(IncrSmallStrainModel):
    """Shared body of the 3D -> 1D/2D adapters (reference models/utils.py:211-412): the mapped
    components of grad_del_u / stress are written into persistent 3D scratch arrays, the FULL
    model is evaluated on them, and the mapped components of the 3D stress / tangent are
    copied back.  Host numpy arrays take numpy slicing (as in the reference); CUDA tensors
    take the fcx_embed_3d / fcx_extract_from_3d kernels and never leave HBM."""

    _constraint: StressStrainConstraint

    def __init__(self, model) -> None:
        assert model.constraint == _C.FULL
        self.model = model
        self.stress_3d = None
        self.tangent_3d = None
        self.grad_del_u_3d = None

    @property
    def constraint(self) -> StressStrainConstraint:
        return self._constraint

    @property
    def history_dim(self):
        return self.model.history_dim

    def update(self) -> None:
        self.model.update()

    def _scratch(self, n: int, like):
        if isinstance(like, np.ndarray):
            make = lambda m: np.zeros(m)  # noqa: E731
        else:
            import torch

            make = lambda m: torch.zeros(m, dtype=torch.float64, device=like.device)  # noqa: E731
        if self.tangent_3d is None:
            self.tangent_3d = make(36 * n)
        if self.stress_3d is None:
            self.stress_3d = make(6 * n)
        if self.grad_del_u_3d is None:
            self.grad_del_u_3d = make(9 * n)

    def evaluate(self, time, del_t, grad_del_u, mandel_stress, tangent, history) -> None:
        g, s = self.geometric_dim, self.stress_strain_dim
        bg = B.as_buf(grad_del_u, "grad_del_u")
        n = bg.size // (g * g)
        self._scratch(n, grad_del_u)
        if bg.kind == B.HOST:
            g3, s3, t3 = self.grad_del_u_3d.reshape(-1, 9), self.stress_3d.reshape(-1, 6), self.tangent_3d.reshape(-1, 36)
            if g == 1:
                g3[:, 0] = grad_del_u          # :285-287
                s3[:, 0] = mandel_stress       # :289-292
            else:
                g2 = grad_del_u.reshape(-1, 4)
                g3[:, 0:2] = g2[:, 0:2]        # :375-376
                g3[:, 3:5] = g2[:, 2:4]
                s3[:, 0:4] = mandel_stress.reshape(-1, 4)  # :384-386
            self.model.evaluate(time, del_t, self.grad_del_u_3d, self.stress_3d, self.tangent_3d, history)
            tangent.reshape(-1, s, s)[:] = t3.reshape(-1, 6, 6)[:, :s, :s]   # :299-302 / :393-412
            mandel_stress.reshape(-1, s)[:] = s3[:, :s]                     # :293-297 / :388-391
            return
        L = lib()
        bs = B.as_buf(mandel_stress, "stress", writable=True)
        bt = B.as_buf(tangent, "tangent", writable=True)
        b3 = [B.as_buf(a, "scratch", writable=True) for a in (self.grad_del_u_3d, self.stress_3d, self.tangent_3d)]
        B.common_kind([bg, bs, bt] + b3)
        check(L.fcx_set_device(bg.device_index))
        stream = B.current_stream_ptr(bg.device_index)
        check(L.fcx_embed_3d(self._constraint.value, n, bg.ptr, bs.ptr, b3[0].ptr, b3[1].ptr, stream), "fcx_embed_3d")
        self.model.evaluate(time, del_t, self.grad_del_u_3d, self.stress_3d, self.tangent_3d, history)
        check(L.fcx_extract_from_3d(self._constraint.value, n, b3[1].ptr, b3[2].ptr, bs.ptr, bt.ptr, stream),
              "fcx_extract_from_3d")


class UniaxialStrainFrom3D(_From3D):
    """reference models/utils.py:211-297"""

    _constraint = _C.UNIAXIAL_STRAIN


class PlaneStrainFrom3D(_From3D):
    """reference models/utils.py:300-412"""

    _constraint = _C.PLANE_STRAIN
