"""VonMises3D on the GPU: J2 plasticity with saturating isotropic hardening,
radial return (scalar Newton) and the 6x6 consistent tangent.

Reference: src/fenics_constitutive/models/mises_plasticity_isotropic_hardening.py:9-186
Kernel: MisesModel<EPS_SOA> in csrc/fcx_models.cuh via fcx_mises_evaluate[_host].
"""
from __future__ import annotations

import numpy as np

from .. import _buffers as B
from .._lib import check, lib
from ._base import CudaModel
from .interfaces import StressStrainConstraint
from .utils import get_identity

_NEWTON_MSG = "Newton-Raphson method did not converge for plastic multiplier."
_INT_MAX = 2**31 - 1


class VonMises3D(CudaModel):
    """Args:
        param: ``{"p_ka": bulk modulus, "p_mu": shear modulus, "p_y0": initial
            yield stress, "p_y00": saturation yield stress, "p_w": saturation rate}``
            (reference :51-55).

    History ``{"eps_n": 6, "alpha": 1}`` (reference :184-186): plastic strain and
    accumulated plastic multiplier.  FULL constraint only (reference :180-182).

    Extras that do not exist in the reference (all default to its behaviour):
        record_plastic_flag: if True, ``self.plastic_flag`` holds a uint8 array
            (numpy or CUDA tensor, matching the inputs) with 1 where the trial
            state was plastic (``phitr > 0``, reference :98) after ``evaluate``.
        defer_errors: CUDA-tensor path only.  False (default): ``evaluate``
            synchronises and raises RuntimeError on Newton failure exactly like
            the reference (:141-143).  True: enqueue only; call
            ``check_converged()`` when convenient.
        eps_layout: "aos" (reference contract, [n][6]) or "soa" (six planes
            [6][n]) for device-resident plastic strain; CUDA-tensor path only.
    """

    def __init__(self, param: dict[str, float]):
        self.xioi = np.zeros((6, 6), dtype=np.int64)
        self.xioi[:3, :3] = 1  # 1 (x) 1, reference :33-42
        self.I2 = get_identity(self.stress_strain_dim, self.constraint)
        self.I4 = np.eye(self.stress_strain_dim, dtype=np.float64)
        self.xpp = self.I4 - (1 / 3) * self.xioi  # deviatoric projector, reference :48
        self.p_ka = param["p_ka"]
        self.p_mu = param["p_mu"]
        self.p_y0 = param["p_y0"]
        self.p_y00 = param["p_y00"]
        self.p_w = param["p_w"]
        self.record_plastic_flag = False
        self.plastic_flag = None
        self.defer_errors = False
        self.eps_layout = "aos"
        self._status = {}  # device index -> int32[2] tensor [failed points, first failing index]

    def _params(self) -> np.ndarray:
        return np.array([self.p_ka, self.p_mu, self.p_y0, self.p_y00, self.p_w], dtype=np.float64)

    def _status_tensor(self, dev):
        import torch

        if dev not in self._status:
            self._status[dev] = torch.tensor([0, _INT_MAX], dtype=torch.int32, device=f"cuda:{dev}")
        return self._status[dev]

    def check_converged(self) -> None:
        """Raise RuntimeError if any Newton iteration launched since the last check failed
        (synchronises the device).  The device status word is reset when a failure is reported,
        so the count and the first index of the next report belong to later calls only."""
        import torch

        for dev, st in self._status.items():
            count, first = (int(v) for v in st.cpu())
            if count > 0:
                st.copy_(torch.tensor([0, _INT_MAX], dtype=torch.int32))
                raise RuntimeError(f"{_NEWTON_MSG} ({count} point(s), first index {first})")

    def evaluate(self, time, del_t, grad_del_u, mandel_stress, tangent, history) -> None:
        n, kind, bufs, dev = self._collect(
            grad_del_u, mandel_stress, tangent,
            [("eps_n", history["eps_n"], 6), ("alpha", history["alpha"], 1)],
        )
        bg, bs, bt, be, ba = bufs
        P = self._params()
        L = lib()
        if kind == B.HOST:
            if self.eps_layout != "aos":
                raise ValueError("eps_layout='soa' is only available for CUDA tensors")
            flag = np.zeros(n, dtype=np.uint8) if self.record_plastic_flag else None
            rc = L.fcx_mises_evaluate_host(
                P.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr, be.ptr, ba.ptr,
                flag.ctypes.data if flag is not None else None,
            )
            self.plastic_flag = flag
            if check(rc, "VonMises3D.evaluate") > 0:
                raise RuntimeError(_NEWTON_MSG)
            return
        import torch

        stream = self._bind(dev)
        flag = None
        if self.record_plastic_flag:
            flag = torch.zeros(n, dtype=torch.uint8, device=f"cuda:{dev}")
        status = self._status_tensor(dev)
        layout = {"aos": 0, "soa": 1}[self.eps_layout]
        rc = L.fcx_mises_evaluate(
            P.ctypes.data, n, bg.ptr, bs.ptr, bt.ptr, be.ptr, ba.ptr, layout,
            flag.data_ptr() if flag is not None else None, status.data_ptr(), stream,
        )
        self.plastic_flag = flag
        check(rc, "VonMises3D.evaluate")
        if not self.defer_errors:
            self.check_converged()

    @property
    def constraint(self) -> StressStrainConstraint:
        return StressStrainConstraint.FULL

    @property
    def history_dim(self) -> dict[str, int]:
        return {"eps_n": self.constraint.stress_strain_dim, "alpha": 1}
