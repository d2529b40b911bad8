"""Shared plumbing of the CUDA-backed models: argument checks with the
reference's exceptions, host/device dispatch, device binding."""
from __future__ import annotations

from .. import _buffers as B
from .._lib import check, lib
from .interfaces import IncrSmallStrainModel


class CudaModel(IncrSmallStrainModel):
    """Base of the four models.  Subclasses implement `_launch_host` and
    `_launch_device`; `evaluate` keeps the reference's signature and in-place
    semantics (reference models/interfaces.py:82-101)."""

    def _collect(self, grad_del_u, stress, tangent, history_items):
        """Validate sizes like the reference (linear_elasticity_model.py:36-40)
        and classify the arrays.  Returns (n, kind, bufs, device_index)."""
        g, s = self.geometric_dim, self.stress_strain_dim
        bg = B.as_buf(grad_del_u, "grad_del_u")
        bs = B.as_buf(stress, "stress", writable=True)
        # tangent=None: stress-only evaluate, as the reference's compiled models allow
        # (`tangent: Option<..>`, bindings/src/lib.rs:83,109-113; comfe-rs/src/interfaces.rs:368)
        bt = B.as_buf(tangent, "tangent", writable=True) if tangent is not None else B.NullBuf(bg)
        assert bg.size // (g**2) == bs.size // s and (tangent is None or bs.size // s == bt.size // (s**2)), (
            "grad_del_u, stress and tangent disagree on the number of quadrature points"
        )
        n = bg.size // (g**2)
        bufs = [bg, bs, bt]
        for name, arr, dim in history_items:
            bh = B.as_buf(arr, f"history['{name}']", writable=True)
            assert bh.size == n * dim, f"history['{name}'] has {bh.size} entries, expected {n * dim}"
            bufs.append(bh)
        kind = B.common_kind(bufs)
        return n, kind, bufs, bufs[0].device_index

    @staticmethod
    def _bind(device_index):
        """Bind libfcx to the device that owns the tensors; return its stream."""
        check(lib().fcx_set_device(device_index), "fcx_set_device")
        return B.current_stream_ptr(device_index)
