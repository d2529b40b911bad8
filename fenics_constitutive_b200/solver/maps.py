"""Device-side twins of the reference's parent <-> submesh maps for quadrature
arrays (reference src/fenics_constitutive/solver/maps.py:28-178).

A quadrature array is ``[cell][qp][...]`` flat (reference
tests/solver/test_maps.py:119-121), so the map of a law that owns the cell list
``cells`` is a row gather/scatter of the ``[num_cells][row]`` view:
    map_to_sub:     sub[arange(len(cells))] = parent[cells]
    map_to_parent:  parent[cells]           = sub[arange(len(cells))]
``IdentityMap`` (law on all cells) copies like the reference does (:37-59);
``IncrSmallStrainProblem`` avoids even that copy by aliasing the arrays.
"""
from __future__ import annotations

from typing import Protocol, runtime_checkable

import numpy as np

__all__ = ["IdentityMap", "SpaceMap", "SubSpaceMap", "build_subspace_map"]  # reference maps.py:11


@runtime_checkable
class SpaceMap(Protocol):
    """What a parent <-> sub-space map offers (reference solver/maps.py:14-26); arrays are CUDA tensors here."""

    def map_to_parent(self, sub, parent) -> None:
        """sub -> the rows of parent this map covers."""

    def map_to_sub(self, parent, sub) -> None:
        """the rows of parent this map covers -> sub."""


class IdentityMap:
    def map_to_parent(self, sub, parent) -> None:
        assert sub.shape == parent.shape, "Shapes do not match"
        if sub.data_ptr() != parent.data_ptr():
            parent.copy_(sub)

    def map_to_sub(self, parent, sub) -> None:
        assert sub.shape == parent.shape, "Shapes do not match"
        if sub.data_ptr() != parent.data_ptr():
            sub.copy_(parent)


class SubSpaceMap:
    """Rows ``cells`` of the parent <-> all rows of the sub array: the row gather / scatter
    kernels of csrc/fcx_maps.cu (``fcx_map_rows_to_sub`` / ``fcx_map_rows_to_parent``).  CUDA
    tensors only -- there is no CPU fallback."""

    def __init__(self, cells: np.ndarray, num_parent_cells: int, device):
        import torch

        self.cell_map = np.asarray(cells, dtype=np.int64)
        self.num_parent_cells = int(num_parent_cells)
        assert self.cell_map.size == 0 or (0 <= self.cell_map.min() and self.cell_map.max() < self.num_parent_cells)
        self._cells = torch.as_tensor(self.cell_map, dtype=torch.int32, device=device).contiguous()

    def cells_ptr(self) -> int:
        """Device address of the int32 cell list (the ``cells`` argument of fcx_mises_form)."""
        return self._cells.data_ptr()

    def _rows(self, sub, parent, to_parent: bool) -> None:
        from .. import _buffers as B
        from .._lib import check, lib

        n = self.cell_map.size
        if n == 0:
            return
        bs, bp = B.as_buf(sub, "sub", writable=not to_parent), B.as_buf(parent, "parent", writable=to_parent)
        if B.common_kind([bs, bp]) != B.DEVICE:
            raise ValueError("SubSpaceMap needs CUDA tensors")
        assert bs.size % n == 0, "sub array is not a whole number of rows"
        row = bs.size // n
        assert bp.size == row * self.num_parent_cells, "Shapes do not match"
        L = lib()
        check(L.fcx_set_device(bs.device_index), "fcx_set_device")
        stream = B.current_stream_ptr(bs.device_index)
        if to_parent:
            rc = L.fcx_map_rows_to_parent(n, row, self._cells.data_ptr(), bs.ptr, bp.ptr, stream)
        else:
            rc = L.fcx_map_rows_to_sub(n, row, self._cells.data_ptr(), bp.ptr, bs.ptr, stream)
        check(rc, "SubSpaceMap")

    def map_to_parent(self, sub, parent) -> None:
        """reference solver/maps.py:104-123"""
        self._rows(sub, parent, True)

    def map_to_sub(self, parent, sub) -> None:
        """reference solver/maps.py:82-102"""
        self._rows(sub, parent, False)


def build_subspace_map(cells: np.ndarray, num_parent_cells: int, device):
    """IdentityMap if the law owns every cell (reference maps.py:145-146), else a SubSpaceMap."""
    if len(cells) == num_parent_cells:
        return IdentityMap()
    return SubSpaceMap(cells, num_parent_cells, device)
