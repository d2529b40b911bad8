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

import numpy as np


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
    """Rows ``cells`` of the parent <-> all rows of the sub array (device index ops)."""

    def __init__(self, cells: np.ndarray, num_parent_cells: int, device):
        import torch

        self.cell_map = np.asarray(cells, dtype=np.int64)
        self.num_parent_cells = int(num_parent_cells)
        self._idx = torch.as_tensor(self.cell_map, dtype=torch.int64, device=device)

    def map_to_parent(self, sub, parent) -> None:
        n = self.cell_map.size
        if n == 0:
            return
        parent.view(self.num_parent_cells, -1).index_copy_(0, self._idx, sub.view(n, -1))

    def map_to_sub(self, parent, sub) -> None:
        import torch

        n = self.cell_map.size
        if n == 0:
            return
        torch.index_select(parent.view(self.num_parent_cells, -1), 0, self._idx, out=sub.view(n, -1))


def build_subspace_map(cells: np.ndarray, num_parent_cells: int, device):
    """IdentityMap if the law owns every cell (reference maps.py:145-146), else a SubSpaceMap."""
    if len(cells) == num_parent_cells:
        return IdentityMap()
    return SubSpaceMap(cells, num_parent_cells, device)
