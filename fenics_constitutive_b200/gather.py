"""Companion operator: grad_del_u at the quadrature points.

Reference semantics: ``IncrementalDisplacement.evaluate_local_incremental_gradient``
(src/fenics_constitutive/solver/_incrementalunknowns.py:19-27,40-49): a dolfinx
``Expression(ufl.nabla_grad(u - u_prev), q_points)`` interpolated cell by cell, i.e.

    grad[c][q][i][j] = d(u - u_prev)_j / dx_i        (cell-major QP order)

Here the same numbers come from ``fcx_gather_grad`` (csrc/fcx_gather.cu): cell
DOFs gathered against a precomputed reference basis-gradient table and the
per-cell inverse Jacobian of an affine simplex mesh.

In a dolfinx deployment the tables are produced once at set-up from basix
(``element.tabulate(1, q_points)[1:]`` -> dphi_ref, ``mesh.geometry`` -> Jinv,
``V.dofmap.list`` -> dofmap).  basix/dolfinx are not available in the build
image, so this module also carries its own P1/P2 Lagrange tables and simplex
quadrature points for the stand-in driver and the tests (quadrature point
ORDER inside a cell is therefore not pinned against basix -- DESIGN.md).
"""
from __future__ import annotations

import itertools

import numpy as np

from . import _buffers as B
from ._lib import check, lib

# ----------------------------------------------------------- reference tables

# sub-entity edges (vertex pairs) in basix/DOLFINx order
_EDGES = {
    1: [(0, 1)],
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)],
}


def lagrange_gradients(gdim: int, degree: int, points: np.ndarray) -> np.ndarray:
    """d phi_a / d X_k of the P1 / P2 Lagrange basis on the reference simplex at
    `points` [nq][gdim].  Returns [nq][nd][gdim].  Local numbering: vertices
    first, then edge midpoints in basix edge order."""
    points = np.atleast_2d(np.asarray(points, dtype=np.float64))
    nq = points.shape[0]
    lam = np.concatenate([1.0 - points.sum(axis=1, keepdims=True), points], axis=1)  # [nq][d+1]
    dlam = np.concatenate([-np.ones((1, gdim)), np.eye(gdim)], axis=0)              # [d+1][gdim]
    if degree == 1:
        return np.broadcast_to(dlam, (nq, gdim + 1, gdim)).copy()
    if degree != 2:
        raise NotImplementedError("Lagrange degree 1 or 2")
    edges = _EDGES[gdim]
    out = np.zeros((nq, gdim + 1 + len(edges), gdim))
    for i in range(gdim + 1):
        out[:, i, :] = (4.0 * lam[:, i : i + 1] - 1.0) * dlam[i]
    for e, (i, j) in enumerate(edges):
        out[:, gdim + 1 + e, :] = 4.0 * (lam[:, i : i + 1] * dlam[j] + lam[:, j : j + 1] * dlam[i])
    return out


def simplex_quadrature(gdim: int, degree: int) -> tuple[np.ndarray, np.ndarray]:
    """(points [nq][gdim], weights [nq]) exact to `degree` (1 or 2) on the reference simplex.
    The reference takes any degree basix offers (_incrementalunknowns.py:21-22); only the rules the
    fused kernels are specialised for exist here, and a higher degree is refused rather than
    silently integrated with the degree-2 rule (the QP count would differ from the reference's)."""
    if degree not in (0, 1, 2):
        raise NotImplementedError(f"simplex_quadrature: q_degree must be 1 or 2 (got {degree})")
    if gdim == 1:
        if degree <= 1:
            return np.array([[0.5]]), np.array([1.0])
        a = 0.5 - 0.5 / 3**0.5
        return np.array([[a], [1.0 - a]]), np.array([0.5, 0.5])
    if gdim == 2:
        if degree <= 1:
            return np.array([[1 / 3, 1 / 3]]), np.array([0.5])
        return (np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.full(3, 1 / 6))
    if gdim == 3:
        if degree <= 1:
            return np.array([[0.25, 0.25, 0.25]]), np.array([1 / 6])
        a, b = 0.1381966011250105, 0.5854101966249685
        pts = np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]])
        return pts, np.full(4, 1 / 24)
    raise ValueError("gdim must be 1, 2 or 3")


def affine_inverse_jacobians(coords: np.ndarray, cells_vertices: np.ndarray) -> np.ndarray:
    """Jinv[c][k][i] = dX_k/dx_i for affine simplices.  coords [nv][gdim],
    cells_vertices [ncells][gdim+1]."""
    x = coords[cells_vertices]                      # [nc][d+1][gdim]
    J = np.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)  # J[i][k] = dx_i/dX_k
    return np.linalg.inv(J)


# ------------------------------------------------------------------ operator

class IncrementalGradient:
    """Device-resident gather operator.  Build once per mesh/function space:

        op = IncrementalGradient(gdim, dofmap, dphi_ref, Jinv)
        op.evaluate(u, u_prev, grad_del_u)       # CUDA tensors, in place

    dofmap [ncells][nd] int32, dphi_ref [nq][nd][gdim], Jinv [ncells][gdim][gdim]
    (numpy or CUDA tensors; copied to the GPU once)."""

    def __init__(self, gdim: int, dofmap, dphi_ref, Jinv, device=None):
        import torch

        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.gdim = int(gdim)
        self.dofmap = torch.as_tensor(np.ascontiguousarray(dofmap) if isinstance(dofmap, np.ndarray) else dofmap,
                                      dtype=torch.int32, device=dev).contiguous()
        self.dphi_ref = torch.as_tensor(dphi_ref, dtype=torch.float64, device=dev).contiguous()
        self.Jinv = torch.as_tensor(Jinv, dtype=torch.float64, device=dev).contiguous()
        self.ncells, self.nd = self.dofmap.shape
        self.nq = self.dphi_ref.shape[0]
        assert self.dphi_ref.shape == (self.nq, self.nd, self.gdim)
        assert self.Jinv.shape == (self.ncells, self.gdim, self.gdim)
        self.device = dev
        # True: with u_prev given, form du = u - u_prev once as a nodal vector (fcx_nodal_increment) and gather
        # that -- every nodal value is fetched once instead of twice, same bits.  False: the kernel fetches both.
        self.form_increment_first = True
        self._du = None

    @property
    def num_qps(self) -> int:
        return self.ncells * self.nq

    def evaluate(self, u, u_prev, grad_del_u) -> None:
        """grad_del_u[c][q][i][j] <- d(u - u_prev)_j/dx_i; u_prev may be None."""
        bu = B.as_buf(u, "u")
        bp = B.as_buf(u_prev, "u_prev") if u_prev is not None else None
        bg = B.as_buf(grad_del_u, "grad_del_u", writable=True)
        bufs = [bu, bg] + ([bp] if bp is not None else [])
        if B.common_kind(bufs) != B.DEVICE:
            raise ValueError("IncrementalGradient.evaluate needs CUDA tensors")
        assert bg.size == self.num_qps * self.gdim**2, "grad_del_u has the wrong size"
        L = lib()
        check(L.fcx_set_device(bu.device_index))
        stream = B.current_stream_ptr(bu.device_index)
        u_ptr, prev_ptr = bu.ptr, (bp.ptr if bp is not None else None)
        if bp is not None and self.form_increment_first:
            assert bp.size == bu.size, "u and u_prev differ in size"
            import torch

            if self._du is None or self._du.numel() != bu.size or self._du.device.index != bu.device_index:
                self._du = torch.empty(bu.size, dtype=torch.float64, device=f"cuda:{bu.device_index}")
            check(L.fcx_nodal_increment(bu.size, bu.ptr, bp.ptr, self._du.data_ptr(), stream), "fcx_nodal_increment")
            u_ptr, prev_ptr = self._du.data_ptr(), None
        rc = L.fcx_gather_grad(
            self.gdim, self.ncells, self.nq, self.nd, self.dofmap.data_ptr(), u_ptr,
            prev_ptr, self.dphi_ref.data_ptr(), self.Jinv.data_ptr(),
            bg.ptr, stream,
        )
        check(rc, "IncrementalGradient.evaluate")


# ------------------------------------------------------- structured test mesh

def unit_cube_p2_tets(nx: int, ny: int, nz: int):
    """Kuhn triangulation of the unit cube into 6*nx*ny*nz tetrahedra with a P2
    (10-node) dofmap.  Nodes are the points of the doubled grid
    (2nx+1)(2ny+1)(2nz+1): cube vertices, edge midpoints, face and body centres
    are exactly the vertices and edge midpoints of the Kuhn tets.

    Returns (node_coords [nn][3], cells_vertices [nc][4] (node ids of the 4
    vertices), dofmap [nc][10] int32)."""
    gx, gy, gz = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1

    def nid(i, j, k):  # doubled-grid index -> node id
        return (i * gy + j) * gz + k

    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    base = np.stack([2 * ii.ravel(), 2 * jj.ravel(), 2 * kk.ravel()], axis=1)  # [ncube][3]
    tets = []
    for perm in itertools.permutations(range(3)):
        # path 000 -> +e_perm[0] -> +e_perm[1] -> +e_perm[2]
        v = [np.zeros(3, dtype=np.int64)]
        for ax in perm:
            nxt = v[-1].copy()
            nxt[ax] += 2
            v.append(nxt)
        tets.append(np.stack(v))  # [4][3] offsets on the doubled grid
    tets = np.stack(tets)  # [6][4][3]
    # positive orientation for every tet
    for t in range(6):
        e = (tets[t, 1:] - tets[t, 0]).astype(float)
        if np.linalg.det(e) < 0:
            tets[t, [2, 3]] = tets[t, [3, 2]]
    vert = base[:, None, None, :] + tets[None, :, :, :]            # [ncube][6][4][3]
    vert = vert.reshape(-1, 4, 3)
    vid = nid(vert[..., 0], vert[..., 1], vert[..., 2])            # [nc][4]
    edge_ids = []
    for (i, j) in _EDGES[3]:
        mid = (vert[:, i, :] + vert[:, j, :]) // 2
        edge_ids.append(nid(mid[:, 0], mid[:, 1], mid[:, 2]))
    dofmap = np.concatenate([vid, np.stack(edge_ids, axis=1)], axis=1).astype(np.int32)
    gi, gj, gk = np.meshgrid(np.arange(gx), np.arange(gy), np.arange(gz), indexing="ij")
    coords = np.stack([gi.ravel() / (2 * nx), gj.ravel() / (2 * ny), gk.ravel() / (2 * nz)], axis=1)
    return coords, vid.astype(np.int64), dofmap
