"""Minimal dolfinx-free mesh / function-space layer for the device-resident
stand-in of the reference's solver.

dolfinx, basix and PETSc are not available in the build image, so the pieces of
dolfinx that ``IncrSmallStrainProblem`` needs are restated here for affine
simplex meshes with P1/P2 Lagrange vector spaces:

    mesh = create_unit_cube(2, 2, 2)                 # dolfinx.mesh.create_unit_cube
    V = functionspace(mesh, ("CG", 1, (3,)))         # dolfinx.fem.functionspace
    u = Function(V)                                  # dolfinx.fem.Function
    dofs = locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 0.0))
    bc = dirichletbc(Constant(0.0), dofs, V.sub(0))  # dolfinx.fem.dirichletbc

The names and call shapes follow dolfinx so that the restated reference tests
(tests/test_solver_*.py) read like the originals (reference
tests/models/test_elasticity.py:26-87 etc.).  Nodal vectors live in HBM as
torch CUDA tensors, blocked ``[node][component]`` like dolfinx.
"""
from __future__ import annotations

import itertools

import numpy as np

from ..gather import _EDGES, affine_inverse_jacobians, lagrange_gradients, simplex_quadrature


class Mesh:
    """Affine simplex mesh: ``coords [nv][gdim]``, ``cells [nc][gdim+1]`` (vertex ids)."""

    def __init__(self, coords: np.ndarray, cells: np.ndarray):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int64)
        self.gdim = self.coords.shape[1]
        assert self.cells.shape[1] == self.gdim + 1

    @property
    def num_cells(self) -> int:
        return self.cells.shape[0]


def create_unit_interval(n: int) -> Mesh:
    x = np.linspace(0.0, 1.0, n + 1)[:, None]
    cells = np.stack([np.arange(n), np.arange(1, n + 1)], axis=1)
    return Mesh(x, cells)


def create_rectangle(p0, p1, nx: int, ny: int) -> Mesh:
    """Each grid square is split into two triangles along its diagonal."""
    xs, ys = np.linspace(p0[0], p1[0], nx + 1), np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    vid = lambda i, j: i * (ny + 1) + j  # noqa: E731
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    i, j = i.ravel(), j.ravel()
    t0 = np.stack([vid(i, j), vid(i + 1, j), vid(i + 1, j + 1)], axis=1)
    t1 = np.stack([vid(i, j), vid(i + 1, j + 1), vid(i, j + 1)], axis=1)
    cells = np.stack([t0, t1], axis=1).reshape(-1, 3)
    return Mesh(coords, cells)


def create_unit_square(nx: int, ny: int) -> Mesh:
    return create_rectangle((0.0, 0.0), (1.0, 1.0), nx, ny)


def create_box(p0, p1, nx: int, ny: int, nz: int) -> Mesh:
    """Kuhn triangulation: six positively oriented tetrahedra per grid cube."""
    xs = [np.linspace(p0[d], p1[d], n + 1) for d, n in enumerate((nx, ny, nz))]
    X, Y, Z = np.meshgrid(*xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def vid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    base = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, dtype=np.int64)]
        for ax in perm:
            nxt = v[-1].copy()
            nxt[ax] += 1
            v.append(nxt)
        t = np.stack(v)
        if np.linalg.det((t[1:] - t[0]).astype(float)) < 0:
            t[[2, 3]] = t[[3, 2]]
        tets.append(t)
    tets = np.stack(tets)  # [6][4][3]
    vert = (base[:, None, None, :] + tets[None]).reshape(-1, 4, 3)
    cells = vid(vert[..., 0], vert[..., 1], vert[..., 2])
    return Mesh(coords, cells)


def create_unit_cube(nx: int, ny: int, nz: int) -> Mesh:
    return create_box((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), nx, ny, nz)


class _SubSpace:
    """``V.sub(i)``: one component of a vector space (only used to build BCs)."""

    def __init__(self, V: "FunctionSpace", component: int):
        self.parent = V
        self.component = int(component)


class FunctionSpace:
    """Vector-valued P1/P2 Lagrange space, block size = gdim.

    ``node_coords [nn][gdim]`` and ``dofmap [nc][nd]`` (int32): vertices first,
    then (P2) edge midpoints in basix edge order."""

    def __init__(self, mesh: Mesh, degree: int):
        if degree not in (1, 2):
            raise NotImplementedError("Lagrange degree 1 or 2")
        self.mesh = mesh
        self.degree = degree
        g = mesh.gdim
        self.block_size = g
        nv = mesh.coords.shape[0]
        if degree == 1:
            self.node_coords = mesh.coords.copy()
            self.dofmap = mesh.cells.astype(np.int32)
        else:
            edges = _EDGES[g]
            pairs = np.stack([np.sort(mesh.cells[:, list(e)], axis=1) for e in edges], axis=1)  # [nc][ne][2]
            flat = pairs.reshape(-1, 2)
            uniq, inv = np.unique(flat, axis=0, return_inverse=True)
            mid = 0.5 * (mesh.coords[uniq[:, 0]] + mesh.coords[uniq[:, 1]])
            self.node_coords = np.concatenate([mesh.coords, mid], axis=0)
            edge_nodes = (nv + inv.reshape(-1)).reshape(mesh.num_cells, len(edges))
            self.dofmap = np.concatenate([mesh.cells, edge_nodes], axis=1).astype(np.int32)
        self.num_nodes = self.node_coords.shape[0]
        self.num_dofs = self.num_nodes * g

    @classmethod
    def from_arrays(cls, mesh: Mesh, degree: int, node_coords: np.ndarray, dofmap: np.ndarray) -> "FunctionSpace":
        """Space with a prescribed node numbering (``solver/partitioned.py``: the local space of a
        mesh partition numbers its owned nodes first)."""
        self = cls.__new__(cls)
        self.mesh, self.degree = mesh, int(degree)
        self.block_size = mesh.gdim
        self.node_coords = np.ascontiguousarray(node_coords, dtype=np.float64)
        self.dofmap = np.ascontiguousarray(dofmap, dtype=np.int32)
        assert self.dofmap.shape[0] == mesh.num_cells
        self.num_nodes = self.node_coords.shape[0]
        self.num_dofs = self.num_nodes * mesh.gdim
        return self

    def sub(self, i: int) -> _SubSpace:
        return _SubSpace(self, i)

    def tabulate_dof_coordinates(self) -> np.ndarray:
        return self.node_coords


def functionspace(mesh: Mesh, element) -> FunctionSpace:
    """``functionspace(mesh, ("CG", degree))`` or ``("CG", degree, (gdim,))``."""
    family, degree = element[0], int(element[1])
    if family not in ("CG", "Lagrange", "P"):
        raise NotImplementedError(family)
    if len(element) > 2 and tuple(element[2]) != (mesh.gdim,):
        raise NotImplementedError("vector spaces with block size = gdim only")
    return FunctionSpace(mesh, degree)


class _Vector:
    """Stand-in for ``Function.x``: ``.array`` is the flat torch CUDA tensor."""

    def __init__(self, array):
        self.array = array

    def scatter_forward(self) -> None:  # single partition per rank: nothing to do
        return None


class Function:
    def __init__(self, V: FunctionSpace, device=None):
        import torch

        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.function_space = V
        self.x = _Vector(torch.zeros(V.num_dofs, dtype=torch.float64, device=dev))

    def copy(self) -> "Function":
        f = Function(self.function_space, self.x.array.device)
        f.x.array.copy_(self.x.array)
        return f

    def numpy(self) -> np.ndarray:
        return self.x.array.detach().cpu().numpy()


class Constant:
    """Mutable value holder, like ``dolfinx.fem.Constant``: BCs read ``.value`` at solve time."""

    def __init__(self, value, mesh: Mesh | None = None):
        if isinstance(value, Mesh):  # dolfinx argument order: Constant(mesh, value)
            value, mesh = mesh, value
        self.value = value


class DirichletBC:
    """Flat dof indices and the value source.  ``values()`` is evaluated at
    every solve so that ``Constant.value`` updates take effect (reference
    tests/models/test_plasticity.py:103-104)."""

    def __init__(self, value, nodes: np.ndarray, V):
        nodes = np.asarray(nodes, dtype=np.int64).ravel()
        if isinstance(V, _SubSpace):
            space, comp = V.parent, V.component
            self.dofs = nodes * space.block_size + comp
            self._shape = "scalar"
        else:
            space = V
            g = space.block_size
            self.dofs = (nodes[:, None] * g + np.arange(g)[None, :]).ravel()
            self._shape = "vector"
        self.function_space = space
        self._value = value
        self._nodes = nodes

    def values(self) -> np.ndarray:
        v = self._value.value if isinstance(self._value, Constant) else self._value
        if callable(v):
            v = v(self.function_space.node_coords[self._nodes].T)
        v = np.asarray(v, dtype=np.float64)
        g = self.function_space.block_size
        if self._shape == "scalar":
            return np.broadcast_to(v, self.dofs.shape).astype(np.float64)
        if v.ndim == 0:
            return np.full(self.dofs.shape, float(v))
        if v.shape == (g,):
            return np.tile(v, self._nodes.size)
        return np.ascontiguousarray(v.T if v.shape[0] == g and v.ndim == 2 else v).ravel()


def dirichletbc(value, dofs, V) -> DirichletBC:
    return DirichletBC(value, dofs, V)


def locate_dofs_geometrical(V, marker) -> np.ndarray:
    """Node indices whose coordinates satisfy ``marker(x)``, x of shape (3, n) like dolfinx."""
    space = V.parent if isinstance(V, _SubSpace) else V
    x = np.zeros((3, space.num_nodes))
    x[: space.mesh.gdim] = space.node_coords.T
    return np.flatnonzero(marker(x)).astype(np.int64)


def surface_load(V, marker, traction) -> np.ndarray:
    """Consistent nodal load vector of a constant traction on the boundary facets whose vertices all
    satisfy ``marker`` -- what the reference's tests add to the residual as
    ``problem.R_form -= inner(t, v) * ds(tag)`` (tests/models/test_viscoelasticity.py:452-470).
    Assign it to ``problem.f_ext``.  P1: a facet's load is split equally among its vertices; P2: a
    triangular facet loads its three edge midpoints with a third each (vertices get nothing), an edge
    facet its ends with 1/6 and its midpoint with 2/3."""
    space = V.parent if isinstance(V, _SubSpace) else V
    mesh, g = space.mesh, space.mesh.gdim
    if g == 1:
        raise NotImplementedError("point loads in 1D: set f_ext directly")
    t = np.asarray(traction, dtype=np.float64).ravel()
    assert t.size == g
    x = np.zeros((3, mesh.coords.shape[0]))
    x[:g] = mesh.coords.T
    on = np.asarray(marker(x), dtype=bool)
    f = np.zeros((space.num_nodes, g))
    nv = g + 1
    edges = _EDGES[g]
    for skip in range(nv):  # facet opposite to local vertex `skip`
        loc = [i for i in range(nv) if i != skip]
        hit = np.flatnonzero(on[mesh.cells[:, loc]].all(axis=1))
        if hit.size == 0:
            continue
        X = mesh.coords[mesh.cells[hit][:, loc]]  # [nf][g][g]
        if g == 2:
            area = np.linalg.norm(X[:, 1] - X[:, 0], axis=1)
        else:
            area = 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)
        if space.degree == 1:
            for i in loc:
                np.add.at(f, space.dofmap[hit, i], (area / g)[:, None] * t[None, :])
        else:
            mids = [nv + e for e, (a, b) in enumerate(edges) if a in loc and b in loc]
            if g == 2:
                for i in loc:
                    np.add.at(f, space.dofmap[hit, i], (area / 6.0)[:, None] * t[None, :])
                np.add.at(f, space.dofmap[hit, mids[0]], (area * 2.0 / 3.0)[:, None] * t[None, :])
            else:
                for m in mids:
                    np.add.at(f, space.dofmap[hit, m], (area / 3.0)[:, None] * t[None, :])
    return f.ravel()


class ElementTables:
    """Reference tables of the (degree, q_degree) pair on the mesh's simplex and
    the per-cell affine geometry, as numpy arrays (moved to HBM by the problem).
    Quadrature = the default simplex rules of matching degree (1 or 2 points per
    direction equivalent); basix is absent, so the ORDER of the points inside a
    cell is this module's, not basix's."""

    def __init__(self, V: FunctionSpace, q_degree: int):
        g = V.mesh.gdim
        pts, w = simplex_quadrature(g, q_degree)
        self.points, self.weights = pts, w
        self.dphi_ref = lagrange_gradients(g, V.degree, pts)  # [nq][nd][g]
        self.Jinv = affine_inverse_jacobians(V.mesh.coords, V.mesh.cells)  # [nc][g][g]
        self.detJ = 1.0 / np.abs(np.linalg.det(self.Jinv))
        self.nq, self.nd = self.dphi_ref.shape[0], self.dphi_ref.shape[1]


def node_adjacency(dofmap: np.ndarray, num_nodes: int):
    """CSR node -> (cell*nd + local index) in increasing order (deterministic sum)."""
    nc, nd = dofmap.shape
    flat = dofmap.ravel().astype(np.int64)
    order = np.argsort(flat, kind="stable")
    counts = np.bincount(flat, minlength=num_nodes)
    ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    return ptr, order.astype(np.int32)
