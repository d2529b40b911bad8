"""CPU tests of the dolfinx-free mesh / function-space layer (fenics_constitutive_b200/solver/mesh.py):
the pieces of dolfinx the stand-in of IncrSmallStrainProblem relies on, checked against their definitions."""
import numpy as np
import pytest

from fenics_constitutive_b200.solver import mesh as M


@pytest.mark.parametrize("make,vol", [(lambda: M.create_unit_interval(7), 1.0), (lambda: M.create_unit_square(4, 3), 1.0),
                                      (lambda: M.create_box((0, 0, 0), (2.0, 1.0, 0.5), 3, 2, 2), 1.0),
                                      (lambda: M.create_rectangle((1.0, -1.0), (3.0, 2.0), 2, 5), 6.0)])
def test_meshes_tile_their_domain_with_positive_cells(make, vol):
    mesh = make()
    g = mesh.gdim
    X = mesh.coords[mesh.cells]
    J = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)
    det = np.linalg.det(J) if g > 1 else J[:, 0, 0]
    assert np.all(det > 0)  # positively oriented simplices
    assert np.isclose(det.sum() / [1, 1, 2, 6][g], vol)
    # every interior facet is shared by exactly two cells (conforming mesh)
    if g > 1:
        facets = {}
        for c in mesh.cells:
            for skip in range(g + 1):
                key = tuple(sorted(np.delete(c, skip)))
                facets[key] = facets.get(key, 0) + 1
        assert set(facets.values()) <= {1, 2}


@pytest.mark.parametrize("make", [lambda: M.create_unit_square(3, 2), lambda: M.create_unit_cube(2, 2, 3)])
def test_p2_space_nodes_are_vertices_and_edge_midpoints(make):
    mesh = make()
    V = M.FunctionSpace(mesh, 2)
    g = mesh.gdim
    nv = mesh.coords.shape[0]
    edges = M._EDGES[g]
    assert V.dofmap.shape == (mesh.num_cells, g + 1 + len(edges))
    assert np.array_equal(V.dofmap[:, : g + 1], mesh.cells)
    for e, (i, j) in enumerate(edges):
        mid = 0.5 * (mesh.coords[mesh.cells[:, i]] + mesh.coords[mesh.cells[:, j]])
        assert np.allclose(V.node_coords[V.dofmap[:, g + 1 + e]], mid)
    # Euler: one extra node per unique edge
    uniq = {tuple(sorted((c[i], c[j]))) for c in mesh.cells for i, j in edges}
    assert V.num_nodes == nv + len(uniq) and V.num_dofs == V.num_nodes * g
    with pytest.raises(NotImplementedError):
        M.FunctionSpace(mesh, 3)


def test_element_tables_partition_of_unity_and_weights():
    for make, g in ((lambda: M.create_unit_interval(3), 1), (lambda: M.create_unit_square(2, 2), 2), (lambda: M.create_unit_cube(1, 2, 1), 3)):
        for degree in (1, 2):
            for qd in (1, 2):
                V = M.FunctionSpace(make(), degree)
                T = M.ElementTables(V, qd)
                assert np.isclose(T.weights.sum(), [1.0, 0.5, 1.0 / 6.0][g - 1])
                assert np.allclose(T.dphi_ref.sum(axis=1), 0.0, atol=1e-14)  # gradients of a partition of unity
                assert np.isclose((T.detJ * T.weights.sum()).sum(), 1.0)     # cell volumes add up to the domain
                assert T.Jinv.shape == (V.mesh.num_cells, g, g)


def test_dirichlet_values_follow_dolfinx_shapes():
    mesh = M.create_unit_cube(1, 1, 1)
    V = M.functionspace(mesh, ("CG", 1, (3,)))
    right = M.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 1.0))
    assert right.size == 4
    c = M.Constant(mesh, 0.01)
    bc = M.dirichletbc(c, M.locate_dofs_geometrical(V.sub(0), lambda x: np.isclose(x[0], 1.0)), V.sub(0))
    assert np.array_equal(bc.dofs, right * 3) and np.all(bc.values() == 0.01)
    c.value = 0.02  # read at solve time (reference tests/models/test_plasticity.py:103-104)
    assert np.all(bc.values() == 0.02)
    vec = M.dirichletbc(np.array([1.0, 2.0, 3.0]), right, V)
    assert np.array_equal(vec.dofs, (right[:, None] * 3 + np.arange(3)).ravel())
    assert np.array_equal(vec.values(), np.tile([1.0, 2.0, 3.0], 4))
    scalar_on_all = M.dirichletbc(M.Constant(0.5), right, V)
    assert np.all(scalar_on_all.values() == 0.5) and scalar_on_all.values().size == 12
    fn = M.dirichletbc(M.Constant(lambda x: np.stack([x[0], 2 * x[1], 0 * x[2]])), right, V)
    xy = V.node_coords[right]
    assert np.allclose(fn.values().reshape(-1, 3), np.stack([xy[:, 0], 2 * xy[:, 1], 0 * xy[:, 2]], axis=1))
    with pytest.raises(NotImplementedError):
        M.functionspace(mesh, ("DG", 1))
    with pytest.raises(NotImplementedError):
        M.functionspace(mesh, ("CG", 1, (2,)))


def test_node_adjacency_lists_every_contribution_in_order():
    mesh = M.create_unit_square(3, 3)
    V = M.FunctionSpace(mesh, 2)
    ptr, idx = M.node_adjacency(V.dofmap, V.num_nodes)
    assert ptr[0] == 0 and ptr[-1] == V.dofmap.size
    flat = V.dofmap.ravel()
    for node in (0, 5, V.num_nodes - 1):
        slots = idx[ptr[node]:ptr[node + 1]]
        assert np.all(flat[slots] == node) and np.all(np.diff(slots) > 0)
    assert np.array_equal(np.sort(idx), np.arange(flat.size))
