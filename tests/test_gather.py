"""Companion gather.  CPU part: tables, mesh and the oracle
restatement reproduce analytic gradients.  GPU part: fcx_gather_grad vs the oracle.
The reference for this row is a dolfinx Expression (absent here), so parity is
anchored on the mathematical definition: nabla_grad of a P2 field is exact for
quadratic displacement fields."""
import numpy as np
import pytest

from fenics_constitutive_b200 import gather as G
from oracle import models as om


def quad_field(x):
    """A quadratic vector field and its nabla_grad: grad[i][j] = d u_j / d x_i."""
    A = np.array([[0.3, -0.2, 0.5], [0.1, 0.4, -0.6], [-0.7, 0.2, 0.9]])
    Q = np.array([[[0.2, 0.1, 0.0], [0.1, -0.3, 0.4], [0.0, 0.4, 0.5]],
                  [[-0.1, 0.2, 0.3], [0.2, 0.6, 0.0], [0.3, 0.0, -0.2]],
                  [[0.4, 0.0, -0.1], [0.0, 0.1, 0.2], [-0.1, 0.2, 0.3]]])  # symmetric per component
    u = x @ A.T + np.einsum("ni,jik,nk->nj", x, Q, x)
    grad = A.T[None, :, :] + 2.0 * np.einsum("jik,nk->nij", Q, x)          # [n][i][j]
    return u, grad


def test_basis_gradients_sum_to_zero():
    for gdim in (1, 2, 3):
        for degree in (1, 2):
            pts, w = G.simplex_quadrature(gdim, 2)
            d = G.lagrange_gradients(gdim, degree, pts)
            assert np.allclose(d.sum(axis=1), 0.0, atol=1e-14)
            assert np.isclose(w.sum(), [1.0, 0.5, 1 / 6][gdim - 1])


def test_mesh_and_oracle_gather_exact_for_quadratics():
    coords, cv, dofmap = G.unit_cube_p2_tets(3, 2, 2)
    assert dofmap.shape == (6 * 12, 10) and coords.shape == (7 * 5 * 5, 3)
    Jinv = G.affine_inverse_jacobians(coords, cv)
    vol = 1.0 / (6.0 * np.abs(np.linalg.det(Jinv)))
    assert np.isclose(vol.sum(), 1.0) and np.all(np.linalg.det(Jinv) > 0)
    # edge dofs really are the midpoints
    for e, (i, j) in enumerate(G._EDGES[3]):
        assert np.allclose(coords[dofmap[:, 4 + e]], 0.5 * (coords[cv[:, i]] + coords[cv[:, j]]))
    pts, _ = G.simplex_quadrature(3, 2)
    dphi = G.lagrange_gradients(3, 2, pts)
    u, _ = quad_field(coords)
    u_prev = 0.25 * u
    grad = om.gather_grad(3, dofmap, u.reshape(-1), u_prev.reshape(-1), dphi, Jinv).reshape(-1, 4, 3, 3)
    # physical QP coordinates
    xq = np.einsum("qa,cad->cqd", np.concatenate([1 - pts.sum(1, keepdims=True), pts], 1), coords[cv])
    _, exact = quad_field(xq.reshape(-1, 3))
    assert np.allclose(grad.reshape(-1, 3, 3), 0.75 * exact, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("with_prev", [True, False])
def test_gpu_gather_vs_oracle_p2_tets(with_prev):
    import torch

    coords, cv, dofmap = G.unit_cube_p2_tets(7, 5, 6)  # 1260 cells: ragged vs the 128-cell CTA
    Jinv = G.affine_inverse_jacobians(coords, cv)
    pts, _ = G.simplex_quadrature(3, 2)
    dphi = G.lagrange_gradients(3, 2, pts)
    rng = np.random.default_rng(0)
    u = rng.standard_normal(coords.size)
    u_prev = rng.standard_normal(coords.size) if with_prev else None
    ref = om.gather_grad(3, dofmap, u, u_prev, dphi, Jinv)
    op = G.IncrementalGradient(3, dofmap, dphi, Jinv)
    out = torch.full((op.num_qps * 9,), float("nan"), dtype=torch.float64, device="cuda")
    op.evaluate(torch.from_numpy(u).cuda(), torch.from_numpy(u_prev).cuda() if with_prev else None, out)
    got = out.cpu().numpy()
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) <= 1e-12 * scale


@pytest.mark.gpu
@pytest.mark.parametrize("gdim,degree,qdeg", [(3, 1, 1), (2, 2, 2), (2, 1, 1), (1, 2, 2), (1, 1, 1), (3, 2, 1), (2, 2, 1)])
def test_gpu_gather_other_elements(gdim, degree, qdeg):
    """Every compiled specialisation plus the generic fallback ((2,6,1) has none)."""
    import torch

    rng = np.random.default_rng(gdim * 10 + degree)
    pts, _ = G.simplex_quadrature(gdim, qdeg)
    dphi = G.lagrange_gradients(gdim, degree, pts)
    nd = dphi.shape[1]
    ncells, nnodes = 1000 + gdim, 700
    dofmap = rng.integers(0, nnodes, size=(ncells, nd)).astype(np.int32)
    Jinv = rng.standard_normal((ncells, gdim, gdim)) + 3 * np.eye(gdim)
    u, u_prev = rng.standard_normal(nnodes * gdim), rng.standard_normal(nnodes * gdim)
    ref = om.gather_grad(gdim, dofmap, u, u_prev, dphi, Jinv)
    op = G.IncrementalGradient(gdim, dofmap, dphi, Jinv)
    out = torch.zeros(op.num_qps * gdim * gdim, dtype=torch.float64, device="cuda")
    op.evaluate(torch.from_numpy(u).cuda(), torch.from_numpy(u_prev).cuda(), out)
    assert np.max(np.abs(out.cpu().numpy() - ref)) <= 1e-12 * np.abs(ref).max()


@pytest.mark.gpu
def test_gpu_gather_feeds_evaluate():
    """form() order of the reference (solver/_lawonsubmesh.py:72-95): gather, then evaluate."""
    import torch

    from fenics_constitutive_b200 import synthetic
    from fenics_constitutive_b200.models import VonMises3D

    coords, cv, dofmap = G.unit_cube_p2_tets(4, 4, 4)
    Jinv = G.affine_inverse_jacobians(coords, cv)
    pts, _ = G.simplex_quadrature(3, 2)
    dphi = G.lagrange_gradients(3, 2, pts)
    u, _ = quad_field(coords)
    u = (u * 6e-3).reshape(-1)
    op = G.IncrementalGradient(3, dofmap, dphi, Jinv)
    n = op.num_qps
    grad = torch.empty(n * 9, dtype=torch.float64, device="cuda")
    op.evaluate(torch.from_numpy(u).cuda(), None, grad)
    z = lambda m: torch.zeros(m, dtype=torch.float64, device="cuda")  # noqa: E731
    st, tg, ep, al = z(n * 6), z(n * 36), z(n * 6), z(n)
    VonMises3D(synthetic.MISES_PARAMS).evaluate(0.0, 1.0, grad, st, tg, {"eps_n": ep, "alpha": al})
    g_ref = om.gather_grad(3, dofmap, u, None, dphi, Jinv)
    ref = [np.zeros(n * 6), np.zeros(n * 36), np.zeros(n * 6), np.zeros(n)]
    orc = om.VonMises3D(synthetic.MISES_PARAMS)
    orc.evaluate(0, 1, g_ref, ref[0], ref[1], {"eps_n": ref[2], "alpha": ref[3]})
    assert 0.02 < orc.plastic_flag.mean() < 0.98
    from _util import TOL_PLASTIC, assert_close
    assert_close(st.cpu().numpy(), ref[0], 6, TOL_PLASTIC, "stress")
    assert_close(tg.cpu().numpy(), ref[1], 36, TOL_PLASTIC, "tangent")


@pytest.mark.gpu
@pytest.mark.parametrize("gdim,degree,qdeg,ncells", [(3, 2, 2, 40_007), (3, 2, 2, 4 * 32), (3, 2, 2, 4 * 64), (3, 2, 2, 200_000),
                                                      (3, 1, 1, 9_001), (3, 1, 2, 5_000),
                                                      (2, 2, 2, 7_777), (2, 1, 2, 3_333), (1, 2, 2, 2_049), (1, 1, 1, 1_000)])
@pytest.mark.parametrize("with_prev", [True, False])
def test_gpu_gather_staged_kernel_bitwise(gdim, degree, qdeg, ncells, with_prev):
    """gather_staged_kernel (nodal values staged by cp.async one tile ahead, fcx_tune
    "gather_variant" 1, the default) and gather_cell_kernel (thread per cell, table in the constant bank,
    variant 2, 3-D 4-point rules; measured slower, kept as an option) run the same fma chains as gather_kernel: bit-identical
    output, ragged tails and many tiles per CTA included; and both match the oracle."""
    import torch

    from fenics_constitutive_b200._lib import lib

    rng = np.random.default_rng(ncells)
    pts, _ = G.simplex_quadrature(gdim, qdeg)
    dphi = G.lagrange_gradients(gdim, degree, pts)
    nd = dphi.shape[1]
    nnodes = max(50, ncells // 3)
    dofmap = rng.integers(0, nnodes, size=(ncells, nd)).astype(np.int32)
    Jinv = rng.standard_normal((ncells, gdim, gdim)) + 3 * np.eye(gdim)
    u = rng.standard_normal(nnodes * gdim)
    u_prev = rng.standard_normal(nnodes * gdim) if with_prev else None
    op = G.IncrementalGradient(gdim, dofmap, dphi, Jinv)
    L = lib()
    outs = []
    old = L.fcx_tune(b"gather_variant", -1)
    try:
        # 2 = gather_cell_kernel (thread per cell), 3 = gather_wq_kernel (warp-uniform point pair): 3-D 4-point
        # rules only, otherwise as 1
        for variant in (0, 1, 2, 3, 3):
            L.fcx_tune(b"gather_variant", variant)
            # u_prev fetched by the kernel (False) or subtracted first as a nodal vector (True, the default)
            for first in ((False, True) if with_prev else (True,)):
                op.form_increment_first = first
                out = torch.full((op.num_qps * gdim * gdim,), float("nan"), dtype=torch.float64, device="cuda")
                op.evaluate(torch.from_numpy(u).cuda(), torch.from_numpy(u_prev).cuda() if with_prev else None, out)
                outs.append(out.cpu().numpy())
    finally:
        L.fcx_tune(b"gather_variant", old)
    assert all(np.array_equal(outs[0], o) for o in outs[1:])
    ref = om.gather_grad(gdim, dofmap, u, u_prev, dphi, Jinv)
    assert np.max(np.abs(outs[-1] - ref)) <= 1e-12 * np.abs(ref).max()


def _permutation_case(seed=5, ncells=4 * 64 + 37):
    rng = np.random.default_rng(seed)
    pts, _ = G.simplex_quadrature(3, 2)
    dphi = G.lagrange_gradients(3, 2, pts)          # [nq][nd][3]
    nq, nd = dphi.shape[0], dphi.shape[1]
    nnodes = ncells * 2
    dofmap = rng.integers(0, nnodes, size=(ncells, nd)).astype(np.int32)
    Jinv = rng.standard_normal((ncells, 3, 3)) + 3 * np.eye(3)
    u, u_prev = rng.standard_normal(nnodes * 3), rng.standard_normal(nnodes * 3)
    pd, pq = rng.permutation(nd), rng.permutation(nq)
    return dphi, dofmap, Jinv, u, u_prev, pd, pq, nq


def test_gather_permutation_covariance_oracle():
    """The element's dof and point ORDERING (basix's, in the reference: _incrementalunknowns.py:21-27) enters the
    gather only through the caller-supplied tables: permuting the local dofs in `dphi_ref` and `dofmap`
    together changes nothing (to round-off of the reordered sum); permuting the quadrature points permutes
    the output rows identically.  CPU: the oracle restatement."""
    dphi, dofmap, Jinv, u, u_prev, pd, pq, nq = _permutation_case()
    base = om.gather_grad(3, dofmap, u, u_prev, dphi, Jinv).reshape(-1, nq, 9)
    by_dof = om.gather_grad(3, dofmap[:, pd], u, u_prev, dphi[:, pd, :], Jinv).reshape(-1, nq, 9)
    assert np.abs(by_dof - base).max() <= 1e-13 * np.abs(base).max()
    by_q = om.gather_grad(3, dofmap, u, u_prev, dphi[pq], Jinv).reshape(-1, nq, 9)
    assert np.array_equal(by_q, base[:, pq, :])


@pytest.mark.gpu
def test_gather_permutation_covariance():
    """Same property on the CUDA path (fcx_nodal_increment + fcx_gather_grad), every kernel variant."""
    import torch

    from fenics_constitutive_b200._lib import lib

    dphi, dofmap, Jinv, u, u_prev, pd, pq, nq = _permutation_case()
    ud, upd = torch.from_numpy(u).cuda(), torch.from_numpy(u_prev).cuda()

    def run(dm, tab):
        op = G.IncrementalGradient(3, np.ascontiguousarray(dm), np.ascontiguousarray(tab), Jinv)
        out = torch.empty(op.num_qps * 9, dtype=torch.float64, device="cuda")
        op.evaluate(ud, upd, out)
        return out.cpu().numpy().reshape(-1, nq, 9)

    L = lib()
    old = L.fcx_tune(b"gather_variant", -1)
    try:
        for variant in (0, 1, 2, 3):
            L.fcx_tune(b"gather_variant", variant)
            base = run(dofmap, dphi)
            assert np.abs(run(dofmap[:, pd], dphi[:, pd, :]) - base).max() <= 1e-13 * np.abs(base).max()
            assert np.array_equal(run(dofmap, dphi[pq]), base[:, pq, :])
    finally:
        L.fcx_tune(b"gather_variant", old)
