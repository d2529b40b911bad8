"""CPU ORACLE (test infrastructure, NOT the product) for the solver-side rows of
SURVEY.md 8f: a plain numpy/scipy restatement of what the reference's
``IncrSmallStrainProblem`` + dolfinx ``NewtonSolver`` compute on an affine
simplex mesh:

  * grad_del_u at the quadrature points     (solver/_incrementalunknowns.py:19-27,40-49)
  * R(v)  = int eps(v) . sigma dx           (solver/_solver.py:87-89, R_form)
  * dR    = int eps(du) . (C eps(v)) dx     (solver/_solver.py:90-96, dR_form)
    with eps = ufl_mandel_strain            (solver/utils.py:10-62)
  * form() data flow and update()           (solver/_solver.py:130-159,
                                             solver/_lawonsubmesh.py:72-95, solver/_history.py:64-88)
  * Newton loop with Dirichlet lifting as in dolfinx 0.9 NonlinearProblem/NewtonSolver
    (third-party; restated from its published algorithm) and a sparse direct
    solve (scipy SuperLU standing in for PETSc LU).

Parity status: the forms above have no golden vectors in the reference beyond
the closed-form answers of its solver tests (homogeneous-deformation tests,
tests/models/test_elasticity.py, test_plasticity.py, test_viscoelasticity.py);
tests/test_solver_*.py check this oracle against those closed forms and the
GPU stand-in against this oracle.  Quadrature point ORDER inside a cell is not
pinned against basix (absent here).

Only tests/ may import this module.  Constitutive laws are the oracle's own
(``oracle.models``), meshes/tables are passed in as numpy arrays.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

R2 = 1 / 2**0.5  # models/utils.py:202-204


def mandel_strain(grad: np.ndarray, gdim: int) -> np.ndarray:
    """[..., g, g] nabla_grad -> [..., s] Mandel strain (solver/utils.py:27-62)."""
    g = grad
    if gdim == 1:
        return g[..., 0, 0][..., None]
    if gdim == 2:
        z = np.zeros_like(g[..., 0, 0])
        return np.stack([g[..., 0, 0], g[..., 1, 1], z, R2 * (g[..., 0, 1] + g[..., 1, 0])], axis=-1)
    return np.stack([g[..., 0, 0], g[..., 1, 1], g[..., 2, 2], R2 * (g[..., 0, 1] + g[..., 1, 0]),
                     R2 * (g[..., 0, 2] + g[..., 2, 0]), R2 * (g[..., 1, 2] + g[..., 2, 1])], axis=-1)


class FemOracle:
    """Tables: dphi_ref [nq][nd][g], weights [nq], Jinv [nc][g][g] (dX_k/dx_i), detJ [nc], dofmap [nc][nd]."""

    def __init__(self, gdim, dofmap, dphi_ref, weights, Jinv, detJ, num_nodes):
        self.g = gdim
        self.s = {1: 1, 2: 4, 3: 6}[gdim]
        self.dofmap = np.asarray(dofmap, dtype=np.int64)
        self.nc, self.nd = self.dofmap.shape
        self.nq = dphi_ref.shape[0]
        self.num_nodes = num_nodes
        self.w = weights
        self.detJ = detJ
        # physical gradients gphi[c][q][a][i] = sum_k Jinv[c][k][i] dphi_ref[q][a][k]
        self.gphi = np.einsum("cki,qak->cqai", Jinv, dphi_ref)
        # B[c][q][s][a*g+j] = mandel_strain(grad = gphi_a (x) e_j)
        g = gdim
        B = np.zeros((self.nc, self.nq, self.s, self.nd * g))
        for a in range(self.nd):
            for j in range(g):
                grad = np.zeros((self.nc, self.nq, g, g))
                grad[..., :, j] = self.gphi[:, :, a, :]
                B[..., a * g + j] = mandel_strain(grad, g)
        self.B = B
        self.cell_dofs = (self.dofmap[:, :, None] * g + np.arange(g)[None, None, :]).reshape(self.nc, -1)

    def grad(self, u, u_prev=None) -> np.ndarray:
        """[nc*nq][g][g] flat: grad[c][q][i][j] = d(u-u_prev)_j/dx_i"""
        du = u if u_prev is None else u - u_prev
        ue = du.reshape(-1, self.g)[self.dofmap]  # [nc][nd][g]
        return np.einsum("cqai,caj->cqij", self.gphi, ue).reshape(-1)

    def internal_force(self, stress) -> np.ndarray:
        sig = stress.reshape(self.nc, self.nq, self.s)
        fe = np.einsum("q,c,cqsd,cqs->cd", self.w, self.detJ, self.B, sig)
        f = np.zeros(self.num_nodes * self.g)
        np.add.at(f, self.cell_dofs.ravel(), fe.ravel())
        return f

    def tangent_matrix(self, tangent) -> sp.csr_matrix:
        """K[v][du] = sum_q w|J| B_v^T C^T B_du  (dR_form: inner(eps(du), C eps(v)))"""
        C = tangent.reshape(self.nc, self.nq, self.s, self.s)
        Ke = np.einsum("q,c,cqmv,cqmk,cqkd->cvd", self.w, self.detJ, self.B, np.swapaxes(C, -1, -2), self.B)
        n = self.num_nodes * self.g
        rows = np.repeat(self.cell_dofs[:, :, None], self.cell_dofs.shape[1], axis=2)
        cols = np.repeat(self.cell_dofs[:, None, :], self.cell_dofs.shape[1], axis=1)
        return sp.csr_matrix((Ke.ravel(), (rows.ravel(), cols.ravel())), shape=(n, n))


class OracleProblem:
    """One law on the whole mesh (or a list of (law, cells)); the data flow of the
    reference's IncrSmallStrainProblem on numpy arrays, with oracle.models laws."""

    def __init__(self, laws, fem: FemOracle, bc_dofs_values, del_t=1.0):
        if not isinstance(laws, list):
            laws = [(laws, np.arange(fem.nc))]
        self.laws = [(law, np.asarray(cells, dtype=np.int64)) for law, cells in laws]
        self.fem = fem
        self.bc_dofs_values = bc_dofs_values  # callable -> (dofs, values)
        self.dt, self.t = del_t, 0.0
        n, nq, s = fem.nc * fem.nq, fem.nq, fem.s
        self.u = np.zeros(fem.num_nodes * fem.g)
        self.u_prev = np.zeros_like(self.u)
        self.stress_0 = np.zeros(n * s)
        self.stress_1 = np.zeros(n * s)
        self.tangent = np.zeros(n * s * s)
        self.f_ext = np.zeros_like(self.u)
        self.history_0, self.history_1 = [], []
        for law, cells in self.laws:
            hd = law.history_dim
            h = None if hd is None else {k: np.zeros(cells.size * nq * v) for k, v in hd.items()}
            self.history_0.append(h)
            self.history_1.append(None if h is None else {k: v.copy() for k, v in h.items()})

    def form(self) -> None:
        fem = self.fem
        g, s, nq = fem.g, fem.s, fem.nq
        grad = fem.grad(self.u, self.u_prev).reshape(fem.nc, nq * g * g)
        for i, (law, cells) in enumerate(self.laws):
            gl = np.ascontiguousarray(grad[cells]).ravel()
            sl = np.ascontiguousarray(self.stress_0.reshape(fem.nc, -1)[cells]).ravel()
            tl = np.zeros(cells.size * nq * s * s)
            h = None
            if self.history_0[i] is not None:
                for k in self.history_0[i]:
                    self.history_1[i][k][:] = self.history_0[i][k]
                h = self.history_1[i]
            law.evaluate(self.t, self.dt, gl, sl, tl, h)
            self.stress_1.reshape(fem.nc, -1)[cells] = sl.reshape(cells.size, -1)
            self.tangent.reshape(fem.nc, -1)[cells] = tl.reshape(cells.size, -1)

    def update(self) -> None:
        self.u_prev[:] = self.u
        self.stress_0[:] = self.stress_1
        for h0, h1 in zip(self.history_0, self.history_1):
            if h0 is not None:
                for k in h0:
                    h0[k][:] = h1[k]
        self.t += self.dt

    def solve(self, rtol=1e-9, atol=1e-10, max_it=50):
        """dolfinx NewtonSolver semantics ("residual" criterion).  Returns (its, converged)."""
        fem = self.fem
        n = self.u.size
        dofs, vals = self.bc_dofs_values()
        free = np.ones(n, dtype=bool)
        free[dofs] = False
        fidx = np.flatnonzero(free)

        def residual():
            self.form()
            b = fem.internal_force(self.stress_1) - self.f_ext
            K = fem.tangent_matrix(self.tangent)
            z = np.zeros(n)
            z[dofs] = vals - self.u[dofs]
            b = b + K @ z
            b[dofs] = self.u[dofs] - vals
            return b, K

        b, K = residual()
        r0 = r = np.linalg.norm(b)
        it = 0
        converged = r < atol
        while not converged and it < max_it:
            dx = np.zeros(n)
            dx[fidx] = spla.spsolve(K[fidx][:, fidx].tocsc(), b[fidx])
            dx[dofs] = b[dofs]
            self.u -= dx
            it += 1
            b, K = residual()
            r = np.linalg.norm(b)
            converged = r < atol or r / r0 < rtol
        return it, converged
