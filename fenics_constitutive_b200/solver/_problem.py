"""Device-resident ``IncrSmallStrainProblem`` -- the stand-in for the reference's
dolfinx/PETSc problem class (src/fenics_constitutive/solver/_solver.py:30-218)
on an affine simplex mesh, with every quadrature array, the residual and the
Jacobian action living in HBM.

Same data flow and the same public surface as the reference:

    problem = IncrSmallStrainProblem(law | [(law, cells), ...], u, bcs, q_degree, del_t)
    problem.form(x)      # gather grad_del_u, reset trial history, law.evaluate, map back
    problem.F(x, b)      # b = R(u) = int eps(v).sigma dx  (- external forces)
    problem.J_apply(p)   # y = dR(u) p                      (matrix-free)
    problem.update()     # commit u, sigma, history; advance time
    problem.stress_0 / stress_1 / tangent / _history_0 / _history_1 / _del_grad_u / _u / _u0 / _time / _del_t

NOT dolfinx: forms are fixed (no UFL), the mesh layer is ``solver/mesh.py``,
Neumann terms enter through the nodal vector ``problem.f_ext``.  Labelled
"stand-in driver" wherever numbers from it are reported.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .. import _buffers as B
from .._lib import check, lib
from ..gather import IncrementalGradient
from ..models.interfaces import IncrSmallStrainModel
from ..models.mises_plasticity_isotropic_hardening import VonMises3D
from .maps import IdentityMap, build_subspace_map
from .mesh import DirichletBC, ElementTables, Function, _Vector, node_adjacency


@dataclass(slots=True)
class SimulationTime:
    """reference solver/_solver.py:21-27"""

    dt: float
    current: float = 0

    def advance(self) -> None:
        self.current += self.dt


class QuadratureFunction:
    """A flat quadrature array with the ``.x.array`` access of a dolfinx Function."""

    def __init__(self, n: int, device):
        import torch

        self.x = _Vector(torch.zeros(n, dtype=torch.float64, device=device))

    def numpy(self) -> np.ndarray:
        return self.x.array.detach().cpu().numpy()


class IncrementalStress:
    """reference solver/_incrementalunknowns.py:52-79"""

    def __init__(self, n: int, device):
        self._current = QuadratureFunction(n, device)
        self._previous = QuadratureFunction(n, device)

    @property
    def current(self) -> QuadratureFunction:
        return self._current

    @property
    def previous(self) -> QuadratureFunction:
        return self._previous

    def current_array(self):
        return self._current.x.array

    def update_previous(self) -> None:
        self._previous.x.array.copy_(self._current.x.array)

    def update_current(self) -> None:
        self._current.x.array.copy_(self._previous.x.array)

    def scatter_current(self) -> None:
        return None


class IncrementalDisplacement:
    """reference solver/_incrementalunknowns.py:14-49: current / previous displacement and
    nabla_grad(current - previous) at the quadrature points of a cell list."""

    def __init__(self, u: Function, q_degree: int):
        self.u = u
        self.q_degree = q_degree
        self.current = u
        self.previous = u.copy()

    def update_previous(self) -> None:
        self.previous.x.array.copy_(self.current.x.array)

    def update_current(self, x) -> None:
        if x is not None and x.data_ptr() != self.current.x.array.data_ptr():
            self.current.x.array.copy_(x)

    def evaluate_local_incremental_gradient(self, gather_op: IncrementalGradient, displacement_gradient_fn) -> None:
        gather_op.evaluate(self.current.x.array, self.previous.x.array, displacement_gradient_fn.x.array)


class History:
    """reference solver/_history.py:37-88: committed (history_0) and trial (history_1) values."""

    def __init__(self, history_dim: dict, nqp: int, device):
        self._history_dim = history_dim
        size = lambda v: int(np.prod(v)) if not isinstance(v, int) else v  # noqa: E731
        self.history_0 = {k: QuadratureFunction(nqp * size(v), device) for k, v in history_dim.items()}
        self.history_1 = {k: QuadratureFunction(nqp * size(v), device) for k, v in history_dim.items()}

    @staticmethod
    def try_create(law: IncrSmallStrainModel, nqp: int, device):
        if law.history_dim is None:
            return None
        return History(law.history_dim, nqp, device)

    def reset_trial_state(self) -> dict:
        for key in self._history_dim:
            self.history_1[key].x.array.copy_(self.history_0[key].x.array)
        return {key: self.history_1[key].x.array for key in self._history_dim}

    def update(self) -> dict:
        for key in self._history_dim:
            self.history_0[key].x.array.copy_(self.history_1[key].x.array)
        return {key: self.history_0[key].x.array for key in self._history_dim}


class LawOnSubMesh:
    """reference solver/_lawonsubmesh.py:48-100"""

    def __init__(self, law, cells: np.ndarray, problem: "IncrSmallStrainProblem", fused: bool = False):
        s, g = law.stress_strain_dim, law.geometric_dim
        T = problem.tables
        dev = problem.device
        self.law = law
        self.cells = np.asarray(cells, dtype=np.int64)
        nc = self.cells.size
        nqp = nc * T.nq
        self.submesh_map = build_subspace_map(self.cells, problem.num_cells, dev)
        self.identity = isinstance(self.submesh_map, IdentityMap)
        dofmap = problem.V.dofmap if self.identity else problem.V.dofmap[self.cells]
        Jinv = T.Jinv if self.identity else T.Jinv[self.cells]
        self.gather_op = IncrementalGradient(g, dofmap, T.dphi_ref, Jinv, device=dev)
        self.displacement_gradient_fn = QuadratureFunction(nqp * g * g, dev)
        if self.identity:  # alias the global arrays instead of copying through an IdentityMap
            self.stress = None
            self.local_tangent = problem.tangent
        elif fused:  # the fused form() kernel reads / writes the parent rows directly (no local copies)
            self.stress = None
            self.local_tangent = None
        else:
            self.stress = QuadratureFunction(nqp * s, dev)
            self.local_tangent = QuadratureFunction(nqp * s * s, dev)
        self.history = History.try_create(law, nqp, dev)

    def evaluate(self, sim_time, incr_disp, global_stress, global_tangent) -> None:
        incr_disp.evaluate_local_incremental_gradient(self.gather_op, self.displacement_gradient_fn)
        history_input = self.history.reset_trial_state() if self.history is not None else None
        if self.identity:
            # stress.current <- stress.previous, then the law updates it in place
            global_stress.update_current()
            local_stress = global_stress.current.x.array
        else:
            self.submesh_map.map_to_sub(global_stress.previous.x.array, self.stress.x.array)
            local_stress = self.stress.x.array
        self.law.evaluate(
            sim_time.current, sim_time.dt, self.displacement_gradient_fn.x.array, local_stress,
            self.local_tangent.x.array, history_input,
        )
        if not self.identity:
            self.submesh_map.map_to_parent(self.stress.x.array, global_stress.current.x.array)
            self.submesh_map.map_to_parent(self.local_tangent.x.array, global_tangent.x.array)

    def update_history(self) -> None:
        if self.history is not None:
            self.history.update()


class IncrSmallStrainProblem:
    """Args mirror the reference (solver/_solver.py:54-63): ``laws`` is one law
    (homogeneous domain) or a list of ``(law, local cell indices)``; ``u`` the
    displacement Function (the unknown); ``bcs`` Dirichlet conditions;
    ``q_degree`` 1 or 2; ``del_t`` the time increment."""

    def __init__(self, laws, u: Function, bcs: list[DirichletBC], q_degree: int, del_t: float = 1.0,
                 form_compiler_options=None, jit_options=None, fused: bool | None = None) -> None:
        import torch

        self.V = u.function_space
        mesh = self.V.mesh
        self.device = u.x.array.device
        self.num_cells = mesh.num_cells
        if isinstance(laws, IncrSmallStrainModel):
            laws = [(laws, np.arange(0, self.num_cells, dtype=np.int32))]
        constraint = laws[0][0].constraint
        assert all(law[0].constraint == constraint for law in laws), "All laws must have the same constraint"
        assert constraint.geometric_dim == mesh.gdim, "constraint and mesh disagree on the geometric dimension"
        self.constraint = constraint
        self.gdim, self.sdim = constraint.geometric_dim, constraint.stress_strain_dim
        self.q_degree = q_degree
        self.tables = ElementTables(self.V, q_degree)
        T = self.tables
        self.nqp = self.num_cells * T.nq
        dev = self.device
        t64 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)  # noqa: E731
        self._dphi = t64(T.dphi_ref)
        self._weights = t64(T.weights)
        self._Jinv = t64(T.Jinv)
        self._detJ = t64(T.detJ)
        self._dofmap = torch.as_tensor(self.V.dofmap, dtype=torch.int32, device=dev).contiguous()
        ptr, idx = node_adjacency(self.V.dofmap, self.V.num_nodes)
        self._adj_ptr = torch.as_tensor(ptr, dtype=torch.int64, device=dev)
        self._adj_idx = torch.as_tensor(idx, dtype=torch.int32, device=dev)
        # 3-D elements with 4 QPs: element vectors are written in node-major slot order (fe_pos = inverse
        # of adj_idx) so the node-wise sum streams them contiguously (include/fcx.h)
        self._fe_pos = None
        if self.gdim == 3 and T.nq == 4:
            pos = np.empty(idx.size, dtype=np.int32)
            pos[idx] = np.arange(idx.size, dtype=np.int32)
            self._fe_pos = torch.as_tensor(pos, dtype=torch.int32, device=dev)
        # element vectors [ncells][nd][fs]; 3-D slots are padded to one 32-byte sector (include/fcx.h)
        self._fe = torch.empty(self.num_cells * T.nd * lib().fcx_fe_stride(self.gdim), dtype=torch.float64, device=dev)

        self.stress = IncrementalStress(self.nqp * self.sdim, dev)
        self.tangent = QuadratureFunction(self.nqp * self.sdim**2, dev)
        self.sim_time = SimulationTime(dt=del_t)
        self.incr_disp = IncrementalDisplacement(u, q_degree)
        # fused form() kernel: every law a VonMises3D (each on its own cell list: the sub-mesh maps
        # are folded into the kernel's bulk copies), supported element
        # (``fused=False`` forces the generic path: gather, map kernels, law.evaluate, map kernels)
        self.fused = (
            fused is not False
            and all(type(law) is VonMises3D and law.eps_layout == "aos" for law, _ in laws)
            and (T.nd, T.nq) in ((10, 4), (4, 1), (4, 4))
        )
        if fused and not self.fused:
            raise ValueError("fused=True needs VonMises3D laws (AoS history) on P1/P2 tetrahedra")
        self._law_on_submeshs = [LawOnSubMesh(law, cells, self, self.fused) for law, cells in laws]
        self._bcs = list(bcs)
        self.f_ext = torch.zeros(self.V.num_dofs, dtype=torch.float64, device=dev)  # Neumann loads (nodal)
        self.keep_del_grad_u = True  # fused path: also store grad_del_u (72 B/QP) for inspection
        # fused path: form() also emits the tangent as 10-double records (4 coefficients + flow
        # direction) and J_apply reads those -- 80 instead of 288 B per QP per Krylov iteration
        self.use_tangent_records = True
        # fused path: also write the dense [nqp][36] tangent (the reference's `tangent` Function).
        # J_diag reads it; switch off only with use_tangent_records (the Jacobian then lives in the
        # records alone and form() moves 288 B/QP less)
        self.dense_tangent = True
        # (zeros: cells that no law owns keep a zero tangent, like the dense array)
        self._trec = torch.zeros(self.nqp * 10, dtype=torch.float64, device=dev) if self.fused else None
        self._trec_valid = False

    @property
    def symmetric_tangent(self) -> bool:
        """False if a law's consistent tangent is not symmetric (``law.symmetric_tangent`` False: the
        Drucker-Prager models with non-associated flow) -- NewtonSolver then avoids CG."""
        return all(getattr(ctx.law, "symmetric_tangent", True) for ctx in self._law_on_submeshs)

    # ------------------------------------------------------------------ form
    def form(self, x=None) -> None:
        """reference solver/_solver.py:130-147: update the current displacement, then
        evaluate every law (stress, tangent, trial history)."""
        self.incr_disp.update_current(x)
        if self.fused:
            self._form_fused()
            return
        self._trec_valid = False
        for law in self._law_on_submeshs:
            law.evaluate(self.sim_time, self.incr_disp, self.stress, self.tangent)

    def _form_fused(self) -> None:
        T = self.tables
        L = lib()
        dev = self.device.index
        check(L.fcx_set_device(dev), "fcx_set_device")
        stream = B.current_stream_ptr(dev)
        for ctx in self._law_on_submeshs:
            law: VonMises3D = ctx.law
            ncl = int(ctx.cells.size)
            if ncl == 0:
                continue
            h0, h1 = ctx.history.history_0, ctx.history.history_1
            P = law._params()
            status = law._status_tensor(dev)
            flag = None
            if law.record_plastic_flag:
                import torch

                flag = torch.zeros(ncl * T.nq, dtype=torch.uint8, device=self.device)
            op = ctx.gather_op  # dofmap / Jinv rows of the law's cells (sliced once at set-up)
            rc = L.fcx_mises_form(
                P.ctypes.data, ncl, None if ctx.identity else ctx.submesh_map.cells_ptr(), T.nq, T.nd,
                op.dofmap.data_ptr(),
                self.incr_disp.current.x.array.data_ptr(), self.incr_disp.previous.x.array.data_ptr(),
                self._dphi.data_ptr(), op.Jinv.data_ptr(),
                self.stress.previous.x.array.data_ptr(), self.stress.current.x.array.data_ptr(),
                self.tangent.x.array.data_ptr() if self.dense_tangent else None,
                h0["eps_n"].x.array.data_ptr(), h1["eps_n"].x.array.data_ptr(),
                h0["alpha"].x.array.data_ptr(), h1["alpha"].x.array.data_ptr(),
                ctx.displacement_gradient_fn.x.array.data_ptr() if self.keep_del_grad_u else None,
                self._trec.data_ptr() if self.use_tangent_records else None,
                flag.data_ptr() if flag is not None else None, status.data_ptr(), stream,
            )
            law.plastic_flag = flag
            check(rc, "IncrSmallStrainProblem.form (fcx_mises_form)")
        self._trec_valid = self.use_tangent_records
        for ctx in self._law_on_submeshs:
            if not ctx.law.defer_errors:
                ctx.law.check_converged()

    # -------------------------------------------------------- residual / Jacobian
    def _tables_args(self):
        T = self.tables
        return (self.gdim, self.sdim, self.num_cells, T.nq, T.nd)

    def _pos_ptr(self):
        """Device pointer of fe_pos when the element kernels write node-major slots, else None."""
        if self._fe_pos is not None and lib().fcx_tune(b"fem_variant", -1) != 0:
            return self._fe_pos.data_ptr()
        return None

    def _gather_sum(self, out, alpha=1.0, beta=0.0) -> None:
        L = lib()
        idx = None if self._pos_ptr() is not None else self._adj_idx.data_ptr()
        check(L.fcx_gather_sum(self.gdim, self.V.num_nodes, self._adj_ptr.data_ptr(), idx,
                               self._fe.data_ptr(), out.data_ptr(), alpha, beta,
                               B.current_stream_ptr(self.device.index)), "fcx_gather_sum")

    def F(self, x=None, b=None):
        """b = R(u): internal force of stress.current minus ``f_ext``
        (reference R_form, solver/_solver.py:87-89).  No boundary-condition handling here."""
        import torch

        if b is None:
            b = torch.empty(self.V.num_dofs, dtype=torch.float64, device=self.device)
        L = lib()
        g, s, nc, nq, nd = self._tables_args()
        check(L.fcx_set_device(self.device.index))
        check(L.fcx_internal_force(g, s, nc, nq, nd, self._dphi.data_ptr(), self._weights.data_ptr(),
                                   self._Jinv.data_ptr(), self._detJ.data_ptr(),
                                   self.stress.current.x.array.data_ptr(), self._fe.data_ptr(), self._pos_ptr(),
                                   B.current_stream_ptr(self.device.index)), "fcx_internal_force")
        self._gather_sum(b)
        b.sub_(self.f_ext)
        return b

    def J_apply(self, p, out=None):
        """out = dR(u) p with the current tangent (reference dR_form, solver/_solver.py:90-96)."""
        import torch

        if out is None:
            out = torch.empty_like(p)
        L = lib()
        g, s, nc, nq, nd = self._tables_args()
        if self.fused and self._trec_valid and self.use_tangent_records and L.fcx_tune(b"fem_variant", -1) != 0:
            check(L.fcx_tangent_apply_rec(g, s, nc, nq, nd, self._dofmap.data_ptr(), p.data_ptr(),
                                          self._dphi.data_ptr(), self._weights.data_ptr(), self._Jinv.data_ptr(),
                                          self._detJ.data_ptr(), self._trec.data_ptr(), self._fe.data_ptr(),
                                          self._pos_ptr(), B.current_stream_ptr(self.device.index)),
                  "fcx_tangent_apply_rec")
            self._gather_sum(out)
            return out
        check(L.fcx_tangent_apply(g, s, nc, nq, nd, self._dofmap.data_ptr(), p.data_ptr(),
                                  self._dphi.data_ptr(), self._weights.data_ptr(), self._Jinv.data_ptr(),
                                  self._detJ.data_ptr(), self.tangent.x.array.data_ptr(), self._fe.data_ptr(),
                                  self._pos_ptr(), B.current_stream_ptr(self.device.index)), "fcx_tangent_apply")
        self._gather_sum(out)
        return out

    def J_diag(self, out=None):
        import torch

        if out is None:
            out = torch.empty(self.V.num_dofs, dtype=torch.float64, device=self.device)
        if self.fused and not self.dense_tangent:
            raise RuntimeError("J_diag reads the dense tangent array: keep IncrSmallStrainProblem.dense_tangent = True "
                               "(the default) when a solver needs the Jacobi diagonal")
        L = lib()
        g, s, nc, nq, nd = self._tables_args()
        check(L.fcx_tangent_diag(g, s, nc, nq, nd, self._dphi.data_ptr(), self._weights.data_ptr(),
                                 self._Jinv.data_ptr(), self._detJ.data_ptr(), self.tangent.x.array.data_ptr(),
                                 self._fe.data_ptr(), self._pos_ptr(), B.current_stream_ptr(self.device.index)),
              "fcx_tangent_diag")
        self._gather_sum(out)
        return out

    # --------------------------------------------------------------- commit
    def update(self) -> None:
        """reference solver/_solver.py:149-159"""
        self.incr_disp.update_previous()
        self.stress.update_previous()
        for law in self._law_on_submeshs:
            law.update_history()
        self.sim_time.advance()

    # ------------------------------------------------ boundary conditions
    def bc_dofs_values(self):
        """(flat dof indices, prescribed values) of all Dirichlet BCs, evaluated now."""
        if not self._bcs:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        dofs = np.concatenate([bc.dofs for bc in self._bcs])
        vals = np.concatenate([bc.values() for bc in self._bcs])
        return dofs, vals

    # ---------------------------- backward-compatibility properties (reference :165-218)
    @property
    def _time(self) -> float:
        return self.sim_time.current

    @_time.setter
    def _time(self, value: float) -> None:
        self.sim_time.current = value

    @property
    def _del_t(self) -> float:
        return self.sim_time.dt

    @_del_t.setter
    def _del_t(self, value: float) -> None:
        self.sim_time.dt = value

    @property
    def _u(self) -> Function:
        return self.incr_disp.current

    @property
    def _u0(self) -> Function:
        return self.incr_disp.previous

    @property
    def stress_0(self) -> QuadratureFunction:
        return self.stress.previous

    @property
    def stress_1(self) -> QuadratureFunction:
        return self.stress.current

    @property
    def _history_0(self):
        return [law.history.history_0 if law.history else None for law in self._law_on_submeshs]

    @property
    def _history_1(self):
        return [law.history.history_1 if law.history else None for law in self._law_on_submeshs]

    @property
    def _del_grad_u(self):
        return [law.displacement_gradient_fn for law in self._law_on_submeshs]
