"""Newton solver for the device-resident ``IncrSmallStrainProblem`` -- stand-in
for ``dolfinx.nls.petsc.NewtonSolver`` (third-party in the reference, used by
every reference test, e.g. tests/models/test_plasticity.py:87-104).

Same loop and defaults as dolfinx 0.9's NewtonSolver:
    form(x); b = F(x)                      (with Dirichlet lifting, see below)
    while not converged and it < max_it:
        solve J dx = b;  x -= relaxation * dx;  form(x);  b = F(x);  it += 1
    converged:  "residual":    |b| / |b_0| < rtol  or  |b| < atol
                "incremental": |dx| / |dx_0| < rtol or |dx| < atol   (checked after the update)
rtol = 1e-9, atol = 1e-10, max_it = 50, relaxation_parameter = 1.

Dirichlet conditions follow dolfinx's NonlinearProblem.F/J: with z = g - x on
the constrained dofs (0 elsewhere), b <- b + J z on the free dofs and b = x - g
on the constrained ones; J gets identity rows/columns there.

Linear solver: the Jacobian is applied matrix-free from the stored tangents
(``problem.J_apply``).  "cg": Jacobi-preconditioned conjugate gradients on the
free dofs, for SYMMETRIC tangents (every built-in law except the Drucker-Prager
models with non-associated flow, b_flow != b); "bicgstab": Jacobi-preconditioned
BiCGStab for non-symmetric tangents; "dense": the operator is materialised column
by column and solved with LU -- the analogue of the reference tests' default
PETSc LU -- for small single-rank problems.  "auto" picks dense up to 3000 dofs
on one rank, otherwise cg or (``problem.symmetric_tangent`` False) bicgstab; on a
partitioned mesh or with several ranks it never picks dense (a rank-local LU would
ignore the coupling across ranks), and asking for it explicitly raises.  With
several ranks (torchrun), every dot product and norm is summed over ranks (NCCL
all-reduce of one double).  Every Krylov solve records whether it converged
(``krylov_converged``, ``krylov_relres``); hitting ``cg_max_it`` or a CG breakdown
(p.Ap <= 0: the operator is not positive definite on the free dofs) raises
``KrylovError`` unless ``error_on_krylov_failure`` is False (then it warns).
"""
from __future__ import annotations

import numpy as np

import warnings

from ..partition import sum_over_ranks


class KrylovError(RuntimeError):
    """The linear solve of a Newton step did not converge (or CG broke down)."""


class NewtonSolver:
    def __init__(self, comm_or_problem, problem=None):
        # dolfinx signature NewtonSolver(comm, problem); the comm is ignored here
        self.problem = problem if problem is not None else comm_or_problem
        self.rtol = 1e-9
        self.atol = 1e-10
        self.max_it = 50
        self.relaxation_parameter = 1.0
        self.convergence_criterion = "residual"
        self.error_on_nonconvergence = True
        self.report = False
        self.linear_solver = "auto"  # "cg" | "bicgstab" | "dense" | "auto"
        self.error_on_krylov_failure = True
        self.krylov_converged: list[bool] = []
        self.krylov_relres: list[float] = []
        self.cg_rtol = 1e-12
        # Inexact Newton: None = every linear solve to cg_rtol; "eisenstat-walker" = forcing term
        # eta_k = gamma (|r_k| / |r_{k-1}|)^2 (choice 2 of Eisenstat & Walker 1996 with their
        # safeguard), clipped to [cg_rtol, cg_eta_max]: early Newton steps are solved loosely, the
        # last ones tightly, the Newton tolerances (rtol / atol above) are untouched.
        self.cg_forcing = None
        self.cg_eta_max = 1e-2
        self.cg_eta_gamma = 0.9
        self.cg_max_it = 20000
        self.cg_check_every = 10  # host convergence checks (one sync each)
        # "device": the whole Krylov loop runs from C (csrc/fcx_krylov.cu: single-reduction PCG, reduction and
        # ghost exchange over peer memory inside the kernels; no NCCL, no Python per iteration).
        # "python": two-reduction PCG issued kernel by kernel from here, NCCL all-reduces and send/recv.
        self.cg_driver = "device"
        # device driver: True = residual test on the device, blocks of iterations enqueued ahead of its outcome,
        # exact stopping iteration; False = stream drained and tested on the host after every block; None = the
        # driver's default (on; FCX_KRYLOV_LOOKAHEAD=0 switches it off)
        self.cg_lookahead = None
        self._device_krylov = None
        self.reduce_over_ranks = False  # sum norms/dots over torch.distributed ranks
        # solver/partitioned.py MeshPartition (set by MeshPartition.attach): norms and dot products run
        # over OWNED dofs, ghost values of p (before every Jacobian action) and of x (after every Newton
        # update) are refreshed from their owners.  None = every rank holds a whole problem.
        self.partition = None
        self.residual_history: list[float] = []
        self.krylov_iterations: list[int] = []
        self._owned_mask = None
        self._cg_ws = None
        self.profile = False  # accumulate wall time of the linear solves (adds two syncs per solve)
        self.linear_solve_s = 0.0
        self.residual_s = 0.0   # profile: form + F + lifting + norm
        self.update_s = 0.0     # profile: x update + ghost refresh of x
        self.krylov_setup_s = 0.0  # one-off set-up of the device Krylov loop (IPC handles, buffers)

    # ------------------------------------------------------------ helpers
    def _owned(self, like):
        if self.partition is None:
            return None
        if self._owned_mask is None or self._owned_mask.device != like.device:
            self._owned_mask = self.partition.owned_dof_mask(like.device)
        return self._owned_mask

    def _dot(self, a, b) -> float:
        own = self._owned(a)
        v = float((a * b).sum().item()) if own is None else float((a * b)[own].sum().item())
        if self.reduce_over_ranks:
            v = sum_over_ranks(v, a.device)
        return v

    def _norm(self, a) -> float:
        return float(np.sqrt(self._dot(a, a)))

    def _solve_dense(self, apply, rhs, free_mask):
        import torch

        n = rhs.numel()
        idx = torch.nonzero(free_mask, as_tuple=False).ravel()
        m = idx.numel()
        A = torch.empty((m, m), dtype=torch.float64, device=rhs.device)
        e = torch.zeros(n, dtype=torch.float64, device=rhs.device)
        y = torch.empty_like(e)
        for k in range(m):
            e[idx[k]] = 1.0
            apply(e, y)
            A[:, k] = y[idx]
            e[idx[k]] = 0.0
        sol = torch.linalg.solve(A, rhs[idx])
        dx = torch.zeros_like(rhs)
        dx[idx] = sol
        return dx, 0

    def _rsum(self, t):
        """Sum of a 0-dim device tensor over ranks (no host synchronisation)."""
        if self.reduce_over_ranks:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def _cg_workspace(self, like):
        """Persistent PCG vectors and scalars of one size / device (no allocation per linear solve)."""
        import torch

        from .._lib import lib

        ws = self._cg_ws
        if ws is None or ws["n"] != like.numel() or ws["x"].device != like.device:
            dev = like.device
            z = lambda: torch.zeros(like.numel(), dtype=torch.float64, device=dev)  # noqa: E731
            ws = {"n": like.numel(), "x": z(), "r": z(), "p": z(), "Ap": z(), "minv": z(),
                  "scratch": torch.empty(int(lib().fcx_pcg_scratch_doubles()), dtype=torch.float64, device=dev),
                  "ticket": torch.zeros(1, dtype=torch.int32, device=dev),
                  "sc": torch.zeros(4, dtype=torch.float64, device=dev)}  # [rz, pAp, rz_new, rr]
            self._cg_ws = ws
        return ws

    def _solve_cg(self, apply, rhs, free_mask, diag, rtol=None):
        """Jacobi-preconditioned CG on the free dofs (projected operator P J P).  All scalars stay
        on the device; the host looks at the residual norm only every ``cg_check_every``
        iterations, so an iteration is a fixed sequence of enqueued kernels: the two element
        kernels of problem.J_apply + the three fused vector kernels of csrc/fcx_pcg.cu
        (deterministic reductions).  With several ranks the three scalars are summed over ranks
        (NCCL all-reduce of one / two doubles) and, on a partitioned mesh, the ghost values of the
        direction are refreshed.  The returned vector is the solver's workspace: valid until the
        next linear solve."""
        import torch

        from .. import _buffers as B
        from .._lib import check, lib

        L = lib()
        dev = rhs.device
        n = rhs.numel()
        own = self._owned(rhs)
        halo = self.partition.halo_update if self.partition is not None else (lambda v: None)
        ws = self._cg_workspace(rhs)
        x, r, p, Ap, minv, scratch, ticket, sc = (ws[k] for k in ("x", "r", "p", "Ap", "minv", "scratch", "ticket", "sc"))
        # ghost dofs are masked like constrained ones: every reduction runs over owned free dofs
        fm = (free_mask if own is None else free_mask & own).to(torch.float64)
        minv.copy_(fm / torch.where(diag.abs() > 0, diag, torch.ones_like(diag)))
        x.zero_()
        torch.mul(rhs, fm, out=r)
        torch.mul(minv, r, out=p)
        halo(p)
        rz, pAp, new2 = sc[0:1], sc[1:2], sc[2:4]
        rz.copy_(self._rsum(torch.dot(r, p)).reshape(1))
        r0 = float(torch.sqrt(self._rsum(torch.dot(r, r))).item())
        if r0 == 0.0:
            self._krylov_report(True, 0.0, 0, "cg")
            return x, 0
        tol2 = ((self.cg_rtol if rtol is None else rtol) * r0) ** 2
        K = max(1, int(self.cg_check_every))

        def iteration():
            """One PCG iteration, enqueue-only on the current stream (x, r, p, rz updated in place)."""
            apply(p, Ap)
            check(L.fcx_pcg_pap(n, p.data_ptr(), Ap.data_ptr(), minv.data_ptr(), scratch.data_ptr(),
                                ticket.data_ptr(), pAp.data_ptr(), B.current_stream_ptr(dev.index)), "fcx_pcg_pap")
            self._rsum(pAp)
            check(L.fcx_pcg_update_xr(n, x.data_ptr(), r.data_ptr(), p.data_ptr(), Ap.data_ptr(), minv.data_ptr(),
                                      rz.data_ptr(), pAp.data_ptr(), scratch.data_ptr(), ticket.data_ptr(),
                                      new2.data_ptr(), B.current_stream_ptr(dev.index)), "fcx_pcg_update_xr")
            self._rsum(new2)
            # (after the last iteration of a solve this direction update is not used: x is final)
            check(L.fcx_pcg_update_p(n, p.data_ptr(), r.data_ptr(), minv.data_ptr(), sc[2:3].data_ptr(),
                                     rz.data_ptr(), B.current_stream_ptr(dev.index)), "fcx_pcg_update_p")
            halo(p)
            rz.copy_(sc[2:3])

        state = {"rr": r0 * r0, "pAp": 1.0}

        def converged():
            vals = sc.tolist()  # one host synchronisation: [rz, pAp, rz_new, rr]
            state["rr"], state["pAp"] = vals[3], vals[1]
            return vals[3] <= tol2

        # Tried and removed (profiles/r1zr_*): the K iterations between two host checks captured as a CUDA
        # graph (kernels, all-reduces and ghost exchange included).  On one GPU the loop is GPU-bound and the
        # replayed graph was SLOWER (0.76 vs 0.48 ms per iteration); with the mesh split over two GPUs the
        # replays, interleaved with eager NCCL calls on another stream, deadlocked on the 1 M-cell problem.
        it = 0
        ok = False
        while it < self.cg_max_it:
            for _ in range(K):
                iteration()
            it += K
            if converged():
                ok = True
                break
            if not state["pAp"] > 0.0:
                # fcx_pcg_update_xr sets alpha = 0 for p.Ap <= 0: the iteration would stall silently
                self._krylov_report(False, float(np.sqrt(max(state["rr"], 0.0))) / r0, it, "cg",
                                    "breakdown: p.Ap <= 0 (operator not symmetric positive definite on the "
                                    "free dofs; use linear_solver='bicgstab' or 'dense')")
                return x, it
        self._krylov_report(ok, float(np.sqrt(max(state["rr"], 0.0))) / r0, it, "cg",
                            None if ok else f"no convergence in cg_max_it = {self.cg_max_it} iterations")
        return x, it

    def _krylov_report(self, ok: bool, relres: float, it: int, name: str, why: str | None = None) -> None:
        self.krylov_converged.append(bool(ok))
        self.krylov_relres.append(float(relres))
        if ok:
            return
        msg = f"{name}: linear solve failed after {it} iterations (|r|/|r0| = {relres:.3e}): {why}"
        if self.error_on_krylov_failure:
            raise KrylovError(msg)
        warnings.warn(msg, RuntimeWarning, stacklevel=3)

    def _solve_cg_device(self, rhs, free_mask, diag, rtol=None):
        """Jacobi-PCG through the device-resident Krylov loop (solver/_krylov.py)."""
        import torch

        from ._krylov import DeviceKrylov

        own = self._owned(rhs)
        fm = (free_mask if own is None else free_mask & own).to(torch.float64)
        minv = fm / torch.where(diag.abs() > 0, diag, torch.ones_like(diag))
        if self._device_krylov is None or self._device_krylov.problem is not self.problem:
            import time

            from ._krylov import PeerMemoryUnavailable

            t0 = time.perf_counter()
            try:
                self._device_krylov = DeviceKrylov(self.problem, self.partition)
            except PeerMemoryUnavailable as exc:  # raised on every rank together: fall back together
                warnings.warn(f"{exc}; falling back to cg_driver='python' (NCCL)", RuntimeWarning, stacklevel=2)
                self.cg_driver = "python"
                return self._solve_cg(self.problem.J_apply, rhs, free_mask, diag, rtol)
            torch.cuda.synchronize()
            self.krylov_setup_s += time.perf_counter() - t0
        if self.cg_lookahead is not None:
            self._device_krylov.lookahead = bool(self.cg_lookahead)
        tol = self.cg_rtol if rtol is None else rtol
        x, it, ok, relres, brk = self._device_krylov.solve(rhs, minv, tol, self.cg_max_it, self.cg_check_every)
        why = None
        if brk:
            why = ("breakdown: p.Ap <= 0 (operator not symmetric positive definite on the free dofs; "
                   "use linear_solver='bicgstab' or 'dense')")
        elif not ok:
            why = f"no convergence in cg_max_it = {self.cg_max_it} iterations"
        self._krylov_report(ok, relres, it, "cg (device)", why)
        return x, it

    def _solve_bicgstab(self, apply, rhs, free_mask, diag, rtol=None):
        """Jacobi-preconditioned BiCGStab on the free (owned) dofs for NON-symmetric tangents
        (Drucker-Prager with non-associated flow).  Plain torch vector operations with one host
        synchronisation per dot product: a correctness path for the unusual laws, not a tuned one."""
        import torch

        own = self._owned(rhs)
        halo = self.partition.halo_update if self.partition is not None else (lambda v: None)
        fm = (free_mask if own is None else free_mask & own).to(torch.float64)
        minv = fm / torch.where(diag.abs() > 0, diag, torch.ones_like(diag))
        x = torch.zeros_like(rhs)
        r = rhs * fm
        r0n = self._norm(r)
        if r0n == 0.0:
            self._krylov_report(True, 0.0, 0, "bicgstab")
            return x, 0
        tol = (self.cg_rtol if rtol is None else rtol) * r0n
        rhat = r.clone()
        rho = alpha = omega = 1.0
        v = torch.zeros_like(rhs)
        p = torch.zeros_like(rhs)
        y, z, s, t = (torch.empty_like(rhs) for _ in range(4))
        it, rn = 0, r0n
        while it < self.cg_max_it:
            rho_new = self._dot(rhat, r)
            if rho_new == 0.0 or omega == 0.0:
                self._krylov_report(False, rn / r0n, it, "bicgstab", "breakdown (rho or omega = 0)")
                return x, it
            beta = (rho_new / rho) * (alpha / omega)
            p = r + beta * (p - omega * v)
            torch.mul(minv, p, out=y)
            halo(y)
            apply(y, v)
            v.mul_(fm)
            alpha = rho_new / self._dot(rhat, v)
            s = r - alpha * v
            it += 1
            if self._norm(s) <= tol:
                x.add_(y, alpha=alpha)
                rn = self._norm(s)
                break
            torch.mul(minv, s, out=z)
            halo(z)
            apply(z, t)
            t.mul_(fm)
            tt = self._dot(t, t)
            omega = self._dot(t, s) / tt if tt > 0 else 0.0
            x.add_(y, alpha=alpha).add_(z, alpha=omega)
            r = s - omega * t
            rho = rho_new
            rn = self._norm(r)
            if rn <= tol:
                break
        self._krylov_report(rn <= tol, rn / r0n, it, "bicgstab",
                            None if rn <= tol else f"no convergence in cg_max_it = {self.cg_max_it} iterations")
        return x, it

    # -------------------------------------------------------------- solve
    def solve(self, u):
        """Returns (number of Newton iterations, converged) like dolfinx."""
        import torch

        pb = self.problem
        x = u.x.array
        dev = x.device
        n = x.numel()
        dofs_np, vals_np = pb.bc_dofs_values()
        bc_dofs = torch.as_tensor(dofs_np, dtype=torch.int64, device=dev)
        bc_vals = torch.as_tensor(vals_np, dtype=torch.float64, device=dev)
        free = torch.ones(n, dtype=torch.bool, device=dev)
        free[bc_dofs] = False
        b = torch.empty(n, dtype=torch.float64, device=dev)
        z = torch.zeros(n, dtype=torch.float64, device=dev)
        Jz = torch.empty(n, dtype=torch.float64, device=dev)

        def residual():
            pb.form(x)
            pb.F(x, b)
            # lifting: b_free += (J z)_free with z = g - x on the constrained dofs; b_bc = x - g
            z.zero_()
            z[bc_dofs] = bc_vals - x[bc_dofs]
            if bc_dofs.numel() > 0 and float(z.abs().max().item()) > 0.0:
                pb.J_apply(z, Jz)
                b.add_(Jz)
            b[bc_dofs] = x[bc_dofs] - bc_vals
            return self._norm(b)

        import time as _time

        def _timed(fn, attr):
            if not self.profile:
                return fn()
            torch.cuda.synchronize()
            t0 = _time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            setattr(self, attr, getattr(self, attr) + _time.perf_counter() - t0)
            return out

        self.residual_history, self.krylov_iterations = [], []
        self.krylov_converged, self.krylov_relres = [], []
        multi_rank = self.partition is not None or self.reduce_over_ranks
        if self.linear_solver == "dense" and multi_rank:
            raise ValueError("linear_solver='dense' builds a rank-local matrix and cannot be used on a "
                             "partitioned mesh / with several ranks: use 'cg' or 'bicgstab'")
        if self.linear_solver not in ("auto", "cg", "bicgstab", "dense"):
            raise ValueError(f"unknown linear_solver {self.linear_solver!r}")
        symmetric = bool(getattr(pb, "symmetric_tangent", True))
        r = _timed(residual, "residual_s")
        r0 = r
        self.residual_history.append(r)
        dx0 = None
        it = 0
        eta, r_prev = None, None
        self.forcing_terms = []
        converged = (r < self.atol) if self.convergence_criterion == "residual" else False
        if self.convergence_criterion == "residual" and r0 > 0 and r / r0 < self.rtol:
            converged = True
        while not converged and it < self.max_it:
            use_dense = self.linear_solver == "dense" or (
                self.linear_solver == "auto" and n <= 3000 and not multi_rank)
            use_bicgstab = self.linear_solver == "bicgstab" or (self.linear_solver == "auto" and not symmetric)
            rhs = b.clone()
            if use_dense:
                dx, kit = self._solve_dense(pb.J_apply, rhs, free)
            elif use_bicgstab:
                dx, kit = self._solve_bicgstab(pb.J_apply, rhs, free, pb.J_diag(), None)
            else:
                if self.profile:
                    import time

                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                if self.cg_forcing == "eisenstat-walker":
                    if r_prev is None:
                        eta = self.cg_eta_max
                    else:
                        new = self.cg_eta_gamma * (r / r_prev) ** 2
                        safe = self.cg_eta_gamma * eta ** 2
                        eta = max(new, safe) if safe > 0.1 else new
                    # ... and never tighter than the Newton tolerances can still use (no over-solving)
                    target = max(self.atol, self.rtol * r0)
                    eta = min(self.cg_eta_max, max(eta, self.cg_rtol, 0.5 * target / r if r > 0 else self.cg_rtol))
                    r_prev = r
                    self.forcing_terms.append(eta)
                elif self.cg_forcing is not None:
                    raise ValueError(f"unknown cg_forcing {self.cg_forcing!r}")
                if self.cg_driver == "device" and rhs.is_cuda:
                    dx, kit = self._solve_cg_device(rhs, free, pb.J_diag(), eta)
                elif self.cg_driver in ("device", "python"):
                    dx, kit = self._solve_cg(pb.J_apply, rhs, free, pb.J_diag(), eta)
                else:
                    raise ValueError(f"unknown cg_driver {self.cg_driver!r}")
                if self.profile:
                    torch.cuda.synchronize()
                    self.linear_solve_s += time.perf_counter() - t0
            dx[bc_dofs] = b[bc_dofs]  # identity rows: dx_bc = x_bc - g
            self.krylov_iterations.append(kit)

            def _update():
                x.add_(dx, alpha=-self.relaxation_parameter)
                if self.partition is not None:
                    if self._device_krylov is not None and self._device_krylov.problem is self.problem and x.is_cuda:
                        self._device_krylov.halo_update(x)  # peer-memory push, no NCCL
                    else:
                        self.partition.halo_update(x)

            _timed(_update, "update_s")
            it += 1
            r = _timed(residual, "residual_s")
            self.residual_history.append(r)
            if self.convergence_criterion == "incremental":
                dxn = self._norm(dx)
                dx0 = dxn if dx0 is None else dx0
                converged = dxn < self.atol or (dx0 > 0 and dxn / dx0 < self.rtol)
            else:
                converged = r < self.atol or (r0 > 0 and r / r0 < self.rtol)
            if self.report:
                print(f"Newton iteration {it}: r (abs) = {r:.6e} (tol = {self.atol:.1e}) "
                      f"r (rel) = {r / r0 if r0 > 0 else 0.0:.6e} (tol = {self.rtol:.1e}) krylov its = {kit}")
        if not converged and self.error_on_nonconvergence:
            raise RuntimeError(f"Newton solver did not converge in {it} iterations (|r| = {r:.3e}).")
        return it, converged
