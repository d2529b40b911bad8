#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
SANITIZE_ONLY=dp,krylov restricts the run to the named sections (models, dp, fem, krylov, gather)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200.models import (  # noqa: E402
    LinearElasticityModel, SpringKelvinModel, SpringMaxwellModel, StressStrainConstraint, VonMises3D)

C = StressStrainConstraint
dev = torch.device("cuda", 0)
_only = [w for w in os.environ.get("SANITIZE_ONLY", "").split(",") if w]
want = lambda name: not _only or name in _only  # noqa: E731
for n in (1, 128, 1000, 5 * 128 * 148 + 77) if want("models") else ():
    g, s0, e0, a0 = synthetic.mises_inputs_numpy(n, seed=3)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    law = VonMises3D(synthetic.MISES_PARAMS)
    law.record_plastic_flag = True
    tg = torch.empty(n * 36, dtype=torch.float64, device=dev)
    law.evaluate(0.0, 1.0, t(g), t(s0), tg, {"eps_n": t(e0), "alpha": t(a0)})
    for cons in (C.UNIAXIAL_STRESS, C.PLANE_STRAIN, C.FULL):
        gd, sd = cons.geometric_dim, cons.stress_strain_dim
        gr = torch.randn(n * gd * gd, dtype=torch.float64, device=dev) * 1e-3
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        tg = torch.empty(n * sd * sd, dtype=torch.float64, device=dev)
        LinearElasticityModel(synthetic.ELASTIC_PARAMS, cons).evaluate(0.0, 1.0, gr, z(n * sd), tg, None)
        for cls in (SpringKelvinModel, SpringMaxwellModel):
            cls(synthetic.VISCO_PARAMS, cons).evaluate(
                0.0, 2.0, gr, z(n * sd), tg, {"strain_visco": z(n * sd), "strain": z(n * sd)})
from fenics_constitutive_b200.models import DruckerPrager3D, DruckerPragerHyperbolic3D  # noqa: E402
import numpy as _np  # noqa: E402

for cls, extra in ((DruckerPrager3D, {}), (DruckerPragerHyperbolic3D, {"d": _np.array([40.0])})) if want("dp") else ():
    prm = {"mu": _np.array([80769.0]), "kappa": _np.array([175000.0]), "a": _np.array([300.0]),
           "b": _np.array([0.05]), "b_flow": _np.array([0.02]), **extra}
    for n in (1, 1000, 3 * 128 * 148 + 5):
        gr = torch.randn(n * 9, dtype=torch.float64, device=dev) * 1.7e-3
        gr.view(n, 9)[:, [0, 4, 8]] *= 0.2
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        cls(prm).evaluate(0.0, 1.0, gr, z(n * 6), z(n * 36), {"history": z(n * 7)})
        cls(prm).evaluate(0.0, 1.0, gr, z(n * 6), None, {"history": z(n * 7)})  # stress-only instantiation

# FEM kernels (gather, fused form, residual / Jacobian action / diagonal, gather-sum), one CTA
# per SM so that every CTA walks several tiles (prefetch + ticket paths) and meets a ragged tile
from fenics_constitutive_b200 import solver as S  # noqa: E402
from fenics_constitutive_b200._lib import lib  # noqa: E402

lib().fcx_tune(b"ctas_per_sm", 1)
for mk, degree, qd in (((lambda: S.create_unit_cube(9, 9, 8), 2, 2), (lambda: S.create_unit_cube(13, 12, 11), 1, 1))
                       if want("fem") else ()):
    V = S.FunctionSpace(mk(), degree)
    u = S.Function(V)
    pb = S.IncrSmallStrainProblem(VonMises3D(synthetic.MISES_PARAMS), u, [], qd)
    u.x.array.copy_(torch.randn(V.num_dofs, dtype=torch.float64, device=dev) * 3e-4)
    for fused in (True, False):
        pb.fused = fused
        pb.form(u.x.array)
    p = torch.randn(V.num_dofs, dtype=torch.float64, device=dev)
    for variant in (1, 0):
        lib().fcx_tune(b"fem_variant", variant)
        pb.F()
        pb.J_apply(p)
        pb.J_diag()
lib().fcx_tune(b"fem_variant", 1)
# device-resident Krylov loop: reduction tails of the whole last CTA, residual test on the device (the solve freezes
# itself: gated element / reduction / update kernels), look-ahead block enqueue, alternating tile-ticket counters
if want("krylov"):
    left = lambda x: np.isclose(x[0], 0.0)   # noqa: E731
    right = lambda x: np.isclose(x[0], 1.0)  # noqa: E731
    for lookahead in (True, False):
        mesh = S.create_unit_cube(9, 9, 8)
        V = S.functionspace(mesh, ("CG", 2, (3,)))
        u = S.Function(V)
        ux = S.Constant(mesh, 0.0)
        bcs = [S.dirichletbc(S.Constant(mesh, 0.0), S.locate_dofs_geometrical(V, left), V),
               S.dirichletbc(ux, S.locate_dofs_geometrical(V, right), V.sub(0))]
        pb = S.IncrSmallStrainProblem(VonMises3D(synthetic.MISES_PARAMS), u, bcs, 2)
        solver = S.NewtonSolver(None, pb)
        solver.linear_solver, solver.cg_driver, solver.cg_lookahead = "cg", "device", lookahead
        solver.cg_rtol, solver.cg_forcing = 1e-6, "eisenstat-walker"
        for step in (1, 2):
            ux.value = 0.006 * step
            solver.solve(u)
            pb.update()
        print("krylov iterations", solver.krylov_iterations, flush=True)
# companion gather alone: staged (cp.async) and register-load kernels, several tiles per CTA, ragged tail
from fenics_constitutive_b200 import gather as G  # noqa: E402

for gdim, degree, qdeg, ncells in (((3, 2, 2, 148 * 32 * 3 + 17), (3, 1, 1, 148 * 64 * 2 + 5), (2, 2, 2, 148 * 32 * 2 + 3))
                                   if want("gather") else ()):
    rng = np.random.default_rng(ncells)
    pts, _ = G.simplex_quadrature(gdim, qdeg)
    dphi = G.lagrange_gradients(gdim, degree, pts)
    nnodes = ncells // 3
    dofmap = rng.integers(0, nnodes, size=(ncells, dphi.shape[1])).astype(np.int32)
    Jinv = rng.standard_normal((ncells, gdim, gdim)) + 3 * np.eye(gdim)
    op = G.IncrementalGradient(gdim, dofmap, dphi, Jinv)
    uu = torch.randn(nnodes * gdim, dtype=torch.float64, device=dev)
    out = torch.empty(op.num_qps * gdim * gdim, dtype=torch.float64, device=dev)
    for variant in (1, 0):
        lib().fcx_tune(b"gather_variant", variant)
        op.evaluate(uu, None, out)
        op.evaluate(uu, uu * 0.5, out)
lib().fcx_tune(b"gather_variant", 1)
lib().fcx_tune(b"ctas_per_sm", 0)
torch.cuda.synchronize()
print("sanitize_small: done")
