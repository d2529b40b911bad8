"""GPU tests of the device-resident solver stand-in:
fused form() kernel, residual / Jacobian-action kernels, sub-mesh maps and the
Newton driver, against the CPU oracle (oracle/fem.py + oracle/models.py) and the
closed-form answers of the reference's solver tests."""
import numpy as np
import pytest

from fenics_constitutive_b200.models import (
    LinearElasticityModel, SpringKelvinModel, SpringMaxwellModel, VonMises3D)
from fenics_constitutive_b200.models import StressStrainConstraint as C
from fenics_constitutive_b200 import solver as S
from oracle import fem as F
from oracle import models as om
from _util import rel_err

pytestmark = pytest.mark.gpu

E, NU = 42.0, 0.3
MISES = {"p_ka": 175000.0, "p_mu": 80769.0, "p_y0": 1200.0, "p_y00": 2500.0, "p_w": 200.0}
VISCO = {"E0": 42.0, "E1": 10.0, "tau": 10.0, "nu": 0.2}

left = lambda x: np.isclose(x[0], 0.0)    # noqa: E731
right = lambda x: np.isclose(x[0], 1.0)   # noqa: E731
y0b = lambda x: np.isclose(x[1], 0.0)     # noqa: E731
z0b = lambda x: np.isclose(x[2], 0.0)     # noqa: E731


def oracle_for(problem):
    V, T = problem.V, problem.tables
    return F.FemOracle(V.mesh.gdim, V.dofmap, T.dphi_ref, T.weights, T.Jinv, T.detJ, V.num_nodes)


def np_(t):
    return t.detach().cpu().numpy()


MESHES = [
    ("tet_p2_q2", lambda: S.create_box((0, 0, 0), (1.0, 0.7, 1.3), 3, 2, 2), 2, 2, C.FULL),
    ("tet_p1_q1", lambda: S.create_unit_cube(3, 3, 2), 1, 1, C.FULL),
    ("tet_p1_q2", lambda: S.create_unit_cube(2, 2, 2), 1, 2, C.FULL),
    ("tri_p2_q2", lambda: S.create_rectangle((0, 0), (2.0, 1.0), 5, 3), 2, 2, C.PLANE_STRAIN),
    ("tri_p1_q1", lambda: S.create_unit_square(4, 3), 1, 1, C.PLANE_STRESS),
    ("int_p2_q2", lambda: S.create_unit_interval(9), 2, 2, C.UNIAXIAL_STRAIN),
    ("int_p1_q1", lambda: S.create_unit_interval(5), 1, 1, C.UNIAXIAL_STRESS),
]


@pytest.mark.parametrize("name,mk,degree,qd,cons", MESHES, ids=[m[0] for m in MESHES])
def test_residual_and_jacobian_kernels_vs_oracle(name, mk, degree, qd, cons):
    import torch

    mesh = mk()
    V = S.FunctionSpace(mesh, degree)
    u = S.Function(V)
    pb = S.IncrSmallStrainProblem(LinearElasticityModel({"E": E, "nu": NU}, cons), u, [], qd)
    fem = oracle_for(pb)
    rng = np.random.default_rng(5)
    s = cons.stress_strain_dim
    stress = rng.standard_normal(pb.nqp * s)
    tangent = rng.standard_normal(pb.nqp * s * s)  # deliberately non-symmetric: checks the C^T convention
    p = rng.standard_normal(V.num_dofs)
    pb.stress.current.x.array.copy_(torch.from_numpy(stress))
    pb.tangent.x.array.copy_(torch.from_numpy(tangent))
    b = np_(pb.F())
    ref = fem.internal_force(stress)
    assert np.abs(b - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    K = fem.tangent_matrix(tangent)
    y = np_(pb.J_apply(torch.from_numpy(p).to(pb.device)))
    assert np.abs(y - K @ p).max() <= 1e-12 * np.abs(K @ p).max()
    d = np_(pb.J_diag())
    assert np.abs(d - K.diagonal()).max() <= 1e-12 * np.abs(K.diagonal()).max()
    # gather against the oracle as well
    uu = rng.standard_normal(V.num_dofs) * 1e-3
    up = rng.standard_normal(V.num_dofs) * 1e-3
    u.x.array.copy_(torch.from_numpy(uu))
    pb.incr_disp.previous.x.array.copy_(torch.from_numpy(up))
    ctx = pb._law_on_submeshs[0]
    pb.incr_disp.evaluate_local_incremental_gradient(ctx.gather_op, ctx.displacement_gradient_fn)
    g = np_(ctx.displacement_gradient_fn.x.array)
    gref = fem.grad(uu, up)
    assert np.abs(g - gref).max() <= 1e-13 * np.abs(gref).max()


@pytest.mark.parametrize("name,mk,degree,qd,cons", [
    ("tet_p2_q2", lambda: S.create_unit_cube(10, 9, 8), 2, 2, C.FULL),
    ("tri_p2_q2", lambda: S.create_unit_square(53, 47), 2, 2, C.PLANE_STRAIN),
    ("tet_p1_q1", lambda: S.create_unit_cube(13, 12, 11), 1, 1, C.FULL),
], ids=["tet_p2_q2", "tri_p2_q2", "tet_p1_q1"])
@pytest.mark.parametrize("ctas_per_sm", [1, 0])
def test_fem_kernels_many_tiles_per_cta(name, mk, degree, qd, cons, ctas_per_sm):
    """Persistent-CTA paths of the QP-parallel FEM kernels: with one CTA per SM every CTA walks
    several tiles (bulk prefetch one tile ahead, atomic tile tickets, ragged last tile).  Both
    element-kernel variants against the oracle, and against each other."""
    import torch

    from fenics_constitutive_b200._lib import lib

    mesh = mk()
    V = S.FunctionSpace(mesh, degree)
    u = S.Function(V)
    pb = S.IncrSmallStrainProblem(LinearElasticityModel({"E": E, "nu": NU}, cons), u, [], qd)
    fem = oracle_for(pb)
    rng = np.random.default_rng(17)
    s = cons.stress_strain_dim
    stress = rng.standard_normal(pb.nqp * s)
    tangent = rng.standard_normal(pb.nqp * s * s)
    p = rng.standard_normal(V.num_dofs)
    uu = rng.standard_normal(V.num_dofs) * 1e-3
    up = rng.standard_normal(V.num_dofs) * 1e-3
    pb.stress.current.x.array.copy_(torch.from_numpy(stress))
    pb.tangent.x.array.copy_(torch.from_numpy(tangent))
    u.x.array.copy_(torch.from_numpy(uu))
    pb.incr_disp.previous.x.array.copy_(torch.from_numpy(up))
    pd = torch.from_numpy(p).to(pb.device)
    L = lib()
    old = L.fcx_tune(b"ctas_per_sm", ctas_per_sm)
    try:
        got = {}
        for variant in (1, 0):
            L.fcx_tune(b"fem_variant", variant)
            got[variant] = (np_(pb.F()).copy(), np_(pb.J_apply(pd)).copy(), np_(pb.J_diag()).copy())
        ctx = pb._law_on_submeshs[0]
        pb.incr_disp.evaluate_local_incremental_gradient(ctx.gather_op, ctx.displacement_gradient_fn)
        g = np_(ctx.displacement_gradient_fn.x.array)
    finally:
        L.fcx_tune(b"fem_variant", 1)
        L.fcx_tune(b"ctas_per_sm", old)
    K = fem.tangent_matrix(tangent)
    refs = (fem.internal_force(stress), K @ p, K.diagonal())
    for variant in (1, 0):
        for x, r in zip(got[variant], refs):
            assert np.abs(x - r).max() <= 1e-12 * np.abs(r).max(), f"variant {variant}"
    gref = fem.grad(uu, up)
    assert np.abs(g - gref).max() <= 1e-13 * np.abs(gref).max()


@pytest.mark.parametrize("degree,qd,n,scale", [(2, 2, (9, 8, 7), 1.0e-4), (1, 2, (6, 5, 5), 3.0e-4)])
def test_jacobian_action_from_tangent_records(degree, qd, n, scale):
    """Fused form() also emits the VonMises3D tangent as 10-double records; the matrix-free Jacobian
    action computed from them (80 B/QP) equals the one from the dense 6x6 tangents (288 B/QP) and the
    oracle's assembled matrix, on a mixed elastic/plastic state."""
    import torch

    mesh = S.create_unit_cube(*n)
    V = S.FunctionSpace(mesh, degree)
    u = S.Function(V)
    law = VonMises3D(MISES)
    law.record_plastic_flag = True
    pb = S.IncrSmallStrainProblem(law, u, [], qd)
    assert pb.fused and pb.use_tangent_records
    rng = np.random.default_rng(23)
    u.x.array.copy_(torch.from_numpy(rng.standard_normal(V.num_dofs) * scale).to(pb.device))
    pb.form(u.x.array)
    frac = float(law.plastic_flag.double().mean().item())
    assert 0.02 < frac < 0.98, frac
    p = torch.from_numpy(rng.standard_normal(V.num_dofs)).to(pb.device)
    assert pb._trec_valid
    y_rec = np_(pb.J_apply(p)).copy()
    pb.use_tangent_records = False
    y_full = np_(pb.J_apply(p)).copy()
    pb.use_tangent_records = True
    assert np.abs(y_rec - y_full).max() <= 1e-12 * np.abs(y_full).max()
    K = oracle_for(pb).tangent_matrix(np_(pb.tangent.x.array))
    ref = K @ np_(p)
    assert np.abs(y_rec - ref).max() <= 1e-12 * np.abs(ref).max()
    # the unfused path invalidates the records
    pb.fused = False
    pb.form(u.x.array)
    assert not pb._trec_valid


@pytest.mark.parametrize("degree,qd,n,scale", [(2, 2, (5, 4, 3), 2e-4), (1, 1, (6, 5, 5), 4e-4), (1, 2, (3, 3, 3), 7e-4)])
def test_fused_form_equals_unfused_and_oracle(degree, qd, n, scale):
    """fcx_mises_form == gather + trial reset + evaluate, bit for bit, and both match the oracle.
    Two consecutive increments so that the second starts from non-zero sigma/eps_n/alpha."""
    import torch

    mesh = S.create_unit_cube(*n)
    V = S.FunctionSpace(mesh, degree)
    rng = np.random.default_rng(11)
    incs = [rng.standard_normal(V.num_dofs) * scale for _ in range(2)]  # mixed elastic/plastic
    results = []
    for fused in (True, False):
        u = S.Function(V)
        law = VonMises3D(MISES)
        law.record_plastic_flag = True
        pb = S.IncrSmallStrainProblem(law, u, [], qd)
        assert pb.fused
        pb.fused = fused
        out = []
        for inc in incs:
            u.x.array.add_(torch.from_numpy(inc).to(pb.device))
            pb.form(u.x.array)
            out.append([np_(pb.stress_1.x.array).copy(), np_(pb.tangent.x.array).copy(),
                        np_(pb._history_1[0]["eps_n"].x.array).copy(), np_(pb._history_1[0]["alpha"].x.array).copy(),
                        np_(law.plastic_flag).copy(), np_(pb._del_grad_u[0].x.array).copy()])
            pb.update()
        results.append(out)
    for a, b in zip(results[0], results[1]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # oracle
    pbf = S.IncrSmallStrainProblem(VonMises3D(MISES), S.Function(V), [], qd)
    fem = oracle_for(pbf)
    opb = F.OracleProblem(om.VonMises3D(MISES), fem, lambda: (np.zeros(0, dtype=int), np.zeros(0)))
    for k, inc in enumerate(incs):
        opb.u += inc
        opb.form()
        sg, tg, ep, al, fl, _ = results[0][k]
        assert rel_err(sg, opb.stress_1, 6) <= 1e-10
        assert rel_err(tg, opb.tangent, 36) <= 1e-10
        assert rel_err(ep, opb.history_1[0]["eps_n"], 6) <= 1e-10
        assert rel_err(al, opb.history_1[0]["alpha"], 1) <= 1e-10
        assert 0.01 < fl.mean() < 0.99
        opb.update()


def gpu_bcs(V, specs):
    return [S.dirichletbc(val, S.locate_dofs_geometrical(V, marker), V if comp is None else V.sub(comp))
            for marker, comp, val in specs]


def test_uniaxial_stress_1d_elastic():
    """reference tests/models/test_elasticity.py:26-87"""
    mesh = S.create_unit_interval(10)
    V = S.functionspace(mesh, ("CG", 1))
    u = S.Function(V)
    law = LinearElasticityModel(parameters={"E": E, "nu": NU}, constraint=C.UNIAXIAL_STRESS)
    displacement = S.Constant(mesh, 0.01)
    bcs = gpu_bcs(V, [(left, None, S.Constant(mesh, 0.0)), (right, None, displacement)])
    problem = S.IncrSmallStrainProblem(law, u, bcs, 1)
    solver = S.NewtonSolver(None, problem)
    n, converged = solver.solve(u)
    assert converged
    assert np.abs(problem.stress_1.numpy() - E * 0.01).max() < 1e-12
    problem.update()
    assert np.abs(problem.stress_0.numpy() - E * 0.01).max() < 1e-12
    assert problem._u0.numpy().max() == pytest.approx(displacement.value, abs=1e-15)
    displacement.value = 0.02
    solver.solve(u)
    assert np.abs(problem.stress_1.numpy() - E * 0.02).max() < 1e-12


def test_two_law_bar():
    """reference tests/models/test_elasticity.py:90-154: two materials on disjoint cell sets
    (SubSpaceMap): homogeneous stress, strain ratio = inverse stiffness ratio."""
    mesh = S.create_unit_interval(10)
    V = S.functionspace(mesh, ("CG", 1))
    u = S.Function(V)
    laws = [
        (LinearElasticityModel({"E": E, "nu": NU}, C.UNIAXIAL_STRESS), np.arange(0, 5, dtype=np.int32)),
        (LinearElasticityModel({"E": 2 * E, "nu": NU}, C.UNIAXIAL_STRESS), np.arange(5, 10, dtype=np.int32)),
    ]
    bcs = gpu_bcs(V, [(left, None, S.Constant(mesh, 0.0)), (right, None, S.Constant(mesh, 0.01))])
    problem = S.IncrSmallStrainProblem(laws, u, bcs, 1)
    solver = S.NewtonSolver(None, problem)
    n, converged = solver.solve(u)
    problem.update()
    sig = problem.stress_0.numpy()
    assert np.abs(sig - sig[0]).max() < 1e-13
    g = [f.numpy() for f in problem._del_grad_u]
    assert np.allclose(g[0], g[0][0]) and np.allclose(g[1], g[1][0])
    assert g[0][0] / g[1][0] == pytest.approx(2.0, rel=1e-12)
    # series springs: sigma = eps_total / (0.5/E + 0.5/(2E))
    assert sig[0] == pytest.approx(0.01 / (0.5 / E + 0.25 / E), rel=1e-12)


def test_plane_stress_and_plane_strain_2d():
    """reference tests/models/test_elasticity.py:239-333"""
    mesh = S.create_unit_square(3, 2)
    V = S.functionspace(mesh, ("CG", 2, (2,)))
    for cons in (C.PLANE_STRESS, C.PLANE_STRAIN):
        u = S.Function(V)
        bcs = gpu_bcs(V, [(left, 0, S.Constant(mesh, 0.0)), (right, 0, S.Constant(mesh, 0.01)),
                          (y0b, 1, S.Constant(mesh, 0.0))])
        problem = S.IncrSmallStrainProblem(LinearElasticityModel({"E": E, "nu": NU}, cons), u, bcs, 2)
        S.NewtonSolver(None, problem).solve(u)
        s = problem.stress_1.numpy().reshape(-1, 4)
        if cons == C.PLANE_STRESS:
            assert np.abs(s[:, 0] - E * 0.01).max() < 1e-12 and np.abs(s[:, 1:]).max() < 1e-12
        else:
            assert np.abs(s[:, 0] - E / (1 - NU**2) * 0.01).max() < 1e-12   # free in y, eps_zz = 0
            assert np.abs(s[:, 2] - NU * s[:, 0]).max() < 1e-12             # sigma_zz != 0
            assert np.abs(s[:, 1]).max() < 1e-12


@pytest.mark.parametrize("degree,qd", [(2, 2), (1, 1)], ids=["readme_cg2", "test_3d_cg1"])
@pytest.mark.parametrize("kind", ["python", "rust"])
def test_readme_example_3d_elasticity(kind, degree, qd):
    """BASELINE.json configs[0] = the reference's README example (README.md:47-78) and its test twin
    tests/models/test_elasticity.py:336-402: unit cube 2x2x2, FULL LinearElasticityModel (python) /
    LinearElasticity3D (rust, mu / kappa), left face clamped, right face moved by (0.01, 0, 0), ONE
    NewtonSolver solve + update().  The reference checks against a pure-FEniCS linear solve; here the twin
    is the oracle's sparse direct solve of the same discrete problem.  (The README asks for q_degree 1 with
    CG2, which under-integrates the stiffness; the well-posed q_degree 2 is used for CG2.)"""
    from fenics_constitutive_b200.models import LinearElasticity3D

    mesh = S.create_unit_cube(2, 2, 2)
    V = S.functionspace(mesh, ("CG", degree, (3,)))
    u = S.Function(V)
    if kind == "python":
        law = LinearElasticityModel(parameters={"E": E, "nu": NU}, constraint=C.FULL)
        olaw = om.LinearElasticityModel({"E": E, "nu": NU}, C.FULL)
    else:
        prm = {"mu": np.array([E / (2 * (1 + NU))]), "kappa": np.array([E / (3 * (1 - 2 * NU))])}
        law, olaw = LinearElasticity3D(prm), om.RustLinearElasticity3D(prm)
    zero, disp = S.Constant(mesh, 0.0), S.Constant(mesh, 0.01)
    bcs = gpu_bcs(V, [(left, 0, zero), (left, 1, zero), (left, 2, zero),
                      (right, 0, disp), (right, 1, zero), (right, 2, zero)])
    problem = S.IncrSmallStrainProblem(law, u, bcs, qd)
    solver = S.NewtonSolver(None, problem)
    n, converged = solver.solve(u)
    assert converged and n == 1  # linear problem: one Newton step
    problem.update()
    opb = F.OracleProblem(olaw, oracle_for(problem), problem.bc_dofs_values)
    on, ook = opb.solve()
    opb.update()
    assert ook
    assert np.abs(u.numpy() - opb.u).max() <= 1e-10 * np.abs(opb.u).max()
    assert rel_err(problem.stress_0.numpy(), opb.stress_0, 6) <= 1e-9
    # clamped faces: sigma_xx is positive everywhere and the mean axial stress lies between the
    # uniaxial-stress (E eps) and uniaxial-strain (E (1 - nu) / ((1 + nu)(1 - 2 nu)) eps) bounds
    sxx = problem.stress_0.numpy()[::6]
    assert E * 0.01 < sxx.mean() < E * (1 - NU) / ((1 + NU) * (1 - 2 * NU)) * 0.01


def run_mises_uniaxial(n_steps, mesh_n, degree, qd, linear_solver="auto", forcing=None, krylov=None):
    mesh = S.create_unit_cube(*mesh_n)
    V = S.functionspace(mesh, ("CG", degree, (3,)))
    u = S.Function(V)
    law = VonMises3D(MISES)
    scalar_x = S.Constant(mesh, 0.0)
    zero = S.Constant(mesh, 0.0)
    bcs = gpu_bcs(V, [(left, 0, zero), (right, 0, scalar_x), (y0b, 1, zero), (z0b, 2, zero)])
    problem = S.IncrSmallStrainProblem(law, u, bcs, q_degree=qd)
    solver = S.NewtonSolver(None, problem)
    solver.linear_solver = linear_solver
    solver.cg_forcing = forcing
    # oracle twin
    fem = oracle_for(problem)
    opb = F.OracleProblem(om.VonMises3D(MISES), fem, problem.bc_dofs_values)
    displacement, load = [0.0], [0.0]
    worst = 0.0
    for t in np.linspace(0, 1, n_steps + 1)[1:]:
        scalar_x.value = t * 0.05
        niter, converged = solver.solve(u)
        assert converged
        if krylov is not None:
            krylov.append(sum(solver.krylov_iterations))
        problem.update()
        on, ook = opb.solve()
        opb.update()
        assert ook
        worst = max(worst, rel_err(problem.stress_0.numpy(), opb.stress_0, 6),
                    np.abs(u.numpy() - opb.u).max() / np.abs(opb.u).max())
        displacement.append(scalar_x.value)
        load.append(problem.stress_0.numpy()[::6][0])
    return np.array(displacement), np.array(load), worst, problem, opb


def test_mises_uniaxial_stress_3d_100_steps():
    """reference tests/models/test_plasticity.py:13-137 on the device stand-in (fused form kernel,
    one cube = 6 P1 tets, q_degree 1), and step-by-step against the CPU oracle."""
    displacement, load, worst, problem, opb = run_mises_uniaxial(100, (1, 1, 1), 1, 1)
    assert problem.fused
    tol = 1e-8
    assert np.max(load) - MISES["p_y00"] <= tol
    ind = load + tol < MISES["p_y0"]
    ka, mu = MISES["p_ka"], MISES["p_mu"]
    v = (3 * ka - 2 * mu) / (2 * (3 * ka + mu))
    trace = displacement[ind][1] - 2 * v * displacement[ind][1]
    dev = displacement[ind][1] - trace / 3
    slope = (ka * trace + 2 * mu * dev) / displacement[ind][1]
    assert np.all(np.abs(np.ediff1d(load[ind]) / np.ediff1d(displacement[ind]) - slope) < 1e-7)
    assert worst < 1e-9
    assert rel_err(problem._history_0[0]["alpha"].numpy(), opb.history_0[0]["alpha"], 1) < 1e-9


def test_mises_p2_mesh_cg_vs_oracle():
    """P2 tets, q_degree 2, matrix-free PCG linear solver vs the oracle's sparse LU."""
    displacement, load, worst, problem, opb = run_mises_uniaxial(6, (2, 2, 2), 2, 2, linear_solver="cg")
    assert problem.fused and load[-1] > 1500.0
    # both Newton loops stop at rtol 1e-9 from different linear solvers: iterates agree to ~1e-8
    assert worst < 1e-7


def test_mises_p2_mesh_inexact_newton():
    """Eisenstat-Walker forcing terms (NewtonSolver.cg_forcing): early Newton steps are solved loosely,
    the Newton tolerances are untouched -- same answer as the oracle's sparse LU, fewer Krylov iterations
    than with every linear solve taken to cg_rtol."""
    k_fixed, k_ew = [], []
    run_mises_uniaxial(4, (3, 3, 3), 2, 2, linear_solver="cg", krylov=k_fixed)
    _, load, worst, problem, _ = run_mises_uniaxial(4, (3, 3, 3), 2, 2, linear_solver="cg",
                                                    forcing="eisenstat-walker", krylov=k_ew)
    assert load[-1] > 1500.0 and worst < 1e-7
    assert sum(k_ew) < 0.8 * sum(k_fixed), (k_fixed, k_ew)


@pytest.mark.parametrize("cls,ocls", [(SpringKelvinModel, om.SpringKelvinModel), (SpringMaxwellModel, om.SpringMaxwellModel)])
def test_relaxation_3d_visco(cls, ocls):
    """reference tests/models/test_viscoelasticity.py:128-288 (3D relaxation under a step strain with
    free lateral faces): sigma_xx(0+) and sigma_xx(inf) closed forms, and oracle parity."""
    mesh = S.create_unit_cube(1, 1, 1)
    V = S.functionspace(mesh, ("CG", 1, (3,)))
    u = S.Function(V)
    zero = S.Constant(mesh, 0.0)
    eps = 0.001
    bcs = gpu_bcs(V, [(left, 0, zero), (right, 0, S.Constant(mesh, eps)), (y0b, 1, zero), (z0b, 2, zero)])
    problem = S.IncrSmallStrainProblem(cls(VISCO, C.FULL), u, bcs, 1, del_t=1e-8)
    solver = S.NewtonSolver(None, problem)
    opb = F.OracleProblem(ocls(VISCO, C.FULL), oracle_for(problem), problem.bc_dofs_values, del_t=1e-8)
    solver.solve(u)
    problem.update()
    opb.solve()
    opb.update()
    s0 = problem.stress_0.numpy()[0]
    problem._del_t = 2.0
    opb.dt = 2.0
    for _ in range(100):
        solver.solve(u)
        problem.update()
        opb.solve()
        opb.update()
    s_inf = problem.stress_0.numpy()[0]
    E0, E1 = VISCO["E0"], VISCO["E1"]
    if cls is SpringKelvinModel:
        assert abs(s0 - E0 * eps) < 1e-8 and abs(s_inf - E0 * E1 / (E0 + E1) * eps) < 1e-8
    else:
        assert abs(s0 - (E0 + E1) * eps) < 1e-8 and abs(s_inf - E0 * eps) < 1e-8
    assert rel_err(problem.stress_0.numpy(), opb.stress_0, 6) < 1e-9
    assert problem._time == pytest.approx(1e-8 + 200.0)


y1b = lambda x: np.isclose(x[1], 1.0)   # noqa: E731
z1b = lambda x: np.isclose(x[2], 1.0)   # noqa: E731


def test_mises_uniaxial_cyclic_strain_3d():
    """reference tests/models/test_plasticity.py:140-287 on the device stand-in: one full sine cycle of the
    right-face displacement; the reference's own assertions (tests/test_solver_oracle.py::cyclic_checks) and
    step-by-step parity with the CPU oracle."""
    from test_solver_oracle import cyclic_checks

    mesh = S.create_unit_cube(1, 1, 1)
    V = S.functionspace(mesh, ("CG", 1, (3,)))
    u = S.Function(V)
    scalar_x, zero = S.Constant(mesh, 0.0), S.Constant(mesh, 0.0)
    bcs = gpu_bcs(V, [(left, 0, zero), (right, 0, scalar_x), (y0b, 1, zero), (z0b, 2, zero)])
    problem = S.IncrSmallStrainProblem(VonMises3D(MISES), u, bcs, q_degree=1)
    solver = S.NewtonSolver(None, problem)
    opb = F.OracleProblem(om.VonMises3D(MISES), oracle_for(problem), problem.bc_dofs_values)
    nT, max_disp = 100, 0.05
    displacement, load, worst = [0.0], [0.0], 0.0
    for time in np.linspace(np.pi, -np.pi, num=nT + 1):
        scalar_x.value = np.sin(time) * max_disp
        n, converged = solver.solve(u)
        assert converged
        problem.update()
        on, ook = opb.solve()
        assert ook
        opb.update()
        worst = max(worst, rel_err(problem.stress_0.numpy(), opb.stress_0, 6))
        displacement.append(scalar_x.value)
        load.append(problem.stress_0.numpy()[::6][0])
    cyclic_checks(np.array(displacement), np.array(load), nT)
    assert worst < 1e-8
    assert rel_err(problem._history_0[0]["eps_n"].numpy(), opb.history_0[0]["eps_n"], 6) < 1e-8


@pytest.mark.parametrize("cls,ocls", [(SpringKelvinModel, om.SpringKelvinModel), (SpringMaxwellModel, om.SpringMaxwellModel)])
@pytest.mark.parametrize("dim", [2, 3])
def test_creep_under_traction(dim, cls, ocls):
    """reference tests/models/test_viscoelasticity.py:369-526: constant traction on the right face
    (``problem.f_ext`` = the consistent nodal load, ``S.surface_load``), nearly elastic first step, then
    dt = 2 up to 20 tau; 1D chain formulas for the initial and final strain, oracle parity at the end."""
    import torch

    f_max = 0.1
    if dim == 2:
        mesh, cons, load = S.create_unit_square(2, 2), C.PLANE_STRESS, (f_max, 0.0)
    else:
        mesh, cons, load = S.create_unit_cube(2, 2, 2), C.FULL, (f_max, 0.0, 0.0)
    V = S.functionspace(mesh, ("CG", 1, (dim,)))
    u = S.Function(V)
    zero = S.Constant(mesh, 0.0)
    specs = [(left, 0, zero), (y0b, 1, zero)] + ([(z0b, 2, zero)] if dim == 3 else [])
    problem = S.IncrSmallStrainProblem(cls(VISCO, cons), u, gpu_bcs(V, specs), 1, del_t=1e-8)
    fl = S.surface_load(V, right, load)
    problem.f_ext.copy_(torch.from_numpy(fl).to(problem.f_ext.device))
    solver = S.NewtonSolver(None, problem)
    opb = F.OracleProblem(ocls(VISCO, cons), oracle_for(problem), problem.bc_dofs_values, del_t=1e-8)
    opb.f_ext[:] = fl
    solver.solve(u)
    problem.update()
    opb.solve()
    opb.update()
    strain = [problem._history_1[0]["strain"].numpy().max()]
    visco = [problem._history_1[0]["strain_visco"].numpy().max()]
    stress = [problem.stress_1.numpy().max()]
    problem._del_t = 2.0
    opb.dt = 2.0
    while problem._time < 20 * VISCO["tau"]:
        n, converged = solver.solve(u)
        assert converged
        problem.update()
        opb.solve()
        opb.update()
        strain.append(problem._history_1[0]["strain"].numpy().max())
        visco.append(problem._history_1[0]["strain_visco"].numpy().max())
        stress.append(problem.stress_1.numpy().max())
    E0, E1 = VISCO["E0"], VISCO["E1"]
    if cls is SpringKelvinModel:
        s0, s_inf = f_max / E0, f_max / E0 + f_max / E1
    else:
        s0, s_inf = f_max / (E0 + E1), f_max / E0
    assert abs(strain[0] - s0) < 1e-8 and abs(strain[-1] - s_inf) < 1e-8
    assert abs(stress[0] - f_max) < 1e-8
    assert np.sum(np.diff(stress)) < 1e-8 and abs(visco[0]) < 1e-8 and visco[-1] > 0
    assert np.abs(u.numpy() - opb.u).max() <= 1e-9 * np.abs(opb.u).max()
    assert rel_err(problem.stress_0.numpy(), opb.stress_0, cons.stress_strain_dim) < 1e-8


@pytest.mark.parametrize("cls", [SpringKelvinModel, SpringMaxwellModel])
def test_visco_plane_strain_equals_3d(cls):
    """reference tests/models/test_viscoelasticity.py:550-696: 2D PLANE_STRAIN against 3D with the z faces
    held, dt = 5 up to 20 tau, compared at every step."""
    runs = []
    for dim in (2, 3):
        if dim == 2:
            mesh, cons = S.create_unit_square(1, 1), C.PLANE_STRAIN
        else:
            mesh, cons = S.create_unit_cube(1, 1, 1), C.FULL
        V = S.functionspace(mesh, ("CG", 1, (dim,)))
        u = S.Function(V)
        zero = S.Constant(mesh, 0.0)
        specs = [(left, c, zero) for c in range(dim)] + [(y1b, 1, zero), (y0b, 1, zero), (right, 0, S.Constant(mesh, 0.01))]
        if dim == 3:
            specs += [(z0b, 2, zero), (z1b, 2, zero)]
        problem = S.IncrSmallStrainProblem(cls(VISCO, cons), u, gpu_bcs(V, specs), 1, del_t=5.0)
        runs.append((u, problem, S.NewtonSolver(None, problem)))
    while runs[0][1]._time < 20 * VISCO["tau"]:
        for u, problem, solver in runs:
            solver.solve(u)
            problem.update()
        (u2, p2, _), (u3, p3, _) = runs
        s2, s3 = p2.stress_1.numpy(), p3.stress_1.numpy()
        assert abs(s2[0] - s3[0]) < 1e-8 and abs(s2[1] - s3[1]) < 1e-8
        assert abs(u2.numpy().max() - u3.numpy().max()) < 1e-8


def test_kelvin_vs_maxwell_1d():
    """reference tests/models/test_viscoelasticity.py:291-366: Kelvin chain vs the Maxwell chain with
    transferred parameters, uniaxial stress, ten steps of dt = 0.1."""
    E0, E1, tau, nu = VISCO["E0"], VISCO["E1"], VISCO["tau"], VISCO["nu"]
    maxwell = {"E0": E0 * E1 / (E0 + E1), "E1": E0**2 / (E0 + E1), "tau": E1 / (E0 + E1) * tau, "nu": nu}
    hist = []
    for law in (SpringKelvinModel(VISCO, C.UNIAXIAL_STRESS), SpringMaxwellModel(maxwell, C.UNIAXIAL_STRESS)):
        mesh = S.create_unit_interval(2)
        V = S.functionspace(mesh, ("CG", 1))
        u = S.Function(V)
        bcs = gpu_bcs(V, [(left, None, S.Constant(mesh, 0.0)), (right, None, S.Constant(mesh, 0.001))])
        problem = S.IncrSmallStrainProblem(law, u, bcs, 2, del_t=0.1)
        solver = S.NewtonSolver(None, problem)
        stress = []
        while problem._time < 10 * 0.1 - 1e-12:
            solver.solve(u)
            problem.update()
            stress.append(problem.stress_1.numpy()[-1])
        hist.append(np.array(stress))
    assert len(hist[0]) == 10 and np.linalg.norm(hist[0] - hist[1]) < 1e-8


def test_partition_hooks_single_rank():
    """NewtonSolver with a MeshPartition attached (solver/partitioned.py) on ONE rank: every node is owned,
    there is nothing to exchange, and the solve is the plain one (the multi-rank exchange itself is covered by
    tests/test_mesh_partition.py on gloo and scripts/check_partitioned_newton.py on 2 GPUs)."""
    mesh = S.create_unit_cube(4, 3, 2)
    sols = []
    for use_partition in (False, True):
        part = S.MeshPartition(mesh, 2, 0, 1) if use_partition else None
        V = part.V if use_partition else S.functionspace(mesh, ("CG", 2, (3,)))
        u = S.Function(V)
        zero, ux = S.Constant(mesh, 0.0), S.Constant(mesh, 0.0)
        bcs = gpu_bcs(V, [(left, 0, zero), (left, 1, zero), (left, 2, zero), (right, 0, ux)])
        problem = S.IncrSmallStrainProblem(VonMises3D(MISES), u, bcs, q_degree=2)
        solver = S.NewtonSolver(None, problem)
        solver.linear_solver = "cg"
        if use_partition:
            part.attach(solver)
            assert part.num_owned_nodes == V.num_nodes and part.neighbours == [] and not solver.reduce_over_ranks
        its = []
        for k in (1, 2):
            ux.value = 0.006 * k
            n, converged = solver.solve(u)
            assert converged
            problem.update()
            its.append(n)
        x = u.numpy() if not use_partition else part.gather_global(u.numpy())
        sols.append((x, its, problem.stress_0.numpy()))
    assert sols[0][1] == sols[1][1]
    assert np.abs(sols[0][0] - sols[1][0]).max() <= 1e-9 * np.abs(sols[0][0]).max()
    assert rel_err(sols[1][2], sols[0][2], 6) <= 1e-8
    assert float(np.abs(sols[0][2]).max()) > 1200.0  # the second step went plastic


@pytest.mark.parametrize("forcing,lookahead", [(None, True), (None, False), ("eisenstat-walker", False),
                                               ("eisenstat-walker", True)],
                         ids=["fixed-lookahead", "fixed-drained", "ew-drained", "ew-lookahead"])
def test_device_krylov_driver_equals_python_driver(forcing, lookahead):
    """The device-resident Krylov loop (csrc/fcx_krylov.cu: single-reduction PCG driven from C) against the
    kernel-by-kernel two-reduction PCG issued from Python: same displacement and stresses to the Krylov tolerance,
    Mises plasticity across the yield point on a P2 mesh (CG path: more than 3000 dofs); every linear solve reports
    convergence.  lookahead = the residual test runs on the device and the host enqueues blocks of iterations
    ahead of knowing its outcome (the solve stops at the exact iteration); drained = the stream is drained and the
    test is made on the host after every block, like the Python driver does (iteration counts in blocks of 10)."""
    sols = []
    for driver in ("device", "python"):
        mesh = S.create_unit_cube(7, 6, 5)
        V = S.functionspace(mesh, ("CG", 2, (3,)))
        u = S.Function(V)
        ux = S.Constant(mesh, 0.0)
        bcs = [S.dirichletbc(S.Constant(mesh, 0.0), S.locate_dofs_geometrical(V, left), V),
               S.dirichletbc(ux, S.locate_dofs_geometrical(V, right), V.sub(0))]
        pb = S.IncrSmallStrainProblem(VonMises3D(MISES), u, bcs, 2)
        solver = S.NewtonSolver(None, pb)
        solver.linear_solver = "cg"
        solver.cg_driver = driver
        solver.cg_lookahead = lookahead
        solver.cg_rtol = 1e-11
        solver.cg_forcing = forcing
        its, kits = [], []
        for step in (1, 2):
            ux.value = 0.006 * step
            k, ok = solver.solve(u)
            assert ok and all(solver.krylov_converged)
            its.append(k)
            kits.append(list(solver.krylov_iterations))
            pb.update()
        sols.append((its, np_(u.x.array).copy(), np_(pb.stress_0.x.array).copy(),
                     float((pb._history_0[0]["alpha"].x.array > 0).double().mean().item()), kits))
    (ia, ua, sa, pa, ka), (ib, ub, sb, _, kb) = sols
    if forcing is None or not lookahead:
        assert ia == ib
    else:  # loose linear solves stopped at the exact iteration instead of the end of a block: a Newton step may move
        assert all(abs(a - b) <= 1 for a, b in zip(ia, ib))
    if not lookahead:  # host test after every block of cg_check_every = 10 iterations
        assert all(k % 10 == 0 for ks in ka for k in ks)
    assert 0.05 < pa < 1.0
    # fixed Krylov tolerance (1e-11): the two solves agree to it.  Inexact Newton (Eisenstat-Walker): the two PCG
    # variants stop their loose linear solves at different iterates, so the Newton paths differ and agree only to
    # the Newton stop rule (residual rtol 1e-9) -- the bar scripts/check_partitioned_newton.py uses as well
    assert np.abs(ua - ub).max() <= (1e-8 if forcing is None else 1e-7) * np.abs(ub).max()
    assert np.abs(sa - sb).max() <= 1e-6 * np.abs(sb).max()


def test_device_krylov_reports_breakdown_on_indefinite_operator():
    """p.Ap <= 0 is reported (KrylovError), not iterated on silently: a sign-flipped tangent."""
    mesh = S.create_unit_cube(6, 6, 5)
    V = S.functionspace(mesh, ("CG", 2, (3,)))
    u = S.Function(V)
    bcs = [S.dirichletbc(S.Constant(mesh, 0.0), S.locate_dofs_geometrical(V, left), V),
           S.dirichletbc(S.Constant(mesh, 0.01), S.locate_dofs_geometrical(V, right), V.sub(0))]
    law = LinearElasticityModel({"E": E, "nu": NU}, C.FULL)
    law.D = -law.D
    pb = S.IncrSmallStrainProblem(law, u, bcs, 2)
    solver = S.NewtonSolver(None, pb)
    solver.linear_solver = "cg"
    with pytest.raises(S.KrylovError):
        solver.solve(u)
