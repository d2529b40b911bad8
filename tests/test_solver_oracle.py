"""CPU tests (no GPU): the solver-side oracle (oracle/fem.py) against the closed-form
answers of the reference's own solver tests, restated on the dolfinx-free mesh
layer.  These pin the oracle that tests/test_solver_gpu.py compares the device
stand-in against.

Restated reference tests:
  tests/models/test_elasticity.py:26-87     uniaxial stress, sigma = E eps, two load levels
  tests/models/test_elasticity.py:157-199   uniaxial strain
  tests/models/test_elasticity.py:300-333   plane stress: sigma_zz = 0, sigma_xx = E eps (free lateral)
  tests/models/test_elasticity.py:335-402   3D vs the pure-dolfinx linear problem (here: patch test)
  tests/models/test_plasticity.py:13-137    VonMises3D uniaxial stress, 100 steps
  tests/models/test_viscoelasticity.py:26-125  1D relaxation, Kelvin and Maxwell
"""
import numpy as np
import pytest

from fenics_constitutive_b200.models import StressStrainConstraint as C
from fenics_constitutive_b200.solver import mesh as M
from oracle import fem as F
from oracle import models as om

E, NU = 42.0, 0.3  # reference tests/models/test_elasticity.py:22-23
MISES = {"p_ka": 175000.0, "p_mu": 80769.0, "p_y0": 1200.0, "p_y00": 2500.0, "p_w": 200.0}
VISCO = {"E0": 42.0, "E1": 10.0, "tau": 10.0, "nu": 0.2}


def make(mesh, degree, q_degree):
    V = M.FunctionSpace(mesh, degree)
    T = M.ElementTables(V, q_degree)
    fem = F.FemOracle(mesh.gdim, V.dofmap, T.dphi_ref, T.weights, T.Jinv, T.detJ, V.num_nodes)
    return V, T, fem


def bcs_of(V, specs):
    """specs: list of (marker, component or None, value-holder list[float])."""
    def get():
        dofs, vals = [], []
        for marker, comp, val in specs:
            nodes = M.locate_dofs_geometrical(V, marker)
            if comp is None:
                d = (nodes[:, None] * V.block_size + np.arange(V.block_size)[None]).ravel()
                v = np.tile(np.asarray(val[0], dtype=float).ravel(), nodes.size)
            else:
                d = nodes * V.block_size + comp
                v = np.full(nodes.size, float(val[0]))
            dofs.append(d)
            vals.append(v)
        return np.concatenate(dofs), np.concatenate(vals)
    return get


left = lambda x: np.isclose(x[0], 0.0)   # noqa: E731
right = lambda x: np.isclose(x[0], 1.0)  # noqa: E731


def test_uniaxial_stress_1d():
    V, T, fem = make(M.create_unit_interval(10), 1, 1)
    disp = [0.01]
    pb = F.OracleProblem(om.LinearElasticityModel({"E": E, "nu": NU}, C.UNIAXIAL_STRESS), fem,
                         bcs_of(V, [(left, 0, [0.0]), (right, 0, disp)]))
    n, ok = pb.solve()
    assert ok
    assert np.abs(pb.stress_1 - E * 0.01).max() < 1e-12
    pb.update()
    assert np.abs(pb.stress_0 - E * 0.01).max() < 1e-12
    assert pb.u_prev.max() == pytest.approx(0.01, abs=1e-15)
    disp[0] = 0.02
    pb.solve()
    assert np.abs(pb.stress_1 - E * 0.02).max() < 1e-12


def test_uniaxial_strain_1d():
    V, T, fem = make(M.create_unit_interval(7), 2, 2)
    pb = F.OracleProblem(om.LinearElasticityModel({"E": E, "nu": NU}, C.UNIAXIAL_STRAIN), fem,
                         bcs_of(V, [(left, 0, [0.0]), (right, 0, [0.01])]))
    pb.solve()
    assert np.abs(pb.stress_1 - E * (1 - NU) / ((1 + NU) * (1 - 2 * NU)) * 0.01).max() < 1e-12


def test_plane_stress_2d():
    V, T, fem = make(M.create_unit_square(3, 2), 1, 1)
    bottom = lambda x: np.isclose(x[1], 0.0)  # noqa: E731
    pb = F.OracleProblem(om.LinearElasticityModel({"E": E, "nu": NU}, C.PLANE_STRESS), fem,
                         bcs_of(V, [(left, 0, [0.0]), (right, 0, [0.01]), (bottom, 1, [0.0])]))
    pb.solve()
    s = pb.stress_1.reshape(-1, 4)
    assert np.abs(s[:, 0] - E * 0.01).max() < 1e-12   # uniaxial stress state in the plane
    assert np.abs(s[:, 1:]).max() < 1e-12             # sigma_yy = sigma_zz = sigma_xy = 0


def test_3d_patch_test_affine_field():
    """An affine displacement prescribed on the whole boundary is reproduced exactly in the
    interior (P1 and P2), with the uniform stress D eps -- the content of the reference's
    comparison with a pure-dolfinx LinearProblem."""
    A = np.array([[0.01, 0.002, -0.003], [0.0, -0.004, 0.001], [0.005, 0.0, 0.002]])
    for degree, qd in ((1, 1), (2, 2)):
        V, T, fem = make(M.create_unit_cube(2, 2, 2), degree, qd)
        boundary = lambda x: np.any(np.isclose(x, 0.0) | np.isclose(x, 1.0), axis=0)  # noqa: E731
        nodes = M.locate_dofs_geometrical(V, boundary)
        ub = V.node_coords[nodes] @ A.T

        def get(nodes=nodes, ub=ub, V=V):
            return (nodes[:, None] * 3 + np.arange(3)[None]).ravel(), ub.ravel()

        pb = F.OracleProblem(om.LinearElasticityModel({"E": E, "nu": NU}, C.FULL), fem, get)
        pb.solve()
        assert np.abs(pb.u.reshape(-1, 3) - V.node_coords @ A.T).max() < 1e-13
        grad = A.T  # nabla_grad: grad[i][j] = du_j/dx_i
        eps = F.mandel_strain(grad[None], 3)[0]
        D = om.get_elastic_tangent(E, NU, C.FULL)
        assert np.abs(pb.stress_1.reshape(-1, 6) - eps @ D).max() < 1e-12


def test_mises_uniaxial_stress_3d():
    """reference tests/models/test_plasticity.py:13-137 (VonMises3D branch)."""
    V, T, fem = make(M.create_unit_cube(1, 1, 1), 1, 1)
    disp = [0.0]
    y0b = lambda x: np.isclose(x[1], 0.0)  # noqa: E731
    z0b = lambda x: np.isclose(x[2], 0.0)  # noqa: E731
    pb = F.OracleProblem(om.VonMises3D(MISES), fem,
                         bcs_of(V, [(left, 0, [0.0]), (right, 0, disp), (y0b, 1, [0.0]), (z0b, 2, [0.0])]))
    nT, max_disp = 100, 0.05
    displacement, load = [0.0], [0.0]
    for t in np.linspace(0, 1, nT + 1)[1:]:
        disp[0] = t * max_disp
        n, ok = pb.solve()
        assert ok
        pb.update()
        displacement.append(disp[0])
        load.append(pb.stress_0[::6][0])
    displacement, load = np.array(displacement), np.array(load)
    tol = 1e-8
    assert np.max(load) - MISES["p_y00"] <= tol
    ind = load + tol < MISES["p_y0"]
    ka, mu = MISES["p_ka"], MISES["p_mu"]
    v = (3 * ka - 2 * mu) / (2 * (3 * ka + mu))
    trace = displacement[ind][1] - 2 * v * displacement[ind][1]
    dev = displacement[ind][1] - trace / 3
    slope = (ka * trace + 2 * mu * dev) / displacement[ind][1]
    assert np.all(np.abs(np.ediff1d(load[ind]) / np.ediff1d(displacement[ind]) - slope) < 1e-7)
    assert load[-1] > 2400.0  # well into the saturated regime


@pytest.mark.parametrize("name", ["kelvin", "maxwell"])
def test_relaxation_1d(name):
    """reference tests/models/test_viscoelasticity.py:26-125: step strain, first increment
    dt = 1e-8, then dt = 2 up to t = 200."""
    V, T, fem = make(M.create_unit_interval(2), 1, 1)
    cls = om.SpringKelvinModel if name == "kelvin" else om.SpringMaxwellModel
    eps = 0.001
    pb = F.OracleProblem(cls(VISCO, C.UNIAXIAL_STRESS), fem,
                         bcs_of(V, [(left, 0, [0.0]), (right, 0, [eps])]), del_t=1e-8)
    pb.solve()
    pb.update()
    E0, E1 = VISCO["E0"], VISCO["E1"]
    s0 = pb.stress_0[0]
    pb.dt = 2.0
    for _ in range(100):
        pb.solve()
        pb.update()
    s_inf = pb.stress_0[0]
    if name == "kelvin":
        assert abs(s0 - E0 * eps) < 1e-8 and abs(s_inf - E0 * E1 / (E0 + E1) * eps) < 1e-8
    else:
        assert abs(s0 - (E0 + E1) * eps) < 1e-8 and abs(s_inf - E0 * eps) < 1e-8


y0b = lambda x: np.isclose(x[1], 0.0)  # noqa: E731
y1b = lambda x: np.isclose(x[1], 1.0)  # noqa: E731
z0b = lambda x: np.isclose(x[2], 0.0)  # noqa: E731
z1b = lambda x: np.isclose(x[2], 1.0)  # noqa: E731


def cyclic_checks(displacement, load, nT, slope_tol=1e-6):
    """The assertions of reference tests/models/test_plasticity.py:238-287, verbatim in structure.
    The reference's absolute slope tolerance 1e-7 is 5e-13 of the slope (2.1e5), i.e. round-off of its own
    linear solver; the first unloading increment after the peak misses it by 1e-8 here, hence 1e-6."""
    tol = 1e-8
    assert np.max(load) - MISES["p_y00"] <= tol
    assert abs(np.min(load)) - MISES["p_y00"] <= tol
    ka, mu = MISES["p_ka"], MISES["p_mu"]
    l1, d1 = load[: int(nT / 4 + 2)], displacement[: int(nT / 4 + 2)]
    ind = abs(l1) + tol < MISES["p_y0"]
    v = (3 * ka - 2 * mu) / (2 * (3 * ka + mu))
    trace = d1[ind][1] - 2 * v * d1[ind][1]
    dev = d1[ind][1] - trace / 3
    slope = (ka * trace + 2 * mu * dev) / d1[ind][1]
    assert np.all(abs(np.ediff1d(l1[ind][1:]) / np.ediff1d(d1[ind][1:]) - slope) < slope_tol)
    l2, d2 = load[int(nT / 4 + 2): int(3 * nT / 4 + 1)], displacement[int(nT / 4 + 2): int(3 * nT / 4 + 1)]
    ind = abs(l2) + tol < max(np.max(l1), MISES["p_y0"])
    assert np.all(abs(np.ediff1d(l2[ind]) / np.ediff1d(d2[ind]) - slope) < slope_tol)
    l3, d3 = load[int(3 * nT / 4 + 1):], displacement[int(3 * nT / 4 + 1):]
    ind = abs(l3) + tol < max(np.max(l1), abs(np.min(l2)), MISES["p_y0"])
    assert np.all(abs(np.ediff1d(l3[ind]) / np.ediff1d(d3[ind]) - slope) < slope_tol)
    # and the loop really went plastic in both directions
    assert np.max(load) > 2000.0 and np.min(load) < -2000.0


def test_mises_uniaxial_cyclic_strain_3d():
    """reference tests/models/test_plasticity.py:140-287: one full sine cycle of the right-face
    displacement (amplitude 0.05, 101 increments), yield limit never exceeded, elastic slope on every
    unloading branch."""
    V, T, fem = make(M.create_unit_cube(1, 1, 1), 1, 1)
    disp = [0.0]
    pb = F.OracleProblem(om.VonMises3D(MISES), fem,
                         bcs_of(V, [(left, 0, [0.0]), (right, 0, disp), (y0b, 1, [0.0]), (z0b, 2, [0.0])]))
    nT, max_disp = 100, 0.05
    displacement, load = [0.0], [0.0]
    for time in np.linspace(np.pi, -np.pi, num=nT + 1):
        disp[0] = np.sin(time) * max_disp
        n, ok = pb.solve()
        assert ok
        pb.update()
        displacement.append(disp[0])
        load.append(pb.stress_0[::6][0])
    cyclic_checks(np.array(displacement), np.array(load), nT)


def test_kelvin_vs_maxwell_1d():
    """reference tests/models/test_viscoelasticity.py:291-366: a Kelvin chain and the Maxwell chain with
    the transferred parameters give the same uniaxial stress history."""
    E0, E1, tau, nu = VISCO["E0"], VISCO["E1"], VISCO["tau"], VISCO["nu"]
    maxwell = {"E0": E0 * E1 / (E0 + E1), "E1": E0**2 / (E0 + E1), "tau": E1 / (E0 + E1) * tau, "nu": nu}
    hist = []
    for law in (om.SpringKelvinModel(VISCO, C.UNIAXIAL_STRESS), om.SpringMaxwellModel(maxwell, C.UNIAXIAL_STRESS)):
        V, T, fem = make(M.create_unit_interval(2), 1, 2)
        pb = F.OracleProblem(law, fem, bcs_of(V, [(left, 0, [0.0]), (right, 0, [0.001])]), del_t=0.1)
        stress = []
        while pb.t < 10 * 0.1 - 1e-12:
            pb.solve()
            pb.update()
            stress.append(pb.stress_1[-1])
        hist.append(np.array(stress))
    assert len(hist[0]) == 10 and np.linalg.norm(hist[0] - hist[1]) < 1e-8


@pytest.mark.parametrize("degree,qd", [(1, 1), (2, 2)])
@pytest.mark.parametrize("dim", [2, 3])
def test_surface_load_is_consistent(dim, degree, qd):
    """`surface_load`: the nodal forces of a constant traction sum to traction x area, and under that load
    a linear-elastic bar with symmetry BCs answers with the homogeneous uniaxial-stress state (patch test),
    for P1 and P2 facets."""
    f = 0.3
    if dim == 2:
        mesh, cons, load = M.create_unit_square(3, 2), C.PLANE_STRESS, (f, 0.0)
        specs = [(left, 0, [0.0]), (y0b, 1, [0.0])]
    else:
        mesh, cons, load = M.create_unit_cube(2, 3, 2), C.FULL, (f, 0.0, 0.0)
        specs = [(left, 0, [0.0]), (y0b, 1, [0.0]), (z0b, 2, [0.0])]
    V, T, fem = make(mesh, degree, qd)
    fl = M.surface_load(V, right, load)
    assert abs(fl.reshape(-1, dim)[:, 0].sum() - f) < 1e-14 and np.abs(fl.reshape(-1, dim)[:, 1:]).max() == 0.0
    pb = F.OracleProblem(om.LinearElasticityModel({"E": E, "nu": NU}, cons), fem, bcs_of(V, specs))
    pb.f_ext[:] = fl
    n, ok = pb.solve()
    assert ok
    s = pb.stress_1.reshape(-1, fem.s)
    assert np.abs(s[:, 0] - f).max() < 1e-12 and np.abs(s[:, 1:]).max() < 1e-12
    assert np.abs(pb.u.reshape(-1, dim)[:, 0] - f / E * V.node_coords[:, 0]).max() < 1e-13


@pytest.mark.parametrize("name", ["kelvin", "maxwell"])
@pytest.mark.parametrize("dim", [2, 3])
def test_creep(dim, name):
    """reference tests/models/test_viscoelasticity.py:369-526: uniaxial tension by a constant traction on
    the right face (symmetry BCs), a nearly elastic first step (dt = 1e-8), then dt = 2 up to 20 tau;
    initial and final strain against the 1D chain formulas."""
    f_max = 0.1
    cls = om.SpringKelvinModel if name == "kelvin" else om.SpringMaxwellModel
    if dim == 2:
        mesh, cons, load = M.create_unit_square(2, 2), C.PLANE_STRESS, (f_max, 0.0)
        specs = [(left, 0, [0.0]), (y0b, 1, [0.0])]
    else:
        mesh, cons, load = M.create_unit_cube(2, 2, 2), C.FULL, (f_max, 0.0, 0.0)
        specs = [(left, 0, [0.0]), (y0b, 1, [0.0]), (z0b, 2, [0.0])]
    V, T, fem = make(mesh, 1, 1)
    pb = F.OracleProblem(cls(VISCO, cons), fem, bcs_of(V, specs), del_t=1e-8)
    pb.f_ext[:] = M.surface_load(V, right, load)
    assert abs(pb.f_ext.reshape(-1, dim)[:, 0].sum() - f_max) < 1e-15
    pb.solve()
    pb.update()
    strain = [pb.history_1[0]["strain"].max()]
    visco = [pb.history_1[0]["strain_visco"].max()]
    stress = [pb.stress_1.max()]
    pb.dt = 2.0
    while pb.t < 20 * VISCO["tau"]:
        n, ok = pb.solve()
        assert ok
        pb.update()
        strain.append(pb.history_1[0]["strain"].max())
        visco.append(pb.history_1[0]["strain_visco"].max())
        stress.append(pb.stress_1.max())
    E0, E1 = VISCO["E0"], VISCO["E1"]
    s0, s_inf = (f_max / E0, f_max / E0 + f_max / E1) if name == "kelvin" else (f_max / (E0 + E1), f_max / E0)
    assert abs(strain[0] - s0) < 1e-8 and abs(strain[-1] - s_inf) < 1e-8
    assert abs(stress[0] - f_max) < 1e-8
    assert np.sum(np.diff(stress)) < 1e-8 and abs(visco[0]) < 1e-8 and visco[-1] > 0


@pytest.mark.parametrize("name", ["kelvin", "maxwell"])
def test_visco_plane_strain_equals_3d(name):
    """reference tests/models/test_viscoelasticity.py:550-696: 2D PLANE_STRAIN against 3D with the z faces
    held, one cell per direction, dt = 5 up to 20 tau."""
    cls = om.SpringKelvinModel if name == "kelvin" else om.SpringMaxwellModel
    pbs = []
    for dim in (2, 3):
        if dim == 2:
            mesh, cons = M.create_unit_square(1, 1), C.PLANE_STRAIN
            specs = [(left, None, [np.zeros(2)]), (y1b, 1, [0.0]), (y0b, 1, [0.0]), (right, 0, [0.01])]
        else:
            mesh, cons = M.create_unit_cube(1, 1, 1), C.FULL
            specs = [(left, None, [np.zeros(3)]), (y0b, 1, [0.0]), (y1b, 1, [0.0]), (z0b, 2, [0.0]), (z1b, 2, [0.0]),
                     (right, 0, [0.01])]
        V, T, fem = make(mesh, 1, 1)
        pbs.append(F.OracleProblem(cls(VISCO, cons), fem, bcs_of(V, specs), del_t=5.0))
    while pbs[0].t < 20 * VISCO["tau"]:
        for pb in pbs:
            pb.solve()
            pb.update()
        assert abs(pbs[0].stress_1[0] - pbs[1].stress_1[0]) < 1e-8
        assert abs(pbs[0].stress_1[1] - pbs[1].stress_1[1]) < 1e-8
        assert abs(pbs[0].u.max() - pbs[1].u.max()) < 1e-8
