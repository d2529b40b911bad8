"""The reference's Rust-backed models (models/rust_models.py) and 3D->1D/2D adapters
(models/utils.py:211-412) on the B200 path.

CPU part: the C restatement of comfe-rs MisesPlasticity3D (oracle_rs_mises_linear_hardening)
has NO golden vectors in the reference (the crate cannot be compiled here and the reference's
tests pin only its elastic slope, tests/models/test_plasticity.py:32-37,124-137), so it is pinned
by the properties the algorithm must have: elastic slope, yield consistency after the return
map, and tangent == d(stress)/d(strain) by central differences.
GPU part: CUDA kernels vs that oracle; adapters vs the native low-dimensional models (reference
tests/models/test_elasticity.py:201-236,273-297)."""
import numpy as np
import pytest

from fenics_constitutive_b200.models import StressStrainConstraint as C
from oracle import models as om

PRM = {"mu": np.array([80769.0]), "kappa": np.array([175000.0]), "y_0": np.array([1200.0]), "h": np.array([200.0])}
R2 = 0.70710678118654752440


def mandel_to_grad(e):
    """A (symmetric) grad_del_u [n][9] whose Mandel strain is e [n][6]."""
    g = np.zeros((e.shape[0], 9))
    g[:, 0], g[:, 4], g[:, 8] = e[:, 0], e[:, 1], e[:, 2]
    g[:, 1] = g[:, 3] = e[:, 3] / (2 * R2)
    g[:, 2] = g[:, 6] = e[:, 4] / (2 * R2)
    g[:, 5] = g[:, 7] = e[:, 5] / (2 * R2)
    return g


def run_oracle(e, sig0=None, hist0=None):
    n = e.shape[0]
    law = om.RustMisesPlasticityLinearHardening3D(PRM)
    sig = np.zeros(n * 6) if sig0 is None else sig0.copy()
    hist = np.zeros(n * 7) if hist0 is None else hist0.copy()
    tan = np.zeros(n * 36)
    law.evaluate(0.0, 1.0, mandel_to_grad(e).ravel(), sig, tan, {"history": hist})
    return sig.reshape(n, 6), tan.reshape(n, 6, 6), hist.reshape(n, 7), law.plastic_flag


def test_oracle_linear_hardening_elastic_and_yield_consistency():
    rng = np.random.default_rng(0)
    mu, ka, y0, h = (float(PRM[k][0]) for k in ("mu", "kappa", "y_0", "h"))
    e = rng.standard_normal((2000, 6)) * 3e-3
    sig, tan, hist, flag = run_oracle(e)
    assert 0.2 < flag.mean() < 0.8
    el = flag == 0
    tr = e[:, :3].sum(1)
    dev = e.copy()
    dev[:, :3] -= tr[:, None] / 3
    sig_el = 2 * mu * dev
    sig_el[:, :3] += ka * tr[:, None]
    assert np.abs(sig[el] - sig_el[el]).max() < 1e-9
    # plastic points sit on the updated yield surface: sqrt(3/2) |dev sigma| = y0 + h alpha
    s = sig.copy()
    s[:, :3] -= sig[:, :3].sum(1, keepdims=True) / 3
    seq = np.sqrt(1.5 * (s**2).sum(1))
    pl = ~el
    assert np.abs(seq[pl] - (y0 + h * hist[pl, 0])).max() < 1e-8
    assert np.all(seq[el] < y0) and np.all(hist[el] == 0.0)
    # plastic strain is deviatoric; as written in mises_plasticity.rs:103-110 the flow direction is
    # n = s_tr / q with |n| = sqrt(2/3) and del_gamma = sqrt(3/2) del_alpha, hence alpha = |eps_p|
    assert np.abs(hist[pl, 1:4].sum(1)).max() < 1e-15
    assert np.abs(hist[pl, 0] - np.linalg.norm(hist[pl, 1:], axis=1)).max() < 1e-15


def test_oracle_linear_hardening_tangent_as_written_and_stress_derivative():
    """The stress update is pinned through its derivative: central differences of the oracle stress
    equal the textbook consistent tangent kappa 1x1 + 2 mu theta P_dev - 2 mu theta_bar nhat nhat^T
    (nhat the UNIT flow direction).  The tangent the crate WRITES (mises_plasticity.rs:115-121) uses
    n = s_tr / q (|n|^2 = 2/3) and a plus sign; no reference test pins it, so it
    is reproduced as written and checked here against an independent numpy evaluation of that line."""
    rng = np.random.default_rng(1)
    mu, ka, y0, h = (float(PRM[k][0]) for k in ("mu", "kappa", "y_0", "h"))
    e = rng.standard_normal((200, 6)) * 6e-3
    sig, tan, hist, flag = run_oracle(e)
    pl = flag == 1
    assert pl.sum() > 50
    tr = e[:, :3].sum(1)
    s_tr = 2 * mu * e
    s_tr[:, :3] -= 2 * mu * tr[:, None] / 3
    q = np.sqrt(1.5 * (s_tr**2).sum(1))
    dal = (q - y0) / (3 * mu + h)
    theta = 1 - 3 * mu * dal / q
    theta_bar = 1 / (1 + h / (3 * mu)) - (1 - theta)
    n = s_tr / q[:, None]
    oo = np.zeros((6, 6))
    oo[:3, :3] = 1.0
    pdev = np.eye(6) - oo / 3
    base = ka * oo[None] + 2 * mu * theta[:, None, None] * pdev[None]
    nn = n[:, :, None] * n[:, None, :]
    written = base + 2 * mu * theta_bar[:, None, None] * nn
    consistent = base - 3 * mu * theta_bar[:, None, None] * nn
    assert np.abs(tan[pl] - written[pl]).max() < 1e-9 * np.abs(tan).max()
    assert np.abs(tan[~pl] - (ka * oo + 2 * mu * pdev)[None]).max() == 0.0
    d = 1e-7
    for k in range(6):
        ep, em = e.copy(), e.copy()
        ep[:, k] += d
        em[:, k] -= d
        sp, _, _, fp = run_oracle(ep)
        sm, _, _, fm = run_oracle(em)
        ok = pl & (fp == flag) & (fm == flag)
        col = (sp - sm) / (2 * d)
        assert np.abs(col[ok] - consistent[ok][:, :, k]).max() < 1e-6 * np.abs(tan).max()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["host", "device"])
@pytest.mark.parametrize("n", [1, 129, 100_003])
def test_gpu_linear_hardening_vs_oracle(mode, n):
    import torch

    from fenics_constitutive_b200.models import MisesPlasticityLinearHardening3D

    rng = np.random.default_rng(n)
    grad = rng.standard_normal(n * 9) * 2.9e-3
    law = MisesPlasticityLinearHardening3D(PRM)
    law.record_plastic_flag = True
    orc = om.RustMisesPlasticityLinearHardening3D(PRM)
    s_ref, t_ref, h_ref = np.zeros(n * 6), np.zeros(n * 36), np.zeros(n * 7)
    s, t, h = s_ref.copy(), t_ref.copy(), h_ref.copy()
    if mode == "device":
        s, t, h = (torch.from_numpy(a).cuda() for a in (s, t, h))
    for step in range(2):  # second step starts from non-zero stress and history
        g = grad * (1.0 + 0.5 * step)
        orc.evaluate(0.0, 1.0, g, s_ref, t_ref, {"history": h_ref})
        law.evaluate(0.0, 1.0, torch.from_numpy(g).cuda() if mode == "device" else g, s, t, {"history": h})
        out = [a.cpu().numpy() if mode == "device" else a for a in (s, t, h)]
        flag = law.plastic_flag.cpu().numpy() if mode == "device" else law.plastic_flag
        assert np.array_equal(flag, orc.plastic_flag)
        from _util import rel_err

        assert rel_err(out[0], s_ref, 6) <= 1e-10
        assert rel_err(out[1], t_ref, 36) <= 1e-10
        assert rel_err(out[2], h_ref, 7) <= 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["host", "device"])
def test_gpu_rust_linear_elasticity3d_vs_oracle(mode):
    import torch

    from _util import rel_err
    from fenics_constitutive_b200.models import LinearElasticity3D

    E, nu = 42.0, 0.3
    prm = {"mu": np.array([E / (2 * (1 + nu))]), "kappa": np.array([E / (3 * (1 - 2 * nu))])}
    n = 4099
    rng = np.random.default_rng(3)
    grad, s0 = rng.standard_normal(n * 9) * 1e-3, rng.standard_normal(n * 6) * 0.1
    s_ref, t_ref = s0.copy(), np.zeros(n * 36)
    om.RustLinearElasticity3D(prm).evaluate(0.0, 1.0, grad, s_ref, t_ref)
    s, t = s0.copy(), np.zeros(n * 36)
    if mode == "device":
        g, s, t = (torch.from_numpy(a).cuda() for a in (grad, s, t))
    else:
        g = grad
    LinearElasticity3D(prm).evaluate(0.0, 1.0, g, s, t, None)
    s, t = (a.cpu().numpy() if mode == "device" else a for a in (s, t))
    assert rel_err(s, s_ref, 6) <= 1e-12 and rel_err(t, t_ref, 36) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["host", "device"])
@pytest.mark.parametrize("cons", [C.UNIAXIAL_STRAIN, C.PLANE_STRAIN])
def test_gpu_adapters_equal_native_low_dimensional_models(cons, mode):
    """UniaxialStrainFrom3D(FULL law) == UNIAXIAL_STRAIN law, PlaneStrainFrom3D(FULL law) ==
    PLANE_STRAIN law (reference tests/models/test_elasticity.py:201-236, 273-297), two calls so the
    persistent 3D scratch state is exercised."""
    import torch

    from _util import rel_err
    from fenics_constitutive_b200.models import LinearElasticityModel, PlaneStrainFrom3D, UniaxialStrainFrom3D

    prm = {"E": 42.0, "nu": 0.3}
    g, s = cons.geometric_dim, cons.stress_strain_dim
    n = 1000
    rng = np.random.default_rng(7)
    wrap = (UniaxialStrainFrom3D if cons == C.UNIAXIAL_STRAIN else PlaneStrainFrom3D)(LinearElasticityModel(prm, C.FULL))
    native = LinearElasticityModel(prm, cons)
    assert wrap.constraint == cons and wrap.stress_strain_dim == s and wrap.history_dim is None
    sa, sb = np.zeros(n * s), np.zeros(n * s)
    ta, tb = np.zeros(n * s * s), np.zeros(n * s * s)
    if mode == "device":
        sa, sb, ta, tb = (torch.from_numpy(a).cuda() for a in (sa, sb, ta, tb))
    for step in range(2):
        grad = rng.standard_normal(n * g * g) * 1e-3
        gg = torch.from_numpy(grad).cuda() if mode == "device" else grad
        wrap.evaluate(0.0, 1.0, gg, sa, ta, None)
        native.evaluate(0.0, 1.0, gg, sb, tb, None)
        xa, xb, ya, yb = (a.cpu().numpy() if mode == "device" else a for a in (sa, sb, ta, tb))
        assert rel_err(xa, xb, s) <= 1e-12 and rel_err(ya, yb, s * s) <= 1e-12


@pytest.mark.gpu
def test_linear_hardening_uniaxial_stress_through_solver():
    """reference tests/models/test_plasticity.py:13-137, MisesPlasticityLinearHardening3D branch
    (test_max_stress = False): the elastic range has the slope of uniaxial stress; afterwards the
    hardening slope is E h / (E + h)."""
    from fenics_constitutive_b200 import solver as S
    from fenics_constitutive_b200.models import MisesPlasticityLinearHardening3D

    mesh = S.create_unit_cube(1, 1, 1)
    V = S.functionspace(mesh, ("CG", 1, (3,)))
    u = S.Function(V)
    zero, sx = S.Constant(mesh, 0.0), S.Constant(mesh, 0.0)
    f = lambda k, v: (lambda x: np.isclose(x[k], v))  # noqa: E731
    bcs = [S.dirichletbc(zero, S.locate_dofs_geometrical(V, f(0, 0.0)), V.sub(0)),
           S.dirichletbc(sx, S.locate_dofs_geometrical(V, f(0, 1.0)), V.sub(0)),
           S.dirichletbc(zero, S.locate_dofs_geometrical(V, f(1, 0.0)), V.sub(1)),
           S.dirichletbc(zero, S.locate_dofs_geometrical(V, f(2, 0.0)), V.sub(2))]
    problem = S.IncrSmallStrainProblem(MisesPlasticityLinearHardening3D(PRM), u, bcs, q_degree=1)
    solver = S.NewtonSolver(None, problem)
    disp, load = [0.0], [0.0]
    for t in np.linspace(0, 1, 51)[1:]:
        sx.value = t * 0.02
        solver.solve(u)
        problem.update()
        disp.append(sx.value)
        load.append(problem.stress_0.numpy()[::6][0])
    disp, load = np.array(disp), np.array(load)
    mu, ka, y0, h = (float(PRM[k][0]) for k in ("mu", "kappa", "y_0", "h"))
    E = 9 * ka * mu / (3 * ka + mu)
    slopes = np.ediff1d(load) / np.ediff1d(disp)
    el = load[1:] + 1e-8 < y0
    assert el.sum() >= 5 and np.abs(slopes[el] - E).max() < 1e-6
    pl = load[:-1] > y0 + 1.0
    assert pl.sum() >= 5 and np.abs(slopes[pl] - E * h / (E + h)).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["host", "device"])
@pytest.mark.parametrize("lname", ["elastic", "mises"])
@pytest.mark.parametrize("aname", ["uniaxial_strain", "plane_strain"])
def test_gpu_adapters_match_reference_goldens(aname, lname, mode):
    """UniaxialStrainFrom3D / PlaneStrainFrom3D around the CUDA FULL models against fixtures produced by the
    reference's own adapters around its own models (oracle/gen_golden_adapters.py), three consecutive calls
    on one adapter object: the persistent 3D scratch arrays behave like the reference's
    (models/utils.py:252-273, 341-359), on the numpy path and on the device path
    (fcx_embed_3d / fcx_extract_from_3d)."""
    import torch

    from _util import TOL_ELASTIC, TOL_PLASTIC, assert_close, golden
    from fenics_constitutive_b200.models import LinearElasticityModel, PlaneStrainFrom3D, UniaxialStrainFrom3D, VonMises3D

    G = golden("adapters.npz")
    n, ncalls = int(G["n"]), int(G["ncalls"])
    cls, s = (UniaxialStrainFrom3D, 1) if aname == "uniaxial_strain" else (PlaneStrainFrom3D, 4)
    inner = (LinearElasticityModel({"E": 42.0, "nu": 0.3}, C.FULL) if lname == "elastic"
             else VonMises3D({"p_ka": 175000.0, "p_mu": 80769.0, "p_y0": 1200.0, "p_y00": 2500.0, "p_w": 200.0}))
    wrap = cls(inner)
    key = f"{aname}_{lname}"
    to = (lambda a: torch.from_numpy(a).cuda()) if mode == "device" else (lambda a: a)
    back = (lambda a: a.cpu().numpy()) if mode == "device" else (lambda a: a)
    stress, tangent = to(np.zeros(n * s)), to(np.full(n * s * s, np.nan))
    history = {"eps_n": to(np.zeros(n * 6)), "alpha": to(np.zeros(n))} if lname == "mises" else None
    tol = TOL_ELASTIC if lname == "elastic" else TOL_PLASTIC
    for k in range(ncalls):
        wrap.evaluate(0.0, 1.0, to(G[f"{key}_grad{k}"].copy()), stress, tangent, history)
        assert_close(back(stress), G[f"{key}_stress{k}"], s, tol, f"{key} stress call {k}")
        assert_close(back(tangent), G[f"{key}_tangent{k}"], s * s, tol, f"{key} tangent call {k}")
        if history is not None:
            assert_close(back(history["eps_n"]), G[f"{key}_eps_n{k}"], 6, tol, f"{key} eps_n call {k}")
            assert_close(back(history["alpha"]), G[f"{key}_alpha{k}"], 1, tol, f"{key} alpha call {k}")
            assert np.array_equal(back(history["alpha"]) > 0, G[f"{key}_alpha{k}"] > 0)
