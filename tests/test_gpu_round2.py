"""GPU parity tests added in round 2 (all through the C ABI):

* stress-only ``evaluate`` -- ``tangent=None``, the reference's ``tangent: Option<..>``
  (bindings/src/lib.rs:83,109-113; comfe-rs/src/interfaces.rs:368,441-455) -- for every
  model, device and host entry points, every download wire;
* full-size parity: a 1 M-QP strided sample against the oracle at n = 16 M for
  the four elastic variants of BASELINE config 2, for Kelvin / Maxwell after 1, 10 and 100
  increments (config 4) and for the second Mises step from the hardened state (config 3);
* the fused form() kernel with cell lists (several VonMises3D laws on disjoint cell sets: the
  sub-mesh maps folded into the kernel) against the unfused path through the map kernels.
Tolerances are BASELINE.json's: 1e-12 (elastic, visco), 1e-10 (plastic), per-QP norm-wise."""
import numpy as np
import pytest
import torch

from fenics_constitutive_b200 import models as M
from fenics_constitutive_b200 import solver as S
from fenics_constitutive_b200 import synthetic
from fenics_constitutive_b200._lib import lib
from fenics_constitutive_b200.models import StressStrainConstraint as C
from oracle import models as om

from _util import CONSTRAINT_NAMES, TOL_ELASTIC, TOL_PLASTIC, assert_close

pytestmark = pytest.mark.gpu
ELASTIC, MISES, VISCO = synthetic.ELASTIC_PARAMS, synthetic.MISES_PARAMS, synthetic.VISCO_PARAMS
RUST_PRM = {"mu": np.array([80769.0]), "kappa": np.array([175000.0])}
MODES = ["host", "device"]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _law_case(name, n, seed):
    """(law, oracle law or None, dt, grad, stress0, {history}) for a named model on n points."""
    rng = np.random.default_rng(seed)
    if name.startswith(("elastic_", "kelvin_", "maxwell_")):
        kind, cname = name.split("_", 1)
        c = C[cname]
        g, s = c.geometric_dim, c.stress_strain_dim
        grad = rng.standard_normal(n * g * g) * 1e-3
        stress = rng.standard_normal(n * s) * 0.05
        if kind == "elastic":
            return M.LinearElasticityModel(ELASTIC, c), om.LinearElasticityModel(ELASTIC, c), 1.0, grad, stress, None
        hist = {"strain_visco": rng.standard_normal(n * s) * 1e-4, "strain": rng.standard_normal(n * s) * 1e-3}
        cls, ocls = ((M.SpringKelvinModel, om.SpringKelvinModel) if kind == "kelvin"
                     else (M.SpringMaxwellModel, om.SpringMaxwellModel))
        return cls(VISCO, c), ocls(VISCO, c), 2.0, grad, stress, hist
    if name == "mises":
        grad, s0, e0, a0 = synthetic.mises_inputs_numpy(n, seed=seed)
        # a hardened, stressed start for half of the points
        half = n // 2
        a0[:half] = np.abs(rng.standard_normal(half)) * 2e-3
        s0.reshape(n, 6)[:half] = rng.standard_normal((half, 6)) * 300.0
        return M.VonMises3D(MISES), om.VonMises3D(MISES), 1.0, grad, s0, {"eps_n": e0, "alpha": a0}
    if name == "rust_elastic":
        grad = rng.standard_normal(n * 9) * 1e-3
        return M.LinearElasticity3D(RUST_PRM), None, 1.0, grad, rng.standard_normal(n * 6) * 0.05, None
    if name == "mises_lin":
        law = M.MisesPlasticityLinearHardening3D({**RUST_PRM, "y_0": np.array([1200.0]), "h": np.array([200.0])})
        return law, None, 1.0, rng.standard_normal(n * 9) * 2.9e-3, np.zeros(n * 6), {"history": np.zeros(n * 7)}
    prm = {**RUST_PRM, "a": np.array([300.0]), "b": np.array([0.05]), "b_flow": np.array([0.01])}
    g = rng.standard_normal((n, 9)) * 1.7e-3  # deviator-dominated: far from the cone's apex
    g[:, [0, 4, 8]] = rng.standard_normal((n, 3)) * 4e-4
    if name == "dp_hyperbolic":
        prm["d"] = np.array([40.0])
        law = M.DruckerPragerHyperbolic3D(prm)
    else:
        law = M.DruckerPrager3D(prm)
    return law, None, 1.0, g.ravel(), np.zeros(n * 6), {"history": np.zeros(n * 7)}


def _evaluate(law, dt, grad, stress, tangent, hist, mode):
    """In-place evaluate through the host (numpy) or the device (CUDA tensor) entry point."""
    if mode == "host":
        law.evaluate(0.0, dt, grad, stress, tangent, hist)
        return
    g, s = dev(grad), dev(stress)
    t = dev(tangent) if tangent is not None else None
    h = {k: dev(v) for k, v in hist.items()} if hist is not None else None
    law.evaluate(0.0, dt, g, s, t, h)
    torch.cuda.synchronize()
    stress[:] = s.cpu().numpy()
    if tangent is not None:
        tangent[:] = t.cpu().numpy()
    if hist is not None:
        for k in hist:
            hist[k][:] = h[k].cpu().numpy()


ALL_MODELS = ([f"elastic_{c}" for c in CONSTRAINT_NAMES] + [f"kelvin_{c}" for c in CONSTRAINT_NAMES]
              + [f"maxwell_{c}" for c in CONSTRAINT_NAMES]
              + ["mises", "rust_elastic", "mises_lin", "dp_classic", "dp_hyperbolic"])


@pytest.mark.parametrize("n", [1, 129, 3000, 100_003])
@pytest.mark.parametrize("name", ALL_MODELS)
def test_stress_only_equals_full_evaluate(name, n):
    """tangent=None: stress, history and the plastic flag are bit-identical to the call WITH a
    tangent, on both entry points (n = 100 003: output-staged tiles + ragged tail, chunked host
    pipeline with the download wire; n = 3000: the plain host path); and match the oracle."""
    law, orc, dt, grad, s0, h0 = _law_case(name, n, seed=n + 17)
    s = law.stress_strain_dim
    if hasattr(law, "record_plastic_flag"):
        law.record_plastic_flag = True
    for mode in MODES:
        full = [s0.copy(), np.full(n * s * s, np.nan), {k: v.copy() for k, v in h0.items()} if h0 else None]
        _evaluate(law, dt, grad, full[0], full[1], full[2], mode)
        flag_full = None if not hasattr(law, "plastic_flag") else np.asarray(
            law.plastic_flag if mode == "host" else law.plastic_flag.cpu().numpy()).copy()
        only = [s0.copy(), {k: v.copy() for k, v in h0.items()} if h0 else None]
        _evaluate(law, dt, grad, only[0], None, only[1], mode)
        assert np.array_equal(only[0], full[0]), f"stress {mode}"
        if h0:
            for k in h0:
                assert np.array_equal(only[1][k], full[2][k]), f"history[{k}] {mode}"
        if flag_full is not None:
            flag = law.plastic_flag if mode == "host" else law.plastic_flag.cpu().numpy()
            assert np.array_equal(flag, flag_full), f"plastic flag {mode}"
        if orc is not None:
            ref_s, ref_t = s0.copy(), np.zeros(n * s * s)
            ref_h = {k: v.copy() for k, v in h0.items()} if h0 else None
            orc.evaluate(0.0, dt, grad, ref_s, ref_t, ref_h)
            tol = TOL_PLASTIC if name == "mises" else TOL_ELASTIC
            assert_close(only[0], ref_s, s, tol, f"stress vs oracle {mode}")
            if h0:
                for k in h0:
                    assert_close(only[1][k], ref_h[k], ref_h[k].size // n, tol, f"history[{k}] vs oracle {mode}")


@pytest.mark.parametrize("wire", [2, 1, 0], ids=["direct_wire", "slot_wire", "plain_d2h"])
@pytest.mark.parametrize("kind", ["pageable", "pinned"])
@pytest.mark.parametrize("name", ["mises", "dp_classic", "kelvin_FULL"])
def test_stress_only_host_wires(name, kind, wire):
    """Stress-only host calls on every download wire and memory kind, small chunks (ring slots wrap,
    ragged last chunk): bit-identical to the device path; nothing is written anywhere else."""
    n = 150_001
    law, _, dt, grad, s0, h0 = _law_case(name, n, seed=3)
    g, st = dev(grad), dev(s0)
    h = {k: dev(v) for k, v in h0.items()}
    law.evaluate(0.0, dt, g, st, None, h)
    torch.cuda.synchronize()
    keep = []

    def host(a):
        if kind == "pageable":
            return a.copy()
        t = torch.from_numpy(a.copy()).pin_memory()
        keep.append(t)
        return t.numpy()

    hs, hh = host(s0), {k: host(v) for k, v in h0.items()}
    hg = host(grad)
    L = lib()
    old_chunk, old_wire = L.fcx_host_chunk_qps(40_000), L.fcx_host_wire(wire)
    try:
        law.evaluate(0.0, dt, hg, hs, None, hh)
    finally:
        L.fcx_host_chunk_qps(old_chunk)
        L.fcx_host_wire(old_wire)
    assert np.array_equal(hs, st.cpu().numpy())
    for k in h0:
        assert np.array_equal(hh[k], h[k].cpu().numpy()), k
    assert np.array_equal(hg, grad)


def test_stress_only_c_abi_null_tangent_is_not_an_error():
    """Straight through ctypes: NULL tangent returns FCX_OK (0) for the device entry points and
    the other required pointers are still checked (FCX_ERR_NULL = -3)."""
    L = lib()
    n = 1000
    grad, s0, e0, a0 = (dev(a) for a in synthetic.mises_inputs_numpy(n, seed=1))
    P = np.array([MISES[k] for k in ("p_ka", "p_mu", "p_y0", "p_y00", "p_w")])
    st = torch.cuda.current_stream().cuda_stream
    assert L.fcx_mises_evaluate(P.ctypes.data, n, grad.data_ptr(), s0.data_ptr(), None, e0.data_ptr(),
                                a0.data_ptr(), 0, None, None, st) == 0
    assert L.fcx_mises_evaluate(P.ctypes.data, n, grad.data_ptr(), None, None, e0.data_ptr(),
                                a0.data_ptr(), 0, None, None, st) == -3
    D = np.ascontiguousarray(M.LinearElasticityModel(ELASTIC, C.FULL).D)
    assert L.fcx_elastic_evaluate(5, D.ctypes.data, n, grad.data_ptr(), s0.data_ptr(), None, st) == 0
    torch.cuda.synchronize()


# --------------------------------------------------------------- full size (16 M QPs)

N_FULL = 16_000_000
SAMPLE_STRIDE = 16  # 1 M of the 16 M points go to the oracle (pointwise independence makes sampling exact)


def _sample(t, width, sel):
    return t.view(-1, width)[sel].reshape(-1).cpu().numpy()


@pytest.mark.parametrize("name", ["UNIAXIAL_STRESS", "PLANE_STRAIN", "PLANE_STRESS", "FULL"])
def test_elastic_16m_sample_vs_oracle(name):
    """BASELINE config 2 at full size: E = 42, nu = 0.3, grad ~ N(0,1) 1e-3, non-zero initial
    stress; a 1 M-QP strided sample against the oracle, 1e-12; the tangent of the sample bit-exact."""
    c = C[name]
    g, s = c.geometric_dim, c.stress_strain_dim
    n = N_FULL
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)
    grad = torch.randn(n * g * g, dtype=torch.float64, device="cuda", generator=gen) * 1e-3
    stress = torch.randn(n * s, dtype=torch.float64, device="cuda", generator=gen) * 0.1
    tangent = torch.empty(n * s * s, dtype=torch.float64, device="cuda")
    sel = torch.arange(0, n, SAMPLE_STRIDE, device="cuda")
    m = sel.numel()
    g_s, s_s = _sample(grad, g * g, sel), _sample(stress, s, sel)
    M.LinearElasticityModel(ELASTIC, c).evaluate(0.0, 1.0, grad, stress, tangent, None)
    ref_t = np.zeros(m * s * s)
    orc = om.LinearElasticityModel(ELASTIC, c)
    orc.nthreads = 8
    orc.evaluate(0.0, 1.0, g_s, s_s, ref_t, None)
    assert_close(_sample(stress, s, sel), s_s, s, TOL_ELASTIC, "stress sample")
    assert np.array_equal(_sample(tangent, s * s, sel), ref_t)


@pytest.mark.parametrize("cls,ocls", [(M.SpringKelvinModel, om.SpringKelvinModel),
                                      (M.SpringMaxwellModel, om.SpringMaxwellModel)])
def test_visco_16m_100_increments_sample_vs_oracle(cls, ocls):
    """BASELINE config 4 at full size: FULL, dt = 2, 100 increments with the same increment buffer, state
    carried in place on the device; a 1 M-QP strided sample against the oracle after 1, 10 and 100
    increments, 1e-12."""
    n = N_FULL
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    grad = torch.randn(n * 9, dtype=torch.float64, device="cuda", generator=gen) * 1e-4
    z = lambda w: torch.zeros(n * w, dtype=torch.float64, device="cuda")  # noqa: E731
    st, ev, et, tg = z(6), z(6), z(6), torch.empty(n * 36, dtype=torch.float64, device="cuda")
    sel = torch.arange(0, n, SAMPLE_STRIDE, device="cuda")
    m = sel.numel()
    g_s = _sample(grad, 9, sel)
    ref = [np.zeros(m * 6), np.zeros(m * 6), np.zeros(m * 6)]
    ref_t = np.zeros(m * 36)
    orc = ocls(VISCO, C.FULL)
    orc.nthreads = 8
    law = cls(VISCO, C.FULL)
    for step in range(1, 101):
        law.evaluate(0.0, 2.0, grad, st, tg, {"strain_visco": ev, "strain": et})
        orc.evaluate(0.0, 2.0, g_s, ref[0], ref_t, {"strain_visco": ref[1], "strain": ref[2]})
        if step in (1, 10, 100):
            assert_close(_sample(st, 6, sel), ref[0], 6, TOL_ELASTIC, f"stress step {step}")
            assert_close(_sample(ev, 6, sel), ref[1], 6, TOL_ELASTIC, f"strain_visco step {step}")
            assert_close(_sample(et, 6, sel), ref[2], 6, TOL_ELASTIC, f"strain step {step}")
            assert_close(_sample(tg, 36, sel), ref_t, 36, TOL_ELASTIC, f"tangent step {step}")


def test_mises_16m_second_step_sample_vs_oracle():
    """BASELINE config 3 at full size, SECOND step: from the hardened state the first increment leaves
    (non-zero sigma, eps_n, alpha), a second increment of half the size; 1 M-QP strided sample against
    the oracle (1e-10, identical classification), with and without the tangent."""
    n = N_FULL
    grad, st, ep, al = synthetic.mises_inputs_torch(n, "cuda", seed=77)
    tg = torch.empty(n * 36, dtype=torch.float64, device="cuda")
    law = M.VonMises3D(MISES)
    law.record_plastic_flag = True
    sel = torch.arange(0, n, SAMPLE_STRIDE, device="cuda")
    m = sel.numel()
    g_s = _sample(grad, 9, sel)
    ref = [np.zeros(m * 6), np.zeros(m * 36), np.zeros(m * 6), np.zeros(m)]
    orc = om.VonMises3D(MISES)
    orc.nthreads = 8
    for step, scale in enumerate((1.0, 0.5)):
        g = grad * scale
        if step == 1:
            st2, ep2, al2 = st.clone(), ep.clone(), al.clone()
        law.evaluate(0.0, 1.0, g, st, tg, {"eps_n": ep, "alpha": al})
        orc.evaluate(0.0, 1.0, g_s * scale, ref[0], ref[1], {"eps_n": ref[2], "alpha": ref[3]})
        flag = law.plastic_flag.clone()
        assert_close(_sample(st, 6, sel), ref[0], 6, TOL_PLASTIC, f"stress step {step}")
        assert_close(_sample(tg, 36, sel), ref[1], 36, TOL_PLASTIC, f"tangent step {step}")
        assert_close(_sample(ep, 6, sel), ref[2], 6, TOL_PLASTIC, f"eps_n step {step}")
        assert_close(_sample(al, 1, sel), ref[3], 1, TOL_PLASTIC, f"alpha step {step}")
        assert np.array_equal(flag[sel].cpu().numpy(), orc.plastic_flag), f"classification step {step}"
    frac2 = flag.double().mean().item()
    assert 0.2 < frac2 < 0.95, frac2  # the second step is a genuine elastic / plastic mix
    # the same second step stress-only: identical stress / history / flag at every one of the 16 M points
    law.evaluate(0.0, 1.0, grad * 0.5, st2, None, {"eps_n": ep2, "alpha": al2})
    assert torch.equal(st2, st) and torch.equal(ep2, ep) and torch.equal(al2, al)
    assert torch.equal(law.plastic_flag, flag)


# ------------------------------------------------- several laws on disjoint cell lists

MISES_B = {"p_ka": 150000.0, "p_mu": 70000.0, "p_y0": 900.0, "p_y00": 2000.0, "p_w": 150.0}


def np_(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("degree,qd,n,scale", [(2, 2, (5, 4, 3), 2e-4), (1, 1, (6, 5, 5), 4e-4), (1, 2, (3, 3, 3), 7e-4)])
def test_fused_form_with_cell_lists_equals_map_kernels(degree, qd, n, scale):
    """Two VonMises3D laws with different parameters on interleaved, disjoint cell lists
    (reference: laws = [(law, cells), ...], solver/_solver.py:64-85; SubSpaceMap, solver/maps.py:62-123).
    Fused path: fcx_mises_form reads / writes the parent rows cells[c] directly.  Unfused path: gather
    on the sliced tables, fcx_map_rows_to_sub, evaluate, fcx_map_rows_to_parent x 2.  Two increments;
    every parent and sub-mesh array must agree bit for bit, and cells of law A must match a
    single-law problem with law A alone."""
    mesh = S.create_unit_cube(*n)
    V = S.FunctionSpace(mesh, degree)
    nc = mesh.num_cells
    rng = np.random.default_rng(11)
    perm = rng.permutation(nc)
    cells_a, cells_b = np.sort(perm[: nc // 3]).astype(np.int32), perm[nc // 3:].astype(np.int32)  # b unsorted
    incs = [rng.standard_normal(V.num_dofs) * scale for _ in range(2)]
    results = []
    for fused in (True, False):
        u = S.Function(V)
        la, lb = M.VonMises3D(MISES), M.VonMises3D(MISES_B)
        la.record_plastic_flag = lb.record_plastic_flag = True
        # fused=False: the generic path (local sub arrays + map kernels)
        pb = S.IncrSmallStrainProblem([(la, cells_a), (lb, cells_b)], u, [], qd, fused=fused)
        assert pb.fused == fused
        out = []
        for inc in incs:
            u.x.array.add_(torch.from_numpy(inc).to(pb.device))
            pb.form(u.x.array)
            rec = [np_(pb.stress_1.x.array).copy(), np_(pb.tangent.x.array).copy()]
            for k, law in enumerate((la, lb)):
                rec += [np_(pb._history_1[k]["eps_n"].x.array).copy(), np_(pb._history_1[k]["alpha"].x.array).copy(),
                        np_(law.plastic_flag).copy(), np_(pb._del_grad_u[k].x.array).copy()]
            out.append(rec)
            pb.update()
        results.append(out)
    for a, b in zip(results[0], results[1]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert 0.01 < results[0][1][4].mean() < 0.99  # law A's cells: mixed elastic / plastic
    # law A alone on the whole mesh gives the same rows on law A's cells
    u = S.Function(V)
    pa = S.IncrSmallStrainProblem(M.VonMises3D(MISES), u, [], qd)
    nq = pa.tables.nq
    for k, inc in enumerate(incs):
        u.x.array.add_(torch.from_numpy(inc).to(pa.device))
        pa.form(u.x.array)
        sg = np_(pa.stress_1.x.array).reshape(nc, nq * 6)
        tg = np_(pa.tangent.x.array).reshape(nc, nq * 36)
        assert np.array_equal(results[0][k][0].reshape(nc, nq * 6)[cells_a], sg[cells_a])
        assert np.array_equal(results[0][k][1].reshape(nc, nq * 36)[cells_a], tg[cells_a])
        pa.update()


def test_two_mises_laws_newton_solve_fused_equals_unfused():
    """A bar of two VonMises3D materials pulled into the plastic range: the Newton solve through the
    fused cell-list form() + tangent records equals the solve through the map kernels + dense tangents
    (same iteration counts, displacement and stresses to round-off of the Krylov solves)."""
    left = lambda x: np.isclose(x[0], 0.0)   # noqa: E731
    right = lambda x: np.isclose(x[0], 1.0)  # noqa: E731
    sols = []
    for fused in (True, False):
        mesh = S.create_box((0, 0, 0), (1.0, 0.25, 0.25), 8, 2, 2)
        V = S.functionspace(mesh, ("CG", 2, (3,)))
        u = S.Function(V)
        nc = mesh.num_cells
        xs = mesh.coords[mesh.cells].mean(axis=1)[:, 0]  # cell midpoints
        cells_a = np.nonzero(xs < 0.5)[0].astype(np.int32)
        cells_b = np.nonzero(xs >= 0.5)[0].astype(np.int32)
        assert cells_a.size + cells_b.size == nc and cells_a.size > 0 and cells_b.size > 0
        disp = S.Constant(mesh, 0.0)
        bcs = [S.dirichletbc(S.Constant(mesh, 0.0), S.locate_dofs_geometrical(V, left), V.sub(k)) for k in range(3)]
        bcs.append(S.dirichletbc(disp, S.locate_dofs_geometrical(V, right), V.sub(0)))
        pb = S.IncrSmallStrainProblem([(M.VonMises3D(MISES), cells_a), (M.VonMises3D(MISES_B), cells_b)],
                                      u, bcs, 2, fused=fused)
        assert pb.fused == fused
        solver = S.NewtonSolver(None, pb)
        solver.linear_solver = "cg"
        its = []
        for step in range(1, 4):
            disp.value = 0.004 * step
            k, ok = solver.solve(u)
            assert ok
            its.append(k)
            pb.update()
        sols.append((its, np_(u.x.array).copy(), np_(pb.stress_0.x.array).copy(),
                     [np_(h["alpha"].x.array).copy() for h in pb._history_0]))
    (ia, ua, sa, ha), (ib, ub, sb, hb) = sols
    assert ia == ib
    assert np.abs(ua - ub).max() <= 1e-9 * np.abs(ub).max()
    assert np.abs(sa - sb).max() <= 1e-7 * np.abs(sb).max()
    assert max(h.max() for h in ha) > 0.0  # the bar did yield
    for x, y in zip(ha, hb):
        assert np.abs(x - y).max() <= 1e-9
