"""comfe-rs Drucker-Prager models (DruckerPrager3D, DruckerPragerHyperbolic3D; reference
models/rust_models.py:96-141 over comfe-rs/src/plasticity/{general,drucker_prager_classic,
drucker_prager_hyperbolic}.rs) -- SURVEY.md 8f row 4.

PARITY UNPINNED against the reference: the Rust crate cannot be compiled here (no rustc) and
the reference has no test of these models.  CPU part: the C restatement of the generic 8x8
return mapper (oracle_rs_drucker_prager, dense LU as in general.rs:176-190) is pinned by what
the algorithm must produce -- the closed-form return of the classic cone, f(sigma_1) = 0, the
flow direction, the plastic-strain bookkeeping, tangent == d(stress)/d(strain) by central
differences -- and by the reference's documented quirk (alpha grows by sqrt(2/3)|g|).
GPU part: the CUDA kernels (closed-form structured Newton step, csrc/fcx_models.cuh) against
that oracle to 1e-10 with identical elastic/plastic classification."""
import numpy as np
import pytest

from oracle import models as om

MU, KA = 80769.0, 175000.0
R2 = 0.70710678118654752440


def params(a=300.0, b=0.05, b_flow=None, d=None):
    p = {"mu": np.array([MU]), "kappa": np.array([KA]), "a": np.array([a]), "b": np.array([b]),
         "b_flow": np.array([b if b_flow is None else b_flow])}
    if d is not None:
        p["d"] = np.array([d])
    return p


def make_grad(n, seed, vol=4e-4, shear=1.7e-3):
    """Deviator-dominated increments: ~50 % plastic for a = 300, far from the apex of the cone."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, 9)) * shear
    g[:, [0, 4, 8]] = rng.standard_normal((n, 3)) * vol
    return g.ravel()


def mandel(grad):
    g = grad.reshape(-1, 9)
    return np.stack([g[:, 0], g[:, 4], g[:, 8], R2 * (g[:, 1] + g[:, 3]), R2 * (g[:, 2] + g[:, 6]),
                     R2 * (g[:, 5] + g[:, 7])], axis=1)


def mandel_to_grad(e):
    g = np.zeros((e.shape[0], 9))
    g[:, 0], g[:, 4], g[:, 8] = e[:, 0], e[:, 1], e[:, 2]
    g[:, 1] = g[:, 3] = e[:, 3] / (2 * R2)
    g[:, 2] = g[:, 6] = e[:, 4] / (2 * R2)
    g[:, 5] = g[:, 7] = e[:, 5] / (2 * R2)
    return g.ravel()


def C_apply(x):
    tr = x[:, :3].sum(1, keepdims=True)
    out = 2 * MU * x
    out[:, :3] += (KA - 2 * MU / 3) * tr
    return out


def Cinv_apply(x):
    tr = x[:, :3].sum(1, keepdims=True)
    out = x / (2 * MU)
    out[:, :3] += (1 / (9 * KA) - 1 / (6 * MU)) * tr
    return out


def run(cls, prm, grad, sig0=None, hist0=None):
    n = grad.size // 9
    law = cls(prm)
    sig = np.zeros(n * 6) if sig0 is None else sig0.copy()
    hist = np.zeros(n * 7) if hist0 is None else hist0.copy()
    tan = np.full(n * 36, np.nan)
    law.evaluate(0.0, 1.0, grad, sig, tan, {"history": hist})
    return sig.reshape(n, 6), tan.reshape(n, 6, 6), hist.reshape(n, 7), np.asarray(law.plastic_flag).astype(bool)


def invariants(sig, d2=0.0):
    i1 = sig[:, :3].sum(1)
    s = sig.copy()
    s[:, :3] -= i1[:, None] / 3
    j2 = 0.5 * (s**2).sum(1)
    return i1, s, np.sqrt(j2 + d2)


# ----------------------------------------------------------------------------- oracle (CPU)

@pytest.mark.parametrize("b_flow", [0.05, 0.01, 0.0])
def test_oracle_classic_matches_closed_form_return(b_flow):
    """For the classic cone the return is radial in s and linear in I1:
    del_lambda = f_tr / (mu + 9 kappa b b_flow)."""
    a, b = 300.0, 0.05
    grad = make_grad(20000, 1)
    sig, tan, hist, pl = run(om.RustDruckerPrager3D, params(a, b, b_flow), grad)
    assert 0.3 < pl.mean() < 0.7
    sig_tr = C_apply(mandel(grad))
    i1, s, r = invariants(sig_tr)
    f_tr = r + b * i1 - a
    assert np.array_equal(pl, f_tr > 0)
    dl = f_tr / (MU + 9 * KA * b * b_flow)
    ref = s * ((r - dl * MU) / r)[:, None]
    ref[:, :3] += ((i1 - 9 * KA * b_flow * dl) / 3)[:, None]
    assert np.all((r - dl * MU)[pl] > 0)
    err = np.linalg.norm(sig[pl] - ref[pl], axis=1) / np.linalg.norm(ref[pl], axis=1)
    assert err.max() < 1e-12
    assert np.abs(sig[~pl] - sig_tr[~pl]).max() < 1e-10
    # elastic points: tangent is the elastic one, history untouched
    Cmat = np.stack([C_apply(np.eye(6)[k:k + 1])[0] for k in range(6)], axis=1)
    assert np.abs(tan[~pl] - Cmat).max() < 1e-9
    assert np.all(hist[~pl] == 0.0)
    # documented quirk of general.rs:208: alpha_1 = alpha_0 + sqrt(2/3) |g|, no del_lambda factor
    gnorm = np.sqrt(3 * b_flow**2 + 0.5)
    assert np.abs(hist[pl, 0] - np.sqrt(2 / 3) * gnorm).max() < 1e-9
    # plastic strain increment = de - C^-1 (sigma_1 - sigma_0) = del_lambda g
    dpl = mandel(grad) - Cinv_apply(sig)
    assert np.abs(hist[:, 1:] - dpl).max() < 1e-15


@pytest.mark.parametrize("b_flow", [0.05, 0.01])
def test_oracle_hyperbolic_return_properties(b_flow):
    a, b, d = 300.0, 0.05, 40.0
    grad = make_grad(20000, 2)
    sig0 = np.random.default_rng(3).standard_normal(20000 * 6) * 30.0
    sig, tan, hist, pl = run(om.RustDruckerPragerHyperbolic3D, params(a, b, b_flow, d), grad, sig0=sig0)
    assert 0.3 < pl.mean() < 0.7
    sig_tr = C_apply(mandel(grad)) + sig0.reshape(-1, 6)
    i1, s, r = invariants(sig, d * d)
    f1 = r + b * i1 - a
    assert np.abs(f1[pl]).max() < 1e-7      # on the yield surface (Newton atol 1e-8 on f)
    assert np.all(f1[~pl] <= 0)
    # return direction: C^-1 (sigma_tr - sigma_1) = del_lambda g(sigma_1), del_lambda > 0
    g1 = 0.5 * s / r[:, None]
    g1[:, :3] += b_flow
    dep = Cinv_apply(sig_tr - sig)
    dl = (dep * g1).sum(1) / (g1 * g1).sum(1)
    assert np.all(dl[pl] > 0)
    resid = np.linalg.norm(dep - dl[:, None] * g1, axis=1)
    assert resid[pl].max() < 1e-8 * np.linalg.norm(dep[pl], axis=1).max()


@pytest.mark.parametrize("cls,prm", [
    (om.RustDruckerPrager3D, params(300.0, 0.05)),
    (om.RustDruckerPrager3D, params(300.0, 0.05, 0.01)),
    (om.RustDruckerPragerHyperbolic3D, params(300.0, 0.05, None, 40.0)),
    (om.RustDruckerPragerHyperbolic3D, params(300.0, 0.05, 0.0, 40.0)),
], ids=["classic_assoc", "classic_nonassoc", "hyper_assoc", "hyper_nonassoc"])
def test_oracle_tangent_is_consistent(cls, prm):
    """tangent[i][j] = d sigma_i / d eps_j (the transposed product of general.rs:255-262 as stored
    by the bindings) against central differences of the stress update."""
    n = 200
    e = mandel(make_grad(n, 5))
    sig0 = np.random.default_rng(6).standard_normal(n * 6) * 20.0
    sig, tan, hist, pl = run(cls, prm, mandel_to_grad(e), sig0=sig0)
    assert pl.sum() > 40
    h = 1e-7
    fd = np.zeros((n, 6, 6))
    for j in range(6):
        ep, em = e.copy(), e.copy()
        ep[:, j] += h
        em[:, j] -= h
        sp = run(cls, prm, mandel_to_grad(ep), sig0=sig0)[0]
        sm = run(cls, prm, mandel_to_grad(em), sig0=sig0)[0]
        fd[:, :, j] = (sp - sm) / (2 * h)
    err = np.linalg.norm((tan - fd).reshape(n, -1), axis=1) / np.linalg.norm(fd.reshape(n, -1), axis=1)
    assert err.max() < 2e-6
    if float(prm["b"][0]) != float(prm["b_flow"][0]):  # non-associated flow: unsymmetric tangent
        asym = np.abs(tan[pl] - tan[pl].transpose(0, 2, 1)).max()
        assert asym > 1.0


def test_oracle_classic_apex_raises():
    """drucker_prager_classic.rs:86 asserts i_1 < a/b; the Rust code panics there."""
    e = np.zeros((1, 6))
    e[0, :3] = 0.02  # I1_tr = 9 kappa * 0.02 >> a/b
    with pytest.raises(RuntimeError):
        run(om.RustDruckerPrager3D, params(300.0, 0.05), mandel_to_grad(e))
    run(om.RustDruckerPragerHyperbolic3D, params(300.0, 0.05, None, 40.0), mandel_to_grad(e))  # smooth tip: fine


# ----------------------------------------------------------------------------- CUDA vs oracle (GPU)

CASES = [
    ("classic_assoc", "DruckerPrager3D", om.RustDruckerPrager3D, params(300.0, 0.05)),
    ("classic_nonassoc", "DruckerPrager3D", om.RustDruckerPrager3D, params(300.0, 0.05, 0.01)),
    ("hyper_assoc", "DruckerPragerHyperbolic3D", om.RustDruckerPragerHyperbolic3D, params(300.0, 0.05, None, 40.0)),
    ("hyper_nonassoc", "DruckerPragerHyperbolic3D", om.RustDruckerPragerHyperbolic3D, params(300.0, 0.05, 0.02, 40.0)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 129, 4097, 200_003])
@pytest.mark.parametrize("name,gcls,ocls,prm", CASES, ids=[c[0] for c in CASES])
def test_gpu_drucker_prager_vs_oracle(name, gcls, ocls, prm, n):
    import torch

    from fenics_constitutive_b200 import models as M
    from _util import TOL_PLASTIC, assert_close

    grad, grad2 = make_grad(n, 100 + n), make_grad(n, 200 + n) * 0.5
    orc = ocls(prm)
    orc.nthreads = 8
    law = getattr(M, gcls)(prm)
    law.record_plastic_flag = True
    assert law.history_dim == {"history": 7} and law.constraint == M.StressStrainConstraint.FULL
    ref = [np.zeros(n * 6), np.zeros(n * 36), np.zeros(n * 7)]
    host = [np.zeros(n * 6), np.full(n * 36, np.nan), np.zeros(n * 7)]
    dev = [torch.zeros(n * 6, dtype=torch.float64, device="cuda"),
           torch.full((n * 36,), float("nan"), dtype=torch.float64, device="cuda"),
           torch.zeros(n * 7, dtype=torch.float64, device="cuda")]
    for step in range(2):  # the second increment starts from a stressed state with history
        g = grad if step == 0 else grad2  # independent second increment: unloading and further yielding
        orc.evaluate(0.0, 1.0, g, ref[0], ref[1], {"history": ref[2]})
        if n >= 4097:
            assert 0.1 < orc.plastic_flag.mean() < 0.9, f"step {step}"
        law.evaluate(0.0, 1.0, g, host[0], host[1], {"history": host[2]})
        assert np.array_equal(law.plastic_flag, orc.plastic_flag), f"classification host step {step}"
        law.evaluate(0.0, 1.0, torch.from_numpy(g).cuda(), dev[0], dev[1], {"history": dev[2]})
        assert np.array_equal(law.plastic_flag.cpu().numpy(), orc.plastic_flag), f"classification device step {step}"
        for label, got in (("host", host), ("device", [t.cpu().numpy() for t in dev])):
            assert_close(got[0], ref[0], 6, TOL_PLASTIC, f"stress {label} step {step}")
            assert_close(got[1], ref[1], 36, TOL_PLASTIC, f"tangent {label} step {step}")
            assert_close(got[2], ref[2], 7, TOL_PLASTIC, f"history {label} step {step}")


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 0], ids=["shipped", "reference_spelling"])
@pytest.mark.parametrize("name,gcls,ocls,prm", CASES, ids=[c[0] for c in CASES])
def test_gpu_drucker_prager_tangent_finite_difference(name, gcls, ocls, prm, variant):
    """The CUDA path itself: the dense tangent the tile kernel stores against central differences
    of the kernel's own stress update (stress-only calls), from a pre-stressed state, both spellings of the slow
    operations (fcx_tune dp_variant); non-associated flow gives an unsymmetric tangent."""
    import torch

    from fenics_constitutive_b200 import models as M
    from fenics_constitutive_b200._lib import lib

    n = 3000
    e = mandel(make_grad(n, 5))
    sig0 = np.random.default_rng(6).standard_normal(n * 6) * 20.0
    law = getattr(M, gcls)(prm)
    law.record_plastic_flag = True
    old = lib().fcx_tune(b"dp_variant", variant)
    try:
        def run(strain, with_tangent):
            g = torch.from_numpy(mandel_to_grad(strain)).cuda()
            sig = torch.from_numpy(sig0.copy()).cuda()
            hist = torch.zeros(n * 7, dtype=torch.float64, device="cuda")
            tan = torch.full((n * 36,), float("nan"), dtype=torch.float64, device="cuda") if with_tangent else None
            law.evaluate(0.0, 1.0, g, sig, tan, {"history": hist})
            return (sig.cpu().numpy().reshape(n, 6), tan.cpu().numpy().reshape(n, 6, 6) if with_tangent else None,
                    law.plastic_flag.cpu().numpy().astype(bool))

        _, tan, pl = run(e, True)
        assert pl.sum() > 600
        h = 1e-7
        fd = np.zeros((n, 6, 6))
        for j in range(6):
            ep, em = e.copy(), e.copy()
            ep[:, j] += h
            em[:, j] -= h
            fd[:, :, j] = (run(ep, False)[0] - run(em, False)[0]) / (2 * h)
    finally:
        lib().fcx_tune(b"dp_variant", old)
    # points whose classification flips inside [e - h, e + h] have no derivative there: leave them out
    steady = np.ones(n, dtype=bool)
    for j in range(6):
        for sgn in (1.0, -1.0):
            ej = e.copy()
            ej[:, j] += sgn * h
            steady &= run(ej, False)[2] == pl
    assert steady.mean() > 0.99
    err = np.linalg.norm((tan - fd).reshape(n, -1), axis=1) / np.linalg.norm(fd.reshape(n, -1), axis=1)
    assert err[steady].max() < 2e-6
    if float(prm["b"][0]) != float(prm["b_flow"][0]):
        assert np.abs(tan[pl] - tan[pl].transpose(0, 2, 1)).max() > 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1000, 20_000], ids=["plain_d2h", "download_wire"])
def test_gpu_drucker_prager_failure_reporting(n):
    """Points where the Rust code would panic (apex assert of the classic model) raise RuntimeError on
    both paths and keep their input stress / history; the other points are still updated.  (The host
    path sends n >= 4096 over the download wire of csrc/fcx_host.cu.)"""
    import torch

    from fenics_constitutive_b200 import models as M

    grad = make_grad(n, 9).reshape(n, 9)
    grad[17, [0, 4, 8]] = 0.02  # far beyond the apex
    grad = grad.ravel()
    prm = params(300.0, 0.05)
    law = M.DruckerPrager3D(prm)
    sig, tan, hist = np.zeros(n * 6), np.zeros(n * 36), np.zeros(n * 7)
    with pytest.raises(RuntimeError, match="1 point"):
        law.evaluate(0.0, 1.0, grad, sig, tan, {"history": hist})
    assert np.all(sig.reshape(n, 6)[17] == 0.0) and np.all(hist.reshape(n, 7)[17] == 0.0)
    assert np.abs(sig.reshape(n, 6)[16]).max() > 0.0
    d = [torch.from_numpy(grad).cuda()] + [torch.zeros(m, dtype=torch.float64, device="cuda") for m in (n * 6, n * 36, n * 7)]
    with pytest.raises(RuntimeError, match="first index 17"):
        law.evaluate(0.0, 1.0, d[0], d[1], d[2], {"history": d[3]})
    # the status word is reset: a clean batch passes afterwards
    ok = make_grad(n, 10)
    law.evaluate(0.0, 1.0, torch.from_numpy(ok).cuda(), d[1].zero_(), d[2], {"history": d[3].zero_()})
    with pytest.raises(ValueError):
        law.evaluate(0.0, 1.0, ok, sig, tan, None)
