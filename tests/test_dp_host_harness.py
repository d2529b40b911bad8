"""The PRODUCT's Drucker-Prager point update compiled for the host (tests/host_harness/dp_host.cu instantiates
DruckerPragerModel<HYP>::qp / entry of csrc/fcx_models.cuh, the very functions the CUDA tile kernel runs per
quadrature point) against the C restatement of comfe-rs/src/plasticity/general.rs:105-266 -- a check of the
kernel's own arithmetic (structured Newton step, stop rule, tangent record and its expansion) that needs no GPU.
The harness is test infrastructure: it is never part of libfcx.so.  PARITY UNPINNED against the reference
(tests/test_drucker_prager.py says why); tolerance 1e-10, identical classification."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from _util import TOL_PLASTIC, assert_close
from test_drucker_prager import CASES, make_grad, mandel, mandel_to_grad

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_harness")
SO = os.path.join(HERE, "libdp_host.so")
SRC = os.path.join(HERE, "dp_host.cu")
CSRC = os.path.join(os.path.dirname(HERE), os.pardir, "fenics_constitutive_b200", "csrc")


def _build():
    import shutil

    deps = [SRC] + [os.path.join(CSRC, f) for f in ("fcx_models.cuh", "fcx_tile.cuh", "fcx_ptx.cuh")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return
    if shutil.which("nvcc") is None:
        if os.path.exists(SO):
            return  # a prebuilt harness travelled with the tree; nothing to rebuild it with
        pytest.skip("nvcc not found: the host harness of the Drucker-Prager point update cannot be built")
    cmd = ["nvcc", "-O2", "-std=c++17", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "-o", SO, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.fixture(scope="module")
def harness():
    _build()
    L = ctypes.CDLL(SO)
    dp = ctypes.POINTER(ctypes.c_double)
    L.dp_host_evaluate.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_size_t, dp, dp, dp, dp, ctypes.POINTER(ctypes.c_ubyte)]
    L.dp_host_evaluate.restype = ctypes.c_int
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@pytest.mark.parametrize("variant", [0, 1], ids=["reference_spelling", "shipped"])
@pytest.mark.parametrize("with_tangent", [True, False], ids=["tangent", "stress_only"])
@pytest.mark.parametrize("name,gcls,ocls,prm", CASES, ids=[c[0] for c in CASES])
def test_product_point_update_on_host_vs_oracle(harness, name, gcls, ocls, prm, with_tangent, variant):
    n = 60_001
    grad, grad2 = make_grad(n, 100 + n), make_grad(n, 200 + n) * 0.5
    orc = ocls(prm)
    orc.nthreads = 8
    hyp = int("d" in prm)
    keys = ("mu", "kappa", "a", "b", "d", "b_flow") if hyp else ("mu", "kappa", "a", "b", "b_flow")
    pv = np.array([float(prm[k][0]) for k in keys])
    ref = [np.zeros(n * 6), np.zeros(n * 36), np.zeros(n * 7)]
    got = [np.zeros(n * 6), np.full(n * 36, np.nan), np.zeros(n * 7)]
    flag = np.zeros(n, dtype=np.uint8)
    for step in range(2):  # the second increment starts from a stressed state with history
        g = grad if step == 0 else grad2
        orc.evaluate(0.0, 1.0, g, ref[0], ref[1], {"history": ref[2]})
        assert 0.1 < orc.plastic_flag.mean() < 0.9
        rc = harness.dp_host_evaluate(hyp, variant, _p(pv), n, _p(g), _p(got[0]), _p(got[1]) if with_tangent else None,
                                      _p(got[2]), flag.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
        assert rc == 0
        assert np.array_equal(flag, orc.plastic_flag), f"classification step {step}"
        assert_close(got[0], ref[0], 6, TOL_PLASTIC, f"stress step {step}")
        assert_close(got[2], ref[2], 7, TOL_PLASTIC, f"history step {step}")
        if with_tangent:
            assert_close(got[1], ref[1], 36, TOL_PLASTIC, f"tangent step {step}")


def test_product_point_update_on_host_apex_failure(harness):
    """A point beyond the apex of the classic cone (the Rust code's assert) is reported and left untouched."""
    name, gcls, ocls, prm = CASES[0]
    n = 100
    grad = make_grad(n, 9).reshape(n, 9)
    grad[17, [0, 4, 8]] = 0.02
    grad = grad.ravel()
    pv = np.array([float(prm[k][0]) for k in ("mu", "kappa", "a", "b", "b_flow")])
    sig, tan, hist = np.zeros(n * 6), np.zeros(n * 36), np.zeros(n * 7)
    rc = harness.dp_host_evaluate(0, 1, _p(pv), n, _p(grad), _p(sig), _p(tan), _p(hist), None)
    assert rc == 1
    assert np.all(sig.reshape(n, 6)[17] == 0.0) and np.all(hist.reshape(n, 7)[17] == 0.0)
    assert np.abs(sig.reshape(n, 6)[16]).max() > 0.0


@pytest.mark.parametrize("variant", [0, 1], ids=["reference_spelling", "shipped"])
@pytest.mark.parametrize("name,gcls,ocls,prm", CASES, ids=[c[0] for c in CASES])
def test_product_tangent_is_consistent_on_host(harness, name, gcls, ocls, prm, variant):
    """tangent[i][j] = d sigma_i / d eps_j of the PRODUCT's point update (record + entry(), what the tile kernel
    expands into the dense 6x6 block) against central differences of its own stress update -- no oracle involved."""
    n = 200
    e = mandel(make_grad(n, 5))
    sig0 = np.random.default_rng(6).standard_normal(n * 6) * 20.0
    hyp = int("d" in prm)
    keys = ("mu", "kappa", "a", "b", "d", "b_flow") if hyp else ("mu", "kappa", "a", "b", "b_flow")
    pv = np.array([float(prm[k][0]) for k in keys])

    def run(strain, with_tangent):
        g = mandel_to_grad(strain)
        sig, hist, tan = sig0.copy(), np.zeros(n * 7), np.full(n * 36, np.nan)
        flag = np.zeros(n, dtype=np.uint8)
        rc = harness.dp_host_evaluate(hyp, variant, _p(pv), n, _p(g), _p(sig), _p(tan) if with_tangent else None,
                                      _p(hist), flag.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
        assert rc == 0
        return sig.reshape(n, 6), tan.reshape(n, 6, 6), flag.astype(bool)

    _, tan, pl = run(e, True)
    assert pl.sum() > 40
    h = 1e-7
    fd = np.zeros((n, 6, 6))
    for j in range(6):
        ep, em = e.copy(), e.copy()
        ep[:, j] += h
        em[:, j] -= h
        fd[:, :, j] = (run(ep, False)[0] - run(em, False)[0]) / (2 * h)
    err = np.linalg.norm((tan - fd).reshape(n, -1), axis=1) / np.linalg.norm(fd.reshape(n, -1), axis=1)
    assert err.max() < 2e-6
    if float(prm["b"][0]) != float(prm["b_flow"][0]):  # non-associated flow: unsymmetric tangent
        assert np.abs(tan[pl] - tan[pl].transpose(0, 2, 1)).max() > 1.0
