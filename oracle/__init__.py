"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (parity pinned, see fcx_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package.  The product package
(fenics_constitutive_b200) never does.

`oracle.models` exposes the C restatement (oracle/fcx_oracle.c) behind classes
with the reference's names and `evaluate` signature, so parity tests read like
the reference's own tests.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/fcx_oracle.c -> oracle/liboracle.so (gcc, see Makefile)."""
    src = os.path.join(_HERE, "fcx_oracle.c")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        u8p = ctypes.POINTER(ctypes.c_ubyte)
        sz, ci, cd = ctypes.c_size_t, ctypes.c_int, ctypes.c_double
        L.oracle_max_threads.restype = ci
        L.oracle_has_openmp.restype = ci
        L.oracle_strain_from_grad_u.argtypes = [ci, sz, dp, dp]
        L.oracle_strain_from_grad_u.restype = None
        L.oracle_lame_parameters.argtypes = [cd, cd, dp, dp]
        L.oracle_lame_parameters.restype = None
        L.oracle_get_elastic_tangent.argtypes = [cd, cd, ci, dp]
        L.oracle_get_elastic_tangent.restype = None
        L.oracle_get_identity.argtypes = [ci, dp]
        L.oracle_get_identity.restype = None
        L.oracle_elastic_evaluate.argtypes = [ci, dp, sz, dp, dp, dp, ci]
        L.oracle_elastic_evaluate.restype = ci
        L.oracle_rs_linear_elasticity3d.argtypes = [cd, cd, sz, dp, dp, dp]
        L.oracle_rs_linear_elasticity3d.restype = ci
        L.oracle_rs_mises_linear_hardening.argtypes = [dp, sz, dp, dp, dp, dp, u8p]
        L.oracle_rs_mises_linear_hardening.restype = ci
        L.oracle_rs_drucker_prager.argtypes = [ci, dp, sz, dp, dp, dp, dp, u8p, ci]
        L.oracle_rs_drucker_prager.restype = ci
        L.oracle_mises_evaluate.argtypes = [dp, sz, dp, dp, dp, dp, dp, u8p, ci]
        L.oracle_mises_evaluate.restype = ci
        L.oracle_kelvin_evaluate.argtypes = [ci, dp, dp, cd, cd, cd, cd, cd, sz, dp, dp, dp, dp, dp, ci]
        L.oracle_kelvin_evaluate.restype = ci
        L.oracle_maxwell_evaluate.argtypes = [ci, dp, dp, cd, cd, cd, sz, dp, dp, dp, dp, dp, ci]
        L.oracle_maxwell_evaluate.restype = ci
        L.oracle_gather_grad.argtypes = [ci, sz, ci, ci, ip, dp, dp, dp, dp, dp]
        L.oracle_gather_grad.restype = None
        _lib = L
    return _lib


def max_threads() -> int:
    """Host threads the oracle may use: the CPUs this process is allowed to run on
    (OMP_NUM_THREADS is often pinned to 1 in ML images; `num_threads(n)` in
    fcx_oracle.c overrides it).  1 if the library was built without OpenMP."""
    import os

    if int(lib().oracle_has_openmp()) == 0:
        return 1
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)
