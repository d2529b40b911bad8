"""ctypes binding of libfcx.so -- the C ABI declared in include/fcx.h.

The product path has NO CPU fallback: if the CUDA library cannot be loaded the
import-time error is re-raised on first use, loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfcx.so")

_dp = ctypes.c_void_p  # double* (host or device address)
_sz = ctypes.c_size_t
_ci = ctypes.c_int
_cd = ctypes.c_double
_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/fcx.h one to one
SIGNATURES = {
    "fcx_version": (_ci, []),
    "fcx_strerror": (ctypes.c_char_p, [_ci]),
    "fcx_last_cuda_error": (ctypes.c_char_p, []),
    "fcx_stress_strain_dim": (_ci, [_ci]),
    "fcx_geometric_dim": (_ci, [_ci]),
    "fcx_set_device": (_ci, [_ci]),
    "fcx_elastic_evaluate": (_ci, [_ci, _dp, _sz, _dp, _dp, _dp, _vp]),
    "fcx_mises_evaluate": (_ci, [_dp, _sz, _dp, _dp, _dp, _dp, _dp, _ci, _vp, _vp, _vp]),
    "fcx_kelvin_evaluate": (_ci, [_ci, _dp, _dp, _cd, _cd, _cd, _cd, _cd, _sz, _dp, _dp, _dp, _dp, _dp, _vp]),
    "fcx_maxwell_evaluate": (_ci, [_ci, _dp, _dp, _cd, _cd, _cd, _sz, _dp, _dp, _dp, _dp, _dp, _vp]),
    "fcx_strain_from_grad_u": (_ci, [_ci, _sz, _dp, _dp, _vp]),
    "fcx_embed_3d": (_ci, [_ci, _sz, _dp, _dp, _dp, _dp, _vp]),
    "fcx_extract_from_3d": (_ci, [_ci, _sz, _dp, _dp, _dp, _dp, _vp]),
    "fcx_nodal_increment": (_ci, [_sz, _dp, _dp, _dp, _vp]),
    "fcx_gather_grad": (_ci, [_ci, _sz, _ci, _ci, _vp, _dp, _dp, _dp, _dp, _dp, _vp]),
    "fcx_mises_linear_hardening_evaluate": (_ci, [_dp, _sz, _dp, _dp, _dp, _dp, _vp, _vp]),
    "fcx_mises_linear_hardening_evaluate_host": (_ci, [_dp, _sz, _dp, _dp, _dp, _dp, _vp]),
    "fcx_drucker_prager_evaluate": (_ci, [_ci, _dp, _sz, _dp, _dp, _dp, _dp, _vp, _vp, _vp]),
    "fcx_drucker_prager_evaluate_host": (_ci, [_ci, _dp, _sz, _dp, _dp, _dp, _dp, _vp]),
    "fcx_map_rows_to_sub": (_ci, [_sz, _sz, _vp, _dp, _dp, _vp]),
    "fcx_map_rows_to_parent": (_ci, [_sz, _sz, _vp, _dp, _dp, _vp]),
    "fcx_mises_form": (_ci, [_dp, _sz, _vp, _ci, _ci, _vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                             _dp, _dp, _vp, _vp, _vp]),
    "fcx_tangent_apply_rec": (_ci, [_ci, _ci, _sz, _ci, _ci, _vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _vp, _vp]),
    "fcx_fe_stride": (_ci, [_ci]),
    "fcx_internal_force": (_ci, [_ci, _ci, _sz, _ci, _ci, _dp, _dp, _dp, _dp, _dp, _dp, _vp, _vp]),
    "fcx_tangent_apply": (_ci, [_ci, _ci, _sz, _ci, _ci, _vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _vp, _vp]),
    "fcx_tangent_diag": (_ci, [_ci, _ci, _sz, _ci, _ci, _dp, _dp, _dp, _dp, _dp, _dp, _vp, _vp]),
    "fcx_gather_sum": (_ci, [_ci, _sz, _vp, _vp, _dp, _dp, _cd, _cd, _vp]),
    "fcx_pcg_scratch_doubles": (_sz, []),
    "fcx_pcg_pap": (_ci, [_sz, _dp, _dp, _dp, _dp, _vp, _dp, _vp]),
    "fcx_pcg_update_xr": (_ci, [_sz, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _vp, _dp, _vp]),
    "fcx_pcg_update_p": (_ci, [_sz, _dp, _dp, _dp, _dp, _dp, _vp]),
    "fcx_krylov_create": (_ci, [_ci, _ci, _ci, _sz, _sz, _vp, _vp, _vp]),
    "fcx_krylov_connect": (_ci, [_vp, _vp]),
    "fcx_krylov_set_halo": (_ci, [_vp, _ci, _vp, _vp, _vp, _vp]),
    "fcx_krylov_set_operator": (_ci, [_vp, _ci, _ci, _sz, _ci, _ci, _vp, _dp, _dp, _dp, _dp, _dp, _dp, _vp, _vp, _vp, _sz]),
    "fcx_krylov_begin": (_ci, [_vp, _dp, _dp, _vp]),
    "fcx_krylov_iterate": (_ci, [_vp, _ci, _vp]),
    "fcx_krylov_status": (_ci, [_vp, _vp]),
    "fcx_krylov_solution": (_ci, [_vp, _dp, _vp]),
    "fcx_krylov_set_tolerance": (_ci, [_vp, _cd]),
    "fcx_krylov_snapshot": (_ci, [_vp, _ci, _vp]),
    "fcx_krylov_wait_snapshot": (_ci, [_vp, _ci, _vp]),
    "fcx_krylov_halo_update": (_ci, [_vp, _dp, _vp]),
    "fcx_krylov_destroy": (None, [_vp]),
    "fcx_elastic_evaluate_host": (_ci, [_ci, _dp, _sz, _dp, _dp, _dp]),
    "fcx_mises_evaluate_host": (_ci, [_dp, _sz, _dp, _dp, _dp, _dp, _dp, _vp]),
    "fcx_kelvin_evaluate_host": (_ci, [_ci, _dp, _dp, _cd, _cd, _cd, _cd, _cd, _sz, _dp, _dp, _dp, _dp, _dp]),
    "fcx_maxwell_evaluate_host": (_ci, [_ci, _dp, _dp, _cd, _cd, _cd, _sz, _dp, _dp, _dp, _dp, _dp]),
    "fcx_host_register": (_ci, [_vp, _sz]),
    "fcx_host_unregister": (_ci, [_vp]),
    "fcx_host_staging": (_ci, [_ci]),
    "fcx_host_threads": (_ci, [_ci]),
    "fcx_host_wire": (_ci, [_ci]),
    "fcx_host_wire_used": (_ci, []),
    "fcx_host_wire_mix": (_ci, [_ci]),
    "fcx_host_wire_mix_used": (_ci, []),
    "fcx_host_numa": (_ci, [_ci]),
    "fcx_host_numa_info": (_ci, [_vp, _ci]),
    "fcx_diag_host_bandwidth": (_ci, [_ci, _sz, _vp, _ci]),
    "fcx_diag_pcie": (_ci, [_sz, _vp, _ci]),
    "fcx_host_trace": (_ci, [_ci]),
    "fcx_host_slots": (_ci, [_ci]),
    "fcx_host_debug_skip": (_ci, [_ci]),
    "fcx_host_timeline": (_ci, [_vp, _ci]),
    "fcx_host_stats": (_ci, [_vp, _ci]),
    "fcx_host_chunk_qps": (_sz, [_sz]),
    "fcx_host_release": (None, []),
    "fcx_launch_count": (ctypes.c_ulonglong, []),
    "fcx_tune": (_ci, [ctypes.c_char_p, _ci]),
    "fcx_diag_dfma_peak": (_cd, []),
    "fcx_diag_stream_mix": (ctypes.c_longlong, [_dp, _dp, _sz, _vp]),
}

_lib = None


class FcxLibraryError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load libfcx.so (once).  Raises FcxLibraryError if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FcxLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C fenics_constitutive_b200/csrc`. "
                "fenics_constitutive_b200 has no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing: loud
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "fcx") -> int:
    """Map a negative C status to a Python exception (SURVEY.md 8b 'Errors')."""
    if rc >= 0:
        return rc
    L = lib()
    msg = L.fcx_strerror(rc).decode()
    if rc == -2:  # FCX_ERR_TIMESTEP: the reference asserts (spring_kelvin_model.py:72)
        raise AssertionError(msg)
    if rc == -4:
        raise RuntimeError(f"{what}: {msg}: {L.fcx_last_cuda_error().decode()}")
    raise ValueError(f"{what}: {msg}")
