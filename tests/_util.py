"""Shared helpers for the parity tests."""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONSTRAINT_NAMES = ["UNIAXIAL_STRAIN", "UNIAXIAL_STRESS", "PLANE_STRAIN", "PLANE_STRESS", "FULL"]

# tolerances stated by BASELINE.json north_star
TOL_ELASTIC = 1e-12  # elasticity and viscoelasticity, relative
TOL_PLASTIC = 1e-10  # plastic stress and tangent, relative


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


def rel_err(x: np.ndarray, ref: np.ndarray, dim: int) -> float:
    """Worst per-QP norm-wise relative error  max_q |x_q - ref_q|_2 / max(|ref_q|_2, tiny)
    (element-wise relative error is ill-posed on the
    structural zeros of the tangent)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, dim)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, dim)
    assert x.shape == ref.shape
    if x.size == 0:
        return 0.0
    num = np.linalg.norm(x - ref, axis=1)
    den = np.maximum(np.linalg.norm(ref, axis=1), 1e-300)
    scale = max(float(np.max(den)), 1e-300)
    # points whose reference norm is ~0 are judged against the batch scale
    den = np.maximum(den, 1e-6 * scale)
    return float(np.max(num / den))


def assert_close(x, ref, dim, tol, what=""):
    e = rel_err(x, ref, dim)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"
