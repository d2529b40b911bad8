"""Import the reference's UNMODIFIED Python models in the build container.

ORACLE / TEST INFRASTRUCTURE ONLY.  `/root/reference` exists only in the build
container (not on the GPU box), so this module is used solely
  * by oracle/gen_golden.py to generate tests/golden/ fixtures, and
  * by `-m "not gpu"` tests that pin the C oracle live against the reference
    (skipped when the reference tree is absent).

The reference's `models` package cannot be imported as-is because
`models/utils.py:3` imports dolfinx (used only for `@df.common.timed`
decorators, utils.py:279-294,365-393) and `models/__init__.py:11` imports the
compiled pyo3 module `fenics_constitutive._bindings` (rust_models.py:5-10).
Two stub modules satisfy those imports; the model arithmetic runs untouched.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FCX_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "fenics_constitutive", "models"))


def load():
    """Return the reference's `fenics_constitutive.models` module."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if "dolfinx" not in sys.modules:
        df = types.ModuleType("dolfinx")
        common = types.ModuleType("dolfinx.common")
        common.timed = lambda name: (lambda f: f)
        df.common = common
        sys.modules["dolfinx"] = df
        sys.modules["dolfinx.common"] = common
    if "fenics_constitutive._bindings" not in sys.modules:
        b = types.ModuleType("fenics_constitutive._bindings")
        for name in (
            "PyLinearElasticity3D",
            "PyMisesPlasticity3D",
            "PyDruckerPrager3D",
            "PyDruckerPragerHyperbolic3D",
        ):
            setattr(b, name, type(name, (), {}))
        sys.modules["fenics_constitutive._bindings"] = b
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    return importlib.import_module("fenics_constitutive.models")
