"""CPU tests of the host-side mirror: interface, constants, argument checks and
the C-ABI surface.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from fenics_constitutive_b200 import _lib
from fenics_constitutive_b200.models import (
    IncrSmallStrainModel,
    LinearElasticityModel,
    SpringKelvinModel,
    SpringMaxwellModel,
    StressStrainConstraint,
    VonMises3D,
)
from fenics_constitutive_b200.models.utils import get_elastic_tangent, get_identity, lame_parameters
from fenics_constitutive_b200.partition import shard_range

from _util import CONSTRAINT_NAMES, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C = StressStrainConstraint


def test_constraint_enum_matches_reference():
    """reference models/interfaces.py:23-27,30-73"""
    assert [c.name for c in C] == CONSTRAINT_NAMES
    assert [c.value for c in C] == [1, 2, 3, 4, 5]
    assert [c.stress_strain_dim for c in C] == [1, 1, 4, 4, 6]
    assert [c.geometric_dim for c in C] == [1, 1, 2, 2, 3]


@pytest.mark.parametrize("name", CONSTRAINT_NAMES)
def test_constants_bit_identical_to_reference(name):
    d = golden("conversions.npz")
    c = C[name]
    assert np.array_equal(get_elastic_tangent(42.0, 0.3, c), d[f"{name}_D"])
    assert np.array_equal(get_identity(c.stress_strain_dim, c), d[f"{name}_I2"])
    assert np.array_equal(np.array(lame_parameters(42.0, 0.3)), d["lame_42_0.3"])


def test_model_surface():
    el = LinearElasticityModel({"E": 42.0, "nu": 0.3}, C.PLANE_STRESS)
    assert isinstance(el, IncrSmallStrainModel)
    assert el.history_dim is None and el.constraint is C.PLANE_STRESS
    assert el.stress_strain_dim == 4 and el.geometric_dim == 2 and el.D.shape == (4, 4)
    vm = VonMises3D({"p_ka": 1.0, "p_mu": 2.0, "p_y0": 3.0, "p_y00": 4.0, "p_w": 5.0})
    assert vm.constraint is C.FULL and vm.history_dim == {"eps_n": 6, "alpha": 1}
    assert (vm.p_ka, vm.p_mu, vm.p_y0, vm.p_y00, vm.p_w) == (1.0, 2.0, 3.0, 4.0, 5.0)
    assert np.allclose(vm.xpp @ vm.xpp, vm.xpp) and np.array_equal(vm.I2, [1, 1, 1, 0, 0, 0])
    for cls in (SpringKelvinModel, SpringMaxwellModel):
        v = cls({"E0": 42.0, "E1": 10.0, "tau": 10.0, "nu": 0.2}, C.FULL)
        assert v.history_dim == {"strain_visco": 6, "strain": 6}
        u = cls({"E0": 42.0, "E1": 10.0, "tau": 10.0}, C.UNIAXIAL_STRESS)  # nu forced to 0
        assert u.nu == 0.0 and u.D_0[0, 0] == 42.0


def test_argument_errors_match_reference():
    el = LinearElasticityModel({"E": 42.0, "nu": 0.3}, C.FULL)
    with pytest.raises(AssertionError):  # linear_elasticity_model.py:36-40
        el.evaluate(0.0, 1.0, np.zeros(18), np.zeros(6), np.zeros(72), None)
    kv = SpringKelvinModel({"E0": 42.0, "E1": 10.0, "tau": 10.0, "nu": 0.2}, C.FULL)
    with pytest.raises(ValueError):  # spring_kelvin_model.py:63-65
        kv.evaluate(0.0, 1.0, np.zeros(9), np.zeros(6), np.zeros(36), None)
    h = {"strain_visco": np.zeros(6), "strain": np.zeros(6)}
    with pytest.raises(AssertionError):  # spring_kelvin_model.py:72
        kv.evaluate(0.0, 0.0, np.zeros(9), np.zeros(6), np.zeros(36), h)
    mx = SpringMaxwellModel({"E0": 42.0, "E1": 10.0, "tau": 10.0, "nu": 0.2}, C.FULL)
    with pytest.raises(AssertionError):
        mx.evaluate(0.0, -1.0, np.zeros(9), np.zeros(6), np.zeros(36), h)
    with pytest.raises(TypeError):
        el.evaluate(0.0, 1.0, np.zeros(9, dtype=np.float32), np.zeros(6), np.zeros(36), None)
    with pytest.raises(ValueError):
        el.evaluate(0.0, 1.0, np.zeros(18)[::2], np.zeros(6), np.zeros(36), None)


def test_abi_exports_every_declared_symbol():
    """libfcx.so loads and exports exactly what include/fcx.h declares."""
    header = open(os.path.join(ROOT, "include", "fcx.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fcx_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in fcx.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table out of sync with fcx.h"
    lib = _lib.lib()
    assert lib.fcx_version() == 200
    assert [lib.fcx_stress_strain_dim(c) for c in range(0, 7)] == [-1, 1, 1, 4, 4, 6, -1]
    assert [lib.fcx_geometric_dim(c) for c in range(0, 7)] == [-1, 1, 1, 2, 2, 3, -1]
    assert lib.fcx_strerror(-2).decode() == "Time step must be defined and positive."
    assert "Newton-Raphson" in lib.fcx_strerror(3).decode()


def test_no_cpu_fallback_when_library_missing(monkeypatch, tmp_path):
    """The product fails loudly without libfcx.so."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libfcx.so"))
    el = LinearElasticityModel({"E": 42.0, "nu": 0.3}, C.FULL)
    with pytest.raises(_lib.FcxLibraryError):
        el.evaluate(0.0, 1.0, np.zeros(9), np.zeros(6), np.zeros(36), None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fenics_constitutive_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (10, 3), (16_000_000, 8), (1_000_003, 4)])
def test_shard_range_partitions_exactly(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b
    assert all(lo % 2 == 0 for lo, _ in spans)
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 3


def test_solver_maps_surface_matches_reference():
    """The names of the reference's solver/maps.py:11 exist here and both maps satisfy its SpaceMap protocol."""
    from fenics_constitutive_b200.solver import maps

    assert sorted(maps.__all__) == ["IdentityMap", "SpaceMap", "SubSpaceMap", "build_subspace_map"]
    assert isinstance(maps.IdentityMap(), maps.SpaceMap)
    assert issubclass(maps.SubSpaceMap, maps.SpaceMap)
