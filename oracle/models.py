"""Reference-shaped classes over the C oracle (TEST INFRASTRUCTURE ONLY).

Same constructor arguments, attribute names, history keys and `evaluate`
signature as the reference's models (fc = src/fenics_constitutive):
  LinearElasticityModel  fc/models/linear_elasticity_model.py:10-53
  VonMises3D             fc/models/mises_plasticity_isotropic_hardening.py:9-186
  SpringKelvinModel      fc/models/spring_kelvin_model.py:9-99
  SpringMaxwellModel     fc/models/spring_maxwell_model.py:8-99
Constraints are passed as anything with a `.value` in 1..5 (or a plain int),
matching StressStrainConstraint (fc/models/interfaces.py:23-27).
All arrays are flat float64 numpy arrays mutated in place.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import lib

_dp = ctypes.POINTER(ctypes.c_double)


def _code(constraint) -> int:
    return int(getattr(constraint, "value", constraint))


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def sdim(constraint) -> int:
    return int(lib().oracle_stress_strain_dim(_code(constraint)))


def gdim(constraint) -> int:
    return int(lib().oracle_geometric_dim(_code(constraint)))


def strain_from_grad_u(grad_u: np.ndarray, constraint) -> np.ndarray:
    g = np.ascontiguousarray(grad_u, dtype=np.float64).reshape(-1)
    n = g.size // gdim(constraint) ** 2
    out = np.zeros(n * sdim(constraint))
    lib().oracle_strain_from_grad_u(_code(constraint), n, _p(g), _p(out))
    return out


def lame_parameters(E: float, nu: float):
    mu, lam = ctypes.c_double(), ctypes.c_double()
    lib().oracle_lame_parameters(E, nu, ctypes.byref(mu), ctypes.byref(lam))
    return mu.value, lam.value


def get_elastic_tangent(E: float, nu: float, constraint) -> np.ndarray:
    s = sdim(constraint)
    D = np.zeros(s * s)
    lib().oracle_get_elastic_tangent(E, nu, _code(constraint), _p(D))
    return D.reshape(s, s)


def get_identity(constraint) -> np.ndarray:
    I2 = np.zeros(sdim(constraint))
    lib().oracle_get_identity(_code(constraint), _p(I2))
    return I2


class _Base:
    nthreads = 1

    @property
    def stress_strain_dim(self):
        return sdim(self.constraint)

    @property
    def geometric_dim(self):
        return gdim(self.constraint)

    def _check_sizes(self, grad_del_u, stress, tangent):
        g, s = self.geometric_dim, self.stress_strain_dim
        assert grad_del_u.size // g**2 == stress.size // s == tangent.size // s**2
        return grad_del_u.size // g**2


class LinearElasticityModel(_Base):
    def __init__(self, parameters, constraint):
        self.constraint = constraint
        self.D = get_elastic_tangent(parameters["E"], parameters["nu"], constraint)

    history_dim = None

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history=None):
        n = self._check_sizes(grad_del_u, stress, tangent)
        D = np.ascontiguousarray(self.D, dtype=np.float64)
        lib().oracle_elastic_evaluate(
            _code(self.constraint), _p(D), n, _p(grad_del_u), _p(stress), _p(tangent), self.nthreads
        )


class RustLinearElasticity3D(_Base):
    """comfe-rs LinearElasticity3D (comfe-rs/src/linear_elasticity.rs:49-74)."""

    def __init__(self, parameters):
        self.mu = float(np.asarray(parameters["mu"]).reshape(-1)[0])
        self.kappa = float(np.asarray(parameters["kappa"]).reshape(-1)[0])
        self.constraint = 5

    history_dim = None

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history=None):
        n = self._check_sizes(grad_del_u, stress, tangent)
        lib().oracle_rs_linear_elasticity3d(
            self.mu, self.kappa, n, _p(grad_del_u), _p(stress), _p(tangent)
        )


class RustMisesPlasticityLinearHardening3D(_Base):
    """comfe-rs MisesPlasticity3D (comfe-rs/src/mises_plasticity.rs:58-126); history is the single
    key "history" [n][7] = [alpha, plastic_strain[6]] (bindings/src/lib.rs:90-100,131-136)."""

    def __init__(self, parameters):
        get = lambda k: float(np.asarray(parameters[k]).reshape(-1)[0])  # noqa: E731
        self.params = np.array([get("mu"), get("kappa"), get("y_0"), get("h")], dtype=np.float64)
        self.constraint = 5
        self.plastic_flag = None

    @property
    def history_dim(self):
        return {"history": 7}

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history):
        n = self._check_sizes(grad_del_u, stress, tangent)
        self.plastic_flag = np.zeros(n, dtype=np.uint8)
        lib().oracle_rs_mises_linear_hardening(
            _p(self.params), n, _p(grad_del_u), _p(stress), _p(tangent), _p(history["history"]),
            self.plastic_flag.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
        )


class _RustDruckerPrager(_Base):
    """comfe-rs IsotropicPlasticityModel3D<DruckerPrager...> (comfe-rs/src/plasticity/general.rs:105-266)."""

    _keys: tuple = ()
    _hyperbolic = 0

    def __init__(self, parameters):
        get = lambda k: float(np.asarray(parameters[k]).reshape(-1)[0])  # noqa: E731
        self.params = np.array([get(k) for k in self._keys], dtype=np.float64)
        self.constraint = 5
        self.plastic_flag = None
        self.nthreads = 1

    @property
    def history_dim(self):
        return {"history": 7}

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history):
        n = self._check_sizes(grad_del_u, stress, tangent)
        self.plastic_flag = np.zeros(n, dtype=np.uint8)
        rc = lib().oracle_rs_drucker_prager(
            self._hyperbolic, _p(self.params), n, _p(grad_del_u), _p(stress), _p(tangent), _p(history["history"]),
            self.plastic_flag.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), int(self.nthreads),
        )
        if rc > 0:  # the Rust code panics (general.rs:186,236; drucker_prager_classic.rs:86)
            raise RuntimeError(f"Plasticity3D: Newton-Raphson did not converge ({rc} point(s))")


class RustDruckerPrager3D(_RustDruckerPrager):
    """drucker_prager_classic.rs:46-166; params mu, kappa, a, b, b_flow."""

    _keys = ("mu", "kappa", "a", "b", "b_flow")
    _hyperbolic = 0


class RustDruckerPragerHyperbolic3D(_RustDruckerPrager):
    """drucker_prager_hyperbolic.rs:48-164; params mu, kappa, a, b, d, b_flow."""

    _keys = ("mu", "kappa", "a", "b", "d", "b_flow")
    _hyperbolic = 1


class VonMises3D(_Base):
    def __init__(self, param):
        self.constraint = 5
        self.p_ka = param["p_ka"]
        self.p_mu = param["p_mu"]
        self.p_y0 = param["p_y0"]
        self.p_y00 = param["p_y00"]
        self.p_w = param["p_w"]
        self.plastic_flag = None  # filled by evaluate (phitr > 0 per QP)

    @property
    def history_dim(self):
        return {"eps_n": 6, "alpha": 1}

    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history):
        n = self._check_sizes(grad_del_u, stress, tangent)
        params = np.array([self.p_ka, self.p_mu, self.p_y0, self.p_y00, self.p_w], dtype=np.float64)
        flag = np.zeros(n, dtype=np.uint8)
        failed = lib().oracle_mises_evaluate(
            _p(params),
            n,
            _p(grad_del_u),
            _p(stress),
            _p(tangent),
            _p(history["eps_n"]),
            _p(history["alpha"]),
            flag.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
            self.nthreads,
        )
        self.plastic_flag = flag
        if failed:
            raise RuntimeError("Newton-Raphson method did not converge for plastic multiplier.")


class _Visco(_Base):
    def __init__(self, parameters, constraint):
        self.constraint = constraint
        self.E0 = parameters["E0"]
        self.E1 = parameters["E1"]
        self.tau = parameters["tau"]
        self.nu = 0.0 if _code(constraint) == 2 else parameters["nu"]
        self.D_0 = get_elastic_tangent(self.E0, self.nu, constraint)
        self.D_1 = get_elastic_tangent(self.E1, self.nu, constraint)
        self.I2 = get_identity(constraint)
        self.mu0, self.lam0 = lame_parameters(self.E0, self.nu)
        self.mu1, _ = lame_parameters(self.E1, self.nu)

    @property
    def history_dim(self):
        s = self.stress_strain_dim
        return {"strain_visco": s, "strain": s}


class SpringKelvinModel(_Visco):
    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history):
        n = self._check_sizes(grad_del_u, stress, tangent)
        if history is None:
            raise ValueError("history must not be None")
        assert del_t > 0, "Time step must be defined and positive."
        rc = lib().oracle_kelvin_evaluate(
            _code(self.constraint), _p(np.ascontiguousarray(self.D_0)), _p(self.I2),
            self.mu0, self.lam0, self.mu1, self.tau, del_t, n,
            _p(grad_del_u), _p(stress), _p(tangent),
            _p(history["strain_visco"]), _p(history["strain"]), self.nthreads,
        )
        assert rc == 0


class SpringMaxwellModel(_Visco):
    def evaluate(self, t, del_t, grad_del_u, stress, tangent, history):
        n = self._check_sizes(grad_del_u, stress, tangent)
        if history is None:
            raise ValueError("history must not be None")
        assert del_t > 0, "Time step must be defined and positive."
        rc = lib().oracle_maxwell_evaluate(
            _code(self.constraint), _p(np.ascontiguousarray(self.D_0)),
            _p(np.ascontiguousarray(self.D_1)), self.mu1, self.tau, del_t, n,
            _p(grad_del_u), _p(stress), _p(tangent),
            _p(history["strain_visco"]), _p(history["strain"]), self.nthreads,
        )
        assert rc == 0


def gather_grad(gdim_, dofmap, u, u_prev, dphi_ref, Jinv):
    """oracle_gather_grad wrapper: returns flat grad [ncells*nq*g*g]."""
    dofmap = np.ascontiguousarray(dofmap, dtype=np.int32)
    ncells, nd = dofmap.shape
    dphi_ref = np.ascontiguousarray(dphi_ref, dtype=np.float64)
    nq = dphi_ref.shape[0]
    assert dphi_ref.shape == (nq, nd, gdim_)
    Jinv = np.ascontiguousarray(Jinv, dtype=np.float64)
    out = np.zeros(ncells * nq * gdim_ * gdim_)
    lib().oracle_gather_grad(
        gdim_, ncells, nq, nd,
        dofmap.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
        _p(np.ascontiguousarray(u)), _p(np.ascontiguousarray(u_prev)) if u_prev is not None else None,
        _p(dphi_ref), _p(Jinv), _p(out),
    )
    return out
