"""Deterministic synthetic inputs for the benchmark configurations of
BASELINE.json: parameters follow the reference's own tests.

Host (numpy) generators are used by the tests and the CPU baseline; the torch
generators create the same *distribution* directly in HBM for 16M-point runs.
"""
from __future__ import annotations

import numpy as np

# reference tests/models/test_elasticity.py:22-23
ELASTIC_PARAMS = {"E": 42.0, "nu": 0.3}
# reference tests/models/test_plasticity.py:19-25
MISES_PARAMS = {"p_ka": 175000.0, "p_mu": 80769.0, "p_y0": 1200.0, "p_y00": 2500.0, "p_w": 200.0}
# reference tests/models/test_viscoelasticity.py:20-23
VISCO_PARAMS = {"E0": 42.0, "E1": 10.0, "tau": 10.0, "nu": 0.2}
# std of grad_del_u that makes ~50 % of virgin points plastic
MISES_GRAD_STD = 2.906e-3


def mises_inputs_numpy(n: int, seed: int = 1234, step_scale: float = 1.0):
    """Virgin-state Mises batch: (grad, stress, eps_n, alpha) flat float64."""
    rng = np.random.default_rng(seed)
    grad = rng.standard_normal(n * 9) * (MISES_GRAD_STD * step_scale)
    return grad, np.zeros(n * 6), np.zeros(n * 6), np.zeros(n)


def elastic_inputs_numpy(n: int, g: int, s: int, seed: int = 1234):
    rng = np.random.default_rng(seed)
    grad = rng.standard_normal(n * g * g) * 1e-3
    stress = rng.standard_normal(n * s) * 0.1
    return grad, stress


def visco_inputs_numpy(n: int, g: int, s: int, seed: int = 1234):
    rng = np.random.default_rng(seed)
    grad = rng.standard_normal(n * g * g) * 1e-4
    return grad, np.zeros(n * s), np.zeros(n * s), np.zeros(n * s)


def mises_inputs_torch(n: int, device, seed: int = 1234):
    """Same distribution as mises_inputs_numpy, generated in HBM."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    grad = torch.randn(n * 9, dtype=torch.float64, device=device, generator=gen) * MISES_GRAD_STD
    z = lambda m: torch.zeros(m, dtype=torch.float64, device=device)  # noqa: E731
    return grad, z(n * 6), z(n * 6), z(n)
