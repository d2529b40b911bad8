#!/usr/bin/env python
"""Secondary benchmark: device-resident throughput and HBM-roofline fraction of
EVERY kernel on the hot path (BASELINE.json configs 2-4 + the gather), one JSON
line per case.  The headline Mises number is bench.py's; this script fills the
per-kernel table of DESIGN.md / profiles/.

    python scripts/bench_models.py [--qps 16000000] [--steps 10] [--out file]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200.models import (  # noqa: E402
    LinearElasticityModel, SpringKelvinModel, SpringMaxwellModel, StressStrainConstraint, VonMises3D)

C = StressStrainConstraint


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_steps(fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qps", type=int, default=16_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    n, K = args.qps, args.steps
    dev = torch.device("cuda", 0)
    P = peak()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    rows = []

    def rnd(m, scale):
        return torch.randn(m, dtype=torch.float64, device=dev, generator=gen) * scale

    def report(name, bytes_per_qp, ms, extra=None):
        gbs = bytes_per_qp * n / (ms * 1e-3) / 1e9
        row = {"kernel": name, "qps": n, "ms": ms, "qp_per_s": n / (ms * 1e-3),
               "bytes_per_qp": bytes_per_qp, "GBps": gbs, "frac_of_measured_hbm": gbs / P}
        if extra:
            row.update(extra)
        rows.append(row)
        print(json.dumps(row), flush=True)

    # --- config 2: LinearElasticityModel, four (five) constraints ---
    for c in (C.UNIAXIAL_STRAIN, C.UNIAXIAL_STRESS, C.PLANE_STRAIN, C.PLANE_STRESS, C.FULL):
        g, s = c.geometric_dim, c.stress_strain_dim
        law = LinearElasticityModel(synthetic.ELASTIC_PARAMS, c)
        grad, stress = rnd(n * g * g, 1e-3), rnd(n * s, 0.1)
        tangent = torch.empty(n * s * s, dtype=torch.float64, device=dev)
        ms = time_steps(lambda i: law.evaluate(0.0, 1.0, grad, stress, tangent, None), K)
        report(f"elastic_{c.name}", 8 * (g * g + 2 * s + s * s), ms)
        del grad, stress, tangent

    # --- config 3: VonMises3D, AoS and SoA plastic-strain layouts ---
    for layout in ("aos", "soa"):
        law = VonMises3D(synthetic.MISES_PARAMS)
        law.defer_errors = True
        law.eps_layout = layout
        grad = rnd(n * 9, synthetic.MISES_GRAD_STD)
        tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        states = [(z(n * 6), z(n * 6), z(n)) for _ in range(K + 3)]

        def step(i):
            st, ep, al = states[i]
            law.evaluate(0.0, 1.0, grad, st, tangent, {"eps_n": ep, "alpha": al})

        ms = time_steps(step, K)
        law.check_converged()
        frac = float((states[3][2] > 0).double().mean().item())
        report(f"mises_{layout}", 568, ms, {"plastic_fraction": frac})
        del states, grad, tangent

    # --- config 4: Kelvin / Maxwell FULL, 100 increments with history carry-over ---
    for cls in (SpringKelvinModel, SpringMaxwellModel):
        law = cls(synthetic.VISCO_PARAMS, C.FULL)
        grad = rnd(n * 9, 1e-4)
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        stress, ev, et = z(n * 6), z(n * 6), z(n * 6)
        tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
        h = {"strain_visco": ev, "strain": et}
        ms = time_steps(lambda i: law.evaluate(0.0, 2.0, grad, stress, tangent, h), 100, warmup=3)
        report(f"{cls.__name__}_FULL_100inc", 648, ms, {"increments": 100, "total_ms": ms * 100})
        del grad, stress, ev, et, tangent
    for cls in (SpringKelvinModel, SpringMaxwellModel):
        for c in (C.UNIAXIAL_STRESS, C.PLANE_STRAIN):
            g, s = c.geometric_dim, c.stress_strain_dim
            law = cls(synthetic.VISCO_PARAMS, c)
            grad = rnd(n * g * g, 1e-4)
            z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
            stress, ev, et = z(n * s), z(n * s), z(n * s)
            tangent = torch.empty(n * s * s, dtype=torch.float64, device=dev)
            h = {"strain_visco": ev, "strain": et}
            ms = time_steps(lambda i: law.evaluate(0.0, 2.0, grad, stress, tangent, h), K)
            report(f"{cls.__name__}_{c.name}", 8 * (g * g + 6 * s + s * s), ms)
            del grad, stress, ev, et, tangent

    # --- comfe-rs models: linear-hardening Mises, Drucker-Prager classic / hyperbolic ---
    from fenics_constitutive_b200.models import (DruckerPrager3D, DruckerPragerHyperbolic3D,
                                                 MisesPlasticityLinearHardening3D)
    import numpy as np

    A = lambda v: np.array([v])  # noqa: E731
    rs_cases = [
        ("rs_mises_linear_hardening", MisesPlasticityLinearHardening3D,
         {"mu": A(80769.0), "kappa": A(175000.0), "y_0": A(1200.0), "h": A(200.0)}, synthetic.MISES_GRAD_STD, None),
        ("rs_drucker_prager", DruckerPrager3D,
         {"mu": A(80769.0), "kappa": A(175000.0), "a": A(300.0), "b": A(0.05), "b_flow": A(0.05)}, 1.7e-3, 4e-4),
        ("rs_drucker_prager_hyperbolic", DruckerPragerHyperbolic3D,
         {"mu": A(80769.0), "kappa": A(175000.0), "a": A(300.0), "b": A(0.05), "d": A(40.0), "b_flow": A(0.02)},
         1.7e-3, 4e-4),
    ]
    for name, cls, prm, shear, vol in rs_cases:
        law = cls(prm)
        law.record_plastic_flag = True
        grad = rnd(n * 9, shear)
        if vol is not None:  # deviator-dominated increments (stay away from the apex of the cone)
            grad.view(n, 9)[:, [0, 4, 8]] = rnd(n * 3, vol).view(n, 3)
        tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
        z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
        states = [(z(n * 6), z(n * 7)) for _ in range(K + 3)]

        def step(i):
            st, hi = states[i]
            law.evaluate(0.0, 1.0, grad, st, tangent, {"history": hi})

        ms = time_steps(step, K)
        frac = float(law.plastic_flag.double().mean().item())
        report(name, 8 * (9 + 6 + 6 + 36 + 7 + 7) + 1, ms, {"plastic_fraction": frac})
        del states, grad, tangent

    # --- companion gather: ~1M P2 tets (BASELINE config 5 mesh size), q_degree 2 ---
    from fenics_constitutive_b200 import gather as G

    coords, cv, dofmap = G.unit_cube_p2_tets(55, 55, 55)
    Jinv = G.affine_inverse_jacobians(coords, cv)
    pts, _ = G.simplex_quadrature(3, 2)
    op = G.IncrementalGradient(3, dofmap, G.lagrange_gradients(3, 2, pts), Jinv)
    u = rnd(coords.size, 1e-3)
    u_prev = rnd(coords.size, 1e-3)
    gout = torch.empty(op.num_qps * 9, dtype=torch.float64, device=dev)
    nq_total = op.num_qps
    for tag, prev in (("u_and_u_prev", u_prev), ("du_only", None)):
        ms = time_steps(lambda i: op.evaluate(u, prev, gout), K)
        # compulsory bytes per cell: dofmap 40 + Jinv 72 + out 288 (+ nodal values, L2-resident)
        bytes_cell = 40 + 72 + 288
        gbs = bytes_cell * op.ncells / (ms * 1e-3) / 1e9
        row = {"kernel": f"gather_P2tet_q2_{tag}", "cells": op.ncells, "qps": nq_total, "ms": ms,
               "qp_per_s": nq_total / (ms * 1e-3), "bytes_per_cell_compulsory": bytes_cell,
               "GBps_compulsory": gbs, "frac_of_measured_hbm": gbs / P}
        rows.append(row)
        print(json.dumps(row), flush=True)
    del gout, u, u_prev, op

    # --- reference point: torch copy bandwidth on this very GPU ---
    a = torch.empty(1 << 29, dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    ms = time_steps(lambda i: b.copy_(a), 10)
    row = {"kernel": "torch_copy_4GiB", "ms": ms, "GBps": 2 * a.numel() * 8 / (ms * 1e-3) / 1e9}
    rows.append(row)
    print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
