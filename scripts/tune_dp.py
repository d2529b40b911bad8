#!/usr/bin/env python
"""Drucker-Prager kernels (comfe-rs DruckerPrager3D / DruckerPragerHyperbolic3D mirrors), 16 M points, ~52 % plastic:
    dp_variant 0 = the reference's spelling of every division / square root (DruckerPragerModel VAR 0)
    dp_variant 1 = slow fp64 operations spelled for latency (VAR 1, the default)
x resident CTAs per SM (0 = what the occupancy query gives), with and without the tangent.
One JSON line per configuration (best of 3 passes of `--steps` launches, fresh virgin state per launch)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenics_constitutive_b200._lib import lib  # noqa: E402
from fenics_constitutive_b200.models import DruckerPrager3D, DruckerPragerHyperbolic3D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qps", type=int, default=16_000_000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--variants", default="0,1")
ap.add_argument("--ctas", default="0")
args = ap.parse_args()
L = lib()
n, K = args.qps, args.steps
dev = torch.device("cuda", 0)
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
gen = torch.Generator(device=dev)
gen.manual_seed(1234)
grad = torch.randn(n * 9, dtype=torch.float64, device=dev, generator=gen) * 1.7e-3
grad.view(n, 9)[:, [0, 4, 8]] = (torch.randn(n * 3, dtype=torch.float64, device=dev, generator=gen) * 4e-4).view(n, 3)
z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
states = [(z(n * 6), z(n * 7)) for _ in range(K + 3)]
tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
A = lambda v: np.array([v])  # noqa: E731
LAWS = {
    "classic": DruckerPrager3D({"mu": A(80769.0), "kappa": A(175000.0), "a": A(300.0), "b": A(0.05), "b_flow": A(0.05)}),
    "hyperbolic": DruckerPragerHyperbolic3D({"mu": A(80769.0), "kappa": A(175000.0), "a": A(300.0), "b": A(0.05),
                                             "d": A(40.0), "b_flow": A(0.02)}),
}
ref = {}
for name, law in LAWS.items():
    law.record_plastic_flag = True
    for with_tan in (True, False):
        for var in [int(v) for v in args.variants.split(",")]:
            for ctas in [int(v) for v in args.ctas.split(",")]:
                L.fcx_tune(b"dp_variant", var)
                L.fcx_tune(b"ctas_per_sm", ctas)
                best = None
                for rep in range(3):
                    for st, hi in states:
                        st.zero_(); hi.zero_()
                    for i in range(3):
                        st, hi = states[i]
                        law.evaluate(0.0, 1.0, grad, st, tangent if with_tan else None, {"history": hi})
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(K):
                        st, hi = states[3 + i]
                        law.evaluate(0.0, 1.0, grad, st, tangent if with_tan else None, {"history": hi})
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / K
                    best = ms if best is None else min(best, ms)
                bytes_qp = 8 * (9 + 6 + 6 + 7 + 7 + (36 if with_tan else 0)) + 1
                st, hi = states[3]
                key = (name, with_tan)
                if key not in ref:
                    ref[key] = (st.clone(), hi.clone(), tangent[: 36 * 100_000].clone())
                    dmax = 0.0
                else:  # variants against the first one measured: relative difference of stress / history / tangent
                    r = ref[key]
                    dmax = max(float((st - r[0]).abs().max() / r[0].abs().max()),
                               float((hi - r[1]).abs().max() / r[1].abs().max()),
                               float((tangent[: 36 * 100_000] - r[2]).abs().max() / r[2].abs().max()) if with_tan else 0.0)
                print(json.dumps({"model": name, "tangent": with_tan, "dp_variant": var, "ctas_per_sm": ctas, "ms": round(best, 4),
                                  "GQPps": round(n / best / 1e6, 3), "GBps": round(bytes_qp * n / best / 1e6, 1),
                                  "frac_of_hbm": round(bytes_qp * n / best / 1e6 / PEAK, 3),
                                  "plastic_fraction": round(float(law.plastic_flag.double().mean().item()), 4),
                                  "max_rel_diff_vs_first": dmax}), flush=True)
L.fcx_tune(b"dp_variant", 1)
L.fcx_tune(b"ctas_per_sm", 0)
