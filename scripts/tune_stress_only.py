#!/usr/bin/env python
"""Stress-only VonMises3D evaluate (tangent = NULL, 280 B/QP): kernel variants and grid sizes.
    mises_so_variant 1 = output-staged kernel <64,8,false> / <128,4,false> (single stage, 22 doubles per QP)
    mises_so_variant 0 = double-buffered tile pipeline: stress-only instantiation at tile 128 (no tangent
                         record in shared memory, one more CTA per SM), run-time skip at tiles 64 / 256
One JSON line per configuration (best of 3 passes of `--steps` launches, fresh virgin state per launch)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200._lib import lib  # noqa: E402
from fenics_constitutive_b200.models import VonMises3D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qps", type=int, default=16_000_000)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
L = lib()
n, K = args.qps, args.steps
dev = torch.device("cuda", 0)
grad, _, _, _ = synthetic.mises_inputs_torch(n, dev)
z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
states = [(z(n * 6), z(n * 6), z(n)) for _ in range(K + 3)]
law = VonMises3D(synthetic.MISES_PARAMS)
law.defer_errors = True


def run_cfg(tag):
    best = None
    for rep in range(3):
        for st, ep, al in states:
            st.zero_(); ep.zero_(); al.zero_()
        for i in range(3):
            st, ep, al = states[i]
            law.evaluate(0.0, 1.0, grad, st, None, {"eps_n": ep, "alpha": al})
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            st, ep, al = states[3 + i]
            law.evaluate(0.0, 1.0, grad, st, None, {"eps_n": ep, "alpha": al})
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        best = ms if best is None else min(best, ms)
    gbs = 280 * n / (best * 1e-3) / 1e9
    print(json.dumps({"cfg": tag, "ms": round(best, 4), "GBps_280B_per_qp": round(gbs, 1),
                      "GQPps": round(n / best / 1e6, 3)}), flush=True)


for variant, tiles in ((1, (64, 128)), (0, (64, 128, 256))):
    L.fcx_tune(b"mises_so_variant", variant)
    for tile in tiles:
        L.fcx_tune(b"mises_tile" if variant == 1 else b"tile", tile)
        for ctas in (0, 2, 3, 4, 6, 8, 12, 16):
            if ctas * tile > 2048:
                continue
            L.fcx_tune(b"ctas_per_sm", ctas)
            run_cfg(f"variant={variant} tile={tile} ctas_per_sm={ctas or 'occ'}")
L.fcx_tune(b"mises_so_variant", 0)
L.fcx_tune(b"mises_tile", 64)
L.fcx_tune(b"tile", 128)
L.fcx_tune(b"ctas_per_sm", 0)
law.check_converged()
