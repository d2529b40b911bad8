#!/usr/bin/env python
"""A/B of the companion gather kernels on ~1M P2 tets (BASELINE config 5 mesh size), q_degree 2:
gather_kernel (register loads, fcx_tune gather_variant 0) vs gather_staged_kernel (nodal values
staged by cp.async one tile ahead, variant 1), over CTAs per SM.  One JSON line per run."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenics_constitutive_b200 import gather as G  # noqa: E402
from fenics_constitutive_b200._lib import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=55)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--ctas", default="0,1,2,3,4,6")
ap.add_argument("--variants", default="0,1,2,3")
args = ap.parse_args()
dev = torch.device("cuda", 0)
L = lib()
coords, cv, dofmap = G.unit_cube_p2_tets(args.grid, args.grid, args.grid)
Jinv = G.affine_inverse_jacobians(coords, cv)
pts, _ = G.simplex_quadrature(3, 2)
op = G.IncrementalGradient(3, dofmap, G.lagrange_gradients(3, 2, pts), Jinv)
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn(coords.size, dtype=torch.float64, device=dev, generator=g) * 1e-3
u_prev = torch.randn(coords.size, dtype=torch.float64, device=dev, generator=g) * 1e-3
out = torch.empty(op.num_qps * 9, dtype=torch.float64, device=dev)
flush = torch.empty(1 << 26, dtype=torch.float64, device=dev)  # 512 MB > L2


def timed(prev):
    ts = []
    for _ in range(args.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        op.evaluate(u, prev, out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


ref = {}
for variant in [int(v) for v in args.variants.split(',')]:
    L.fcx_tune(b"gather_variant", variant)
    for ctas in [int(c) for c in args.ctas.split(",")]:
        L.fcx_tune(b"ctas_per_sm", ctas)
        for tag, prev in (("du_only", None), ("u_and_u_prev", u_prev)):
            best, med = timed(prev)
            if tag not in ref:
                ref[tag] = out.clone()
            same = bool(torch.equal(out, ref[tag]))
            print(json.dumps({"variant": variant, "ctas_per_sm": ctas, "case": tag, "cells": op.ncells,
                              "ms_best": round(best, 4), "ms_median": round(med, 4),
                              "GBps_compulsory_400B_per_cell": round(400 * op.ncells / (best * 1e-3) / 1e9, 1),
                              "bitwise_equal_to_first": same}), flush=True)
L.fcx_tune(b"ctas_per_sm", 0)
L.fcx_tune(b"gather_variant", 1)
