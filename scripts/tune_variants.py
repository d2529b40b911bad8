#!/usr/bin/env python
"""A/B sweep of kernel variants on one GPU (device-resident, CUDA events):
  * Mises: generic tile pipeline (variant 0) vs output-staged kernel (variant 1),
    tile 64/128, CTAs per SM;
  * constant-tangent models (elastic FULL / PLANE_STRAIN, Kelvin FULL): thread
    stores vs bulk stores from a constant shared-memory block (l2_hints bit 3);
  * DRAM reference points on the same box: copy, pure write (fill), pure read (sum).
    python scripts/tune_variants.py [--qps 16000000]"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--qps", type=int, default=16_000_000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--lib", default=None)
ap.add_argument("--mises-only", action="store_true")
args = ap.parse_args()

from fenics_constitutive_b200 import _lib, synthetic  # noqa: E402
from fenics_constitutive_b200.models import (  # noqa: E402
    LinearElasticityModel, SpringKelvinModel, StressStrainConstraint, VonMises3D)

if args.lib:
    _lib.LIB_PATH = os.path.abspath(args.lib)
L = _lib.lib()
n, K = args.qps, args.steps
dev = torch.device("cuda", 0)
C = StressStrainConstraint


def timeit(fn, reps=3):
    best = None
    for _ in range(reps):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(3 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        best = ms if best is None else min(best, ms)
    return best


def report(tag, ms, bpq):
    print(json.dumps({"cfg": tag, "ms": round(ms, 4), "GBps": round(bpq * n / (ms * 1e-3) / 1e9, 1),
                      "GQPps": round(n / ms / 1e6, 3)}), flush=True)


z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731

# ---------------------------------------------------------------- Mises
grad, _, _, _ = synthetic.mises_inputs_torch(n, dev)
tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
states = [(z(n * 6), z(n * 6), z(n)) for _ in range(K + 3)]
law = VonMises3D(synthetic.MISES_PARAMS)
law.defer_errors = True


def mises_step(i):
    st, ep, al = states[i]
    law.evaluate(0.0, 1.0, grad, st, tangent, {"eps_n": ep, "alpha": al})


def mises_cfg(tag):
    def fresh(i):
        mises_step(i)
    # every repetition must start from the virgin state
    best = None
    for _ in range(3):
        for st, ep, al in states:
            st.zero_(); ep.zero_(); al.zero_()
        ms = timeit(fresh, reps=1)
        best = ms if best is None else min(best, ms)
    report(tag, best, 568)


for variant, dyn in ((0, 0), (1, 0), (1, 1)):
    L.fcx_tune(b"mises_variant", variant)
    L.fcx_tune(b"dynamic_tiles", dyn)
    for tile in (64, 128):
        L.fcx_tune(b"tile", tile)
        L.fcx_tune(b"mises_tile", tile)
        for ctas in (0,) + ((7,) if tile == 64 else ()):
            L.fcx_tune(b"ctas_per_sm", ctas)
            mises_cfg(f"mises variant={variant} dyn={dyn} tile={tile} ctas_per_sm={ctas or 'occ'}")
L.fcx_tune(b"dynamic_tiles", 1)
L.fcx_tune(b"mises_tile", 64)
L.fcx_tune(b"mises_variant", 1)
L.fcx_tune(b"tile", 128)
L.fcx_tune(b"ctas_per_sm", 0)
del states
if args.mises_only:
    sys.exit(0)

# ------------------------------------------------ constant-tangent models
gen = torch.Generator(device=dev)
gen.manual_seed(7)
for name, cons, g, s in (("elastic_FULL", C.FULL, 3, 6), ("elastic_PLANE_STRAIN", C.PLANE_STRAIN, 2, 4)):
    lawe = LinearElasticityModel(synthetic.ELASTIC_PARAMS, cons)
    gr = torch.randn(n * g * g, dtype=torch.float64, device=dev, generator=gen) * 1e-3
    st = z(n * s)
    tg = torch.empty(n * s * s, dtype=torch.float64, device=dev)
    for hints, dyn in ((0, 1), (8, 0), (8, 1)):
        L.fcx_tune(b"l2_hints", hints)
        L.fcx_tune(b"dynamic_tiles", dyn)
        for tile in (64, 128, 256):
            L.fcx_tune(b"tile", tile)
            ms = timeit(lambda i: lawe.evaluate(0.0, 1.0, gr, st, tg, None))
            report(f"{name} tangent_bulk={hints // 8} dyn={dyn} tile={tile}", ms, 8 * (g * g + 2 * s + s * s))
    del gr, st, tg
lawk = SpringKelvinModel(synthetic.VISCO_PARAMS, C.FULL)
gr = torch.randn(n * 9, dtype=torch.float64, device=dev, generator=gen) * 1e-4
st, ev, et = z(n * 6), z(n * 6), z(n * 6)
tg = torch.empty(n * 36, dtype=torch.float64, device=dev)
for hints, dyn in ((0, 1), (8, 0), (8, 1)):
    L.fcx_tune(b"l2_hints", hints)
    L.fcx_tune(b"dynamic_tiles", dyn)
    for tile in (64, 128, 256):
        L.fcx_tune(b"tile", tile)
        ms = timeit(lambda i: lawk.evaluate(0.0, 2.0, gr, st, tg, {"strain_visco": ev, "strain": et}))
        report(f"kelvin_FULL tangent_bulk={hints // 8} dyn={dyn} tile={tile}", ms, 648)
L.fcx_tune(b"l2_hints", 8)
L.fcx_tune(b"tile", 128)
del gr, st, ev, et, tg

# ------------------------------------------------------- DRAM reference points
a = torch.empty(1 << 29, dtype=torch.float64, device=dev)
b = torch.empty_like(a)
ms = timeit(lambda i: b.copy_(a))
print(json.dumps({"cfg": "torch copy 4 GiB (R+W)", "ms": round(ms, 4), "GBps": round(2 * a.numel() * 8 / (ms * 1e-3) / 1e9, 1)}), flush=True)
ms = timeit(lambda i: b.fill_(1.0))
print(json.dumps({"cfg": "torch fill 4 GiB (W only)", "ms": round(ms, 4), "GBps": round(a.numel() * 8 / (ms * 1e-3) / 1e9, 1)}), flush=True)
ms = timeit(lambda i: a.sum())
print(json.dumps({"cfg": "torch sum 4 GiB (R only)", "ms": round(ms, 4), "GBps": round(a.numel() * 8 / (ms * 1e-3) / 1e9, 1)}), flush=True)
