#!/usr/bin/env python
"""Tuning sweep for the Mises tile kernel (tile size, CTAs/SM, L2 hints) plus the
read:write-mix DRAM ceiling (fcx_diag_stream_mix) on the same GPU.
    python scripts/tune_mises.py [--qps 16000000] [--lib path/to/libfcx.so]"""
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
args = ap.parse_args()

from fenics_constitutive_b200 import _lib  # noqa: E402

if args.lib:
    _lib.LIB_PATH = os.path.abspath(args.lib)
from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200.models import VonMises3D  # noqa: E402

L = _lib.lib()
n, K = args.qps, args.steps
dev = torch.device("cuda", 0)
grad, _, _, _ = synthetic.mises_inputs_torch(n, dev)
tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
z = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)  # noqa: E731
states = [(z(n * 6), z(n * 6), z(n)) for _ in range(K + 3)]
law = VonMises3D(synthetic.MISES_PARAMS)
law.defer_errors = True


def run_cfg(tag):
    for st, ep, al in states:
        st.zero_(); ep.zero_(); al.zero_()
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        for st, ep, al in states:
            st.zero_(); ep.zero_(); al.zero_()
        for i in range(3):
            st, ep, al = states[i]
            law.evaluate(0.0, 1.0, grad, st, tangent, {"eps_n": ep, "alpha": al})
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            st, ep, al = states[3 + i]
            law.evaluate(0.0, 1.0, grad, st, tangent, {"eps_n": ep, "alpha": al})
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        best = ms if best is None else min(best, ms)
    gbs = 568 * n / (best * 1e-3) / 1e9
    print(json.dumps({"cfg": tag, "ms": round(best, 4), "GBps": round(gbs, 1), "GQPps": round(n / best / 1e6, 3)}), flush=True)


for tile in (64, 128, 256):
    L.fcx_tune(b"tile", tile)
    for ctas in (0,) + tuple(c for c in (2, 3, 4, 6, 8) if c * tile <= 512):
        L.fcx_tune(b"ctas_per_sm", ctas)
        run_cfg(f"tile={tile} ctas_per_sm={ctas or 'occ'}")
L.fcx_tune(b"tile", 128)
L.fcx_tune(b"ctas_per_sm", 0)
for hints in (2, 4, 6):
    L.fcx_tune(b"l2_hints", hints)
    run_cfg(f"tile=128 l2_hints={hints}")
L.fcx_tune(b"l2_hints", 0)

# DRAM ceiling for the same read:write mix
del states
src = torch.randn(n * 22, dtype=torch.float64, device=dev)
dst = torch.empty(n * 49, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    done = L.fcx_diag_stream_mix(src.data_ptr(), dst.data_ptr(), n, stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    L.fcx_diag_stream_mix(src.data_ptr(), dst.data_ptr(), n, stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"cfg": "diag_stream_mix (coalesced 176R/392W per QP)", "qps": done, "ms": round(ms, 4),
                  "GBps": round(568 * done / (ms * 1e-3) / 1e9, 1)}), flush=True)
a = torch.empty(1 << 29, dtype=torch.float64, device=dev)
b = torch.empty_like(a)
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
e0.record()
for _ in range(K):
    b.copy_(a)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"cfg": "torch copy 4 GiB", "ms": round(ms, 4), "GBps": round(2 * a.numel() * 8 / (ms * 1e-3) / 1e9, 1)}), flush=True)
