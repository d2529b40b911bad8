#!/usr/bin/env python
"""Per-chunk timeline of VonMises3D.evaluate(pinned host arrays): what each chunk of the host
pipeline was doing when (fcx_host_timeline).  Writes one CSV per configuration and prints the
busy fraction of every phase (union of the chunks' intervals / wall time) -- the phase whose
union covers the wall time is the bottleneck."""
from __future__ import annotations

import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200._lib import lib  # noqa: E402
from fenics_constitutive_b200.models import VonMises3D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
out_dir = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
L = lib()
pin = lambda m: torch.empty(m, dtype=torch.float64).pin_memory()  # noqa: E731
h = [pin(n * 9), pin(n * 6), pin(n * 6), pin(n), pin(n * 36)]
h[0].numpy()[:] = np.random.default_rng(99).standard_normal(n * 9) * synthetic.MISES_GRAD_STD
law = VonMises3D(synthetic.MISES_PARAMS)
L.fcx_host_trace(1)
COLS = ["chunk", "slot", "acquired", "enqueued", "g_start", "g_h2d", "g_kernel", "g_pack", "g_d2h", "drain_woke", "expanded"]


def union(iv):
    iv = sorted(iv)
    tot, cur_a, cur_b = 0.0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    return tot + (cur_b - cur_a if cur_b is not None else 0.0)


CONFIGS = [(2, 6, 1 << 16, 0), (1, 6, 1 << 16, 0)]
if os.environ.get("FCX_TIMELINE_SKIPS"):  # diagnostic: leave phases out (fcx_host_debug_skip), results garbage
    CONFIGS += [(w, 6, 1 << 16, m) for w in (2, 1) for m in (1, 2, 4, 8, 1 | 8, 4 | 8, 1 | 4 | 8, 2 | 4 | 8, 1 | 2 | 4)]
for wire, slots, chunk, skip in CONFIGS:
    L.fcx_host_debug_skip(skip)
    L.fcx_host_wire(wire)
    L.fcx_host_slots(slots)
    L.fcx_host_chunk_qps(chunk)
    for rep in range(2):
        for a in h[1:4]:
            a.zero_()
        try:
            law.evaluate(0.0, 1.0, h[0].numpy(), h[1].numpy(), h[4].numpy(), {"eps_n": h[2].numpy(), "alpha": h[3].numpy()})
        except RuntimeError:
            pass  # garbage inputs under a skip mask may not converge
    rows = L.fcx_host_timeline(None, 0)
    buf = (ctypes.c_double * (rows * 11))()
    L.fcx_host_timeline(buf, rows)
    T = np.array(buf).reshape(rows, 11)
    if skip == 0:
        np.savetxt(os.path.join(out_dir, f"host_timeline_w{wire}_s{slots}_c{chunk}.csv"), T, delimiter=",", header=",".join(COLS), fmt="%.6f")
    wall = T[:, 10].max()
    phases = {"h2d": (4, 5), "kernel": (5, 6), "pack": (6, 7), "d2h": (7, 8), "event->drain": (8, 9), "expand": (9, 10),
              "slot idle (expanded -> next acquire)": None}
    rep = {"wire": wire, "slots": slots, "chunk": chunk, "skip_mask": skip, "wall_ms": round(1e3 * wall, 2), "MQPs": round(n / wall / 1e6, 1)}
    for name, cols in phases.items():
        if cols is None:
            continue
        iv = [(r[cols[0]], r[cols[1]]) for r in T]
        rep[name + " busy"] = round(union(iv) / wall, 3)
        rep[name + " mean_ms"] = round(1e3 * float(np.mean(T[:, cols[1]] - T[:, cols[0]])), 3)
    print(json.dumps(rep), flush=True)
L.fcx_host_debug_skip(0)
