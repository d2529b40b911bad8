#!/usr/bin/env python
"""Where the wall time of VonMises3D.evaluate(host arrays) goes: fcx_host_stats phase timings
for the download-wire modes, pinned and pageable caller arrays.  One JSON line per run."""
from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200._lib import lib  # noqa: E402
from fenics_constitutive_b200.models import VonMises3D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
L = lib()
pin = lambda m: torch.empty(m, dtype=torch.float64).pin_memory()  # noqa: E731
h = [pin(n * 9), pin(n * 6), pin(n * 6), pin(n), pin(n * 36)]
pg = [torch.from_numpy(np.zeros(m)) for m in (n * 9, n * 6, n * 6, n, n * 36)]
gr = np.random.default_rng(99).standard_normal(n * 9) * synthetic.MISES_GRAD_STD
h[0].numpy()[:] = gr
pg[0].numpy()[:] = gr
law = VonMises3D(synthetic.MISES_PARAMS)
names = ["total_s", "main_wait_slot_s", "main_stage_in_s", "main_enqueue_s", "drain_event_wait_s", "drain_expand_s",
         "gpu_h2d_s", "gpu_kernel_s", "gpu_pack_s", "gpu_d2h_s", "chunks", "chunk_qps"]
L.fcx_host_trace(1)
for mem, arrs in (("pinned", h), ("pageable", pg)):
    for wire in (1, 2):
        for thr, chunk, slots in ((14, 1 << 16, 3), (14, 1 << 16, 6), (14, 1 << 16, 8), (14, 1 << 15, 8), (14, 1 << 17, 6), (8, 1 << 16, 6)):
            L.fcx_host_wire(wire)
            L.fcx_host_threads(thr)
            L.fcx_host_chunk_qps(chunk)
            L.fcx_host_slots(slots)
            for rep in range(2):
                for a in arrs[1:4]:
                    a.zero_()
                t0 = time.perf_counter()
                law.evaluate(0.0, 1.0, arrs[0].numpy(), arrs[1].numpy(), arrs[4].numpy(),
                             {"eps_n": arrs[2].numpy(), "alpha": arrs[3].numpy()})
                dt = time.perf_counter() - t0
            st = (ctypes.c_double * 12)()
            L.fcx_host_stats(st, 12)
            row = {"memory": mem, "wire": wire, "threads": thr, "slots": slots, "MQPs": round(n / dt / 1e6, 1)}
            row.update({k: round(v, 4) for k, v in zip(names, st)})
            print(json.dumps(row), flush=True)
