#!/usr/bin/env python
"""Host-array (e2e) path of VonMises3D.evaluate over its knobs -- memory kind x download wire x pool
threads x chunk size x slots -- one JSON line per configuration, median of `--steps` calls at
`--qps` points per rank (max over ranks per call).  Run under torchrun for N > 1: the ranks share the
host's memory system, which is what this sweep is about.

    python scripts/e2e_sweep.py [--qps 16000000] [--steps 3] [--quick]
"""
from __future__ import annotations

import argparse
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
from fenics_constitutive_b200.partition import env_rank_world, max_over_ranks  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qps", type=int, default=16_000_000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--quick", action="store_true", help="defaults + wires only")
ap.add_argument("--mini", action="store_true", help="defaults, wires and a few chunk / thread points (multi-GPU sessions)")
ap.add_argument("--kinds", default="pageable,pinned")
ap.add_argument("--stress-only", action="store_true")
ap.add_argument("--mix", default=None,
                help="only this: record wire (1) with these shares (percent, comma separated) of the chunks by plain "
                     "DMA (fcx_host_wire_mix), optionally x --mix-threads pool threads")
ap.add_argument("--mix-threads", default="0")
ap.add_argument("--custom", default=None,
                help="only these: semicolon-separated slots:chunk_qps[:threads] points (0 = leave the default), after a "
                     "run with the defaults and with one more of it at the end (drift of the box)")
args = ap.parse_args()
rank, local_rank, world = env_rank_world()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
L = lib()
n = args.qps
law = VonMises3D(synthetic.MISES_PARAMS)
rng = np.random.default_rng(99 + rank)


def arrays(kind):
    if kind == "pinned":
        mk = lambda m: torch.empty(m, dtype=torch.float64).pin_memory()  # noqa: E731
    else:
        mk = lambda m: torch.from_numpy(np.zeros(m))  # noqa: E731
    a = [mk(n * 9), mk(n * 6), mk(n * 6), mk(n), mk(n * 36)]
    a[0].numpy()[:] = rng.standard_normal(n * 9) * synthetic.MISES_GRAD_STD
    return a


def run(kind, A, tag, **knobs):
    old = {}
    for k, v in knobs.items():
        fn = getattr(L, f"fcx_host_{k}")
        old[k] = fn(v)
    g, st, ep, al, tg = A
    times = []
    try:
        for i in range(args.steps + 1):
            for a in (st, ep, al):
                a.zero_()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            law.evaluate(0.0, 1.0, g.numpy(), st.numpy(), None if args.stress_only else tg.numpy(),
                         {"eps_n": ep.numpy(), "alpha": al.numpy()})
            dt = time.perf_counter() - t0
            if i > 0:
                times.append(max_over_ranks(dt, dev))
    finally:
        for k, v in old.items():
            getattr(L, f"fcx_host_{k}")(v)
    if rank == 0:
        med = float(np.median(times))
        print(json.dumps({"n_gpus": world, "memory": kind, **knobs, "wire_used": int(L.fcx_host_wire_used()),
                          "wire_mix_used": int(L.fcx_host_wire_mix_used()),
                          "stress_only": args.stress_only, "tag": tag,
                          "MQPps_aggregate": round(world * n / med / 1e6, 1), "step_s": [round(t, 4) for t in times]}),
              flush=True)


for kind in args.kinds.split(",") if args.custom else ():
    A = arrays(kind)
    run(kind, A, "defaults (wire auto)")
    for pt in args.custom.split(";"):
        v = [int(x) for x in pt.split(":")]
        kn = {}
        if v[0] > 0:
            kn["slots"] = v[0]
        if len(v) > 1 and v[1] > 0:
            kn["chunk_qps"] = v[1]
        if len(v) > 2 and v[2] > 0:
            kn["threads"] = v[2]
        run(kind, A, "custom", **kn)
    run(kind, A, "defaults again")
    del A
for kind in args.kinds.split(",") if args.mix else ():
    A = arrays(kind)
    for threads in [int(t) for t in args.mix_threads.split(",")]:
        for mix in [int(m) for m in args.mix.split(",")]:
            kn = {"wire": 1, "wire_mix": mix}
            if threads > 0:
                kn["threads"] = threads
            run(kind, A, "mix", **kn)
    del A
for kind in args.kinds.split(",") if not (args.mix or args.custom) else ():
    A = arrays(kind)
    run(kind, A, "defaults (wire auto)")
    for wire in (0, 1, 2):
        run(kind, A, "wire", wire=wire)
    if args.mini:
        for chunk in (8192, 16384, 65536):
            run(kind, A, "chunk", chunk_qps=chunk)
        for threads in (2, 4, 8):
            run(kind, A, "threads", threads=threads)
    elif not args.quick:
        for chunk in (1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18):
            run(kind, A, "chunk", chunk_qps=chunk)
        for slots in (3, 4, 6, 8):
            run(kind, A, "slots", slots=slots)
        # pool threads LAST and ascending: the pool only grows, so every point runs with at most its own count
        for threads in (4, 6, 8, 10, 12, 14, 16):
            run(kind, A, "threads", threads=threads)
            if threads in (10, 12, 16):
                for chunk in (1 << 15, 1 << 17):
                    run(kind, A, "threads+chunk", threads=threads, chunk_qps=chunk)
    del A
if world > 1:
    dist.destroy_process_group()
