#!/usr/bin/env python
"""CPU baselines of the headline workload (VonMises3D, ~50 % plastic), no GPU needed:
  reference  the reference's UNMODIFIED Python class through oracle/ref_shim.py (only where /root/reference
             exists, i.e. the build container), one process, small sample -- it is a per-point CPython loop;
  c_port     oracle/fcx_oracle.c (gcc -O2, OpenMP), all host threads;
  numba      oracle/numba_models.py (@njit(parallel=True)), all numba threads.
One JSON line per baseline.  Test infrastructure (it lives under tests/ because only tests/, smoke() and
bench.py's CPU legs may use oracle/): nothing here is part of the product path.
    python tests/cpu_baselines.py [--qps N --seconds S]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from fenics_constitutive_b200 import synthetic  # noqa: E402
from oracle import models as om  # noqa: E402
from oracle import numba_models as nm  # noqa: E402
from oracle import ref_shim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qps", type=int, default=4_000_000)
ap.add_argument("--reference-qps", type=int, default=100_000)
ap.add_argument("--seconds", type=float, default=8.0)
args = ap.parse_args()


def run(name, law, n, extra):
    grad, s0, e0, a0 = synthetic.mises_inputs_numpy(n, seed=1234)
    tangent = np.zeros(n * 36)
    state = (s0.copy(), e0.copy(), a0.copy())

    def one():
        for dst, src in zip(state, (s0, e0, a0)):
            np.copyto(dst, src)
        t0 = time.perf_counter()
        law.evaluate(0.0, 1.0, grad, state[0], tangent, {"eps_n": state[1], "alpha": state[2]})
        return time.perf_counter() - t0

    one()  # warm-up (JIT compilation, thread pools, page faults)
    total, passes = 0.0, 0
    while total < args.seconds and passes < 200:
        total += one()
        passes += 1
    row = {"baseline": name, "qp_per_s": n * passes / total, "sample_qps": n, "passes": passes,
           "plastic_fraction": float((state[2] > 0).mean()), "host_cpus": os.cpu_count()}
    row.update(extra)
    print(json.dumps(row), flush=True)


if ref_shim.available():
    m = ref_shim.load()
    run("reference (unmodified Python class, 1 process)", m.VonMises3D(synthetic.MISES_PARAMS), args.reference_qps,
        {"threads": 1})
law = om.VonMises3D(synthetic.MISES_PARAMS)
law.nthreads = oracle.max_threads()
run("c_port (gcc -O2, OpenMP)", law, args.qps, {"threads": law.nthreads})
if nm.available():
    import numba

    run("numba (@njit(parallel=True))", nm.VonMises3D(synthetic.MISES_PARAMS), args.qps,
        {"threads": numba.get_num_threads()})
