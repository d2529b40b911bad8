#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric):
"Mises fp64 QP updates/s at 1/2/4/8 B200 and % of HBM roofline vs host CPU".

Workload (BASELINE.json configs[2]): VonMises3D with nonlinear isotropic
hardening, 16M synthetic quadrature points per GPU, ~50 % plastic from the
virgin state, radial return + 6x6 consistent tangent.  A "step" is ONE
`evaluate` over the whole batch = one kernel launch.  Every step starts from
its own virgin state set, so each of the K timed steps does identical work.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Prints ONE JSON line on rank 0.  `value` is device-resident throughput (inputs
already in HBM); `e2e` is the same metric through the public host-array API
(`VonMises3D.evaluate` with pinned numpy arrays: H2D + kernel + D2H inside the
timed region).  `--impl reference` times the CPU oracle port (the reference is
pure Python and cannot travel to the GPU box) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_QP = 16_000_000          # QPs per GPU (BASELINE.json: "16M synthetic QPs")
BYTES_PER_QP = 568         # algorithmic bytes per Mises update
H2D_PER_QP = 176           # grad 72 + stress 48 + eps_n 48 + alpha 8
D2H_PER_QP = 392           # stress 48 + tangent 288 + eps_n 48 + alpha 8
METRIC = "mises_fp64_qp_updates_per_s"
UNIT = "QP/s"
WORKLOAD = "VonMises3D nonlinear isotropic hardening, 16M synthetic QPs per GPU, ~50% plastic"


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the Mises kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "mises_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU legs

def _cpu_law_and_inputs(sample_qps: int):
    import oracle
    from oracle import models as om
    from fenics_constitutive_b200 import synthetic

    threads = oracle.max_threads()
    law = om.VonMises3D(synthetic.MISES_PARAMS)
    law.nthreads = threads
    grad, s0, e0, a0 = synthetic.mises_inputs_numpy(sample_qps, seed=1234)
    return law, threads, grad, (s0, e0, a0), np.zeros(sample_qps * 36)


def _cpu_pass(law, grad, virgin, state, tangent) -> float:
    """One timed evaluate over the sample from the virgin state (state reset untimed)."""
    for dst, src in zip(state, virgin):
        np.copyto(dst, src)
    t0 = time.perf_counter()
    law.evaluate(0.0, 1.0, grad, state[0], tangent, {"eps_n": state[1], "alpha": state[2]})
    return time.perf_counter() - t0


def cpu_baseline_leg(sample_qps: int, budget_s: float = 10.0, max_passes: int = 400):
    """Time the CPU oracle port (all host threads) on a bounded sample of the workload:
    repeated passes over the first `sample_qps` QPs until ~budget_s of CPU time."""
    law, threads, grad, virgin, tangent = _cpu_law_and_inputs(sample_qps)
    state = tuple(a.copy() for a in virgin)
    _cpu_pass(law, grad, virgin, state, tangent)  # warm-up (page faults, thread pool)
    total, passes = 0.0, 0
    while total < budget_s and passes < max_passes:
        total += _cpu_pass(law, grad, virgin, state, tangent)
        passes += 1
    return sample_qps * passes / total, threads, total, passes


def run_reference(args):
    """`--impl reference`: the CPU arm.  The reference is pure Python/numpy (it cannot
    travel to the GPU box), so this times the pinned C oracle port of its algorithm
    on all host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample
    law, threads, grad, virgin, tangent = _cpu_law_and_inputs(sample)
    state = tuple(a.copy() for a in virgin)
    for _ in range(args.warmup):
        _cpu_pass(law, grad, virgin, state, tangent)
    dt = 0.0
    for _ in range(args.steps):
        dt += _cpu_pass(law, grad, virgin, state, tangent)
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} QPs per step (bounded sample of the 16M workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} QPs x {args.steps} steps, C oracle port (gcc -O2, OpenMP {threads} threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm

def run_native(args):
    import torch
    import torch.distributed as dist

    from fenics_constitutive_b200 import synthetic
    from fenics_constitutive_b200._lib import lib
    from fenics_constitutive_b200.models import VonMises3D
    from fenics_constitutive_b200.partition import env_rank_world, max_over_ranks

    rank, local_rank, world = env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    n = args.qps
    K, W = args.steps, args.warmup
    L = lib()
    law = VonMises3D(synthetic.MISES_PARAMS)
    law.defer_errors = True  # enqueue-only inside the timed region; checked after it

    # ---- device-resident leg: inputs already in HBM ----
    grad, _, _, _ = synthetic.mises_inputs_torch(n, device, seed=1234 + rank)
    tangent = torch.empty(n * 36, dtype=torch.float64, device=device)
    z = lambda m: torch.zeros(m, dtype=torch.float64, device=device)  # noqa: E731
    # a fresh virgin state set per step (1.66 GB each); beyond 32 sets they are recycled, i.e. later
    # steps continue from an already loaded state (same bytes moved, same kernel)
    nsets = min(K + W, 32)
    states = [(z(n * 6), z(n * 6), z(n)) for _ in range(nsets)]

    def step(i):
        st, ep, al = states[i % nsets]
        law.evaluate(0.0, 1.0, grad, st, tangent, {"eps_n": ep, "alpha": al})

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = L.fcx_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(K):
        step(W + i)
    ev1.record()
    torch.cuda.synchronize()
    launches = int(L.fcx_launch_count() - launches0)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_local = ev0.elapsed_time(ev1)
    ms_total = max_over_ranks(ms_local, device)
    law.check_converged()
    plastic_frac = float((states[W % nsets][2] > 0).double().mean().item())
    value = world * n * K / (ms_total * 1e-3)
    kernel_ms = ms_local / K  # one launch per step
    achieved = BYTES_PER_QP * n / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    del states
    torch.cuda.empty_cache()

    # ---- end-to-end leg: public API, pinned host arrays, H2D + D2H in the timed region ----
    ne = args.e2e_qps
    Ke = args.e2e_steps
    reg_s = None
    if args.e2e_memory == "pinned":
        pin = lambda m: torch.empty(m, dtype=torch.float64).pin_memory()  # noqa: E731
    else:  # ordinary (pageable) numpy memory, as dolfinx hands it over
        pin = lambda m: torch.from_numpy(np.zeros(m))  # noqa: E731
    h_grad, h_st, h_ep, h_al, h_tg = pin(ne * 9), pin(ne * 6), pin(ne * 6), pin(ne), pin(ne * 36)
    if args.e2e_memory == "registered":  # pageable arrays page-locked once with fcx_host_register
        t0 = time.perf_counter()
        for a in (h_grad, h_st, h_ep, h_al, h_tg):
            rc = L.fcx_host_register(a.data_ptr(), a.numel() * 8)
            assert rc == 0, f"fcx_host_register rc={rc}"
        reg_s = time.perf_counter() - t0
    rng = np.random.default_rng(99 + rank)
    h_grad.numpy()[:] = rng.standard_normal(ne * 9) * synthetic.MISES_GRAD_STD
    law_h = VonMises3D(synthetic.MISES_PARAMS)
    e2e_s = 0.0
    for i in range(Ke + 1):  # first call is the warm-up (allocates the staging buffers)
        for a in (h_st, h_ep, h_al):
            a.zero_()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        law_h.evaluate(0.0, 1.0, h_grad.numpy(), h_st.numpy(), h_tg.numpy(),
                       {"eps_n": h_ep.numpy(), "alpha": h_al.numpy()})
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_s += max_over_ranks(dt, device)
    e2e_value = world * ne * Ke / e2e_s
    e2e_plastic = float((h_al.numpy() > 0).mean())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, threads, secs, passes = cpu_baseline_leg(args.cpu_sample)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {args.cpu_sample} QPs of the workload x {passes} passes from the virgin state, "
                         f"C oracle port (gcc -O2, OpenMP, {threads} threads), {secs:.1f} s timed"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "qps_per_gpu": n, "plastic_fraction": round(plastic_frac, 4),
                       "history_layout": "aos (reference contract)",
                       "l2": "inputs (9.1 GB touched per step) larger than L2; fresh state set per step"
                             + ("" if K + W <= nsets else f" (recycled after {nsets} steps)"),
                       "kernel": "fcx_mises_ostage_kernel<64,8>, atomic tile tickets",
                       "parallelism": f"qp-shard x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_qp": BYTES_PER_QP, "kernel_ms": kernel_ms},
            "cpu_baseline": cpu,
            # d2h_bytes_per_step: the result arrays delivered into host memory (392 B/QP);
            # d2h_wire_bytes_per_step: what actually crosses PCIe with the packed download wire
            # (stress + flag byte for every point, 224 B only for plastic points; include/fcx.h fcx_host_wire)
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": H2D_PER_QP * ne,
                    "d2h_bytes_per_step": D2H_PER_QP * ne,
                    "d2h_wire_bytes_per_step": (int(ne * (49 + 224 * e2e_plastic)) if L.fcx_host_wire(-1)
                                                else D2H_PER_QP * ne),
                    "qps_per_gpu": ne, "steps": Ke,
                    "plastic_fraction": round(e2e_plastic, 4),
                    "host_memory": args.e2e_memory, "register_s": reg_s,
                    "api": f"VonMises3D.evaluate(numpy {args.e2e_memory} host arrays) -> fcx_mises_evaluate_host"},
            "gpu_launches": launches, "clocks": clocks, "host_cpus": os.cpu_count(),
        }
        print(json.dumps(line), flush=True)
    if args.e2e_memory == "registered":
        for a in (h_grad, h_st, h_ep, h_al, h_tg):
            L.fcx_host_unregister(a.data_ptr())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--qps", type=int, default=N_QP, help="QPs per GPU (default: the 16M workload)")
    ap.add_argument("--e2e-qps", type=int, default=N_QP)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-memory", choices=["pinned", "pageable", "registered"], default="pinned",
                    help="host memory of the e2e leg's arrays (default: page-locked)")
    ap.add_argument("--cpu-sample", type=int, default=4_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
