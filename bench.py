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
pure Python and cannot travel to the GPU box) with all host threads, on the same
16M QPs per step as the GPU arm.  Further keys of the native line: `e2e` holds the
pageable-array result (what dolfinx hands over) as `value` next to `pinned`, per-step medians over
>= 5 steps and a measured host roofline (`e2e.roofline`); `models` carries every other BASELINE
config (scripts/bench_blocks.py); `newton` the full-Newton stand-in of config 5.
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
    sample = args.ref_qps
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
        "config": {"workload": WORKLOAD, "qps_per_gpu": sample,
                   "sample": (f"{sample} QPs per step" + (" (the whole 16M workload)" if sample == N_QP
                              else " (bounded sample of the 16M workload)"))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} QPs x {args.steps} steps, C oracle port (gcc -O2, OpenMP {threads} threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm

def _e2e_leg(kind, law_h, ne, Ke, rank, world, device, dist, L):
    """Ke timed `VonMises3D.evaluate` calls on host arrays of one memory kind (+ 1 warm-up call that
    allocates the pipeline's buffers and faults the pages in).  Every call starts from the virgin
    state (reset untimed); per-step time = max over ranks; reported value = median step."""
    import torch

    from fenics_constitutive_b200 import synthetic
    from fenics_constitutive_b200.partition import max_over_ranks

    reg_s = None
    if kind == "pinned":
        mk = lambda m: torch.empty(m, dtype=torch.float64).pin_memory()  # noqa: E731
    else:  # ordinary (pageable) numpy memory, as dolfinx hands it over
        mk = lambda m: torch.from_numpy(np.zeros(m))  # noqa: E731
    h_grad, h_st, h_ep, h_al, h_tg = mk(ne * 9), mk(ne * 6), mk(ne * 6), mk(ne), mk(ne * 36)
    arrays = (h_grad, h_st, h_ep, h_al, h_tg)
    if kind == "registered":  # pageable arrays page-locked once with fcx_host_register (a solver's set-up step)
        t0 = time.perf_counter()
        for a in arrays:
            rc = L.fcx_host_register(a.data_ptr(), a.numel() * 8)
            assert rc == 0, f"fcx_host_register rc={rc}"
        reg_s = time.perf_counter() - t0
    rng = np.random.default_rng(99 + rank)
    h_grad.numpy()[:] = rng.standard_normal(ne * 9) * synthetic.MISES_GRAD_STD
    times = []
    for i in range(Ke + 1):
        for a in (h_st, h_ep, h_al):
            a.zero_()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        law_h.evaluate(0.0, 1.0, h_grad.numpy(), h_st.numpy(), h_tg.numpy(),
                       {"eps_n": h_ep.numpy(), "alpha": h_al.numpy()})
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(max_over_ranks(dt, device))
    plastic = float((h_al.numpy() > 0).mean())
    wire = int(L.fcx_host_wire_used())
    if kind == "registered":
        for a in arrays:
            L.fcx_host_unregister(a.data_ptr())
    med = float(np.median(times))
    return {"value": world * ne / med, "best": world * ne / min(times), "steps": Ke,
            "step_s": [round(t, 4) for t in times], "plastic_fraction": round(plastic, 4),
            "wire": wire, "register_s": reg_s}


def run_native(args):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import bench_blocks as BB
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
    del states, grad, tangent
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, device-resident (rank 0 only: per-kernel numbers, not a scaling claim) ----
    models = None
    if rank == 0 and not args.no_models:
        try:
            models = BB.models_block(device, n, peak, steps=max(4, min(K, 10)))
        except Exception as exc:  # a secondary block must never cost the headline line
            models = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()

    # ---- host roofline: what this box's memory system and PCIe link deliver (rank 0 probes alone) ----
    roof = None
    if rank == 0:
        try:
            roof = BB.host_roofline(L)
        except Exception as exc:
            roof = {"error": f"{type(exc).__name__}: {exc}"}
    if world > 1:
        dist.barrier()

    # ---- end-to-end legs: public API, host arrays, H2D + D2H in the timed region ----
    ne, Ke = args.e2e_qps, max(1, args.e2e_steps)
    law_h = VonMises3D(synthetic.MISES_PARAMS)
    legs = {}
    for kind in args.e2e_memory.split(","):
        legs[kind] = _e2e_leg(kind, law_h, ne, Ke, rank, world, device, dist, L)
    head_kind = "pageable" if "pageable" in legs else next(iter(legs))
    head = legs[head_kind]

    # ---- full-Newton stand-in, BASELINE config 5: one mesh over all ranks ----
    newton = None
    if not args.no_newton:
        try:
            newton = BB.newton_block(rank, world, device, grid=args.newton_grid)
        except Exception as exc:
            newton = {"error": f"{type(exc).__name__}: {exc}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, threads, secs, passes = cpu_baseline_leg(args.cpu_sample)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {args.cpu_sample} QPs of the workload x {passes} passes from the virgin state, "
                         f"C oracle port (gcc -O2, OpenMP, {threads} threads), {secs:.1f} s timed"}

    if rank == 0:
        def leg_roofline(kind, leg):
            """Aggregate ceilings of the host-array path for this leg's memory kind / wire, from the probes."""
            if not roof or "error" in roof or not roof.get("pcie"):
                return None
            tr = BB.e2e_traffic("pinned" if kind in ("pinned", "registered") else "pageable", leg["wire"],
                                leg["plastic_fraction"])
            caps = {"host_dram": roof["host_memory"]["memcpy_rw_GBps"] * 1e9 / tr["host_dram"],
                    "pcie_h2d": world * roof["pcie"]["h2d_concurrent_GBps"] * 1e9 / tr["pcie_h2d"],
                    "pcie_d2h": world * roof["pcie"]["d2h_concurrent_GBps"] * 1e9 / tr["pcie_d2h"]}
            bound = min(caps, key=caps.get)
            return {"bound": bound, "ceiling_qp_per_s": {k: round(v) for k, v in caps.items()},
                    "bytes_per_qp": tr, "achieved_host_bytes_per_s": leg["value"] * tr["host_dram"],
                    "peak": roof["host_memory"]["memcpy_rw_GBps"] * 1e9,
                    "peak_unit": "B/s, host memcpy read+write, measured at start-up",
                    "frac": leg["value"] / caps[bound],
                    "note": "host_dram counts every pass through a pinned ring slot as DRAM traffic; the part the "
                            "last-level cache serves makes frac exceed 1 (bytes_per_qp.host_dram_if_slots_stay_in_llc "
                            "is the other extreme)"}

        e2e = {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": H2D_PER_QP * ne,
               "d2h_bytes_per_step": D2H_PER_QP * ne,
               # d2h_bytes_per_step: the result arrays delivered into host memory (392 B/QP);
               # d2h_wire_bytes_per_step: what actually crosses PCIe (record wire: stress + flag byte for every
               # point, 224 B only for plastic points; include/fcx.h fcx_host_wire)
               "d2h_wire_bytes_per_step": int(ne * (49 + 224 * head["plastic_fraction"])) if head["wire"]
               else D2H_PER_QP * ne,
               "qps_per_gpu": ne, "steps": Ke, "statistic": "median step, max over ranks per step",
               "host_memory": head_kind, "wire": head["wire"], "plastic_fraction": head["plastic_fraction"],
               "step_s": head["step_s"], "best": head["best"],
               "api": f"VonMises3D.evaluate(numpy {head_kind} host arrays) -> fcx_mises_evaluate_host",
               "roofline": leg_roofline(head_kind, head), "host_probes": roof}
        for kind, leg in legs.items():
            if kind != head_kind:
                e2e[kind] = {**leg, "roofline": leg_roofline(kind, leg)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "qps_per_gpu": n, "plastic_fraction": round(plastic_frac, 4),
                       "history_layout": "aos (reference contract)",
                       "l2": "inputs (9.1 GB touched per step) larger than L2; fresh state set per step"
                             + ("" if K + W <= nsets else f" (recycled after {nsets} steps)"),
                       "kernel": "fcx_mises_ostage_kernel<64,8,true>, atomic tile tickets",
                       "parallelism": f"qp-shard x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_qp": BYTES_PER_QP, "kernel_ms": kernel_ms},
            "cpu_baseline": cpu, "e2e": e2e, "models": models, "newton": newton,
            "gpu_launches": launches, "clocks": clocks, "host_cpus": os.cpu_count(),
        }
        print(json.dumps(line), flush=True)
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
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-memory", default="pageable,pinned",
                    help="comma list of host-memory kinds for the e2e legs: pageable (ordinary numpy arrays, the "
                         "headline e2e), pinned (page-locked), registered (pageable + fcx_host_register)")
    ap.add_argument("--cpu-sample", type=int, default=4_000_000, help="QPs of the cpu_baseline leg of the native arm")
    ap.add_argument("--ref-qps", type=int, default=N_QP, help="QPs per step of --impl reference (default: the 16M workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-models", action="store_true", help="skip the per-config `models` block")
    ap.add_argument("--no-newton", action="store_true", help="skip the full-Newton stand-in block")
    ap.add_argument("--newton-grid", type=int, default=55, help="cubes per direction (6 n^3 P2 tets; 55 = 998 250 cells)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
