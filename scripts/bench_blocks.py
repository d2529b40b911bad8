"""Secondary blocks of bench.py's JSON line: every BASELINE config next to
the headline, measured in the same process on the same box.

    models_block   device-resident time, algorithmic B/QP and fraction of the measured HBM peak for
                   elastic x4 constraints (config 2), Mises second step / stress-only (config 3),
                   Kelvin / Maxwell per increment and over 100 increments (config 4), the companion
                   gather (du once, u and u_prev) and the fused form() kernel
    newton_block   config 5: IncrSmallStrainProblem + NewtonSolver stand-in, VonMises3D, ~1 M P2 tets,
                   ONE mesh partitioned over the ranks (N = 1: the whole mesh on one GPU)
    host_roofline  what the host memory system and the PCIe link of this box deliver (fcx_diag_*)
    e2e_traffic    host-DRAM / PCIe bytes per QP of the host-array path for a memory kind and wire
"""
from __future__ import annotations

import ctypes
import time

import numpy as np


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _time_steps(torch, fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = _events(torch)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def models_block(dev, n: int, peak_gbs: float, steps: int = 10) -> dict:
    """One row per kernel: {"ms", "bytes_per_qp", "GBps", "frac", ...}; n QPs, inputs resident in HBM,
    CUDA events on the launching stream, every array >> L2 (no flush needed at n = 16 M)."""
    import torch

    from fenics_constitutive_b200 import gather as G
    from fenics_constitutive_b200 import solver as S
    from fenics_constitutive_b200 import synthetic
    from fenics_constitutive_b200.models import (LinearElasticityModel, SpringKelvinModel, SpringMaxwellModel,
                                                 StressStrainConstraint as C, VonMises3D)

    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    rows: dict[str, dict] = {}

    def rnd(m, scale):
        return torch.randn(m, dtype=torch.float64, device=dev, generator=gen) * scale

    def z(m):
        return torch.zeros(m, dtype=torch.float64, device=dev)

    def row(name, bytes_per_qp, ms, units=n, **extra):
        gbs = bytes_per_qp * units / (ms * 1e-3) / 1e9
        rows[name] = {"ms": round(ms, 5), "bytes_per_qp": bytes_per_qp, "qp_per_s": units / (ms * 1e-3),
                      "GBps": round(gbs, 1), "frac": round(gbs / peak_gbs, 4), **extra}

    # ---- config 2: LinearElasticityModel, the four constraints BASELINE names
    for c in (C.UNIAXIAL_STRESS, C.PLANE_STRAIN, C.PLANE_STRESS, C.FULL):
        g, s = c.geometric_dim, c.stress_strain_dim
        law = LinearElasticityModel(synthetic.ELASTIC_PARAMS, c)
        grad, stress = rnd(n * g * g, 1e-3), rnd(n * s, 0.1)
        tangent = torch.empty(n * s * s, dtype=torch.float64, device=dev)
        ms = _time_steps(torch, lambda i: law.evaluate(0.0, 1.0, grad, stress, tangent, None), steps)
        row(f"elastic_{c.name}", 8 * (g * g + 2 * s + s * s), ms)
        if c == C.FULL:
            ms = _time_steps(torch, lambda i: law.evaluate(0.0, 1.0, grad, stress, None, None), steps)
            row("elastic_FULL_stress_only", 8 * (g * g + 2 * s), ms)
        del grad, stress, tangent

    # ---- config 3: VonMises3D beyond the headline: second step from the hardened state, stress-only
    law = VonMises3D(synthetic.MISES_PARAMS)
    law.defer_errors = True
    law.record_plastic_flag = False
    grad = rnd(n * 9, synthetic.MISES_GRAD_STD)
    half = grad * 0.5
    tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
    st0, ep0, al0 = z(n * 6), z(n * 6), z(n)
    law.evaluate(0.0, 1.0, grad, st0, tangent, {"eps_n": ep0, "alpha": al0})  # hardened state after step 1
    nsets = 4
    sets = [(st0.clone(), ep0.clone(), al0.clone()) for _ in range(nsets)]

    def reset():
        for st, ep, al in sets:
            st.copy_(st0), ep.copy_(ep0), al.copy_(al0)

    def step2(i, tg):
        st, ep, al = sets[i % nsets]
        law.evaluate(0.0, 1.0, half, st, tg, {"eps_n": ep, "alpha": al})

    # every timed step starts from the SAME hardened state: 3 warm-ups + `nsets` timed steps per pass
    def timed_pass(tg):
        ms = []
        for _ in range(max(1, steps // nsets)):
            reset()
            torch.cuda.synchronize()
            e0, e1 = _events(torch)
            e0.record()
            for i in range(nsets):
                step2(i, tg)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / nsets)
        return float(np.median(ms))

    step2(0, tangent), step2(1, tangent), step2(2, None)  # warm-up (both instantiations)
    ms2 = timed_pass(tangent)
    st, ep, al = sets[0]
    frac2 = float((al > al0).double().mean().item())
    row("mises_second_step", 568, ms2, plastic_fraction=round(frac2, 4))
    ms_so = timed_pass(None)
    row("mises_second_step_stress_only", 280, ms_so, plastic_fraction=round(frac2, 4))
    # stress-only from the virgin state (the headline's inputs)
    vsets = [(z(n * 6), z(n * 6), z(n)) for _ in range(steps + 3)]

    def virgin(i):
        st, ep, al = vsets[i]
        law.evaluate(0.0, 1.0, grad, st, None, {"eps_n": ep, "alpha": al})

    ms = _time_steps(torch, virgin, steps)
    row("mises_virgin_stress_only", 280, ms, plastic_fraction=round(float((vsets[3][2] > 0).double().mean().item()), 4))
    law.check_converged()
    del sets, vsets, st0, ep0, al0, grad, half, tangent

    # ---- config 4: Kelvin / Maxwell FULL, 100 increments with history carry-over
    for cls in (SpringKelvinModel, SpringMaxwellModel):
        law = cls(synthetic.VISCO_PARAMS, C.FULL)
        grad = rnd(n * 9, 1e-4)
        stress, ev, et = z(n * 6), z(n * 6), z(n * 6)
        tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
        h = {"strain_visco": ev, "strain": et}
        ms = _time_steps(torch, lambda i: law.evaluate(0.0, 2.0, grad, stress, tangent, h), 100, warmup=3)
        row(f"{cls.__name__}_FULL_per_increment", 648, ms, increments=100, total_ms_100_increments=round(100 * ms, 3))
        del grad, stress, ev, et, tangent

    # ---- comfe-rs mirrors: linear-hardening Mises, Drucker-Prager classic / hyperbolic, ~52 % plastic
    from fenics_constitutive_b200.models import (DruckerPrager3D, DruckerPragerHyperbolic3D,
                                                 MisesPlasticityLinearHardening3D)

    A1 = lambda v: np.array([v])  # noqa: E731
    rs_cases = [
        ("rs_mises_linear_hardening", MisesPlasticityLinearHardening3D,
         {"mu": A1(80769.0), "kappa": A1(175000.0), "y_0": A1(1200.0), "h": A1(200.0)}, synthetic.MISES_GRAD_STD, None),
        ("rs_drucker_prager", DruckerPrager3D,
         {"mu": A1(80769.0), "kappa": A1(175000.0), "a": A1(300.0), "b": A1(0.05), "b_flow": A1(0.05)}, 1.7e-3, 4e-4),
        ("rs_drucker_prager_hyperbolic", DruckerPragerHyperbolic3D,
         {"mu": A1(80769.0), "kappa": A1(175000.0), "a": A1(300.0), "b": A1(0.05), "d": A1(40.0), "b_flow": A1(0.02)},
         1.7e-3, 4e-4),
    ]
    for name, cls, prm, shear, vol in rs_cases:
        law = cls(prm)
        law.record_plastic_flag = True
        grad = rnd(n * 9, shear)
        if vol is not None:  # deviator-dominated increments (stay away from the apex of the cone)
            grad.view(n, 9)[:, [0, 4, 8]] = rnd(n * 3, vol).view(n, 3)
        tangent = torch.empty(n * 36, dtype=torch.float64, device=dev)
        states = [(z(n * 6), z(n * 7)) for _ in range(steps + 3)]  # a fresh virgin state per launch

        def rs_step(i, law=law, grad=grad, tangent=tangent, states=states):
            st, hi = states[i]
            law.evaluate(0.0, 1.0, grad, st, tangent, {"history": hi})

        ms = _time_steps(torch, rs_step, steps)
        row(name, 8 * (9 + 6 + 6 + 36 + 7 + 7) + 1, ms,
            plastic_fraction=round(float(law.plastic_flag.double().mean().item()), 4),
            parity="vs the C restatement of comfe-rs (unpinned against the reference: no rustc here)")
        del states, grad, tangent

    # ---- companion gather + fused form(): 998 250 P2 tets (BASELINE config 5 mesh size), q_degree 2
    coords, cv, dofmap = G.unit_cube_p2_tets(55, 55, 55)
    Jinv = G.affine_inverse_jacobians(coords, cv)
    pts, _ = G.simplex_quadrature(3, 2)
    op = G.IncrementalGradient(3, dofmap, G.lagrange_gradients(3, 2, pts), Jinv, device=dev)
    u, u_prev = rnd(coords.size, 1e-3), rnd(coords.size, 1e-3)
    gout = torch.empty(op.num_qps * 9, dtype=torch.float64, device=dev)
    flush = torch.empty(1 << 25, dtype=torch.float64, device=dev)  # 256 MB > L2 between launches

    def flushed(fn, reps=10):
        ts = []
        fn()
        for _ in range(reps):
            flush.zero_()
            e0, e1 = _events(torch)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    # compulsory bytes per cell: dofmap 40 + Jinv 72 + grad out 288 = 400 (nodal values L2-resident)
    for tag, prev in (("du_once", None), ("u_and_u_prev", u_prev)):
        ms = flushed(lambda: op.evaluate(u, prev, gout))
        row(f"gather_P2tet_q2_{tag}", 100, ms, units=op.num_qps, cells=op.ncells, l2="flushed between launches",
            bytes_per_cell=400)
    del gout
    # fused form(): state in 104 + state out 104 + tangent 288 per QP + (dofmap 40 + Jinv 72) per cell / 4 QPs
    nq = op.num_qps
    P = np.array([synthetic.MISES_PARAMS[k] for k in ("p_ka", "p_mu", "p_y0", "p_y00", "p_w")])
    from fenics_constitutive_b200._lib import check, lib

    L = lib()
    sp, sc, e0_, e1_, a0_, a1_ = z(nq * 6), z(nq * 6), z(nq * 6), z(nq * 6), z(nq), z(nq)
    tg = torch.empty(nq * 36, dtype=torch.float64, device=dev)
    du = rnd(coords.size, 2.0e-4)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def form(tangent_ptr):
        check(L.fcx_mises_form(P.ctypes.data, op.ncells, None, 4, 10, op.dofmap.data_ptr(), du.data_ptr(), None,
                               op.dphi_ref.data_ptr(), op.Jinv.data_ptr(), sp.data_ptr(), sc.data_ptr(), tangent_ptr,
                               e0_.data_ptr(), e1_.data_ptr(), a0_.data_ptr(), a1_.data_ptr(), None, None, None, None,
                               stream), "fcx_mises_form")

    ms = flushed(lambda: form(tg.data_ptr()))
    row("mises_form_fused_P2tet_q2", 524, ms, units=nq, cells=op.ncells, l2="flushed between launches",
        plastic_fraction=round(float((a1_ > 0).double().mean().item()), 4))
    ms = flushed(lambda: form(None))
    row("mises_form_fused_stress_only", 236, ms, units=nq, cells=op.ncells, l2="flushed between launches")
    return rows


def newton_block(rank: int, world: int, dev, grid: int = 55, load_steps: int = 2) -> dict | None:
    """BASELINE config 5 on the stand-in driver (NOT dolfinx/PETSc): VonMises3D, 6*grid^3 P2 tets, clamped
    left face, right face pulled in x through the yield point; ONE mesh partitioned over the ranks
    (solver/partitioned.py).  Inexact Newton (Eisenstat-Walker) + Jacobi-PCG on the tangent records."""
    import torch

    from fenics_constitutive_b200 import solver as S
    from fenics_constitutive_b200 import synthetic
    from fenics_constitutive_b200.models import VonMises3D
    from fenics_constitutive_b200.partition import max_over_ranks

    t0 = time.perf_counter()
    mesh = S.create_unit_cube(grid, grid, grid)
    part = None
    if world > 1:
        part = S.MeshPartition(mesh, 2, rank, world)
        V = part.V
    else:
        V = S.functionspace(mesh, ("CG", 2, (3,)))
    u = S.Function(V, dev)
    law = VonMises3D(synthetic.MISES_PARAMS)
    law.defer_errors = True
    left = lambda x: np.isclose(x[0], 0.0)   # noqa: E731
    right = lambda x: np.isclose(x[0], 1.0)  # noqa: E731
    zero, ux = S.Constant(0.0), S.Constant(0.0)
    bcs = [S.dirichletbc(zero, S.locate_dofs_geometrical(V, left), V),
           S.dirichletbc(ux, S.locate_dofs_geometrical(V, right), V.sub(0))]
    problem = S.IncrSmallStrainProblem(law, u, bcs, q_degree=2)
    problem.keep_del_grad_u = False
    solver = S.NewtonSolver(None, problem)
    solver.linear_solver = "cg"
    solver.cg_rtol = 1e-8
    solver.cg_forcing = "eisenstat-walker"
    solver.reduce_over_ranks = world > 1
    if part is not None:
        part.attach(solver)
    solver.profile = True
    setup_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    t1 = time.perf_counter()
    newton_its, krylov = [], []
    for k in range(1, load_steps + 1):
        ux.value = 0.012 * k / load_steps
        n_it, _ = solver.solve(u)
        problem.update()
        newton_its.append(n_it)
        krylov.append(int(sum(solver.krylov_iterations)))
    torch.cuda.synchronize()
    solve_s = max_over_ranks(time.perf_counter() - t1, dev)
    law.check_converged()
    lin_s = max_over_ranks(solver.linear_solve_s, dev)
    if rank != 0:
        return None
    kit = max(1, sum(krylov))
    return {"driver": "stand-in (not dolfinx/PETSc)", "n_gpus": world,
            "mode": "one mesh partitioned over the ranks (strong scaling)" if world > 1 else "whole mesh on one GPU",
            "global_cells": int(mesh.num_cells), "cells_this_rank": int(problem.num_cells),
            "dofs_this_rank": int(V.num_dofs), "load_steps": load_steps, "newton_iterations": newton_its,
            "krylov_iterations": krylov, "krylov": getattr(solver, "krylov_method", "pcg"),
            "setup_s": round(setup_s, 2), "solve_s": round(solve_s, 3),
            "linear_solve_s": round(lin_s, 3), "ms_per_krylov_iteration": round(1e3 * lin_s / kit, 4)}


def host_roofline(L, sizes_mb: int = 512) -> dict:
    """Measured ceilings of the host side: memcpy (read + written bytes per second) over the thread counts
    the pool may use, streaming fill, and the PCIe link with both copy engines busy."""
    info = (ctypes.c_int * 4)()
    L.fcx_host_numa_info(info, 4)
    allowed = int(info[2]) if info[2] > 0 else 8
    out = (ctypes.c_double * 4)()
    best = {"memcpy_rw_GBps": 0.0, "stream_fill_GBps": 0.0, "read_GBps": 0.0, "threads": 0}
    for t in sorted({min(8, allowed), min(12, allowed), min(16, allowed), allowed}):
        if L.fcx_diag_host_bandwidth(t, sizes_mb << 20, out, 4) != 0:
            continue
        if out[0] > best["memcpy_rw_GBps"]:
            best = {"memcpy_rw_GBps": round(out[0], 1), "stream_fill_GBps": round(out[1], 1),
                    "read_GBps": round(out[2], 1), "threads": t}
    p = (ctypes.c_double * 4)()
    pcie = None
    if L.fcx_diag_pcie(256 << 20, p, 4) == 0:
        pcie = {"h2d_alone_GBps": round(p[0], 1), "d2h_alone_GBps": round(p[1], 1),
                "h2d_concurrent_GBps": round(p[2], 1), "d2h_concurrent_GBps": round(p[3], 1)}
    return {"host_memory": best, "pcie": pcie, "cpus_usable": allowed,
            "numa": {"gpu_node": int(info[0]), "node_cpus": int(info[1]), "threads_pinned": bool(info[3])}}


def e2e_traffic(memory: str, wire: int, p: float, tangent: bool = True) -> dict:
    """Bytes per QP of a VonMises3D host-array call: what crosses PCIe each way and what
    moves through host DRAM (DMA reads / writes + what the host threads read and write).
    memory: "pageable" (inputs and stress staged through pinned ring slots by the pool) or "pinned";
    wire: 0 plain DMA of every array, 1 record wire (stress + flag for all points, 21 + 7 doubles for
    plastic points, elastic tangents filled from the GPU-computed constant); p = plastic fraction."""
    h2d = 176.0
    if wire == 0:
        d2h = 392.0 if tangent else 104.0
        dram = (3 * h2d + 3 * d2h) if memory == "pageable" else (h2d + d2h)
    else:
        rec = (224.0 if tangent else 56.0) * p
        d2h = 48.0 + 1.0 + rec
        expand_w = (288.0 if tangent else 0.0) + 56.0 * p
        if memory == "pageable":
            dram = 3 * h2d + 2 * d2h + 48.0 + expand_w
        else:
            dram = h2d + d2h + (1.0 + rec) + expand_w
    # lower bound: everything that only passes through a pinned ring slot (staged inputs, the record stream)
    # is served by the last-level cache; only the caller's own arrays are read from / written to DRAM
    if wire == 0:
        dram_lo = h2d + d2h
    else:
        dram_lo = h2d + 48.0 + expand_w
    return {"pcie_h2d": h2d, "pcie_d2h": round(d2h, 1), "host_dram": round(dram, 1),
            "host_dram_if_slots_stay_in_llc": round(dram_lo, 1)}
