#!/usr/bin/env python
"""Full Newton solve on the device-resident stand-in driver (BASELINE.json config 5,
SURVEY.md 8d): IncrSmallStrainProblem with VonMises3D on a P2 tetrahedral mesh of
the unit cube, uniaxial tension driven by Dirichlet BCs, a few load steps that
cross the yield point.  One process per GPU; with torchrun every rank solves its
own block (weak scaling: the dolfinx partition is replaced by independent
blocks) and the residual norms / dot products are summed over ranks with an
NCCL all-reduce of one double, standing in for the MPI residual-norm reduction.

STAND-IN DRIVER: not dolfinx/PETSc (absent from the image).  Reports per phase
(form = fused gather+evaluate kernel, F, Jacobian action) kernel times measured
with CUDA events, and whole-solve QP updates/s.

    python scripts/bench_newton.py [--n 55] [--steps 3] [--degree 2]
    torchrun --nproc-per-node N scripts/bench_newton.py ...
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

from fenics_constitutive_b200 import solver as S  # noqa: E402
from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200.models import VonMises3D  # noqa: E402
from fenics_constitutive_b200.partition import env_rank_world, max_over_ranks  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", "--grid", dest="n", type=int, default=55, help="grid cubes per direction (6 n^3 tets); use --grid under torchrun (its parser treats --n as an abbreviation)")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--degree", type=int, default=2)
ap.add_argument("--cg-rtol", type=float, default=1e-8)
ap.add_argument("--max-disp", type=float, default=0.012)
ap.add_argument("--partition", action="store_true",
                help="ONE mesh partitioned over the ranks (strong scaling; ghost values of nodal vectors exchanged "
                     "over NCCL, solver/partitioned.py) instead of one whole block per rank (weak scaling)")
ap.add_argument("--forcing", choices=["none", "ew"], default="none",
                help="ew = Eisenstat-Walker forcing terms for the Krylov tolerance (inexact Newton)")
ap.add_argument("--newton-steps-only", type=int, default=0,
                help="if > 0: stop each load step's CG after this many iterations (kernel timing runs)")
ap.add_argument("--ab", action="store_true", help="also time the element kernels with fem_variant 0")
ap.add_argument("--driver", choices=["device", "python"], default="device",
                help="Krylov loop: device = csrc/fcx_krylov.cu (peer-memory reduction / ghost push, driven from C), "
                     "python = kernel by kernel from Python with NCCL")
ap.add_argument("--check-every", type=int, default=10)
ap.add_argument("--eta-max", type=float, default=None, help="upper bound of the Eisenstat-Walker forcing term")
ap.add_argument("--eta-gamma", type=float, default=None)
args = ap.parse_args()

rank, local_rank, world = env_rank_world()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)

t0 = time.perf_counter()
mesh = S.create_unit_cube(args.n, args.n, args.n)
part = None
if args.partition:
    part = S.MeshPartition(mesh, args.degree, rank, world)
    V = part.V
else:
    V = S.functionspace(mesh, ("CG", args.degree, (3,)))
u = S.Function(V, dev)
law = VonMises3D(synthetic.MISES_PARAMS)
law.defer_errors = True
left = lambda x: np.isclose(x[0], 0.0)   # noqa: E731
right = lambda x: np.isclose(x[0], 1.0)  # noqa: E731
y0b = lambda x: np.isclose(x[1], 0.0)    # noqa: E731
z0b = lambda x: np.isclose(x[2], 0.0)    # noqa: E731
zero, ux = S.Constant(0.0), S.Constant(0.0)
# clamped left face, right face pulled in x: an inhomogeneous field (plastic zones grow from the clamp)
bcs = [S.dirichletbc(zero, S.locate_dofs_geometrical(V, left), V),
       S.dirichletbc(ux, S.locate_dofs_geometrical(V, right), V.sub(0))]
qd = 2 if args.degree == 2 else 1
problem = S.IncrSmallStrainProblem(law, u, bcs, q_degree=qd)
problem.keep_del_grad_u = False
solver = S.NewtonSolver(None, problem)
solver.linear_solver = "cg"
solver.cg_rtol = args.cg_rtol
solver.cg_forcing = "eisenstat-walker" if args.forcing == "ew" else None
solver.cg_driver = args.driver
solver.cg_check_every = args.check_every
if args.eta_max is not None:
    solver.cg_eta_max = args.eta_max
if args.eta_gamma is not None:
    solver.cg_eta_gamma = args.eta_gamma
solver.reduce_over_ranks = world > 1
if part is not None:
    part.attach(solver)
solver.profile = True
if args.newton_steps_only > 0:
    solver.cg_max_it = args.newton_steps_only
    solver.max_it = 2
    solver.error_on_nonconvergence = False
    solver.error_on_krylov_failure = False  # truncated solves are the point of this mode
setup_s = time.perf_counter() - t0


def time_kernel(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- whole solve ----
form_calls = 0
orig_form = problem.form


def counting_form(x=None):
    global form_calls
    form_calls += 1
    orig_form(x)


problem.form = counting_form
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t1 = time.perf_counter()
newton_its, krylov = [], []
for k in range(1, args.steps + 1):
    ux.value = args.max_disp * k / args.steps
    n_it, conv = solver.solve(u)
    problem.update()
    newton_its.append(n_it)
    krylov.append(list(solver.krylov_iterations))
torch.cuda.synchronize()
solve_s = max_over_ranks(time.perf_counter() - t1, dev)
law.check_converged()
alpha = problem._history_0[0]["alpha"].x.array
plastic_frac = float((alpha > 0).double().mean().item())
sxx = float(problem.stress_0.x.array[::6].mean().item())
if part is not None:  # mean over the OWNED cells of all ranks = the global mean
    import torch.distributed as dist  # noqa: F811

    own = torch.as_tensor(part.cell_owner[part.local_cells] == rank, device=dev).repeat_interleave(problem.nqp // problem.num_cells)
    acc = torch.stack([problem.stress_0.x.array[::6][own].sum(), (alpha[own] > 0).double().sum(), own.double().sum()])
    if world > 1:
        dist.all_reduce(acc)
    sxx, plastic_frac = float(acc[0] / acc[2]), float(acc[1] / acc[2])

# ---- per-kernel timings on the final state ----
p = torch.randn(V.num_dofs, dtype=torch.float64, device=dev)
y = torch.empty_like(p)
from fenics_constitutive_b200._lib import lib  # noqa: E402

ms_form = time_kernel(lambda: orig_form(None), 20)
ms_F = time_kernel(lambda: problem.F(), 20)
ms_J = time_kernel(lambda: problem.J_apply(p, y), 20)
ms_D = time_kernel(lambda: problem.J_diag(y), 20)
ms_gs = time_kernel(lambda: problem._gather_sum(y), 20)
ab = None
if args.ab:
    lib().fcx_tune(b"fem_variant", 0)
    ab = {"F": time_kernel(lambda: problem.F(), 20), "J_apply": time_kernel(lambda: problem.J_apply(p, y), 20),
          "J_diag": time_kernel(lambda: problem.J_diag(y), 20)}
    lib().fcx_tune(b"fem_variant", 1)
nqp = problem.nqp
fs = lib().fcx_fe_stride(3) * 8 * (10 if args.degree == 2 else 4)  # element-vector bytes per cell
form_bytes = nqp * (104 + 392) + problem.num_cells * (40 + 72)  # state in/out + dofmap + Jinv (nodal values L2-resident)
# tangent + dofmap + Jinv + detJ + element vector written and read once + adjacency index + nodal result
J_bytes = nqp * 288 + problem.num_cells * (40 + 72 + 8 + 2 * fs + 40) + V.num_dofs * 8
if rank == 0:
    print(json.dumps({
        "bench": "full Newton solve, stand-in driver (not dolfinx/PETSc)", "n_gpus": world,
        "mode": ("one mesh partitioned over the ranks (strong scaling)" if part is not None else "one block per rank (weak scaling)"),
        "global_cells": mesh.num_cells if part is not None else world * problem.num_cells,
        "owned_cells_this_rank": part.num_owned_cells if part is not None else problem.num_cells,
        "halo_neighbours": [(int(s), int(a.size), int(b.size)) for s, a, b in part.neighbours] if part is not None else [],
        "cells_per_gpu": problem.num_cells, "qps_per_gpu": nqp, "dofs_per_gpu": V.num_dofs,
        "degree": args.degree, "q_degree": qd, "load_steps": args.steps, "newton_iterations": newton_its,
        "krylov_iterations": krylov, "cg_rtol": args.cg_rtol, "cg_forcing": args.forcing, "cg_driver": args.driver,
 "fused_form": problem.fused,
        "setup_s": round(setup_s, 2), "solve_s": round(solve_s, 3),
        "linear_solve_s": round(solver.linear_solve_s, 3), "residual_s": round(solver.residual_s, 3),
        "update_s": round(solver.update_s, 3), "krylov_setup_s": round(solver.krylov_setup_s, 3),
        "ms_per_krylov_iteration": 1e3 * solver.linear_solve_s / max(1, sum(sum(k) for k in krylov)), "form_calls": form_calls,
        "qp_updates_per_s_whole_solve": (mesh.num_cells * (nqp // problem.num_cells) if part is not None else world * nqp) * form_calls / solve_s,
        "plastic_fraction_final": round(plastic_frac, 4), "mean_sigma_xx": sxx,
        "kernel_ms": {"form_fused": ms_form, "F": ms_F, "J_apply": ms_J, "J_diag": ms_D, "gather_sum_alone": ms_gs},
        "kernel_ms_fem_variant0": ab,
        "form_qp_per_s": nqp / (ms_form * 1e-3), "form_GBps": form_bytes / (ms_form * 1e-3) / 1e9,
        "J_apply_GBps": J_bytes / (ms_J * 1e-3) / 1e9,
    }), flush=True)
if world > 1:
    dist.destroy_process_group()
