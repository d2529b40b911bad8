#!/usr/bin/env python
"""torchrun check of the partitioned Newton solve (solver/partitioned.py) on N GPUs: ONE mesh split over
the ranks, Mises plasticity, two load steps across the yield point; rank 0 also solves the whole problem on
its own GPU and compares displacement, Newton iteration counts and the owned stresses.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_partitioned_newton.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenics_constitutive_b200 import solver as S  # noqa: E402
from fenics_constitutive_b200 import synthetic  # noqa: E402
from fenics_constitutive_b200.models import VonMises3D  # noqa: E402
from fenics_constitutive_b200.partition import env_rank_world  # noqa: E402

rank, local_rank, world = env_rank_world()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

left = lambda x: np.isclose(x[0], 0.0)   # noqa: E731
right = lambda x: np.isclose(x[0], 1.0)  # noqa: E731


def run(V, part, steps, forcing, driver="device"):
    u = S.Function(V, dev)
    law = VonMises3D(synthetic.MISES_PARAMS)
    zero, ux = S.Constant(0.0), S.Constant(0.0)
    bcs = [S.dirichletbc(zero, S.locate_dofs_geometrical(V, left), V),
           S.dirichletbc(ux, S.locate_dofs_geometrical(V, right), V.sub(0))]
    problem = S.IncrSmallStrainProblem(law, u, bcs, q_degree=2)
    solver = S.NewtonSolver(None, problem)
    solver.linear_solver = "cg"
    solver.cg_rtol = 1e-11
    solver.cg_forcing = forcing
    solver.cg_driver = driver  # "device": peer-memory Krylov loop (csrc/fcx_krylov.cu); "python": NCCL from Python
    if part is not None:
        part.attach(solver)
    its = []
    for k in range(1, steps + 1):
        ux.value = 0.012 * k / steps
        n_it, conv = solver.solve(u)
        assert conv
        problem.update()
        its.append((n_it, sum(solver.krylov_iterations)))
    return u, problem, its


ONLY_TWIN = "--only-twin" in sys.argv
for degree, n in (() if ONLY_TWIN else ((2, (12, 5, 4)), (1, (16, 6, 5)))):
    mesh = S.create_unit_cube(*n)
    for forcing, driver in ((None, "device"), ("eisenstat-walker", "device"), (None, "python")):
        part = S.MeshPartition(mesh, degree, rank, world)
        u, problem, its = run(part.V, part, 2, forcing, driver)
        glob = part.gather_global(u.x.array.cpu().numpy())
        sig_local = problem.stress_0.x.array.cpu().numpy().reshape(part.local_cells.size, -1)
        if rank == 0:
            Vg = S.functionspace(mesh, ("CG", degree, (3,)))
            ug, pg, its_g = run(Vg, None, 2, forcing, driver)
            ref = ug.x.array.cpu().numpy()
            err = np.abs(glob - ref).max() / np.abs(ref).max()
            sig_g = pg.stress_0.x.array.cpu().numpy().reshape(mesh.num_cells, -1)[part.local_cells]
            serr = np.abs(sig_local - sig_g).max() / np.abs(sig_g).max()
            plastic = float((pg._history_0[0]["alpha"].x.array > 0).double().mean().item())
            print(f"degree {degree} mesh {n} world {world} forcing {forcing} driver {driver}: |u - u_1gpu|/|u| = {err:.2e}, "
                  f"stress (rank 0 cells incl. ghosts) {serr:.2e}, newton/krylov its {its} vs 1 GPU {its_g}, "
                  f"plastic {plastic:.2f}, owned cells {part.num_owned_cells}/{mesh.num_cells}, "
                  f"neighbours {[(s, a.size, b.size) for s, a, b in part.neighbours]}", flush=True)
            assert err < 1e-7 and serr < 1e-6, (err, serr)
            assert [i[0] for i in its] == [i[0] for i in its_g]
        if world > 1:
            dist.barrier()
# ---- twin of the reference's partition-independence test (tests/solver/test_solver_mpi.py:93-121): unit cube 4 x 6 x 7,
# P1, q_degree 1, VonMises3D, left face clamped, right face pulled to 0.05 in load steps; the N-rank solution against the
# one-rank solution in the relative L2 (here: nodal l2) norm.  The reference compares two PETSc LU solves (1e-14); here
# both sides are Krylov solves driven to cg_rtol = 1e-14, so the difference is the round-off of two converged solves.
mesh = S.create_unit_cube(4, 6, 7)
nsteps = 20


def run_ref_test(V, part):
    u = S.Function(V, dev)
    law = VonMises3D({"p_ka": 175000, "p_mu": 80769, "p_y0": 1200, "p_y00": 2500, "p_w": 200})
    zero, ux = S.Constant(0.0), S.Constant(0.0)
    bcs = [S.dirichletbc(zero, S.locate_dofs_geometrical(V, left), V),
           S.dirichletbc(ux, S.locate_dofs_geometrical(V, right), V.sub(0)),
           S.dirichletbc(zero, S.locate_dofs_geometrical(V, right), V.sub(1)),
           S.dirichletbc(zero, S.locate_dofs_geometrical(V, right), V.sub(2))]
    problem = S.IncrSmallStrainProblem(law, u, bcs, q_degree=1)
    solver = S.NewtonSolver(None, problem)
    solver.linear_solver = "cg"
    solver.cg_rtol = 1e-14
    solver.error_on_krylov_failure = False  # at 1e-14 the last solves may stall at round-off: that is the point
    solver.rtol, solver.atol = 1e-13, 1e-11  # Newton driven to round-off too (dolfinx default 1e-9 leaves 1e-11 in u)
    solver.error_on_nonconvergence = False
    if part is not None:
        part.attach(solver)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(1, nsteps + 1):
            ux.value = 0.05 * k / nsteps
            solver.solve(u)
            problem.update()
    return u


part = S.MeshPartition(mesh, 1, rank, world)
u = run_ref_test(part.V, part)
glob = part.gather_global(u.x.array.cpu().numpy())
if rank == 0:
    ref = run_ref_test(S.functionspace(mesh, ("CG", 1, (3,))), None).x.array.cpu().numpy()
    err = np.linalg.norm(glob - ref) / np.linalg.norm(ref)
    print(f"reference test_mpi_solver twin (4x6x7 P1, {nsteps} load steps to 0.05, cg_rtol 1e-14), world {world}: "
          f"|u_world - u_self| / |u_self| = {err:.2e}", flush=True)
    assert err < 1e-14, err  # the reference's bar (test_solver_mpi.py:118-121)
if world > 1:
    dist.barrier()
if rank == 0:
    print("check_partitioned_newton: ok", flush=True)
if world > 1:
    dist.destroy_process_group()
