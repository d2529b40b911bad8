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


for degree, n in ((2, (12, 5, 4)), (1, (16, 6, 5))):
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
if rank == 0:
    print("check_partitioned_newton: ok", flush=True)
if world > 1:
    dist.destroy_process_group()
