"""Mesh partition of the solver stand-in (fenics_constitutive_b200/solver/partitioned.py), on CPU:
ownership, local numbering, completeness of the owned rows (oracle FEM, numpy) and the ghost exchange
over a world_size-2 gloo group.  The GPU kernels are not involved: the partition is host logic, the
exchange runs on whatever tensors it is given (gloo here, NCCL on the GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fenics_constitutive_b200.solver.mesh import ElementTables, FunctionSpace, create_unit_cube, create_unit_square
from fenics_constitutive_b200.solver.partitioned import MeshPartition
from oracle import fem as F


def _oracle(V, qd):
    T = ElementTables(V, qd)
    return F.FemOracle(V.mesh.gdim, V.dofmap, T.dphi_ref, T.weights, T.Jinv, T.detJ, V.num_nodes)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("make,degree,qd", [(lambda: create_unit_cube(5, 3, 2), 2, 2), (lambda: create_unit_cube(9, 2, 2), 1, 1),
                                            (lambda: create_unit_square(7, 4), 2, 2)])
def test_partition_ownership_and_complete_owned_rows(make, degree, qd, world):
    mesh = make()
    Vg = FunctionSpace(mesh, degree)
    g = mesh.gdim
    s = {2: 4, 3: 6}[g]
    rng = np.random.default_rng(3)
    p_glob = rng.standard_normal(Vg.num_dofs)
    fem_g = _oracle(Vg, qd)
    nqp = mesh.num_cells * fem_g.nq
    # a random SPD-ish tangent per QP and a random stress field, defined per GLOBAL cell
    A = rng.standard_normal((nqp, s, s))
    tang_g = (A @ A.transpose(0, 2, 1) + s * np.eye(s)).reshape(mesh.num_cells, fem_g.nq, s * s)
    sig_g = rng.standard_normal((mesh.num_cells, fem_g.nq, s))
    y_glob = fem_g.tangent_matrix(tang_g.ravel()) @ p_glob
    f_glob = fem_g.internal_force(sig_g.ravel())
    owned_seen = np.zeros(Vg.num_nodes, dtype=int)
    cells_seen = np.zeros(mesh.num_cells, dtype=int)
    plans = []
    for r in range(world):
        P = MeshPartition(mesh, degree, r, world)
        no = P.num_owned_nodes
        owned_seen[P.l2g[:no]] += 1
        cells_seen[P.local_cells[P.cell_owner[P.local_cells] == r]] += 1
        assert np.all(P.node_owner[P.l2g[:no]] == r) and np.all(P.node_owner[P.l2g[no:]] != r)
        assert np.all(np.diff(P.l2g[:no]) > 0) and np.all(np.diff(P.l2g[no:]) > 0)
        assert np.allclose(P.V.node_coords, Vg.node_coords[P.l2g])
        assert np.array_equal(P.l2g[P.V.dofmap], Vg.dofmap[P.local_cells])
        # rows of owned nodes are complete on the local mesh (owned + ghost cells), no exchange needed
        fem_l = _oracle(P.V, qd)
        idx = (P.l2g[:, None] * g + np.arange(g)[None, :]).ravel()
        y_loc = fem_l.tangent_matrix(tang_g[P.local_cells].ravel()) @ p_glob[idx]
        f_loc = fem_l.internal_force(sig_g[P.local_cells].ravel())
        own = slice(0, no * g)
        assert np.abs(y_loc[own] - y_glob[idx[own]]).max() <= 1e-12 * np.abs(y_glob).max()
        assert np.abs(f_loc[own] - f_glob[idx[own]]).max() <= 1e-12 * np.abs(f_glob).max()
        plans.append(P)
    assert np.all(owned_seen == 1) and np.all(cells_seen == 1)
    # halo plans are mutually consistent: what r sends to s is what s expects from r, in the same order
    for P in plans:
        for s_rank, snd, rcv in P.neighbours:
            Q = plans[s_rank]
            back = [t for t in Q.neighbours if t[0] == P.rank]
            assert len(back) == 1
            assert np.array_equal(P.l2g[snd], Q.l2g[back[0][2]]) and np.array_equal(P.l2g[rcv], Q.l2g[back[0][1]])
    if world == 2:
        assert all(len(P.neighbours) == 1 for P in plans)  # slabs along x: one interface


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = create_unit_cube(6, 3, 2)
        P = MeshPartition(mesh, 2, rank, world)
        g, no = 3, P.num_owned_nodes
        want = (P.l2g[:, None] * 10.0 + np.arange(g)[None, :]).ravel()  # value = f(global node, component)
        x = torch.full((P.V.num_dofs,), float("nan"), dtype=torch.float64)
        x[: no * g] = torch.from_numpy(want[: no * g])
        P.halo_update(x)
        ok = bool(np.array_equal(x.numpy(), want))
        # reductions over owned dofs sum to the global ones
        tot = torch.tensor([float(want[: no * g].sum())], dtype=torch.float64)
        dist.all_reduce(tot)
        glob = P.gather_global(x.numpy())
        if rank == 0:
            n = P.global_space.num_nodes
            ref = (np.arange(n)[:, None] * 10.0 + np.arange(g)[None, :]).ravel()
            out.put((ok, float(tot.item()), float(ref.sum()), bool(np.array_equal(glob, ref))))
        else:
            assert ok
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_halo_exchange_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    ok, tot, ref_tot, glob_ok = out.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and glob_ok and tot == ref_tot


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("degree", [1, 2])
def test_peer_memory_push_plan_reaches_the_right_slots(degree, world):
    """The plan of the peer-memory ghost push (csrc/fcx_krylov.cu halo_push_kernel; MeshPartition.peer_local):
    for every neighbour, peer_local[k][e] is the NEIGHBOUR's local index of the node this rank sends as entry
    e.  Simulated on the host: every rank pushes the owned values of a global field into its neighbours'
    vectors at those indices; afterwards every ghost slot of every rank holds its owner's value, every slot is
    written exactly once, and no owned slot is touched."""
    mesh = create_unit_cube(7, 3, 2)
    parts = [MeshPartition(mesh, degree, r, world) for r in range(world)]
    rng = np.random.default_rng(0)
    field = rng.standard_normal(parts[0].global_space.num_nodes)
    vecs, hits = [], []
    for P in parts:
        v = np.full(P.l2g.size, np.nan)
        v[: P.num_owned_nodes] = field[P.l2g[: P.num_owned_nodes]]  # owned part known locally
        vecs.append(v)
        hits.append(np.zeros(P.l2g.size, dtype=int))
    for P in parts:
        assert len(P.peer_local) == len(P.neighbours)
        for (s, snd, _), dst in zip(P.neighbours, P.peer_local):
            assert dst.size == snd.size
            assert np.all(snd < P.num_owned_nodes)                       # only owned values are sent
            assert np.all(dst >= parts[s].num_owned_nodes)               # ... into ghost slots of the neighbour
            assert np.array_equal(parts[s].l2g[dst], P.l2g[snd])         # the same mesh node on both sides
            vecs[s][dst] = vecs[P.rank][snd]
            hits[s][dst] += 1
    for P, v, h in zip(parts, vecs, hits):
        assert np.array_equal(v, field[P.l2g])                           # every ghost refreshed from its owner
        assert np.all(h[: P.num_owned_nodes] == 0) and np.all(h[P.num_owned_nodes:] == 1)
