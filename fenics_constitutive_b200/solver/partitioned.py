"""Mesh partition for the device-resident solver stand-in: one rank per GPU, each rank holds its
owned cells plus the ghost cells that touch its owned nodes -- the analogue of dolfinx's MPI mesh
partition that ``IncrSmallStrainProblem`` runs on in the reference (solver/_solver.py:64-68: the
constitutive update is evaluated on owned AND ghost cells, so stress / tangent / history on ghosts
are recomputed locally and never exchanged).

What IS exchanged is what the reference's solver stack (PETSc / dolfinx ``scatter_forward``,
solver/_incrementalunknowns.py:36-38) exchanges: the ghost values of nodal vectors -- once per
Jacobian action for the Krylov direction, once per Newton update for the displacement -- plus the
scalar all-reduces of the dot products.  Nothing of the constitutive hot path crosses ranks.

Construction is deterministic and communication-free: every rank derives the same global
ownership from the (replicated) coarse description of the mesh.

    part = MeshPartition(mesh, degree, rank, world)       # mesh = the GLOBAL mesh
    V = part.V                                            # local FunctionSpace (owned + ghost cells)
    u = Function(V); problem = IncrSmallStrainProblem(law, u, bcs, q_degree)
    solver = NewtonSolver(None, problem); part.attach(solver)
"""
from __future__ import annotations

import numpy as np

from .mesh import FunctionSpace, Mesh


def _balanced_slabs(key: np.ndarray, world: int) -> np.ndarray:
    """owner[c] for cells sorted by `key` (stable) and cut into `world` equal runs."""
    order = np.argsort(key, kind="stable")
    owner = np.empty(key.size, dtype=np.int32)
    bounds = np.linspace(0, key.size, world + 1).astype(np.int64)
    for r in range(world):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


class MeshPartition:
    """Local view of rank `rank` of a `world`-way partition of `mesh` for a P1/P2 vector space.

    Cells are assigned to ranks in slabs along x (equal counts).  A node belongs to the lowest
    rank among the cells that touch it.  The local mesh of a rank = every cell touching one of
    its owned nodes (so the rows of its owned nodes are complete without any exchange of
    element contributions), local nodes = owned nodes first, then ghosts, each in ascending
    global order."""

    def __init__(self, mesh: Mesh, degree: int, rank: int, world: int):
        if not (0 <= rank < world):
            raise ValueError("invalid rank/world")
        self.rank, self.world = int(rank), int(world)
        self.global_space = Vg = FunctionSpace(mesh, degree)
        g = mesh.gdim
        dm = Vg.dofmap.astype(np.int64)  # [nc][nd] global node ids
        centroid_x = mesh.coords[mesh.cells][:, :, 0].mean(axis=1)
        self.cell_owner = cell_owner = _balanced_slabs(centroid_x, world)
        node_owner = np.full(Vg.num_nodes, world, dtype=np.int32)
        np.minimum.at(node_owner, dm.ravel(), np.repeat(cell_owner, dm.shape[1]))
        self.node_owner = node_owner
        owner_of_cell_nodes = node_owner[dm]

        def local_sets(r):
            touches = (owner_of_cell_nodes == r).any(axis=1)
            cells = np.flatnonzero(touches | (cell_owner == r))
            nodes = np.unique(dm[cells])
            owned = nodes[node_owner[nodes] == r]
            ghost = nodes[node_owner[nodes] != r]
            return cells, owned, ghost

        cells, owned, ghost = local_sets(rank)
        # INTERIOR cells (all nodes owned by this rank) first, then the cells that touch a ghost node, each in
        # ascending global order: the Krylov loop runs its element kernel on the interior cells while the
        # neighbours' ghost values are still on their way (csrc/fcx_krylov.cu)
        interior = (owner_of_cell_nodes[cells] == rank).all(axis=1)
        cells = np.concatenate([cells[interior], cells[~interior]])
        self.num_interior_cells = int(interior.sum())
        self.local_cells = cells                      # global cell ids: interior (ascending), boundary (ascending)
        self.num_owned_cells = int((cell_owner[cells] == rank).sum())
        self.l2g = np.concatenate([owned, ghost])     # local node -> global node
        self.num_owned_nodes = int(owned.size)
        g2l = np.full(Vg.num_nodes, -1, dtype=np.int64)
        g2l[self.l2g] = np.arange(self.l2g.size)
        # local mesh (vertices renumbered) and local space with the SAME node numbering as l2g
        verts = np.unique(mesh.cells[cells])
        v2l = np.full(mesh.coords.shape[0], -1, dtype=np.int64)
        v2l[verts] = np.arange(verts.size)
        self.mesh = Mesh(mesh.coords[verts], v2l[mesh.cells[cells]])
        self.V = FunctionSpace.from_arrays(self.mesh, degree, Vg.node_coords[self.l2g],
                                           g2l[dm[cells]].astype(np.int32))
        # halo plan: (peer, local indices to send, local indices to receive), ascending global ids
        # peer_local[k] = the neighbour's LOCAL index of each node this rank sends to it (same order):
        # what a peer-memory push needs to store a value straight into the neighbour's vector
        self.neighbours, self.peer_local = [], []
        for s in range(world):
            if s == rank:
                continue
            _, owned_s, ghost_s = local_sets(s)
            send = np.intersect1d(owned, ghost_s, assume_unique=True)
            recv = np.intersect1d(ghost, owned_s, assume_unique=True)
            if send.size or recv.size:
                self.neighbours.append((s, g2l[send], g2l[recv]))
                # local numbering of rank s = its owned nodes, then its ghosts, each ascending (as above)
                self.peer_local.append(owned_s.size + np.searchsorted(ghost_s, send))
        self._dev_plan = {}
        self.block = g

    # ------------------------------------------------------------------ vectors
    def owned_dof_mask(self, device=None):
        import torch

        m = torch.zeros(self.V.num_dofs, dtype=torch.bool, device=device)
        m[: self.num_owned_nodes * self.block] = True
        return m

    def _plan(self, device):
        import torch

        key = str(device)
        if key not in self._dev_plan:
            self._dev_plan[key] = [(s, torch.as_tensor(snd, dtype=torch.int64, device=device),
                                    torch.as_tensor(rcv, dtype=torch.int64, device=device))
                                   for s, snd, rcv in self.neighbours]
        return self._dev_plan[key]

    def halo_update(self, x) -> None:
        """Overwrite the ghost entries of the blocked nodal vector `x` (flat torch tensor, host or
        device) with their owners' values.  One send + one receive per neighbouring rank."""
        import torch
        import torch.distributed as dist

        if self.world == 1 or not self.neighbours:
            return
        xb = x.view(-1, self.block)
        ops, recvs = [], []
        for s, snd, rcv in self._plan(x.device):
            if snd.numel():
                ops.append(dist.P2POp(dist.isend, xb.index_select(0, snd).contiguous(), s))
            if rcv.numel():
                buf = torch.empty((rcv.numel(), self.block), dtype=x.dtype, device=x.device)
                recvs.append((rcv, buf))
                ops.append(dist.P2POp(dist.irecv, buf, s))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for rcv, buf in recvs:
            xb.index_copy_(0, rcv, buf)

    def attach(self, solver) -> None:
        """Make a NewtonSolver partition-aware: norms / dot products over owned dofs summed over ranks,
        ghost values of the Krylov direction and of the solution refreshed where they are read."""
        solver.reduce_over_ranks = self.world > 1
        solver.partition = self

    def gather_global(self, x_local: np.ndarray) -> np.ndarray | None:
        """All owned values assembled into the global vector on rank 0 (tests / post-processing)."""
        import torch.distributed as dist

        no = self.num_owned_nodes
        mine = (self.l2g[:no], np.asarray(x_local).reshape(-1, self.block)[:no])
        parts = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(parts, mine)
        else:
            parts = [mine]
        if self.rank != 0:
            return None
        out = np.zeros((self.global_space.num_nodes, self.block))
        for ids, vals in parts:
            out[ids] = vals
        return out.ravel()
