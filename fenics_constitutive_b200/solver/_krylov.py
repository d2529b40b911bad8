"""Python handle of the device-resident Krylov loop (csrc/fcx_krylov.cu): Chronopoulos-Gear
single-reduction Jacobi-PCG whose dot-product reduction and ghost exchange are peer-memory
stores over NVLink from inside the kernels (one process per GPU, CUDA IPC), driven from C in
blocks of iterations -- the stand-in for what PETSc does behind dolfinx.nls.petsc.NewtonSolver
on an MPI-partitioned mesh (reference solver/_solver.py:64-68).

The only collective left on the host side is the one-off exchange of the 64-byte IPC handles
(``torch.distributed.all_gather_object``) when the solver is set up.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .. import _buffers as B
from .._lib import check, lib


class PeerMemoryUnavailable(RuntimeError):
    """The ranks' GPUs cannot map each other's memory (raised on every rank together)."""


class DeviceKrylov:
    def __init__(self, problem, partition=None):
        import torch
        import torch.distributed as dist

        self.problem = problem
        self.partition = partition
        self.device = problem.device
        L = self.L = lib()
        check(L.fcx_set_device(self.device.index), "fcx_set_device")
        world = partition.world if partition is not None else 1
        rank = partition.rank if partition is not None else 0
        self.world, self.rank = world, rank
        V = problem.V
        h, comm = ctypes.c_void_p(), ctypes.c_void_p()
        ipc = (ctypes.c_ubyte * 64)()
        owned = partition.num_owned_nodes if partition is not None else V.num_nodes
        check(L.fcx_krylov_create(rank, world, problem.gdim, V.num_nodes, owned, ctypes.byref(h), ctypes.byref(comm),
                                  ipc), "fcx_krylov_create")
        self.handle = h
        self._status = (ctypes.c_double * 4)()
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(ipc))
            flat = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
            rc = L.fcx_krylov_connect(h, flat)
            # a rank that cannot open its peers' blocks (no P2P between the GPUs, IPC disabled) must not leave the
            # others waiting in a spin: the outcome is agreed on collectively, and everybody raises or nobody
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                L.fcx_krylov_destroy(h)
                self.handle = None
                raise PeerMemoryUnavailable(
                    "device Krylov loop: CUDA IPC / peer access between the ranks' GPUs is not available "
                    f"(rank {rank}: {L.fcx_last_cuda_error().decode() or 'ok here, failed on a peer'})")
            nbr = [s for s, _, _ in partition.neighbours]
            send_src = [np.asarray(snd, dtype=np.int32) for _, snd, _ in partition.neighbours]
            send_dst = [np.asarray(pl, dtype=np.int32) for pl in partition.peer_local]
            ptr = np.zeros(len(nbr) + 1, dtype=np.int32)
            ptr[1:] = np.cumsum([a.size for a in send_src])
            src = np.concatenate(send_src) if nbr else np.zeros(0, dtype=np.int32)
            dst = np.concatenate(send_dst) if nbr else np.zeros(0, dtype=np.int32)
            nbr_a = np.asarray(nbr, dtype=np.int32)
            check(L.fcx_krylov_set_halo(h, len(nbr), nbr_a.ctypes.data, ptr.ctypes.data,
                                        np.ascontiguousarray(src).ctypes.data, np.ascontiguousarray(dst).ctypes.data),
                  "fcx_krylov_set_halo")
            dist.barrier()  # every rank has opened every block before anyone stores into one
        # local cells [0, num_interior_cells) touch no ghost node (MeshPartition orders them first)
        self.num_interior_cells = (partition.num_interior_cells if partition is not None and world > 1
                                   else problem.num_cells)
        self._x = torch.empty(V.num_dofs, dtype=torch.float64, device=self.device)
        self._status = (ctypes.c_double * 4)()
        self._snap = (ctypes.c_double * 6)()

    def _set_operator(self) -> None:
        pb, L = self.problem, self.L
        T = pb.tables
        rec = pb.fused and pb._trec_valid and pb.use_tangent_records and L.fcx_tune(b"fem_variant", -1) != 0
        tang = pb._trec if rec else pb.tangent.x.array
        pos = pb._pos_ptr()
        check(L.fcx_krylov_set_operator(
            self.handle, 3 if rec else 1, pb.sdim, pb.num_cells, T.nq, T.nd, pb._dofmap.data_ptr(),
            pb._dphi.data_ptr(), pb._weights.data_ptr(), pb._Jinv.data_ptr(), pb._detJ.data_ptr(), tang.data_ptr(),
            pb._fe.data_ptr(), pos, pb._adj_ptr.data_ptr(), None if pos is not None else pb._adj_idx.data_ptr(),
            self.num_interior_cells), "fcx_krylov_set_operator")

    # True: the residual test runs on the device (fcx_krylov_set_tolerance) and the host enqueues the next block
    # of iterations BEFORE it reads the outcome of the previous one -- no drained stream per check; the answer and
    # the iteration count are those of the exact stopping iteration.  False: drain and test on the host after
    # every block (the first version; iteration counts are multiples of check_every).
    lookahead = os.environ.get("FCX_KRYLOV_LOOKAHEAD", "1") != "0"

    def solve(self, rhs, minv, rtol: float, max_it: int, check_every: int):
        """Solve J x = rhs on the dofs where minv != 0.  Returns (x, iterations, converged, relres, breakdown);
        x is this object's workspace, valid until the next solve."""
        L, h = self.L, self.handle
        check(L.fcx_set_device(self.device.index), "fcx_set_device")
        stream = B.current_stream_ptr(self.device.index)
        self._set_operator()
        K = max(1, int(check_every))
        if not self.lookahead:
            return self._solve_drained(rhs, minv, rtol, max_it, K, stream)
        check(L.fcx_krylov_set_tolerance(h, float(rtol)), "fcx_krylov_set_tolerance")
        check(L.fcx_krylov_begin(h, rhs.data_ptr(), minv.data_ptr(), stream), "fcx_krylov_begin")
        snap = self._snap
        it, ok, relres, brk = 0, False, 1.0, False
        check(L.fcx_krylov_iterate(h, K, stream), "fcx_krylov_iterate")
        check(L.fcx_krylov_snapshot(h, 0, stream), "fcx_krylov_snapshot")
        enq, blk = K, 0
        while True:
            more = enq < max_it
            if more:  # the next block goes out before this one's outcome is known
                check(L.fcx_krylov_iterate(h, K, stream), "fcx_krylov_iterate")
                check(L.fcx_krylov_snapshot(h, (blk + 1) & 1, stream), "fcx_krylov_snapshot")
                enq += K
            check(L.fcx_krylov_wait_snapshot(h, blk & 1, snap), "fcx_krylov_wait_snapshot")
            frozen, stop_it, rr, rr0, flag, live = (float(v) for v in snap)
            if flag == 2.0:
                raise RuntimeError("device Krylov loop: a peer rank never arrived (peer-memory flag timed out)")
            it = int(stop_it) if frozen != 0.0 else min(enq - (K if more else 0), int(live) + 1)
            if rr0 == 0.0:
                ok, relres = True, 0.0
                break
            relres = float(np.sqrt(max(rr, 0.0) / rr0))
            if flag != 0.0:
                brk = True
                break
            if frozen != 0.0 or relres <= rtol:
                ok = True
                break
            if not more:
                break
            blk += 1
        check(L.fcx_krylov_solution(h, self._x.data_ptr(), stream), "fcx_krylov_solution")
        return self._x, it, ok, relres, brk

    def _solve_drained(self, rhs, minv, rtol, max_it, K, stream):
        import torch

        L, h = self.L, self.handle
        check(L.fcx_krylov_set_tolerance(h, 0.0), "fcx_krylov_set_tolerance")
        check(L.fcx_krylov_begin(h, rhs.data_ptr(), minv.data_ptr(), stream), "fcx_krylov_begin")
        it, ok, relres, brk = 0, False, 1.0, False
        while it < max_it:
            check(L.fcx_krylov_iterate(h, K, stream), "fcx_krylov_iterate")
            it += K
            torch.cuda.current_stream(self.device).synchronize()
            check(L.fcx_krylov_status(h, self._status), "fcx_krylov_status")
            _, rr, rr0, flag = (float(v) for v in self._status)
            if flag == 2.0:
                raise RuntimeError("device Krylov loop: a peer rank never arrived (peer-memory flag timed out)")
            if rr0 == 0.0:
                ok, relres = True, 0.0
                break
            relres = float(np.sqrt(max(rr, 0.0) / rr0))
            if relres <= rtol:
                ok = True
                break
            if flag != 0.0:
                brk = True
                break
        check(L.fcx_krylov_solution(h, self._x.data_ptr(), stream), "fcx_krylov_solution")
        return self._x, it, ok, relres, brk

    def halo_update(self, x) -> None:
        """Ghost entries of the nodal vector x <- their owners' values (peer-memory push; collective)."""
        if self.world == 1:
            return
        check(self.L.fcx_set_device(self.device.index), "fcx_set_device")
        check(self.L.fcx_krylov_halo_update(self.handle, x.data_ptr(), B.current_stream_ptr(self.device.index)),
              "fcx_krylov_halo_update")

    def close(self) -> None:
        if self.handle is not None:
            self.L.fcx_krylov_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
