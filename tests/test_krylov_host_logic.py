"""Host side of the device-resident Krylov loop (solver/_krylov.py: DeviceKrylov.solve) against a scripted stand-in
for libfcx.so -- the block / snapshot bookkeeping of the look-ahead loop needs no GPU: how many blocks go out, which
snapshot is waited for, what iteration count and residual are reported, max_it, breakdown, a peer that never arrives.
(The kernels behind these calls are covered by tests/test_solver_gpu.py and scripts/check_partitioned_newton.py.)"""
import math

import pytest
import torch

from fenics_constitutive_b200.solver import _krylov as KR


class FakeLib:
    """Mimics the fcx_krylov_* calls of one solve: r.r falls by `rate` per iteration from rr0; with a tolerance set
    the solve freezes at the first iteration whose r.r passes the test (like gsum_dots_kernel's finishing thread)."""

    def __init__(self, rr0=4.0, rate=0.5, breakdown_at=None, peer_lost_at=None):
        self.rr0, self.rate = rr0, rate
        self.breakdown_at, self.peer_lost_at = breakdown_at, peer_lost_at
        self.rtol2 = 0.0
        self.enq = 0          # iterations enqueued
        self.frozen_at = None
        self.snaps = {}
        self.log = []

    def _rr(self, it):
        return self.rr0 * self.rate ** it

    def _advance(self, k):
        for _ in range(k):
            it = self.enq
            self.enq += 1
            if self.frozen_at is not None:
                continue  # gated kernels: nothing happens
            if self.breakdown_at is not None and it == self.breakdown_at:
                self.frozen_at, self.brk = it, True
            elif self.rtol2 > 0.0 and self._rr(it) <= self.rtol2 * self.rr0:
                self.frozen_at, self.brk = it, False

    def fcx_set_device(self, i):
        return 0

    def fcx_krylov_set_tolerance(self, h, rtol):
        self.rtol2 = rtol * rtol if rtol > 0 else 0.0
        return 0

    def fcx_krylov_begin(self, h, rhs, minv, stream):
        self.enq, self.frozen_at, self.brk = 0, None, False
        return 0

    def fcx_krylov_iterate(self, h, k, stream):
        self.log.append(("iterate", k))
        self._advance(k)
        return 0

    def fcx_krylov_snapshot(self, h, slot, stream):
        live = (self.frozen_at if self.frozen_at is not None else self.enq - 1)
        flag = 1.0 if getattr(self, "brk", False) else 0.0
        if self.peer_lost_at is not None and self.enq > self.peer_lost_at:
            flag = 2.0
        self.snaps[slot] = (1.0 if self.frozen_at is not None else 0.0, float(self.frozen_at or 0),
                            self._rr(live), self.rr0, flag, float(live))
        self.log.append(("snapshot", slot))
        return 0

    def fcx_krylov_wait_snapshot(self, h, slot, out):
        self.log.append(("wait", slot))
        for k, v in enumerate(self.snaps[slot]):
            out[k] = v
        return 0

    def fcx_krylov_status(self, h, out):
        live = self.enq - 1
        out[0], out[1], out[2], out[3] = float(self.enq), self._rr(live), self.rr0, 0.0
        if self.breakdown_at is not None and live >= self.breakdown_at:
            out[3] = 1.0
        return 0

    def fcx_krylov_solution(self, h, x, stream):
        return 0


def make(lib, lookahead=True, monkeypatch=None):
    import ctypes

    k = KR.DeviceKrylov.__new__(KR.DeviceKrylov)
    k.L, k.handle, k.device = lib, 1, torch.device("cpu")
    k._x = torch.zeros(4, dtype=torch.float64)
    k._status = (ctypes.c_double * 4)()
    k._snap = (ctypes.c_double * 6)()
    k._set_operator = lambda: None
    k.lookahead = lookahead
    k.world = 1
    return k


@pytest.fixture(autouse=True)
def _no_cuda(monkeypatch):
    monkeypatch.setattr(KR.B, "current_stream_ptr", lambda idx: 0)
    monkeypatch.setattr(KR, "check", lambda rc, what="": None if rc == 0 else (_ for _ in ()).throw(RuntimeError(what)))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: type("S", (), {"synchronize": lambda self: None})())


def test_lookahead_reports_the_exact_stopping_iteration_and_stays_one_block_ahead():
    lib = FakeLib(rr0=4.0, rate=0.5)
    k = make(lib)
    rtol = 1e-3  # r.r <= 1e-6 * rr0 first at iteration 20
    x, it, ok, relres, brk = k.solve(torch.zeros(4), torch.ones(4), rtol, 1000, 8)
    stop = math.ceil(math.log(rtol * rtol) / math.log(0.5))
    assert (it, ok, brk) == (stop, True, False)
    assert relres == pytest.approx(math.sqrt(0.5 ** stop)) and relres <= rtol
    # blocks of 8: the test passes inside block 2 (iterations 16..23); block 3 was already out when that was read
    assert [e for e in lib.log if e[0] == "iterate"] == [("iterate", 8)] * 4
    # every wait is for the block BEFORE the one enqueued last, alternating slots
    waits = [e[1] for e in lib.log if e[0] == "wait"]
    assert waits == [0, 1, 0]
    first_wait = lib.log.index(("wait", 0))
    assert lib.log[:first_wait].count(("iterate", 8)) == 2  # two blocks out before the first wait


def test_drained_loop_stops_at_the_end_of_a_block():
    lib = FakeLib(rr0=4.0, rate=0.5)
    k = make(lib, lookahead=False)
    x, it, ok, relres, brk = k.solve(torch.zeros(4), torch.ones(4), 1e-3, 1000, 8)
    assert (it, ok, brk) == (24, True, False)  # first block end at which r.r of the last iteration passes
    assert lib.rtol2 == 0.0  # the device never freezes in this mode


def test_lookahead_max_it_without_convergence():
    lib = FakeLib(rr0=4.0, rate=0.999)
    k = make(lib)
    x, it, ok, relres, brk = k.solve(torch.zeros(4), torch.ones(4), 1e-6, 30, 10)
    assert (ok, brk) == (False, False) and it == 30
    assert [e for e in lib.log if e[0] == "iterate"] == [("iterate", 10)] * 3  # never beyond max_it


def test_lookahead_breakdown_and_zero_rhs_and_lost_peer():
    lib = FakeLib(breakdown_at=13)
    x, it, ok, relres, brk = make(lib).solve(torch.zeros(4), torch.ones(4), 1e-12, 1000, 10)
    assert (it, ok, brk) == (13, False, True)
    lib = FakeLib(rr0=0.0)
    x, it, ok, relres, brk = make(lib).solve(torch.zeros(4), torch.ones(4), 1e-8, 1000, 10)
    assert (ok, relres, brk) == (True, 0.0, False)
    lib = FakeLib(rate=0.999, peer_lost_at=15)
    with pytest.raises(RuntimeError, match="peer rank never arrived"):
        make(lib).solve(torch.zeros(4), torch.ones(4), 1e-8, 1000, 10)
