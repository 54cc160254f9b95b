"""In-process checks of the z-slab halo exchange (cp2k_b200/rsgrid.py) at world sizes
up to 8 -- one thread per rank and a mailbox standing in for torch.distributed's grouped
send/recv -- including slabs thinner than the halo (multi-hop exchange) and slab counts
that do not divide the plane count.  The gloo tests cover the real process groups."""
import queue
import threading

import numpy as np
import pytest
import torch

from cp2k_b200 import rsgrid


class _Mail:
    def __init__(self, world):
        self.q = {(s, d): queue.Queue() for s in range(world) for d in range(world)}


class _FakeDist:
    """The subset of torch.distributed that halo_sum / halo_fill use."""

    def __init__(self, mail, rank):
        self.mail, self.rank = mail, rank

    class _Req:
        def __init__(self, fn):
            self.fn = fn

        def wait(self):
            self.fn()

    def isend(self):  # markers only
        pass

    def irecv(self):
        pass

    def P2POp(self, op, tensor, peer):
        return (op.__name__, tensor, peer)

    def batch_isend_irecv(self, ops):
        reqs = []
        for name, t, peer in ops:  # sends never block: post them all first
            if name == "isend":
                self.mail.q[(self.rank, peer)].put(t.clone())
        for name, t, peer in ops:
            if name == "irecv":
                reqs.append(self._Req(lambda t=t, peer=peer: t.copy_(self.mail.q[(peer, self.rank)].get(timeout=20))))
        return reqs

    def all_reduce(self, t):  # replicated levels are not exercised here
        raise AssertionError("unexpected all_reduce")


def _level(nz, world, border):
    owned = [rsgrid.get_limit(nz, world, r) for r in range(world)]
    return rsgrid.SlabLevel(True, np.array([6, 5, nz]), border, owned)


@pytest.mark.parametrize("nz,world,border", [(40, 2, 7), (72, 6, 13), (70, 8, 9), (64, 8, 20), (45, 7, 6),
                                             (200, 8, 18), (120, 8, 20)])
def test_halo_sum_and_fill(nz, world, border):
    sl = _level(nz, world, border)
    assert max(hi - lo for lo, hi in sl.owned) + 2 * border <= nz
    rng = np.random.default_rng(nz + world)
    # every rank's local grid: random contributions on all its local planes (halo included)
    local = [torch.from_numpy(rng.normal(size=(sl.local_planes(r).size, 5, 6))) for r in range(world)]
    want = np.zeros((nz, 5, 6))
    for r in range(world):
        np.add.at(want, sl.local_planes(r), local[r].numpy())
    mail = _Mail(world)
    errs = []

    def run(r):
        try:
            d = _FakeDist(mail, r)
            rsgrid.halo_sum(local[r], sl, r, world, d)
            lo, hi = sl.owned[r]
            own = rsgrid.owned_view(local[r], sl, r).numpy()
            assert np.allclose(own, want[lo:hi], rtol=0, atol=1e-12), "halo sum"
            assert float(local[r][:border].abs().max()) == 0.0 and float(local[r][border + hi - lo:].abs().max()) == 0.0
            rsgrid.halo_fill(local[r], sl, r, world, d)
            assert np.allclose(local[r].numpy(), want[sl.local_planes(r)], rtol=0, atol=1e-12), "halo fill"
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not errs, errs
    assert all(q.empty() for q in mail.q.values())  # every message was consumed


def test_plans_are_cached_and_consistent():
    sl = _level(64, 8, 20)  # slabs of 8 planes, halo 20: three hops each way
    for r in range(8):
        sends, recvs = rsgrid._rank_plan(sl, r, 8)
        assert rsgrid._rank_plan(sl, r, 8) is sl._plans[(r, 8)]
        # what r sends to p is what p expects from r, message by message
        for peer in range(8):
            out = [b - a for (p, a, b) in sends if p == peer]
            inn = [n for (p, n, _) in rsgrid._rank_plan(sl, peer, 8)[1] if p == r]
            assert out == inn
    # every halo plane goes to exactly one owner
    for r in range(8):
        sends, _ = rsgrid._rank_plan(sl, r, 8)
        assert sum(b - a for _, a, b in sends) == 2 * sl.border


def test_library_halo_plan_matches_python_plan():
    """The C-callable halo exchange (grid_b200_halo_sum / _fill) derives its message plan in
    the library; it must be the plan of cp2k_b200.rsgrid (which the gloo tests verify against
    a full-grid reference): same messages in the same order, same owned runs."""
    from cp2k_b200 import rsgrid
    from cp2k_b200.grid_api import load_b200

    lib = load_b200()
    for nz, world, border in ((80, 2, 9), (72, 6, 14), (126, 8, 20), (200, 8, 21), (40, 3, 5)):
        owned = [rsgrid.get_limit(nz, world, r) for r in range(world)]
        sl = rsgrid.SlabLevel(True, np.array([10, 12, nz]), border, owned)
        if max(hi - lo for lo, hi in owned) + 2 * border > nz:
            continue
        for rank in range(world):
            sends, recvs = rsgrid._rank_plan(sl, rank, world)
            msgs = rsgrid.c_halo_plan(lib, sl, rank, world)
            c_sends = [(dst, a, b) for (src, dst, a, b, runs) in msgs if src == rank]
            c_recvs = [(src, b - a, runs) for (src, dst, a, b, runs) in msgs if dst == rank]
            assert c_sends == [(int(p), int(a), int(b)) for p, a, b in sends]
            assert c_recvs == [(int(p), int(n), [tuple(int(x) for x in r) for r in runs]) for p, n, runs in recvs]
