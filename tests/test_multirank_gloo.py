"""World-size-2 tests of the multi-GPU host logic on CPU (gloo): the task split,
the z-slab rs_grid layouts and the halo sum / halo fill, with the oracle playing
the role of the per-rank backend.  The same code drives NCCL on the GPU box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, mode, result_file):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cp2k_b200.grid_api import OffloadBuffer
    from cp2k_b200 import rsgrid
    from oracle import pyref
    from synth import make_workload
    import bench

    ora = pyref.load_oracle()
    wl = make_workload(seed=41, natoms=5, max_tasks=500, edge=(9.0, 10.0, 16.0),
                       npts_list=((40, 45, 72), (20, 24, 36)))
    pab = wl.random_pab(2)
    # full reference result (every rank computes it; small)
    tl = wl.create(ora)
    full = wl.new_grids()
    tl.collocate(100, pab, full)
    hab_full = OffloadBuffer(wl.pab_len)
    tl.integrate(False, None, full, hab_full)
    tl.free()

    errs = {}
    if mode == "replicated":
        mine = bench.split_blocks(wl, world, rank)
        counts = torch.tensor([mine.ntasks], dtype=torch.int64)
        dist.all_reduce(counts)
        assert int(counts) == wl.ntasks  # a partition: nothing lost, nothing duplicated
        # a rank's P/H buffers hold its own blocks only (compacted)
        from cp2k_b200.workload import block_index_map
        idx = block_index_map(mine)
        sizes = torch.tensor([idx.size], dtype=torch.int64)
        dist.all_reduce(sizes)
        assert int(sizes) == wl.pab_len  # the blocks are partitioned too
        my_pab = OffloadBuffer(mine.pab_len)
        my_pab.host[:] = pab.host[idx]
        tl = mine.create(ora)
        grids = mine.new_grids()
        tl.collocate(100, my_pab, grids)
        for g in grids:
            t = torch.from_numpy(g.host)
            dist.all_reduce(t)
        errs["grid"] = max(float(np.abs(g.host - f.host).max()) for g, f in zip(grids, full))
        my_hab = OffloadBuffer(mine.pab_len)
        tl.integrate(False, None, grids, my_hab)
        hab = OffloadBuffer(wl.pab_len)
        hab.host[idx] = my_hab.host
        t = torch.from_numpy(hab.host)
        dist.all_reduce(t)
        errs["hab"] = float(np.abs(hab.host - hab_full.host).max())
        tl.free()
    else:
        levels = rsgrid.make_slab_levels(wl, world)
        assert any(l.distributed for l in levels)
        if world > 2:  # the case this test is for: halo wider than a slab
            assert any(l.distributed and l.border > min(hi - lo for lo, hi in l.owned) for l in levels)
        mine = rsgrid.local_workload(wl, levels, rank, world)
        counts = torch.tensor([mine.ntasks], dtype=torch.int64)
        dist.all_reduce(counts)
        assert int(counts) == wl.ntasks
        tl = mine.create(ora)
        grids = mine.new_grids()
        tl.collocate(100, pab, grids)
        gerr = 0.0
        for lay, sl, g, f in zip(mine.layouts, levels, grids, full):
            n = lay.npts_local
            t = torch.from_numpy(g.host).view(int(n[2]), int(n[1]), int(n[0]))
            rsgrid.halo_sum(t, sl, rank, world, dist)
            ref = f.host.reshape(int(sl.npts_global[2]), int(n[1]), int(n[0]))
            planes = sl.local_planes(rank)
            own = rsgrid.owned_view(t, sl, rank).numpy()
            if sl.distributed:
                lo, hi = sl.owned[rank]
                gerr = max(gerr, float(np.abs(own - ref[lo:hi]).max()))
                # halo fill: afterwards every local plane equals the global one
                rsgrid.halo_fill(t, sl, rank, world, dist)
                gerr = max(gerr, float(np.abs(t.numpy() - ref[planes]).max()))
            else:
                gerr = max(gerr, float(np.abs(own - ref).max()))
        errs["grid"] = gerr
        hab = OffloadBuffer(wl.pab_len)
        tl.integrate(False, None, grids, hab)  # grids now hold the filled potential
        t = torch.from_numpy(hab.host)
        dist.all_reduce(t)
        errs["hab"] = float(np.abs(hab.host - hab_full.host).max())
        tl.free()
        # the same with per-rank compacted P/H blocks and the owner reduction (one all-to-all)
        from cp2k_b200.workload import block_index_map
        cmine = rsgrid.local_workload(wl, levels, rank, world, compact_blocks=True)
        assert cmine.ntasks == mine.ntasks and cmine.pab_len <= wl.pab_len
        tl = cmine.create(ora)
        chab = OffloadBuffer(cmine.pab_len)
        tl.integrate(False, None, grids, chab)
        tl.free()
        ex = rsgrid.HabExchange(wl, levels, rank, world)
        assert block_index_map(cmine).size == ex.local_len
        own = ex.reduce(torch.from_numpy(chab.host), dist).numpy()[ex.owned_local_index]
        want = hab_full.host[ex.owned_global_index]
        errs["hab"] = max(errs["hab"], float(np.abs(own - want).max()) if own.size else 0.0)
        covered = torch.tensor([ex.owned_len], dtype=torch.int64)
        dist.all_reduce(covered)
        touched = np.unique(wl.tasks["block_num_list"] - 1)
        sizes = np.diff(np.append(wl.block_offsets.astype(np.int64), wl.pab_len))
        assert int(covered) == int(sizes[touched].sum())  # every block some task touches has one owner
    if rank == 0:
        np.save(result_file, np.array([errs["grid"], errs["hab"]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,world", [("replicated", 2), ("slab", 2), ("slab", 6)])
def test_ranks(mode, world, tmp_path):
    """world 6 on a 72-plane level: slabs of 12 planes are thinner than the halo,
    so the halo sum / fill reaches beyond the nearest neighbour."""
    port = 29500 + (os.getpid() % 400) + (0 if mode == "replicated" else 401 * world // 2)
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, port, mode, out), nprocs=world, join=True)
    grid_err, hab_err = np.load(out)
    assert grid_err < 1e-11
    assert hab_err < 1e-10


def test_get_limit_partitions():
    from cp2k_b200.rsgrid import get_limit

    for n in (40, 70, 126, 200):
        for parts in (1, 2, 3, 8):
            spans = [get_limit(n, parts, p) for p in range(parts)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
