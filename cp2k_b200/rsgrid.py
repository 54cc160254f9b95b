"""z-slab distribution of the real-space grids over the ranks (GPUs) of one
box, with the halo exchange that brackets the grid hot path.

Restates, for the 1-D slab case, CP2K's rs_grid layer
(`src/pw/realspace_grid_types.F`):

  descriptor      :285-329 (group_dim choice; for cubic grids and <= 8 ranks a
                  1-D slab wins), :413-415 (get_limit), :514-519 (local bounds =
                  owned +- border), :266-267 (border = (max cube width + 1) / 2)
  task ownership  src/task_list_methods.F:2711-2726 (rank owning the cube centre)
  halo sum        realspace_grid_types.F:988-1204 (after collocate: halo planes
                  are sent to their owners and ADDED, n_shifts rounds when the
                  border is wider than a slab)
  halo fill       :1677-1893 (before integrate: owned planes are copied into the
                  neighbours' halos)
  local layout    src/grid/grid_api.F:501-547 (npts_local, shift_local,
                  border_width as handed to grid_create_task_list)

The exchange is expressed with `torch.distributed` point-to-point ops on
contiguous z-plane ranges (z is the slowest grid index), so the same code runs
over NCCL/NVLink on GPUs and over gloo in the CPU tests.  Levels whose slab plus
two borders would not fit into the global grid stay replicated, exactly like the
reference refuses such layouts (:305-314).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import ctypes as C

import numpy as np

from .grid_api import GridLayout
from .workload import Workload


def get_limit(n: int, nparts: int, part: int) -> Tuple[int, int]:
    """Owned index range [lo, hi) of `part` (cf. get_limit, src/common/util.F):
    the first n % nparts parts get one extra plane."""
    base, extra = divmod(n, nparts)
    lo = part * base + min(part, extra)
    return lo, lo + base + (1 if part < extra else 0)


@dataclass
class SlabLevel:
    distributed: bool
    npts_global: np.ndarray
    border: int
    owned: List[Tuple[int, int]]       # per rank: owned global z-planes [lo, hi)

    def local_planes(self, rank: int) -> np.ndarray:
        """Global z index of every local plane of `rank` (halo included)."""
        nz = int(self.npts_global[2])
        if not self.distributed:
            return np.arange(nz)
        lo, hi = self.owned[rank]
        return np.arange(lo - self.border, hi + self.border) % nz


def cube_halfwidth(wl: Workload, level: int) -> int:
    """Largest cube half-width (grid points, z) on a level, from the discretised
    radius rule of src/grid/ref/grid_ref_collint.h:237-245."""
    lay = wl.layouts[level]
    sel = wl.tasks["level_list"] == level + 1
    if not np.any(sel):
        return 1
    h = np.array([lay.dh[0][0], lay.dh[1][1], lay.dh[2][2]])
    drmin = h.min()
    disr = drmin * np.maximum(1.0, np.ceil(wl.tasks["radius_list"][sel] / drmin))
    lb = np.ceil(-1e-8 - disr.max() * lay.dh_inv[2][2])
    return int(1 - lb)


def _centre_planes(wl: Workload, ilev: int):
    """Tasks of a level and the global z-plane of their cube centres."""
    t = wl.tasks
    lay = wl.layouts[ilev]
    sel = np.nonzero(t["level_list"] == ilev + 1)[0]
    ia = t["iatom_list"][sel] - 1
    ja = t["jatom_list"][sel] - 1
    zeta = _exponents(wl, wl.atom_kinds[ia] - 1, t["iset_list"][sel] - 1, t["ipgf_list"][sel] - 1)
    zetb = _exponents(wl, wl.atom_kinds[ja] - 1, t["jset_list"][sel] - 1, t["jpgf_list"][sel] - 1)
    rp_z = wl.atom_positions[ia, 2] + zetb / (zeta + zetb) * t["rab_list"][sel, 2]
    nz = int(lay.npts_global[2])
    return sel, np.floor(lay.dh_inv[2][2] * rp_z).astype(np.int64) % nz


def balanced_limits(cost_per_plane: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous plane ranges of (nearly) equal summed cost, at least one plane each.  The
    reference cuts a level into slabs of equal thickness (get_limit) and then moves tasks
    between neighbours to balance the load (distribute_tasks / load_balance_distributed,
    src/task_list_methods.F:1889-2057); with one rank per GPU on a node it is simpler to cut
    where the work is: a slab's tasks are the ones whose cube centre it owns."""
    nz = cost_per_plane.size
    cum = np.concatenate([[0.0], np.cumsum(cost_per_plane)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        c = int(np.searchsorted(cum, target, side="left"))
        if c > 0 and abs(cum[c - 1] - target) <= abs(cum[min(c, nz)] - target):
            c -= 1
        c = min(max(c, cuts[-1] + 1), nz - (world - r))
        cuts.append(c)
    cuts.append(nz)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def make_slab_levels(wl: Workload, world: int, balance: bool = True) -> List[SlabLevel]:
    levels = []
    for ilev, lay in enumerate(wl.layouts):
        npts = np.asarray(lay.npts_global, dtype=np.int64)
        nz = int(npts[2])
        border = cube_halfwidth(wl, ilev) + 1
        owned = [get_limit(nz, world, r) for r in range(world)]
        if balance and wl.orthorhombic and world > 1 and nz >= world:
            sel, centre = _centre_planes(wl, ilev)
            if sel.size:
                h = float(lay.dh[0][0])
                cost = np.bincount(centre, weights=(wl.tasks["radius_list"][sel] / h) ** 3, minlength=nz)
                cand = balanced_limits(cost, world)
                if max(hi - lo for lo, hi in cand) + 2 * border <= nz:
                    owned = cand
        thickest = max(hi - lo for lo, hi in owned)
        distributed = wl.orthorhombic and world > 1 and thickest + 2 * border <= nz
        levels.append(SlabLevel(distributed, npts, border, owned))
    return levels


def _exponents(wl: Workload, kinds, isets, ipgfs) -> np.ndarray:
    out = np.zeros(kinds.shape[0])
    for k, b in enumerate(wl.basis_sets):
        m = kinds == k
        out[m] = b.zet[isets[m], ipgfs[m]]
    return out


def block_home(wl: Workload, levels: Sequence[SlabLevel], world: int, owner_dist: np.ndarray) -> np.ndarray:
    """Home rank of every matrix block (atom pair): the slab that works on most of the block's
    tasks of the distributed levels (all of them, unless the pair sits at a slab boundary);
    blocks without such tasks are dealt round-robin.  The block's tasks on the REPLICATED
    levels run at home too and the block's H is owned there, so that only the blocks straddling
    a slab boundary take part in the owner reduction of H (rs_gather_matrices' all-to-all,
    src/task_list_methods.F:2419-2667, moves every block because DBCSR's matrix distribution is
    unrelated to the rs_grid's; here the harness chooses the matrix distribution)."""
    t = wl.tasks
    b = t["block_num_list"] - 1
    votes = np.zeros((wl.nblocks, world), dtype=np.int64)
    on_dist = owner_dist >= 0
    np.add.at(votes, (b[on_dist], owner_dist[on_dist]), 1)
    home = np.argmax(votes, axis=1).astype(np.int32)
    none = votes.sum(axis=1) == 0
    home[none] = (np.nonzero(none)[0] % world).astype(np.int32)
    return home


def task_owner(wl: Workload, levels: Sequence[SlabLevel], world: int) -> np.ndarray:
    """Rank that works on every task.  Distributed levels: the rank owning the z-plane of
    the task's cube centre (the halo is wide enough for the whole cube).  Replicated
    levels (every rank holds the full grid): tasks follow their matrix block to its home
    rank (`block_home`)."""
    t = wl.tasks
    owner = np.full(wl.ntasks, -1, dtype=np.int32)
    for ilev, (lay, sl) in enumerate(zip(wl.layouts, levels)):
        sel = np.nonzero(t["level_list"] == ilev + 1)[0]
        if sl.distributed:
            sel, centre = _centre_planes(wl, ilev)
            bounds = np.array([hi for _, hi in sl.owned])
            owner[sel] = np.searchsorted(bounds, centre, side="right")
    home = block_home(wl, levels, world, owner)
    rest = owner < 0
    owner[rest] = home[t["block_num_list"][rest] - 1]
    return owner


def block_owner(wl: Workload, levels: Sequence[SlabLevel], world: int) -> np.ndarray:
    """Rank owning the H (and P) block of every atom pair in the slab decomposition."""
    t = wl.tasks
    owner = np.full(wl.ntasks, -1, dtype=np.int32)
    for ilev, (lay, sl) in enumerate(zip(wl.layouts, levels)):
        if sl.distributed:
            sel = t["level_list"] == ilev + 1
            owner[sel] = task_owner(wl, levels, world)[sel]
    return block_home(wl, levels, world, owner)


def local_workload(wl: Workload, levels: Sequence[SlabLevel], rank: int, world: int,
                   compact_blocks: bool = False) -> Workload:
    """The task list and grid layouts of one rank (see `task_owner`).  With
    `compact_blocks` the rank's P/H buffers hold only the blocks its tasks refer to
    (see `HabExchange` for summing the H blocks into their owners)."""
    keep = task_owner(wl, levels, world) == rank
    layouts = []
    for lay, sl in zip(wl.layouts, levels):
        if sl.distributed:
            lo, hi = sl.owned[rank]
            nloc = np.array(lay.npts_global, dtype=np.int32)
            nloc[2] = (hi - lo) + 2 * sl.border
            shift = np.array([0, 0, lo - sl.border], dtype=np.int32)
            layouts.append(GridLayout(lay.npts_global, nloc, shift, np.array([0, 0, sl.border], np.int32),
                                      lay.dh, lay.dh_inv))
        else:
            layouts.append(lay)
    sub = wl.subset(keep, compact_blocks=compact_blocks)
    sub.layouts = layouts
    return sub


class HabExchange:
    """Sums the ranks' partial H blocks into the blocks' owners -- the counterpart of
    rs_gather_matrices / rs_scatter_matrices' all-to-all (src/task_list_methods.F:2419-2667)
    for ranks whose buffers hold only the blocks they touch (`compact_blocks=True`).

    A block is owned by its home rank (`block_home`: the slab that computes most of it), so a
    rank's compacted H buffer already holds the owned blocks in place; only the blocks it
    touched but does not own travel: their elements are gathered (one index), exchanged with
    one `all_to_all_single` and added into the owners' buffers through a precomputed index.
    All index arithmetic happens once, here."""

    def __init__(self, wl: Workload, levels: Sequence[SlabLevel], rank: int, world: int):
        sizes = np.diff(np.append(wl.block_offsets.astype(np.int64), wl.pab_len))
        offsets = wl.block_offsets.astype(np.int64)
        towner = task_owner(wl, levels, world)
        b_owner = block_owner(wl, levels, world)
        used = [np.unique(wl.tasks["block_num_list"][towner == r] - 1) for r in range(world)]

        def local_offsets(r):  # where rank r's compacted buffer keeps its blocks
            off = np.full(wl.nblocks, -1, dtype=np.int64)
            off[used[r]] = np.concatenate([[0], np.cumsum(sizes[used[r]])[:-1]]) if used[r].size else 0
            return off

        def elements(off, blocks):  # element indices of `blocks` (ascending) in a compacted buffer
            if blocks.size == 0:
                return np.zeros(0, np.int64)
            return np.repeat(off[blocks] - np.concatenate([[0], np.cumsum(sizes[blocks])[:-1]]), sizes[blocks]) \
                + np.arange(int(sizes[blocks].sum()))

        mine = local_offsets(rank)
        self.local_len = int(sizes[used[rank]].sum())
        # what I send: my blocks owned elsewhere, grouped by owner (ascending block inside a group)
        send_blocks = [used[rank][b_owner[used[rank]] == o] if o != rank else np.zeros(0, np.int64) for o in range(world)]
        self.in_split = [int(sizes[bl].sum()) for bl in send_blocks]
        self.send_index = np.concatenate([elements(mine, bl) for bl in send_blocks]).astype(np.int64)
        # what I receive: from every other rank, the blocks it touched that I own
        recv_blocks = [used[src][b_owner[used[src]] == rank] if src != rank else np.zeros(0, np.int64)
                       for src in range(world)]
        self.out_split = [int(sizes[bl].sum()) for bl in recv_blocks]
        for bl in recv_blocks:  # an owner works on its own blocks (it is their majority slab)
            assert np.all(mine[bl] >= 0), "an owned block is missing from the owner's buffer"
        self.recv_index = np.concatenate([elements(mine, bl) for bl in recv_blocks]).astype(np.int64)
        # the blocks I own: where they sit in my buffer and in the global block buffer
        owned = used[rank][b_owner[used[rank]] == rank]
        self.owned_local_index = elements(mine, owned)
        self.owned_global_index = elements(offsets, owned)
        self.owned_len = int(sizes[owned].sum())
        self._dev = {}

    def _idx(self, name, device):
        key = (name, str(device))
        if key not in self._dev:
            import torch

            self._dev[key] = torch.from_numpy(getattr(self, name)).to(device)
        return self._dev[key]

    def reduce(self, my_hab, dist):
        """`my_hab`: the rank's compacted H buffer (torch tensor), updated IN PLACE: afterwards
        the blocks this rank owns hold the sum over all ranks (elements `owned_local_index`,
        which are the elements `owned_global_index` of the global block buffer).  Returns it."""
        import torch

        assert my_hab.numel() == self.local_len
        send = my_hab.index_select(0, self._idx("send_index", my_hab.device)) if self.send_index.size else \
            torch.empty(0, dtype=my_hab.dtype, device=my_hab.device)
        recv = torch.empty(sum(self.out_split), dtype=my_hab.dtype, device=my_hab.device)
        dist.all_to_all_single(recv, send, self.out_split, self.in_split)
        if self.recv_index.size:
            my_hab.index_add_(0, self._idx("recv_index", my_hab.device), recv)
        return my_hab


# ----------------------------------------------------------------------------
# halo exchange
# ----------------------------------------------------------------------------
def _segments(planes: np.ndarray) -> List[Tuple[int, int]]:
    """Split a sorted array of local plane indices into contiguous [a, b) runs."""
    if planes.size == 0:
        return []
    cuts = np.nonzero(np.diff(planes) != 1)[0] + 1
    return [(int(s[0]), int(s[-1]) + 1) for s in np.split(planes, cuts)]


def _exchange_plan(sl: SlabLevel, world: int):
    """For every ordered pair (src, dst): the local plane indices of src's HALO
    that dst OWNS, and the matching local plane indices on dst."""
    plan = {}
    for src in range(world):
        gsrc = sl.local_planes(src)
        lo_s, hi_s = sl.owned[src]
        nown = hi_s - lo_s
        halo_local = np.concatenate([np.arange(0, sl.border), np.arange(sl.border + nown, gsrc.size)])
        for dst in range(world):
            if dst == src:
                continue
            lo_d, hi_d = sl.owned[dst]
            g = gsrc[halo_local]
            m = (g >= lo_d) & (g < hi_d)
            if not np.any(m):
                continue
            src_idx = halo_local[m]
            dst_idx = g[m] - lo_d + sl.border
            order = np.argsort(src_idx, kind="stable")
            plan[(src, dst)] = (src_idx[order], dst_idx[order])
    return plan


def _p2p(ops, dist):
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def _runs(idx: np.ndarray) -> List[Tuple[int, int, int]]:
    """Split an index array into maximal runs of consecutive values:
    (position in idx, first value, length)."""
    if idx.size == 0:
        return []
    cuts = np.concatenate([[0], np.nonzero(np.diff(idx) != 1)[0] + 1, [idx.size]])
    return [(int(a), int(idx[a]), int(b - a)) for a, b in zip(cuts[:-1], cuts[1:])]


def _rank_plan(sl: SlabLevel, rank: int, world: int):
    """This rank's part of the exchange plan, as contiguous plane ranges, computed once
    per (level, rank, world) and cached on the level: the halo exchange runs every SCF
    step and must not redo index arithmetic on the host.

    sends: [(peer, a, b)]            local HALO planes [a, b) that `peer` owns
    recvs: [(peer, n, [(k, d, m)])]  a message of n planes from `peer`; its planes
                                     [k, k+m) belong to my OWNED local planes [d, d+m)
    Messages between a pair of ranks are posted in the same order on both sides."""
    cache = sl.__dict__.setdefault("_plans", {})
    key = (rank, world)
    if key not in cache:
        plan = _exchange_plan(sl, world)
        sends, recvs = [], []
        for (src, dst), (src_idx, dst_idx) in sorted(plan.items()):
            for a, b in _segments(src_idx):
                i0 = int(np.searchsorted(src_idx, a))
                if src == rank:
                    sends.append((dst, a, b))
                if dst == rank:
                    recvs.append((src, b - a, _runs(dst_idx[i0: i0 + (b - a)])))
        cache[key] = (sends, recvs)
    return cache[key]


def halo_sum(grid, sl: SlabLevel, rank: int, world: int, dist) -> None:
    """After collocate: add every rank's halo planes into their owners
    (realspace_grid_types.F:988-1204).  `grid` is the rank's local grid as a torch
    tensor of shape [nz_local, ny, nx] (CPU for gloo, CUDA for nccl).  Replicated
    levels are summed with one all-reduce (:763-825)."""
    import torch

    if not sl.distributed:
        if world > 1:
            dist.all_reduce(grid)
        return
    sends, recvs = _rank_plan(sl, rank, world)
    ops, bufs = [], []
    for peer, a, b in sends:
        ops.append(dist.P2POp(dist.isend, grid[a:b], peer))  # z-plane ranges are contiguous
    for peer, n, runs in recvs:
        buf = torch.empty((n,) + tuple(grid.shape[1:]), dtype=grid.dtype, device=grid.device)
        ops.append(dist.P2POp(dist.irecv, buf, peer))
        bufs.append((buf, runs))
    _p2p(ops, dist)
    for buf, runs in bufs:
        for k, d, m in runs:
            grid[d: d + m] += buf[k: k + m]
    # the halo has been handed over: zero it so that a later sum is idempotent
    lo, hi = sl.owned[rank]
    grid[: sl.border] = 0
    grid[sl.border + (hi - lo):] = 0


def halo_fill(grid, sl: SlabLevel, rank: int, world: int, dist) -> None:
    """Before integrate: copy the owners' planes into every rank's halo
    (realspace_grid_types.F:1677-1893): the halo sum's plan run backwards."""
    import torch

    if not sl.distributed:
        return
    sends, recvs = _rank_plan(sl, rank, world)  # roles swap: owners send, halo holders receive
    ops, bufs = [], []
    for peer, n, runs in recvs:
        if len(runs) == 1:
            _, d, m = runs[0]
            ops.append(dist.P2POp(dist.isend, grid[d: d + m], peer))
        else:  # the owned range wraps around the periodic boundary
            ops.append(dist.P2POp(dist.isend, torch.cat([grid[d: d + m] for _, d, m in runs]), peer))
    for peer, a, b in sends:
        buf = torch.empty((b - a,) + tuple(grid.shape[1:]), dtype=grid.dtype, device=grid.device)
        ops.append(dist.P2POp(dist.irecv, buf, peer))
        bufs.append((a, b, buf))
    _p2p(ops, dist)
    for a, b, buf in bufs:
        grid[a:b] = buf


def owned_view(grid, sl: SlabLevel, rank: int):
    if not sl.distributed:
        return grid
    lo, hi = sl.owned[rank]
    return grid[sl.border: sl.border + (hi - lo)]


# ----------------------------------------------------------------------------
# The C-callable halo exchange of the library (include/grid_b200.h:
# grid_b200_halo_sum / grid_b200_halo_fill, NCCL send/recv grouped per level).
# ----------------------------------------------------------------------------
class _CSlab(C.Structure):
    _fields_ = [("npts_global", C.c_int * 3), ("nranks", C.c_int), ("rank", C.c_int), ("border", C.c_int),
                ("distributed", C.c_bool), ("owned_lo", C.POINTER(C.c_int)), ("owned_hi", C.POINTER(C.c_int))]


def c_slab(sl: SlabLevel, rank: int, world: int):
    """`grid_b200_slab` for one level; returns (struct, keep-alive arrays)."""
    lo = np.ascontiguousarray([o[0] for o in sl.owned], dtype=np.int32)
    hi = np.ascontiguousarray([o[1] for o in sl.owned], dtype=np.int32)
    cs = _CSlab((C.c_int * 3)(*[int(x) for x in sl.npts_global]), int(world), int(rank), int(sl.border),
                bool(sl.distributed), lo.ctypes.data_as(C.POINTER(C.c_int)), hi.ctypes.data_as(C.POINTER(C.c_int)))
    return cs, (lo, hi)


def c_halo_plan(lib, sl: SlabLevel, rank: int, world: int):
    """The library's exchange plan for `rank` (no GPU needed): list of
    (src, dst, a, b, [(k, d, m), ...])."""
    f = lib.lib.grid_b200_halo_plan
    f.restype = C.c_int
    f.argtypes = [C.POINTER(_CSlab), C.POINTER(C.c_int), C.c_int]
    cs, keep = c_slab(sl, rank, world)
    n = f(C.byref(cs), None, 0)
    out = np.zeros(11 * max(n, 1), dtype=np.int32)
    f(C.byref(cs), out.ctypes.data_as(C.POINTER(C.c_int)), n)
    msgs = []
    for i in range(n):
        o = out[11 * i: 11 * i + 11]
        msgs.append((int(o[0]), int(o[1]), int(o[2]), int(o[3]),
                     [(int(o[5 + 3 * r]), int(o[6 + 3 * r]), int(o[7 + 3 * r])) for r in range(int(o[4]))]))
    return msgs


class HaloComm:
    """An NCCL communicator of the library (`grid_b200_comm`) for the ranks of a
    torch.distributed job: rank 0 draws the unique id, torch.distributed carries it."""

    def __init__(self, lib, rank: int, world: int, dist, stream_ptr: int = 0):
        import torch

        L = lib.lib
        L.grid_b200_comm_unique_id.restype = None
        L.grid_b200_comm_unique_id.argtypes = [C.c_void_p]
        L.grid_b200_comm_create.restype = None
        L.grid_b200_comm_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.grid_b200_comm_destroy.restype = None
        L.grid_b200_comm_destroy.argtypes = [C.c_void_p]
        for name in ("grid_b200_halo_sum", "grid_b200_halo_fill"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p, C.POINTER(_CSlab), C.c_void_p]
        self.L, self.rank, self.world = L, rank, world
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            L.grid_b200_comm_unique_id(uid.ctypes.data_as(C.c_void_p))
        t = torch.from_numpy(uid).cuda()
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()
        self.handle = C.c_void_p()
        L.grid_b200_comm_create(world, rank, uid.ctypes.data_as(C.c_void_p), C.c_void_p(stream_ptr), C.byref(self.handle))
        self._slabs = {}

    def _slab(self, sl: SlabLevel):
        key = id(sl)
        if key not in self._slabs:
            self._slabs[key] = c_slab(sl, self.rank, self.world)
        return self._slabs[key][0]

    def allreduce(self, buf, stream_ptr: int = 0) -> None:
        """Sum of a replicated CUDA float64 tensor over the ranks, on the given stream."""
        f = self.L.grid_b200_comm_allreduce
        f.restype, f.argtypes = None, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        f(self.handle, C.c_void_p(buf.data_ptr()), buf.numel(), C.c_void_p(stream_ptr))

    def share_grids(self, npts):
        """Replicated grids in NVLink peer memory (`grid_b200_comm_share_grids`): this rank's
        grids as CUDA float64 tensors over the library's slab, or None when peer memory is not
        available (collective over the ranks)."""
        import torch

        f = self.L.grid_b200_comm_share_grids
        f.restype, f.argtypes = C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
        n = len(npts)
        sizes = (C.c_size_t * n)(*[int(x) for x in npts])
        ptrs = (C.c_void_p * n)()
        if f(self.handle, n, sizes, ptrs) != 0:
            return None

        class _Dev:  # zero-copy view of library-owned device memory
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        return [torch.as_tensor(_Dev(int(ptrs[i]), int(npts[i])), device="cuda") for i in range(n)]

    def allreduce_levels(self, bufs, stream_ptr: int = 0) -> None:
        """Sums of several replicated CUDA float64 tensors as one grouped NCCL operation."""
        f = self.L.grid_b200_comm_allreduce_levels
        f.restype, f.argtypes = None, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[b.data_ptr() for b in bufs])
        counts = (C.c_size_t * n)(*[b.numel() for b in bufs])
        f(self.handle, n, ptrs, counts, C.c_void_p(stream_ptr))

    def set_collocate_reduce(self, on: bool = True) -> None:
        """Replicated rs_grids: `grid_b200_collocate_task_list` sums every level over the ranks
        itself, each level's all-reduce overlapping the other levels' kernels."""
        f = self.L.grid_b200_set_collocate_reduce
        f.restype, f.argtypes = None, [C.c_void_p]
        f(self.handle if on else None)

    def halo_sum(self, grid, sl: SlabLevel) -> None:
        """`grid`: this rank's local grid of the level, a contiguous CUDA float64 tensor."""
        self.L.grid_b200_halo_sum(self.handle, C.byref(self._slab(sl)), C.c_void_p(grid.data_ptr()))

    def halo_fill(self, grid, sl: SlabLevel) -> None:
        self.L.grid_b200_halo_fill(self.handle, C.byref(self._slab(sl)), C.c_void_p(grid.data_ptr()))

    def _arrays(self, grids, levels):
        n = len(levels)
        slabs = (C.POINTER(_CSlab) * n)(*[C.pointer(self._slab(sl)) for sl in levels])
        ptrs = (C.c_void_p * n)(*[g.data_ptr() for g in grids])
        return n, slabs, ptrs

    def halo_sum_levels(self, grids, levels) -> None:
        """All levels in one grouped NCCL operation (`grids`: the local CUDA tensors per level)."""
        f = self.L.grid_b200_halo_sum_levels
        f.restype, f.argtypes = None, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        n, slabs, ptrs = self._arrays(grids, levels)
        f(self.handle, n, slabs, ptrs)

    def halo_fill_levels(self, grids, levels) -> None:
        f = self.L.grid_b200_halo_fill_levels
        f.restype, f.argtypes = None, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        n, slabs, ptrs = self._arrays(grids, levels)
        f(self.handle, n, slabs, ptrs)

    def destroy(self) -> None:
        if self.handle:
            self.L.grid_b200_comm_destroy(self.handle)
            self.handle = C.c_void_p()

