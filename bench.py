#!/usr/bin/env python
"""Benchmark of the grid hot path: one "step" = one SCF iteration's worth of
grid work = grid_collocate_task_list(GRID_FUNC_AB) + grid_integrate_task_list
on the synthetic-but-faithful H2O-N task list (cp2k_b200/workload.py).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--workload H2O-256] [--forces]

Prints ONE JSON line (rank 0).  `value` is model GFLOP/s (algorithmic FP64 flops
of the reference loop nest, SURVEY.md 8(d), divided by the device time of the
step, max over ranks) with everything resident in HBM; `e2e` is the same metric
through the public API with HOST buffers (H2D/D2H inside the timed region);
`ms_per_step` is the seconds-per-SCF-step half of BASELINE.json's metric.
Multi-GPU (strong scaling): the matrix blocks -- and with them the tasks -- are
split over the ranks, every rank collocates onto replicated grids which are
summed with one NCCL all-reduce (the reference's replicated-level mode,
src/pw/realspace_grid_types.F:763-825); integrate needs no exchange.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_PEAK_TFLOPS = 36.1  # measured on this pool's B200 (profiles/microbench/mb.log, DFMA loop)


def load_traffic(workload, direction):
    """DRAM bytes per call of the grid kernels from the newest committed ncu capture
    (profiles/*/traffic_*.json, written by tools/traffic_from_ncu.py); None for
    workloads that were not captured."""
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "traffic_*.json")), reverse=True):
        try:
            t = json.load(open(path))
            if t.get("workload") == workload:  # TZV2P capture
                return t[f"{direction}_bytes_per_call"]
        except Exception:
            continue
    return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0, "_fallback": True}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self.recording = False  # the thread is started early (NVML init); samples count once armed

    def _nvml_loop(self):
        """NVML polled every 10 ms (an nvidia-smi process per sample takes longer than
        a whole step); same fields and the same reason names as the nvidia-smi query."""
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        vmax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [0x8, 0x40, 0x20, 0x4]  # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self._stop_evt.is_set():
            if self.recording:
                r = int(get_reasons(h))
                self.samples.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(vmax)]
                                    + ["Active" if r & b else "Not Active" for b in bits])
            self._stop_evt.wait(0.01)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            pass
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(",")]
                if len(p) >= 6 and self.recording:
                    self.samples.append(p)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def split_blocks(wl, world, rank):
    """Cost-balanced assignment of matrix blocks (atom pairs) to ranks; a task
    follows its block (cf. load_balance_replicated, src/task_list_methods.F:1611)."""
    if world == 1:
        return wl
    import heapq

    t = wl.tasks
    cost = (t["radius_list"] / np.array([np.linalg.norm(wl.layouts[l - 1].dh[0]) for l in t["level_list"]])) ** 3
    bcost = np.bincount(t["block_num_list"] - 1, weights=cost, minlength=wl.nblocks)
    owner = np.zeros(wl.nblocks, dtype=np.int32)
    heap = [(0.0, r) for r in range(world)]
    for b in np.argsort(-bcost):
        load, r = heapq.heappop(heap)
        owner[b] = r
        heapq.heappush(heap, (load + bcost[b], r))
    return wl.subset(owner[t["block_num_list"] - 1] == rank, compact_blocks=True)


def run_reference(args):
    """Reference arm: the reference's own CPU backend (GRID_BACKEND_CPU, OpenMP on
    all host cores) from oracle/_ref, same workload, metric and unit."""
    from cp2k_b200.grid_api import GRID_BACKEND_CPU, OffloadBuffer
    from cp2k_b200.workload import build_h2o_workload
    from oracle import pyref

    if not pyref.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgrid_ref.so not built"}))
        return
    wl = build_h2o_workload(args.workload, basis=args.basis)
    res = cpu_reference_timing(wl, args.steps, args.warmup, args.forces, budget_s=args.cpu_budget, virial=args.virial)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(wl, args, extra={"sample": res["sample"]}),
        "cpu_baseline": {"value": res["value"], "unit": "GFLOP/s", "cores": res["cores"], "kind": "reference",
                         "sample": res["sample"], "same_config": res["same_config"],
                         "dgemm_shim_share": res["dgemm_shim_share"]},
        "e2e": {"value": res["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def reference_gpu_timing(wl, steps, warmup, forces, virial=False, device=0):
    """Times the UNMODIFIED reference CUDA backend (src/grid/gpu compiled for sm_100a,
    oracle/_ref/libgrid_ref_gpu.so) through the reference's public API on the same task list.
    Its API is synchronous and always moves P/H blocks and grids between the caller's pinned
    host buffers and its device buffers (gpu/grid_gpu_context.cu:479-655), so the number is an
    END-TO-END one: compare with our `e2e`, not with the resident `value`."""
    from cp2k_b200.grid_api import OffloadBuffer
    from oracle import pyref

    lib = pyref.load_reference_gpu(device)
    t0 = time.perf_counter()
    tl = wl.create(lib)
    create_s = time.perf_counter() - t0
    pab = wl.random_pab(1, make=OffloadBuffer.with_device)
    grids = wl.new_grids(make=OffloadBuffer.with_device)
    hab = OffloadBuffer.with_device(wl.pab_len)
    f = np.zeros((wl.natoms, 3)) if forces else None
    v = np.zeros((3, 3)) if virial else None
    times, tc, ti = [], [], []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tl.collocate(FUNC, pab, grids)
        t1 = time.perf_counter()
        tl.integrate(TAU, pab if forces else None, grids, hab, f, v)
        t2 = time.perf_counter()
        if i >= warmup:
            times.append(t2 - t0), tc.append(t1 - t0), ti.append(t2 - t1)
    tl.free()
    return {"ms_per_step": float(np.mean(times)) * 1e3, "collocate_ms": float(np.mean(tc)) * 1e3,
            "integrate_ms": float(np.mean(ti)) * 1e3, "create_task_list_s": create_s, "steps": steps,
            "what": "reference CUDA backend (src/grid/gpu, sm_100a build), public API, pinned host buffers in/out"}


def run_reference_gpu(args):
    import torch

    from cp2k_b200.workload import build_h2o_workload
    from oracle import pyref

    if not pyref.have_reference_gpu() or not torch.cuda.is_available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libgrid_ref_gpu.so not built or no GPU"}))
        return
    wl = build_h2o_workload(args.workload, basis=args.basis)
    res = reference_gpu_timing(wl, args.steps, max(args.warmup, 1), args.forces, args.virial)
    ora = pyref.load_oracle()
    flops = model_flops_cpu(wl, ora)
    val = flops / (res["ms_per_step"] * 1e-3) * 1e-9
    grid_bytes = 8 * sum(l.npts_local_total for l in wl.layouts)
    print(json.dumps({
        "impl": "reference-gpu", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(wl, args), "detail": res,
        "e2e": {"value": val, "unit": "GFLOP/s", "ms_per_step": res["ms_per_step"],
                "h2d_bytes_per_step": 8 * wl.pab_len + grid_bytes, "d2h_bytes_per_step": 8 * wl.pab_len + grid_bytes},
    }))


METRIC = "grid collocate+integrate model FP64 GFLOP/s per SCF step (s/SCF-step = ms_per_step/1000)"
FUNC, TAU = 100, False  # GRID_FUNC_AB, no tau; --tau switches to GRID_FUNC_DADB (200) + compute_tau


def bench_config(wl, args, extra=None):
    cfg = {"workload": f"{args.workload} GPW {wl.meta['basis']} cutoff 280 Ry rel_cutoff 30 Ry 4 levels "
                       f"{wl.meta['npts']}; step = collocate({'GRID_FUNC_DADB' if getattr(args, 'tau', False) else 'GRID_FUNC_AB'}) + "
                       f"integrate{'(compute_tau)' if getattr(args, 'tau', False) else ''}"
                       + ("+forces" if args.forces else "") + ("+virial" if getattr(args, "virial", False) else ""),
           "ntasks": wl.ntasks, "nblocks": wl.nblocks, "natoms": wl.natoms,
           "l2": "per-step working set (task records + P/H blocks + grids > 0.5 GB) exceeds the 126 MB L2",
           "parallelism": (f"z-slab rs_grids over {args.gpus} GPU(s), NCCL halo sum/fill ({getattr(args, 'halo', 'library')})"
                           if getattr(args, "decomp", "blocks") == "slab" and args.gpus > 1 else
                           f"blocks/tasks split over {args.gpus} GPU(s), replicated grids, NCCL all-reduce")}
    if extra:
        cfg.update(extra)
    return cfg


def cpu_reference_timing(wl, steps, warmup, forces, budget_s=25.0, virial=False):
    """Times the unmodified reference CPU backend on a bounded sample of `wl`."""
    from cp2k_b200.grid_api import GRID_BACKEND_CPU, OffloadBuffer
    from oracle import pyref

    # All host cores this process may use, whatever OMP_NUM_THREADS says (torchrun exports
    # OMP_NUM_THREADS=1 to its workers); `cores` is what the OpenMP runtime then reports.
    # Set BEFORE the reference library initialises: grid_library_init sizes its per-thread
    # state from omp_get_max_threads() (src/grid/common/grid_library.c:53-67).
    import ctypes

    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    gomp = ctypes.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(int(avail))
    cores = int(gomp.omp_get_max_threads())
    lib = pyref.load_reference(GRID_BACKEND_CPU)
    ora = pyref.load_oracle()

    def one(sample_wl, nsteps, nwarm):
        tl = sample_wl.create(lib)
        pab = sample_wl.random_pab(1)
        grids = sample_wl.new_grids()
        hab = OffloadBuffer(sample_wl.pab_len)
        f = np.zeros((sample_wl.natoms, 3)) if forces else None
        v = np.zeros((3, 3)) if virial else None
        times = []
        for i in range(nwarm + nsteps):
            t0 = time.perf_counter()
            tl.collocate(FUNC, pab, grids)
            tl.integrate(TAU, pab if forces else None, grids, hab, f, v)
            dt = time.perf_counter() - t0
            if i >= nwarm:
                times.append(dt)
        tl.free()
        return float(np.mean(times))

    # calibrate on 1/16 of the blocks, then take the largest sample that fits the budget
    nb = wl.nblocks
    probe = wl.subset(((wl.tasks["block_num_list"] - 1) % 16) == 0)
    t_probe = one(probe, 1, 1)
    per_task = t_probe / max(probe.ntasks, 1)
    full_est = per_task * wl.ntasks
    total_steps = steps + warmup
    frac = min(1.0, max(budget_s, 1e-3) / max(full_est * total_steps, 1e-9))
    if frac >= 0.999:
        sample, desc = wl, f"full task list ({wl.ntasks} tasks)"
    else:
        stride = int(np.ceil(1.0 / frac))
        sample = wl.subset(((wl.tasks["block_num_list"] - 1) % stride) == 0)
        desc = f"every {stride}-th matrix block ({sample.ntasks} of {wl.ntasks} tasks), full-size grids"
    shim = getattr(lib.lib, "ref_shims_dgemm_stats", None)
    sec, calls = ctypes.c_double(0.0), ctypes.c_longlong(0)
    if shim is not None:
        shim.restype = None
        shim(ctypes.byref(sec), ctypes.byref(calls), 1)  # reset
    t_step = one(sample, steps, warmup)
    dgemm_share = None
    if shim is not None:  # thread-seconds inside the link shim over thread-seconds of the timed + warm-up steps
        shim(ctypes.byref(sec), ctypes.byref(calls), 1)
        dgemm_share = float(sec.value) / max(t_step * (steps + warmup) * cores, 1e-12)
    flops = model_flops_cpu(sample, ora)
    return {"value": flops / t_step * 1e-9, "ms_per_step": t_step * 1e3, "cores": cores, "sample": desc,
            "sample_fraction": sample.ntasks / wl.ntasks, "same_config": sample is wl,
            "dgemm_shim_share": dgemm_share}


def model_flops_cpu(wl, ora):
    """Model flops of collocate + integrate (no GPU needed): walks the REF bounds
    with the oracle's counters on a thinned task list and scales up."""
    # the GPU path gets exact counts from the backend; here a 1/64 sample is enough
    n = wl.ntasks
    stride = max(1, n // 20000)
    sub = wl.subset(np.arange(0, n, stride))
    from cp2k_b200.grid_api import OffloadBuffer

    tl = sub.create(ora)
    ora.reset_counters()
    pab = sub.random_pab(1)
    grids = sub.new_grids()
    tl.collocate(100, pab, grids)
    c = ora.counters()
    tl.free()
    # integrate costs one flop per point less than collocate
    coll = c["flops"]
    integ = c["flops"] - c["npts"]
    return (coll + integ) * (n / sub.ntasks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"],
                    help="reference = the reference's CPU backend on the host cores (the contract's reference arm); "
                         "reference-gpu = the reference's own CUDA backend compiled for sm_100a (SURVEY 8(a) a19)")
    ap.add_argument("--workload", default="H2O-256")
    ap.add_argument("--basis", default="TZV2P-GTH", choices=["TZV2P-GTH", "DZVP-MOLOPT-SR-GTH"],
                    help="TZV2P-GTH as shipped in benchmarks/QS/H2O-N.inp; DZVP-MOLOPT-SR-GTH is BASELINE config 2's")
    ap.add_argument("--forces", action="store_true", help="integrate with forces (BASELINE config 3)")
    ap.add_argument("--virial", action="store_true", help="... and the virial (implies --forces)")
    ap.add_argument("--tau", action="store_true",
                    help="meta-GGA step (BASELINE config 4): collocate GRID_FUNC_DADB, integrate with compute_tau")
    ap.add_argument("--cpu-budget", type=float, default=600.0,
                    help="seconds the CPU reference may take over all its steps before its task list is thinned "
                         "(the default covers the full H2O-256 / H2O-1024 lists: same config as the GPU arm)")
    ap.add_argument("--no-verify", action="store_true",
                    help="N > 1: skip the (untimed) comparison of the N-rank grids and H blocks with a "
                         "1-GPU run of the full task list")
    ap.add_argument("--no-reference-gpu", action="store_true",
                    help="skip the reference-CUDA-backend leg of the N=1 line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--allreduce", default="after", choices=["after", "fused", "peer"],
                    help="replicated grids (--decomp blocks, --halo library): sum over the ranks as one grouped NCCL "
                         "all-reduce between collocate and integrate (default), as per-level NCCL all-reduces inside "
                         "the collocate call, or through NVLink peer memory on the copy engines inside the call")
    ap.add_argument("--decomp", default="blocks", choices=["blocks", "slab"],
                    help="multi-GPU decomposition: matrix blocks + replicated grids + all-reduce (default), "
                         "or z-slab rs_grids + NCCL halo sum/fill (cp2k_b200/rsgrid.py)")
    ap.add_argument("--halo", default="library", choices=["library", "torch"],
                    help="with --decomp slab: the library's C-callable NCCL halo exchange (grid_b200_halo_sum / "
                         "_fill, default) or the torch.distributed restatement in cp2k_b200/rsgrid.py")
    ap.add_argument("--slab-compact", action="store_true",
                    help="with --decomp slab: per-rank compacted P/H blocks + all-to-all owner reduction of H")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.forces = args.forces or args.virial
    global FUNC, TAU
    FUNC, TAU = (200, True) if args.tau else (100, False)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    if args.impl == "reference-gpu":
        if rank == 0:
            run_reference_gpu(args)
        return

    import torch
    import torch.distributed as dist

    from cp2k_b200 import OffloadBuffer, load_b200
    from cp2k_b200.workload import build_h2o_workload

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL writes its version banner to stdout while the communicator is created; the
        # contract is ONE JSON line on stdout, so fd 1 points at /dev/null meanwhile
        sys.stdout.flush()
        saved_fd, null_fd = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(null_fd, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
            os.close(null_fd)
    lib = load_b200()
    lib.set_device(local_rank)
    lib.set_kernel_variant(args.variant)

    wl_full = build_h2o_workload(args.workload, basis=args.basis)
    slab_levels, hab_exchange, halo_comm = None, None, None
    if args.decomp == "slab" and world > 1:
        from cp2k_b200 import rsgrid

        slab_levels = rsgrid.make_slab_levels(wl_full, world)
        wl = rsgrid.local_workload(wl_full, slab_levels, rank, world, compact_blocks=args.slab_compact)
        # --slab-compact (opt-in; CPU-tested with gloo, not yet timed on GPUs): the rank's P/H
        # buffers hold its own blocks only and the partial H blocks are summed into their
        # owners with one all-to-all instead of an all-reduce of the whole H buffer
        hab_exchange = rsgrid.HabExchange(wl_full, slab_levels, rank, world) if args.slab_compact else None
        halo_comm = rsgrid.HaloComm(lib, rank, world, dist) if args.halo == "library" else None
    else:
        wl = split_blocks(wl_full, world, rank)
    # replicated grids: the library sums each level over the ranks inside the collocate call
    # (a level's NCCL all-reduce behind its kernels, overlapping the other levels' kernels)
    fused_reduce = (world > 1 and slab_levels is None and args.halo == "library" and args.allreduce in ("fused", "peer"))
    if world > 1 and slab_levels is None and args.halo == "library":
        from cp2k_b200 import rsgrid

        halo_comm = rsgrid.HaloComm(lib, rank, world, dist, stream_ptr=torch.cuda.current_stream().cuda_stream)
        if fused_reduce:
            halo_comm.set_collocate_reduce(True)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    # create is host-heavy (sorting, table building) and the box's host cores are noisy: three
    # creates, the fastest is quoted, all are reported.  The first one of a process also pays CUDA
    # module loading and the OpenMP pool start; CP2K calls create once per MD step, warm.
    create_all_ms, tl = [], None
    for _ in range(3):
        if tl is not None:
            tl.free()
        torch.cuda.synchronize()
        t_create = time.perf_counter()
        tl = wl.create(lib)
        torch.cuda.synchronize()
        create_all_ms.append((time.perf_counter() - t_create) * 1e3)
    create_ms = min(create_all_ms)
    lib.release_cache()  # the builders' scratch, kept for the next create
    table_bytes = free0 - torch.cuda.mem_get_info()[0]  # device memory the task list holds
    st = lib.stats(tl)
    flops_local = st["flops_collocate"] + st["flops_integrate"]
    flops_t = torch.tensor([flops_local, st["flops_collocate"], st["flops_integrate"], st["npts_model"]],
                           dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(flops_t)
    flops_total, flops_coll, flops_int, npts_total = (float(x) for x in flops_t.cpu())

    # ---- resident-in-HBM arm -------------------------------------------------
    pab_h = wl.random_pab(1)
    pab = OffloadBuffer.with_device(wl.pab_len)
    pab.device.copy_(torch.from_numpy(pab_h.host))
    pab.host[:] = pab_h.host
    grids, peer_memory = None, False
    if fused_reduce and args.allreduce == "peer":
        # the grids of all ranks in NVLink peer memory: the sum over the ranks runs on the copy
        # engines beside the other levels' kernels (grid_b200_comm_reduce_grid)
        shared = halo_comm.share_grids([l.npts_local_total for l in wl.layouts])
        if shared is not None:
            grids = [OffloadBuffer(l.npts_local_total, pinned=True, device=t) for l, t in zip(wl.layouts, shared)]
            peer_memory = True
    if grids is None:
        grids = [OffloadBuffer.with_device(l.npts_local_total) for l in wl.layouts]
    hab = OffloadBuffer.with_device(wl.pab_len)
    forces = np.zeros((wl.natoms, 3)) if args.forces else None
    virial = np.zeros((3, 3)) if args.virial else None

    def exchange(gs):
        """The exchange step between collocate and integrate."""
        if world == 1 or fused_reduce:
            return
        if slab_levels is None:
            if halo_comm is not None:  # the four levels as one grouped NCCL operation
                halo_comm.allreduce_levels([g.device for g in gs], torch.cuda.current_stream().cuda_stream)
            else:
                for g in gs:
                    dist.all_reduce(g.device)
            return
        if halo_comm is not None:
            # the library's exchange: ONE grouped NCCL send/recv (+ all-reduce of the replicated
            # levels) for the sum of all levels, the add kernels, one more group for the fill
            ts = [g.device for g in gs]
            halo_comm.halo_sum_levels(ts, slab_levels)
            halo_comm.halo_fill_levels(ts, slab_levels)
            return
        for lay, sl, g in zip(wl.layouts, slab_levels, gs):
            n = lay.npts_local
            t = g.device[: int(n[0]) * int(n[1]) * int(n[2])].view(int(n[2]), int(n[1]), int(n[0]))
            if False:
                pass
            else:
                rsgrid.halo_sum(t, sl, rank, world, dist)   # density: halos -> owners
                rsgrid.halo_fill(t, sl, rank, world, dist)  # potential: owners -> halos

    xt = {"armed": False, "grid_exchange": [], "hab_reduce": []}  # CUDA-event spans of the exchange pieces

    def span(key, fn):
        if not xt["armed"]:
            fn()
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        xt[key].append((a, b))

    def hab_sum():
        if slab_levels is not None:  # a block's tasks may live on several slabs
            if hab_exchange is not None:
                hab_exchange.reduce(hab.device[: wl.pab_len], dist)
            else:
                dist.all_reduce(hab.device)

    def step_resident():
        tl.collocate(FUNC, pab, grids)
        span("grid_exchange", lambda: exchange(grids))
        tl.integrate(TAU, pab if args.forces else None, grids, hab, forces, virial)
        span("hab_reduce", hab_sum)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib.set_device_resident(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.launch_count()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches_per_step = (lib.launch_count() - launches0) // max(args.warmup, 1) if args.warmup else None
    sampler.recording = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.launch_count()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    launches = lib.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    clocks = sampler.finish() if rank == 0 else None

    # ---- per-kernel roofline (device events inside the library) ---------------
    lib.set_timing(True)
    lib.timings()
    xt["armed"] = world > 1
    for _ in range(args.steps):
        step_resident()
    barrier()
    xt["armed"] = False
    tm = lib.timings()
    lib.set_timing(False)
    exchange_ms = {k: sum(a.elapsed_time(b) for a, b in xt[k]) / args.steps for k in ("grid_exchange", "hab_reduce")}
    peaks = load_peaks()
    coll_ms = tm["collocate"][0] / args.steps
    int_ms = tm["integrate"][0] / args.steps
    dom = "collocate" if coll_ms >= int_ms else "integrate"
    dom_ms = max(coll_ms, int_ms)
    dom_flops = st["flops_collocate"] if dom == "collocate" else st["flops_integrate"]
    grid_bytes = 8.0 * sum(l.npts_local_total for l in wl.layouts)
    alg_bytes = grid_bytes + 8.0 * wl.pab_len + 72.0 * wl.ntasks
    achieved_tf = dom_flops / (dom_ms * 1e-3) * 1e-12 if dom_ms > 0 else 0.0
    roofline = {
        "bound": "fp64", "kernel": f"{dom} grid kernels (all levels of one call)",
        "achieved": achieved_tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
        "frac": achieved_tf / FP64_PEAK_TFLOPS,
        "traffic": load_traffic(args.workload, dom) if (world == 1 and args.basis == "TZV2P-GTH") else None,
        "peak_source": "measured DFMA loop, profiles/microbench (MEASURED_PEAKS.json has no FP64 entry)",
        "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (dom_ms * 1e-3) * 1e-9 if dom_ms else 0,
                "peak_gbs": peaks.get("hbm_gbs"), "of": "fallback" if peaks.get("_fallback") else "measured"},
        "phase_ms_per_step": {k: v[0] / args.steps for k, v in tm.items()},
    }

    # ---- end-to-end arm: host buffers through the public API -------------------
    lib.set_device_resident(False)
    pab_e = OffloadBuffer(wl.pab_len, pinned=True)
    pab_e.host[:] = pab_h.host
    grids_e = [OffloadBuffer.with_device(l.npts_local_total) for l in wl.layouts]
    hab_e = OffloadBuffer(wl.pab_len, pinned=True)

    pin_pab = torch.from_numpy(pab_h.host).pin_memory() if world > 1 else None
    pin_hab = torch.empty(wl.pab_len, dtype=torch.float64).pin_memory() if world > 1 else None

    def step_e2e():
        if world == 1:
            # host buffers through the public API: the library pipelines the P/H block
            # copies against its coefficient kernels and copies the grids per level
            tl.collocate(FUNC, pab_e, grids_e)  # H2D pab, kernels, D2H grids
            tl.integrate(TAU, pab_e if args.forces else None, grids_e, hab_e, forces, virial)  # H2D grids, D2H hab
            return
        # N > 1: the grids stay on the device between collocate, the exchange and
        # integrate (device_buffer authoritative); this rank's P blocks come from
        # pinned host memory and its H blocks go back to it every step
        lib.set_device_resident(True)
        if slab_levels is None:
            # through the public API: P and H are host-only buffers (the library pipelines
            # their copies against its coefficient kernels and returns when H has landed)
            tl.collocate(FUNC, pab_e, grids)
            exchange(grids)
            tl.integrate(TAU, pab_e if args.forces else None, grids, hab_e, forces, virial)
            return
        # slab modes: the partial H blocks are summed into their owners on the device first
        pab.device.copy_(pin_pab, non_blocking=True)
        tl.collocate(FUNC, pab, grids)
        exchange(grids)
        tl.integrate(TAU, pab if args.forces else None, grids, hab, forces, virial)
        if slab_levels is not None:
            if hab_exchange is not None:
                hab_exchange.reduce(hab.device[: wl.pab_len], dist)
            else:
                dist.all_reduce(hab.device)
        pin_hab.copy_(hab.device[: wl.pab_len], non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    n_e2e = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        step_e2e()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_ms = float(dt.item()) / n_e2e * 1e3
    # The same definition at every N ("host P in -> host H out, grids wherever the mode keeps
    # them"): at N = 1 also measured with the grids resident, like the N > 1 arm does.
    e2e_ph_ms = None
    if world == 1:
        lib.set_device_resident(True)

        def step_ph():
            # host-only P and H buffers, grids with an authoritative device_buffer: the library
            # pipelines the P / H copies and returns when H has landed
            tl.collocate(FUNC, pab_e, grids)
            tl.integrate(TAU, pab_e if args.forces else None, grids, hab_e, forces, virial)

        step_ph()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_ph()
        e2e_ph_ms = (time.perf_counter() - t0) / n_e2e * 1e3
    if world == 1:
        h2d = 8 * wl.pab_len * (2 if args.forces else 1) + int(grid_bytes)
        d2h = int(grid_bytes) + 8 * wl.pab_len
    else:  # per rank: its P blocks in, its H blocks out (the grids never leave the device)
        h2d = 8 * wl.pab_len
        d2h = 8 * wl.pab_len

    # ---- N > 1: the distributed result against a single-GPU run of the full list (untimed) ----
    parity = None
    if world > 1 and not args.no_verify:
        from cp2k_b200.workload import block_index_map

        lib.set_device_resident(True)
        pab_full = wl_full.random_pab(7)
        idx = block_index_map(wl) if "parent_block_ids" in wl.meta else None
        pab.device[: wl.pab_len].copy_(torch.from_numpy(pab_full.host[idx] if idx is not None else pab_full.host))
        tl.collocate(FUNC, pab, grids)
        exchange(grids)
        tl.integrate(TAU, None, grids, hab, None, None)
        own_hab, own_ref_slice = hab.device[: wl.pab_len], idx
        if slab_levels is not None:
            if hab_exchange is not None:
                own_hab = hab_exchange.reduce(hab.device[: wl.pab_len], dist)[
                    torch.from_numpy(hab_exchange.owned_local_index).to("cuda")]
                own_ref_slice = hab_exchange.owned_global_index
            else:
                dist.all_reduce(hab.device)
        torch.cuda.synchronize()
        # the reference: this rank alone on the whole task list
        if fused_reduce:
            halo_comm.set_collocate_reduce(False)
        tl1 = wl_full.create(lib)
        pab1 = OffloadBuffer.with_device(wl_full.pab_len)
        pab1.device.copy_(torch.from_numpy(pab_full.host))
        grids1 = [OffloadBuffer.with_device(l.npts_local_total) for l in wl_full.layouts]
        hab1 = OffloadBuffer.with_device(wl_full.pab_len)
        tl1.collocate(FUNC, pab1, grids1)
        tl1.integrate(TAU, None, grids1, hab1, None, None)
        torch.cuda.synchronize()

        def rel(a, b):
            if a.numel() == 0:
                return 0.0
            return float(((a - b).abs() / b.abs().clamp(min=1.0)).max().item())

        gerr = 0.0
        for l, (lay, g, g1) in enumerate(zip(wl.layouts, grids, grids1)):
            n, ng = lay.npts_local, wl_full.layouts[l].npts_global
            mine = g.device[: int(n[0]) * int(n[1]) * int(n[2])].view(int(n[2]), int(n[1]), int(n[0]))
            ref = g1.device.view(int(ng[2]), int(ng[1]), int(ng[0]))
            if slab_levels is not None and slab_levels[l].distributed:
                planes = torch.from_numpy(np.asarray(slab_levels[l].local_planes(rank))).to("cuda")
                ref = ref[planes]  # after the halo fill every local plane equals the global one
            gerr = max(gerr, rel(mine, ref))
        ref_h = hab1.device if own_ref_slice is None else hab1.device[
            torch.from_numpy(own_ref_slice).to("cuda") if isinstance(own_ref_slice, np.ndarray) else own_ref_slice]
        herr = rel(own_hab, ref_h)
        errs = torch.tensor([gerr, herr], dtype=torch.float64, device="cuda")
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        parity = {"vs": "1-GPU run of the full task list on every rank, |d|/max(1,|ref|)",
                  "grid_max_rel": float(errs[0]), "hab_max_rel": float(errs[1]), "tol": 1e-10}
        tl1.free()
        del grids1, hab1, pab1
        torch.cuda.empty_cache()
        parity["ok"] = bool(parity["grid_max_rel"] < 1e-10 and parity["hab_max_rel"] < 1e-10)
        if not parity["ok"] and rank == 0:  # reported in the line (and loudly here): the timing of a wrong result is void
            print(f"bench.py: MULTI-GPU PARITY FAILED: {parity}", file=sys.stderr)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyref

        if pyref.have_reference():
            r = cpu_reference_timing(wl_full, 2, 1, args.forces, budget_s=args.cpu_budget, virial=args.virial)
            cpu_baseline = {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": "reference",
                            "sample": r["sample"], "same_config": r["same_config"],
                            "ms_per_step_sample": r["ms_per_step"],
                            "ms_per_step_extrapolated": r["ms_per_step"] / r["sample_fraction"],
                            "dgemm_shim_share": r["dgemm_shim_share"]}

    reference_gpu = None
    if rank == 0 and world == 1 and not args.no_reference_gpu:
        from oracle import pyref

        if pyref.have_reference_gpu():
            tl.free()
            tl = None
            torch.cuda.empty_cache()
            reference_gpu = reference_gpu_timing(wl_full, 3, 1, args.forces, args.virial, device=local_rank)

    if rank == 0:
        line = {
            "metric": METRIC, "value": flops_total / (ms_per_step * 1e-3) * 1e-9, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "s_per_scf_step": ms_per_step * 1e-3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(wl_full, args, extra={"model_gflop_per_step": flops_total * 1e-9,
                                                         "grid_points_per_pass": npts_total,
                                                         "tasks_tiled": st["ntasks_fast"],
                                                         "tasks_generic": st["ntasks_generic"],
                                                         "task_block_pairs": st["npairs"]}),
            "clocks": clocks,
            "e2e": {"value": flops_total / (e2e_ms * 1e-3) * 1e-9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
                    "what": ("host buffers through the public API: P blocks and grids in, grids and H blocks out"
                             if world == 1 else "this rank's P blocks in, its H blocks out, grids stay on the device"),
                    "ms_per_step_p_in_h_out_grids_resident": e2e_ph_ms if world == 1 else e2e_ms},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "reference_gpu": reference_gpu,
            "multi_gpu_parity": parity,
            # what the step spends outside this rank's kernels (rank 0's view): the grid exchange
            # (all-reduce, or halo sum + fill) and the owner reduction of H
            "exchange_ms_per_step": (ms_per_step - sum(v[0] for v in tm.values()) / args.steps) if world > 1 else 0.0,
            "exchange_pieces_ms": exchange_ms if world > 1 else None,
            "grid_sum": (("peer memory (copy engines + step flags), inside collocate" if peer_memory else
                          "NCCL all-reduce per level, inside collocate") if fused_reduce else
                         ("between the calls: " + ("library NCCL, all levels in one group" if halo_comm is not None
                                                   else "torch.distributed") if world > 1 else None)),
            "create_task_list": {"ms": create_ms, "all_ms": create_all_ms,
                                 "device_bytes": int(table_bytes),
                                 "in_steps": create_ms / ms_per_step},
        }
        print(json.dumps(line))
    if tl is not None:
        tl.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
