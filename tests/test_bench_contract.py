"""CPU-side checks of bench.py's contract: the reference arm (the unmodified
reference CPU backend from oracle/_ref) prints one JSON line with the keys the
driver reads, and the product arm refuses to run without a CUDA device (there is
no CPU fallback to fall into)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600, env=env)


def test_reference_arm_line(reference):
    out = _run("--impl", "reference", "--workload", "H2O-64", "--steps", "1", "--warmup", "1", "--cpu-budget", "2")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_uses_all_host_cores_under_torchrun_env(reference):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must still time
    the CPU backend on every core it may use and report that count."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = _run("--impl", "reference", "--gpus", "2", "--workload", "H2O-64", "--steps", "1", "--warmup", "0",
               "--cpu-budget", "1", env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["n_gpus"] == 2
    # the other ranks print nothing and exit 0
    env["RANK"] = "1"
    out = _run("--impl", "reference", "--gpus", "2", "--workload", "H2O-64", env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_product_arm_fails_loudly_without_gpu():
    out = _run("--workload", "H2O-64", "--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert out.returncode != 0
    assert not any(l.startswith("{") for l in out.stdout.splitlines())  # no number without the device


def test_committed_product_line_carries_the_contract_keys():
    """The product arm needs a GPU; its latest line measured on a B200 is committed under
    profiles/ -- the keys the driver and the judge read must all be there."""
    path = os.path.join(ROOT, "profiles", "r02", "bench_h2o256_1gpu_r02e.json")
    d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "create_task_list"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor", "fp64") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["same_config"] is True
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["create_task_list"]["ms"] == min(d["create_task_list"]["all_ms"])
