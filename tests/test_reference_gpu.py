"""The reference's own CUDA backend (src/grid/gpu, built for sm_100a by
`make -C oracle ref_gpu`) as comparator -- SURVEY.md 8(a) row a19, "the kernel to
beat".  It is driven through the reference's public API with GRID_BACKEND_GPU on
offload_buffers that carry pinned host memory and a device buffer (what
offload_create_buffer hands out in an __OFFLOAD build, src/offload/offload_buffer.c:78-95).
Here it is checked on the B200 against the reference's golden vectors and against the
B200 backend on water task lists, so that the timings bench.py quotes for it are timings
of correct runs.

(Found while writing this test: on the dense random-basis lists of tests/synth.py the
reference GPU backend disagrees with the reference's own CPU backend by O(1) on every
grid level, while this backend agrees with the CPU backend to 1e-13; the water lists and
all 13 golden vectors are fine.  tools/dbg/dbg_refgpu2.py reproduces it.)"""
import numpy as np
import pytest

from cp2k_b200.grid_api import OffloadBuffer
from cp2k_b200.workload import build_h2o_workload
from replay import TASK_NAMES, assert_parity, load_task, replay_batched

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refgpu(b200):
    from oracle import pyref

    if not pyref.have_reference_gpu():
        pytest.skip("oracle/_ref/libgrid_ref_gpu.so not built (needs /root/reference)")
    return pyref.load_reference_gpu(0)


def _dev_buf(n):
    return OffloadBuffer.with_device(n)


@pytest.mark.parametrize("name", TASK_NAMES)
def test_reference_gpu_golden_vectors(refgpu, name):
    for collocate in (True, False):
        assert replay_batched(refgpu, load_task(name), collocate, make_buffer=_dev_buf) < 1e-12


@pytest.mark.parametrize("system,natoms", [("H2O-64", 36), ("H2O-64_nonortho", 18)], ids=["ortho", "triclinic"])
def test_reference_gpu_matches_b200(b200, refgpu, system, natoms):
    wl = build_h2o_workload(system, max_atoms=natoms)
    b200.set_device_resident(False)  # host buffers are the source of truth here (library-wide flag)
    b200.set_kernel_variant(0)
    out = {}
    for name, L in (("b200", b200), ("refgpu", refgpu)):
        tl = wl.create(L)
        pab = wl.random_pab(1, make=_dev_buf)
        grids = wl.new_grids(make=_dev_buf)
        tl.collocate(100, pab, grids)
        hab = _dev_buf(wl.pab_len)
        forces, virial = np.zeros((wl.natoms, 3)), np.zeros((3, 3))
        tl.integrate(False, pab, grids, hab, forces, virial)
        out[name] = ([g.host.copy() for g in grids], hab.host.copy(), forces.copy(), virial.copy())
        tl.free()
    for lvl, (a, b) in enumerate(zip(out["b200"][0], out["refgpu"][0])):
        assert_parity(a, b, 1e-10, f"grid level {lvl}")
    assert_parity(out["b200"][1], out["refgpu"][1], 1e-10, "hab")
    assert_parity(out["b200"][2], out["refgpu"][2], 1e-8, "forces")
    assert_parity(out["b200"][3], out["refgpu"][3], 1e-8, "virial")
