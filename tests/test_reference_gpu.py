"""The reference's own CUDA backend (src/grid/gpu, built for sm_100a by
`make -C oracle ref_gpu`) as comparator -- SURVEY.md 8(a) row a19, "the kernel to
beat".  It is driven through the reference's public API with GRID_BACKEND_GPU on
offload_buffers that carry pinned host memory and a device buffer (what
offload_create_buffer hands out in an __OFFLOAD build, src/offload/offload_buffer.c:78-95).
Here it is checked against the B200 backend on the same seeded lists, so that the
timings bench.py quotes for it are timings of a correct run."""
import numpy as np
import pytest

from cp2k_b200.grid_api import OffloadBuffer
from replay import rel_diff
from synth import make_workload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refgpu(b200):
    from oracle import pyref

    if not pyref.have_reference_gpu():
        pytest.skip("oracle/_ref/libgrid_ref_gpu.so not built (needs /root/reference)")
    return pyref.load_reference_gpu(0)


def _dev_buf(n):
    return OffloadBuffer.with_device(n)


@pytest.mark.parametrize("ortho", [True, False], ids=["ortho", "triclinic"])
def test_reference_gpu_matches_b200(b200, refgpu, ortho):
    wl = make_workload(seed=11, natoms=6, max_tasks=1500, orthorhombic=ortho)
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
    for a, b in zip(out["b200"][0], out["refgpu"][0]):
        assert rel_diff(a, b) < 1e-10
    assert rel_diff(out["b200"][1], out["refgpu"][1]) < 1e-10
    assert rel_diff(out["b200"][2], out["refgpu"][2]) < 1e-8
    assert rel_diff(out["b200"][3], out["refgpu"][3]) < 1e-8
