"""Drop-in proof: the reference's OWN dispatcher (src/grid/grid_task_list.c) with
INTEGRATION.md's patch applied (oracle/patch_dispatcher.py) selects this backend
as GRID_BACKEND_B200 = 15 through the reference's public API
(grid_create_task_list / grid_collocate_task_list / grid_integrate_task_list /
grid_free_task_list, src/grid/grid_task_list.h:59-126) -- and with the
library's VALIDATE switch on, the reference itself shadows every call with its
REF backend and aborts on a mismatch beyond 1e-12 (grids, hab) / 1e-8 (forces,
virial) (src/grid/grid_task_list.c:225-260, 352-420)."""
import ctypes
import os

import numpy as np
import pytest

from cp2k_b200.grid_api import OffloadBuffer
from replay import TASK_NAMES, load_task, rel_diff, replay_batched
from synth import make_workload


def _pyref():
    from oracle import pyref

    if not pyref.have_reference_b200():
        pytest.skip("oracle/_ref/libgrid_ref_b200.so not built (needs /root/reference)")
    return pyref


def test_patched_dispatcher_links_the_product_library():
    """CPU check: the patched reference library exports the public API and imports the
    four backend entry points from libgrid_b200.so."""
    pyref = _pyref()
    lib = ctypes.CDLL(pyref.REF_B200_SO)
    for s in ("grid_create_task_list", "grid_free_task_list", "grid_collocate_task_list",
              "grid_integrate_task_list", "grid_library_set_config"):
        assert hasattr(lib, s)
    import subprocess

    undefined = subprocess.run(["nm", "-D", "--undefined-only", pyref.REF_B200_SO], capture_output=True,
                               text=True).stdout
    for s in ("grid_b200_create_task_list", "grid_b200_free_task_list", "grid_b200_collocate_task_list",
              "grid_b200_integrate_task_list"):
        assert s in undefined, s


@pytest.mark.gpu
@pytest.mark.parametrize("name", TASK_NAMES)
@pytest.mark.parametrize("collocate", [True, False], ids=["collocate", "integrate"])
def test_golden_vectors_through_reference_dispatcher(b200, name, collocate):
    lib = _pyref().load_reference_b200(validate=False)
    launches = b200.launch_count()
    err = replay_batched(lib, load_task(name), collocate, cycles=3, cycles_per_block=2)
    assert err < 3e-12, err
    assert b200.launch_count() > launches  # the CUDA backend did the work


@pytest.mark.gpu
@pytest.mark.parametrize("ortho", [True, False], ids=["ortho", "triclinic"])
def test_task_list_through_reference_dispatcher(b200, oracle, ortho):
    """A multi-level synthetic list with forces and virial through the public API."""
    lib = _pyref().load_reference_b200(validate=False)
    wl = make_workload(seed=77, natoms=5, max_tasks=400, orthorhombic=ortho)
    pab = wl.random_pab(3)
    out = {}
    for key, L in (("b200", lib), ("oracle", oracle)):
        tl = wl.create(L)
        grids = wl.new_grids()
        tl.collocate(100, pab, grids)
        hab = OffloadBuffer(wl.pab_len)
        f, v = np.zeros((wl.natoms, 3)), np.zeros((3, 3))
        tl.integrate(False, pab, grids, hab, f, v)
        tl.free()
        out[key] = ([g.host.copy() for g in grids], hab.host.copy(), f, v)
    for a, b in zip(out["b200"][0], out["oracle"][0]):
        assert rel_diff(a, b) < 1e-10
    assert rel_diff(out["b200"][1], out["oracle"][1]) < 1e-10
    assert rel_diff(out["b200"][2], out["oracle"][2]) < 1e-8
    assert rel_diff(out["b200"][3], out["oracle"][3]) < 1e-8


_VALIDATE_SCRIPT = """
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from oracle import pyref
from replay import TASK_NAMES, load_task, replay_batched
lib = pyref.load_reference_b200(validate=True)   # the dispatcher shadows every call with REF
worst = 0.0
for name in TASK_NAMES:
    for collocate in (True, False):
        worst = max(worst, replay_batched(lib, load_task(name), collocate))
print("VALIDATE_OK", len(TASK_NAMES), worst)
"""


@pytest.mark.gpu
def test_reference_validate_mode_accepts_the_backend(b200):
    """The reference's own VALIDATE switch: after every call the dispatcher runs its REF
    backend on the same inputs and aborts beyond 1e-12 / 1e-8.  Run in a subprocess: an
    abort() must fail this test, not the session."""
    import subprocess
    import sys

    _pyref()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = _VALIDATE_SCRIPT.format(root=root, tests=os.path.join(root, "tests"))
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "VALIDATE_OK 13" in out.stdout
    assert out.stdout.count("Validated grid collocate") >= 13 and "Validation failure" not in out.stderr


_REPLAY_SCRIPT = """
import glob, os, sys
sys.path.insert(0, {root!r})
from oracle import pyref
lib = pyref.load_reference_b200(validate=False)
files = sorted(glob.glob(os.path.join({tasks!r}, "*.task")))
bad = []
for f in files:
    for collocate in (True, False):
        # grid_unittest.c:51-57 -- one cycle, batched (the dispatcher path), tolerance 1e-12
        if not lib.lib.grid_replay(f.encode(), 1, collocate, True, 1, 1e-12):
            bad.append((os.path.basename(f), collocate))
print("REPLAY_DONE", len(files), bad)
"""


def _write_sample_tasks(directory):
    """The 13 golden vectors back in the reference's `.task` text format (replay.write_task_file)."""
    from replay import write_task_file

    for name in TASK_NAMES:
        write_task_file(load_task(name), os.path.join(str(directory), name + ".task"))


def test_task_file_writer_feeds_the_reference_harness(reference, tmp_path):
    """CPU check of the fixture writer: the unmodified reference's replay harness parses the
    regenerated files and passes all four legs of its unit test (grid_unittest.c:51-57) with
    its own backends."""
    _write_sample_tasks(tmp_path)
    lib = reference.load_reference()
    for name in TASK_NAMES:
        f = os.path.join(str(tmp_path), name + ".task").encode()
        for collocate in (True, False):
            for batch in (True, False):
                assert lib.lib.grid_replay(f, 1, collocate, batch, 1, 1e-12), (name, collocate, batch)


@pytest.mark.gpu
def test_reference_unit_test_harness_on_the_backend(b200, tmp_path):
    """BASELINE config 1: the reference's OWN replay harness (src/grid/grid_replay.c, the
    engine of grid_unittest.x / grid_miniapp.x) reads the 13 sample tasks (regenerated in
    its file format from tests/golden/*.npz) and checks the batched legs of its unit test
    (src/grid/grid_unittest.c:51-57) against the values stored in the files -- with this
    backend doing the work."""
    import subprocess
    import sys

    _pyref()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _write_sample_tasks(tmp_path)
    out = subprocess.run([sys.executable, "-c", _REPLAY_SCRIPT.format(root=root, tasks=str(tmp_path))],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "REPLAY_DONE 13 []" in out.stdout, out.stdout[-3000:]
