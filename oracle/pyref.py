"""TEST INFRASTRUCTURE ONLY -- loaders for the two checkers.

* ``load_oracle()``    -> oracle/libgrid_oracle.so (our plain-C restatement)
* ``load_reference()`` -> oracle/_ref/libgrid_ref.so (the UNMODIFIED reference
  grid library built from /root/reference by oracle/Makefile; REF or CPU
  backend selected through its own ``grid_library_set_config``)

Both are driven through the product's generic ctypes binding
(cp2k_b200.grid_api.GridLibrary).  Only tests/, bench.py (reference arm and
cpu_baseline) and __graft_entry__.smoke() import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from cp2k_b200.grid_api import (GRID_BACKEND_CPU, GRID_BACKEND_GPU, GRID_BACKEND_REF, GridLibrary,
                                _COffloadBuffer, _dptr, _iptr, _i32, _ip, _dp)

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libgrid_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgrid_ref.so")
# the same reference library with INTEGRATION.md's dispatcher patch (oracle/patch_dispatcher.py):
# its public API reaches the product library as GRID_BACKEND_B200
REF_B200_SO = os.path.join(HERE, "_ref", "libgrid_ref_b200.so")
GRID_BACKEND_B200 = 15
# the same unmodified reference library built WITH its own CUDA backend (src/grid/gpu, -D__OFFLOAD_CUDA,
# sm_100a): the comparator of SURVEY.md 8(a) row a19.  Needs a GPU to run; its offload_buffers must carry
# pinned host memory and a device buffer (cp2k_b200.grid_api.OffloadBuffer(pinned=True, device=...)).
REF_GPU_SO = os.path.join(HERE, "_ref", "libgrid_ref_gpu.so")


def build(quiet: bool = True) -> None:
    """Compile the C restatement and, when /root/reference is present, the
    reference library (building the checker is not using it)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


class _Counters(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("npts", "nrows", "nplanes", "ntasks", "flops")]


class OracleLibrary(GridLibrary):
    def __init__(self):
        super().__init__(ORACLE_SO, "grid_oracle")
        L = self.lib
        L.grid_oracle_reset_counters.restype = None
        L.grid_oracle_get_counters.restype = None
        L.grid_oracle_get_counters.argtypes = [C.POINTER(_Counters)]
        L.grid_oracle_collocate_pgf_product.restype = None
        L.grid_oracle_collocate_pgf_product.argtypes = (
            [C.c_bool, C.c_int, C.c_int] + [C.c_int] * 4 + [C.c_double] * 3 + [_dptr] * 4
            + [_iptr] * 4 + [C.c_double] + [C.c_int] * 4 + [_dptr, _dptr])
        L.grid_oracle_integrate_pgf_product.restype = None
        L.grid_oracle_integrate_pgf_product.argtypes = (
            [C.c_bool, C.c_bool, C.c_int] + [C.c_int] * 4 + [C.c_double] * 2 + [_dptr] * 4
            + [_iptr] * 4 + [C.c_double] + [C.c_int] * 4 + [_dptr] * 5)

    def reset_counters(self):
        self.lib.grid_oracle_reset_counters()

    def counters(self) -> dict:
        c = _Counters()
        self.lib.grid_oracle_get_counters(C.byref(c))
        return {n: getattr(c, n) for n, _ in _Counters._fields_}


class ReferenceLibrary(GridLibrary):
    """The reference's public, dispatching ABI (src/grid/grid_task_list.h:59-126):
    same arguments as the per-backend one plus ``npts_local`` on every call."""

    def __init__(self, backend: int = GRID_BACKEND_REF, path: str = REF_SO):
        super().__init__(path, "grid")
        self.path = path
        L = self.lib
        L.grid_library_init.restype = None
        L.grid_library_set_config.restype = None
        L.grid_library_set_config.argtypes = [C.c_int, C.c_bool, C.c_bool]
        L.grid_library_init()
        self.set_backend(backend)
        self._collocate.argtypes = [C.c_void_p, C.c_int, C.c_int, _iptr, C.POINTER(_COffloadBuffer),
                                    C.POINTER(C.POINTER(_COffloadBuffer))]
        self._integrate.argtypes = [C.c_void_p, C.c_bool, C.c_int, C.c_int, _iptr,
                                    C.POINTER(_COffloadBuffer), C.POINTER(C.POINTER(_COffloadBuffer)),
                                    C.POINTER(_COffloadBuffer), _dptr, _dptr]
        L.grid_replay.restype = C.c_bool
        L.grid_replay.argtypes = [C.c_char_p, C.c_int, C.c_bool, C.c_bool, C.c_int, C.c_double]

    def set_backend(self, backend: int, validate: bool = False) -> None:
        """Affects task lists created afterwards (src/grid/grid_task_list.c:50-59).
        ``validate`` switches on the dispatcher's own shadow run of the REF backend
        after every call (1e-12 on grids and hab, 1e-8 on forces/virial; it aborts on a
        mismatch, src/grid/grid_task_list.c:225-260, 352-420)."""
        assert backend in (GRID_BACKEND_REF, GRID_BACKEND_CPU) or \
            (backend == GRID_BACKEND_B200 and self.path == REF_B200_SO) or \
            (backend == GRID_BACKEND_GPU and self.path == REF_GPU_SO)
        self.backend = backend
        self.lib.grid_library_set_config(backend, bool(validate), False)

    def _call_collocate(self, tl, func, pab, grids_arr):
        npl = _i32([l.npts_local for l in tl.layouts]).reshape(-1)
        self._collocate(tl.handle, func, tl.nlevels, _ip(npl), C.byref(pab.c), grids_arr)

    def _call_integrate(self, tl, compute_tau, pab_ref, grids_arr, hab, f_ptr, v_ptr):
        npl = _i32([l.npts_local for l in tl.layouts]).reshape(-1)
        self._integrate(tl.handle, compute_tau, tl.natoms, tl.nlevels, _ip(npl), pab_ref, grids_arr,
                        C.byref(hab.c), f_ptr, v_ptr)


_ORACLE = None
_REFS = {}


def load_oracle() -> OracleLibrary:
    global _ORACLE
    if _ORACLE is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _ORACLE = OracleLibrary()
    return _ORACLE


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def have_reference_b200() -> bool:
    return os.path.exists(REF_B200_SO)


def load_reference_b200(validate: bool = True) -> ReferenceLibrary:
    """The patched-dispatcher build, with the B200 backend selected."""
    if "b200" not in _REFS:
        _REFS["b200"] = ReferenceLibrary(GRID_BACKEND_B200, REF_B200_SO)
    lib = _REFS["b200"]
    lib.set_backend(GRID_BACKEND_B200, validate)
    return lib


def have_reference_gpu() -> bool:
    return os.path.exists(REF_GPU_SO)


def load_reference_gpu(device: int = 0, backend: int = GRID_BACKEND_GPU) -> ReferenceLibrary:
    """The reference built with its CUDA backend; `device` is what CP2K's
    offload_set_chosen_device would receive (src/offload/offload_library.c:88)."""
    if "gpu" not in _REFS:
        lib = ReferenceLibrary(backend, REF_GPU_SO)
        lib.lib.offload_set_chosen_device.restype = None
        lib.lib.offload_set_chosen_device.argtypes = [C.c_int]
        _REFS["gpu"] = lib
    lib = _REFS["gpu"]
    lib.lib.offload_set_chosen_device(int(device))
    lib.set_backend(backend)
    return lib


def load_reference(backend: int = GRID_BACKEND_REF) -> ReferenceLibrary:
    """One process-wide instance; ``backend`` is switched through the library's
    own config before the next task list is created."""
    if "lib" not in _REFS:
        _REFS["lib"] = ReferenceLibrary(backend)
    lib = _REFS["lib"]
    lib.set_backend(backend)
    return lib
