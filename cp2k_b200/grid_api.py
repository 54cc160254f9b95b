"""Host-side mirror of CP2K's Fortran module ``grid_api`` for the B200 backend.

The reference binds the grid library's C ABI from Fortran
(``src/grid/grid_api.F:70-75``: ``grid_create_basis_set``,
``grid_create_task_list``, ``grid_collocate_task_list``,
``grid_integrate_task_list``, ``grid_free_task_list``).  There is no Fortran
compiler in this image, so the same five operations are bound here with ctypes
over the plain C ABI declared in ``include/grid_b200.h`` -- names, argument
order, 1-based index conventions and overwrite semantics are the reference's.

The binding itself is generic over the symbol prefix, because every backend of
the reference exports the same per-backend signature
(``src/grid/gpu/grid_gpu_task_list.h:25-60``).  The product always loads
``libgrid_b200.so``; the test-suite points the very same class at the oracle
libraries to obtain reference results (that happens only under ``tests/``,
``bench.py --impl reference`` / ``cpu_baseline`` and ``smoke()``).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

# enum grid_func -- src/grid/common/grid_constants.h:10-46
GRID_FUNC_AB = 100
GRID_FUNC_DADB = 200
GRID_FUNC_ADBmDAB_X, GRID_FUNC_ADBmDAB_Y, GRID_FUNC_ADBmDAB_Z = 301, 302, 303
GRID_FUNC_ARDBmDARB_XX, GRID_FUNC_ARDBmDARB_XY, GRID_FUNC_ARDBmDARB_XZ = 411, 412, 413
GRID_FUNC_ARDBmDARB_YX, GRID_FUNC_ARDBmDARB_YY, GRID_FUNC_ARDBmDARB_YZ = 421, 422, 423
GRID_FUNC_ARDBmDARB_ZX, GRID_FUNC_ARDBmDARB_ZY, GRID_FUNC_ARDBmDARB_ZZ = 431, 432, 433
GRID_FUNC_DABpADB_X, GRID_FUNC_DABpADB_Y, GRID_FUNC_DABpADB_Z = 501, 502, 503
GRID_FUNC_DX, GRID_FUNC_DY, GRID_FUNC_DZ = 601, 602, 603
GRID_FUNC_DXDY, GRID_FUNC_DYDZ, GRID_FUNC_DZDX = 701, 702, 703
GRID_FUNC_DXDX, GRID_FUNC_DYDY, GRID_FUNC_DZDZ = 801, 802, 803
GRID_FUNC_DAB_X, GRID_FUNC_DAB_Y, GRID_FUNC_DAB_Z = 901, 902, 903
GRID_FUNC_ADB_X, GRID_FUNC_ADB_Y, GRID_FUNC_ADB_Z = 904, 905, 906
GRID_FUNC_CORE_X, GRID_FUNC_CORE_Y, GRID_FUNC_CORE_Z = 1001, 1002, 1003
ALL_GRID_FUNCS = (
    [100, 200] + [301, 302, 303] + [411, 412, 413, 421, 422, 423, 431, 432, 433]
    + [501, 502, 503] + [601, 602, 603] + [701, 702, 703] + [801, 802, 803]
    + [901, 902, 903, 904, 905, 906] + [1001, 1002, 1003]
)

# enum grid_backend -- src/grid/common/grid_constants.h:48-54 (+ the new value)
GRID_BACKEND_AUTO, GRID_BACKEND_REF, GRID_BACKEND_CPU = 10, 11, 12
GRID_BACKEND_DGEMM, GRID_BACKEND_GPU, GRID_BACKEND_B200 = 13, 14, 15

_dptr = C.POINTER(C.c_double)
_iptr = C.POINTER(C.c_int)


class _CBasisSet(C.Structure):
    """``grid_basis_set`` -- src/grid/common/grid_basis_set.h:14-26."""

    _fields_ = [
        ("nset", C.c_int), ("nsgf", C.c_int), ("maxco", C.c_int), ("maxpgf", C.c_int),
        ("lmin", _iptr), ("lmax", _iptr), ("npgf", _iptr), ("nsgf_set", _iptr),
        ("first_sgf", _iptr), ("sphi", _dptr), ("zet", _dptr),
    ]


class _COffloadBuffer(C.Structure):
    """``offload_buffer`` -- src/offload/offload_buffer.h:16-20."""

    _fields_ = [("size", C.c_size_t), ("host_buffer", _dptr), ("device_buffer", _dptr)]


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_iptr)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_dptr)


class BasisSet:
    """Mirror of ``grid_create_basis_set`` (src/grid/grid_api.F:565-664,
    src/grid/common/grid_basis_set.c:19-69).

    ``sphi`` has shape ``[nsgf][maxco]`` (the Fortran ``sphi(maxco, nsgf)``),
    ``zet`` has shape ``[nset][maxpgf]``, ``first_sgf`` is 1-based.
    """

    def __init__(self, lmin, lmax, npgf, nsgf_set, first_sgf, sphi, zet):
        self.lmin, self.lmax, self.npgf = _i32(lmin), _i32(lmax), _i32(npgf)
        self.nsgf_set, self.first_sgf = _i32(nsgf_set), _i32(first_sgf)
        self.sphi, self.zet = _f64(sphi), _f64(zet)
        self.nset = int(self.lmin.shape[0])
        self.nsgf, self.maxco = (int(x) for x in self.sphi.shape)
        self.maxpgf = int(self.zet.shape[1])
        assert self.zet.shape[0] == self.nset
        for arr in (self.lmax, self.npgf, self.nsgf_set, self.first_sgf):
            assert arr.shape == (self.nset,)
        self.c = _CBasisSet(
            self.nset, self.nsgf, self.maxco, self.maxpgf, _ip(self.lmin), _ip(self.lmax),
            _ip(self.npgf), _ip(self.nsgf_set), _ip(self.first_sgf), _dp(self.sphi), _dp(self.zet),
        )


class OffloadBuffer:
    """Mirror of ``offload_buffer`` as the Fortran side uses it
    (``offload_create_buffer``; src/offload/offload_buffer.c:46-96).

    ``host`` is a float64 numpy array (pinned when it was carved out of a pinned
    torch tensor); ``device`` is an optional CUDA torch tensor playing the role
    of ``device_buffer``.  With ``device is None`` the struct carries a NULL
    ``device_buffer`` exactly like a build without ``__OFFLOAD``.
    """

    def __init__(self, length: int, pinned: bool = False, device=None, host: Optional[np.ndarray] = None):
        self._keep = None
        if host is not None:
            assert host.dtype == np.float64 and host.flags.c_contiguous and host.size == length
            self.host = host
        elif pinned:
            import torch

            t = torch.zeros(max(length, 1), dtype=torch.float64, pin_memory=True)
            self._keep = t
            self.host = t.numpy()[:length]
        else:
            self.host = np.zeros(length, dtype=np.float64)
        self.device = device
        dev_ptr = C.cast(C.c_void_p(device.data_ptr()), _dptr) if device is not None else None
        self.c = _COffloadBuffer(8 * length, _dp(self.host) if length > 0 else None, dev_ptr)

    @classmethod
    def with_device(cls, length: int, device: str = "cuda"):
        import torch

        dev = torch.zeros(max(length, 1), dtype=torch.float64, device=device)
        return cls(length, pinned=True, device=dev)

    def __len__(self):
        return self.host.size


@dataclass
class GridLayout:
    """Per-level real-space grid description handed to ``grid_create_task_list``
    (derived from ``realspace_grid_type`` in src/grid/grid_api.F:501-547)."""

    npts_global: Sequence[int]
    npts_local: Sequence[int]
    shift_local: Sequence[int]
    border_width: Sequence[int]
    dh: np.ndarray      # [3][3], row i = i-th lattice vector / npts
    dh_inv: np.ndarray  # [3][3]

    @property
    def npts_local_total(self) -> int:
        n = self.npts_local
        return int(n[0]) * int(n[1]) * int(n[2])


class GridLibrary:
    """A loaded grid backend exposing ``<prefix>_{create,free,collocate,
    integrate}_task_list`` with the reference's per-backend signature."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing -- build it first (python -c 'import __graft_entry__ as g; g.build()')"
            )
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        self._create = getattr(self.lib, f"{prefix}_create_task_list")
        self._free = getattr(self.lib, f"{prefix}_free_task_list")
        self._collocate = getattr(self.lib, f"{prefix}_collocate_task_list")
        self._integrate = getattr(self.lib, f"{prefix}_integrate_task_list")
        self._create.restype = None
        self._create.argtypes = (
            [C.c_bool] + [C.c_int] * 5 + [_iptr, _dptr, _iptr, C.POINTER(C.POINTER(_CBasisSet))]
            + [_iptr] * 9 + [_dptr, _dptr] + [_iptr] * 4 + [_dptr, _dptr, C.POINTER(C.c_void_p)]
        )
        self._free.restype = None
        self._free.argtypes = [C.c_void_p]
        self._collocate.restype = None
        self._collocate.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.POINTER(_COffloadBuffer),
            C.POINTER(C.POINTER(_COffloadBuffer)),
        ]
        self._integrate.restype = None
        self._integrate.argtypes = [
            C.c_void_p, C.c_bool, C.c_int, C.c_int, C.POINTER(_COffloadBuffer),
            C.POINTER(C.POINTER(_COffloadBuffer)), C.POINTER(_COffloadBuffer), _dptr, _dptr,
        ]

    # -- hooks overridden by the reference adapter (public, dispatching ABI) --
    def _call_collocate(self, tl: "TaskList", func, pab, grids_arr):
        self._collocate(tl.handle, func, tl.nlevels, C.byref(pab.c), grids_arr)

    def _call_integrate(self, tl: "TaskList", compute_tau, pab_ref, grids_arr, hab, f_ptr, v_ptr):
        self._integrate(tl.handle, compute_tau, tl.natoms, tl.nlevels, pab_ref, grids_arr,
                        C.byref(hab.c), f_ptr, v_ptr)

    def create_task_list(self, **kw) -> "TaskList":
        return TaskList(self, **kw)


class TaskList:
    """Mirror of ``grid_create_task_list`` / ``grid_collocate_task_list`` /
    ``grid_integrate_task_list`` / ``grid_free_task_list``
    (src/grid/grid_api.F:721-928, 968-1024, 1039-1129, 936-960;
    C contract in src/grid/grid_task_list.h:59-126).

    All per-task indices are 1-based, ``block_offsets`` is 0-based, exactly as
    the Fortran caller passes them.
    """

    def __init__(self, lib: GridLibrary, *, orthorhombic: bool, natoms: int, block_offsets,
                 atom_positions, atom_kinds, basis_sets: List[BasisSet], level_list, iatom_list,
                 jatom_list, iset_list, jset_list, ipgf_list, jpgf_list, border_mask_list,
                 block_num_list, radius_list, rab_list, layouts: List[GridLayout]):
        self.lib = lib
        self.natoms = int(natoms)
        self.nlevels = len(layouts)
        self.layouts = layouts
        self.basis_sets = basis_sets  # must outlive the list (pointers are retained)
        a = {}
        a["block_offsets"] = _i32(block_offsets)
        a["atom_positions"] = _f64(atom_positions).reshape(-1)
        a["atom_kinds"] = _i32(atom_kinds)
        for name, val in (("level", level_list), ("iatom", iatom_list), ("jatom", jatom_list),
                          ("iset", iset_list), ("jset", jset_list), ("ipgf", ipgf_list),
                          ("jpgf", jpgf_list), ("border_mask", border_mask_list),
                          ("block_num", block_num_list)):
            a[name] = _i32(val)
        a["radius"] = _f64(radius_list)
        a["rab"] = _f64(rab_list).reshape(-1)
        self.ntasks = int(a["level"].shape[0])
        self.nblocks = int(a["block_offsets"].shape[0])
        assert a["atom_positions"].size == 3 * self.natoms and a["rab"].size == 3 * self.ntasks
        a["npts_global"] = _i32([l.npts_global for l in layouts]).reshape(-1)
        a["npts_local"] = _i32([l.npts_local for l in layouts]).reshape(-1)
        a["shift_local"] = _i32([l.shift_local for l in layouts]).reshape(-1)
        a["border_width"] = _i32([l.border_width for l in layouts]).reshape(-1)
        a["dh"] = _f64([l.dh for l in layouts]).reshape(-1)
        a["dh_inv"] = _f64([l.dh_inv for l in layouts]).reshape(-1)
        self._arrays = a
        nkinds = len(basis_sets)
        bs_arr = (C.POINTER(_CBasisSet) * max(nkinds, 1))(*[C.pointer(b.c) for b in basis_sets])
        self._bs_arr = bs_arr
        self.handle = C.c_void_p(None)
        lib._create(
            bool(orthorhombic), self.ntasks, self.nlevels, self.natoms, nkinds, self.nblocks,
            _ip(a["block_offsets"]), _dp(a["atom_positions"]), _ip(a["atom_kinds"]), bs_arr,
            _ip(a["level"]), _ip(a["iatom"]), _ip(a["jatom"]), _ip(a["iset"]), _ip(a["jset"]),
            _ip(a["ipgf"]), _ip(a["jpgf"]), _ip(a["border_mask"]), _ip(a["block_num"]),
            _dp(a["radius"]), _dp(a["rab"]), _ip(a["npts_global"]), _ip(a["npts_local"]),
            _ip(a["shift_local"]), _ip(a["border_width"]), _dp(a["dh"]), _dp(a["dh_inv"]),
            C.byref(self.handle),
        )

    # ------------------------------------------------------------------
    def _grid_array(self, grids: Sequence[OffloadBuffer]):
        assert len(grids) == self.nlevels
        for g, l in zip(grids, self.layouts):
            assert len(g) >= l.npts_local_total, "grid buffer smaller than npts_local"
        return (C.POINTER(_COffloadBuffer) * self.nlevels)(*[C.pointer(g.c) for g in grids])

    def collocate(self, func: int, pab_blocks: OffloadBuffer, grids: Sequence[OffloadBuffer]) -> None:
        """grids[level] <- sum over tasks (overwritten, not accumulated)."""
        self.lib._call_collocate(self, int(func), pab_blocks, self._grid_array(grids))

    def integrate(self, compute_tau: bool, pab_blocks: Optional[OffloadBuffer],
                  grids: Sequence[OffloadBuffer], hab_blocks: OffloadBuffer,
                  forces: Optional[np.ndarray] = None, virial: Optional[np.ndarray] = None) -> None:
        """hab_blocks (and forces[natoms][3], virial[3][3] when given) are overwritten."""
        if forces is not None:
            assert forces.dtype == np.float64 and forces.shape == (self.natoms, 3) and forces.flags.c_contiguous
        if virial is not None:
            assert virial.dtype == np.float64 and virial.shape == (3, 3) and virial.flags.c_contiguous
            assert forces is not None, "virial requires forces (src/grid/ref/grid_ref_integrate.c:58)"
        if forces is not None:
            assert pab_blocks is not None, "forces require pab_blocks (src/grid/grid_task_list.c:321-322)"
        pab_ref = C.byref(pab_blocks.c) if pab_blocks is not None else None
        self.lib._call_integrate(
            self, bool(compute_tau), pab_ref, self._grid_array(grids), hab_blocks,
            _dp(forces) if forces is not None else None, _dp(virial) if virial is not None else None,
        )

    def free(self) -> None:
        if self.handle is not None and self.handle.value:
            self.lib._free(self.handle)
        self.handle = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.free()
        except Exception:
            pass


# ----------------------------------------------------------------------
_B200 = None


def lib_path() -> str:
    override = os.environ.get("GRID_B200_LIB")  # an alternative build of the same library (debugging)
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libgrid_b200.so")


def load_b200() -> GridLibrary:
    """Load the product library.  Fails loudly when the CUDA extension has not
    been built -- there is no CPU fallback."""
    global _B200
    if _B200 is None:
        _B200 = B200Library(lib_path())
    return _B200


class B200Library(GridLibrary):
    """``libgrid_b200.so`` plus its backend-specific controls
    (include/grid_b200.h)."""

    def __init__(self, path: str):
        super().__init__(path, "grid_b200")
        L = self.lib
        L.grid_b200_set_device_resident.argtypes = [C.c_bool]
        L.grid_b200_set_device_resident.restype = None
        L.grid_b200_set_stream.argtypes = [C.c_void_p]
        L.grid_b200_set_stream.restype = None
        L.grid_b200_get_launch_count.argtypes = []
        L.grid_b200_get_launch_count.restype = C.c_longlong
        L.grid_b200_release_cache.argtypes = []
        L.grid_b200_release_cache.restype = None
        L.grid_b200_device_count.argtypes = []
        L.grid_b200_device_count.restype = C.c_int
        L.grid_b200_set_kernel_variant.argtypes = [C.c_int]
        L.grid_b200_set_kernel_variant.restype = None
        L.grid_b200_get_stats.argtypes = [C.c_void_p, _dptr, C.c_int]
        L.grid_b200_get_stats.restype = C.c_int
        L.grid_b200_set_device.argtypes = [C.c_int]
        L.grid_b200_set_device.restype = None
        L.grid_b200_set_timing.argtypes = [C.c_bool]
        L.grid_b200_set_timing.restype = None
        L.grid_b200_get_timings.argtypes = [_dptr, C.c_int]
        L.grid_b200_get_timings.restype = C.c_int
        L.grid_b200_collocate_pgf_product.restype = None
        L.grid_b200_collocate_pgf_product.argtypes = (
            [C.c_bool, C.c_int, C.c_int] + [C.c_int] * 4 + [C.c_double] * 3 + [_dptr] * 4
            + [_iptr] * 4 + [C.c_double] + [C.c_int] * 4 + [_dptr, _dptr])
        L.grid_b200_integrate_pgf_product.restype = None
        L.grid_b200_integrate_pgf_product.argtypes = (
            [C.c_bool, C.c_bool, C.c_int] + [C.c_int] * 4 + [C.c_double] * 2 + [_dptr] * 4
            + [_iptr] * 4 + [C.c_double] + [C.c_int] * 4 + [_dptr] * 8)
        _pp = C.POINTER(_dptr)
        L.grid_b200_collocate_pgf_products.restype = None
        L.grid_b200_collocate_pgf_products.argtypes = (
            [C.c_int, C.c_bool, C.c_int] + [_iptr] * 5 + [_dptr] * 6 + [_iptr] * 4 + [_pp]
            + [_dptr] * 2 + [_iptr] * 4 + [_dptr])
        L.grid_b200_integrate_pgf_products.restype = None
        L.grid_b200_integrate_pgf_products.argtypes = (
            [C.c_int, C.c_bool, C.c_bool] + [_iptr] * 5 + [_dptr] * 5 + [_iptr] * 4
            + [_dptr] * 2 + [_iptr] * 4 + [_dptr, _pp, _pp, _dptr])

    # -- ad-hoc Gaussian products: module grid_api's collocate_pgf_product /
    # -- integrate_pgf_product (src/grid/grid_api.F:110-236, 267-490)
    def collocate_pgf_product(self, *, orthorhombic, border_mask, func, la_max, la_min, lb_max, lb_min,
                              zeta, zetb, rscale, layout: "GridLayout", ra, rab, radius, o1, o2, pab,
                              grid: np.ndarray) -> None:
        """ADDS one Gaussian product to ``grid`` (float64, npts_local points)."""
        pab = _f64(pab)
        n2, n1 = pab.shape
        assert grid.dtype == np.float64 and grid.flags.c_contiguous
        ra, rab = _f64(ra), _f64(rab)
        a = [_f64(layout.dh).reshape(-1), _f64(layout.dh_inv).reshape(-1)]
        n = [_i32(layout.npts_global), _i32(layout.npts_local), _i32(layout.shift_local),
             _i32(layout.border_width)]
        self.lib.grid_b200_collocate_pgf_product(
            bool(orthorhombic), int(border_mask), int(func), int(la_max), int(la_min), int(lb_max),
            int(lb_min), float(zeta), float(zetb), float(rscale), _dp(a[0]), _dp(a[1]), _dp(ra), _dp(rab),
            _ip(n[0]), _ip(n[1]), _ip(n[2]), _ip(n[3]), float(radius), int(o1), int(o2), int(n1), int(n2),
            _dp(pab.reshape(-1)), _dp(grid.reshape(-1)))

    def integrate_pgf_product(self, *, orthorhombic, compute_tau, border_mask, la_max, la_min, lb_max,
                              lb_min, zeta, zetb, layout: "GridLayout", ra, rab, radius, o1, o2,
                              grid: np.ndarray, hab: np.ndarray, pab=None,
                              forces: Optional[np.ndarray] = None) -> None:
        """ADDS the product's integrals to ``hab[n2][n1]`` (and its force
        contributions to ``forces[2][3]`` when given; needs ``pab``)."""
        assert hab.dtype == np.float64 and hab.flags.c_contiguous and hab.ndim == 2
        n2, n1 = hab.shape
        ra, rab, grid = _f64(ra), _f64(rab), _f64(grid)
        pabc = _f64(pab).reshape(-1) if pab is not None else None
        if forces is not None:
            assert forces.dtype == np.float64 and forces.flags.c_contiguous and forces.size == 6
            assert pabc is not None and pabc.size == n1 * n2
        a = [_f64(layout.dh).reshape(-1), _f64(layout.dh_inv).reshape(-1)]
        n = [_i32(layout.npts_global), _i32(layout.npts_local), _i32(layout.shift_local),
             _i32(layout.border_width)]
        null = C.cast(None, _dptr)
        self.lib.grid_b200_integrate_pgf_product(
            bool(orthorhombic), bool(compute_tau), int(border_mask), int(la_max), int(la_min), int(lb_max),
            int(lb_min), float(zeta), float(zetb), _dp(a[0]), _dp(a[1]), _dp(ra), _dp(rab), _ip(n[0]),
            _ip(n[1]), _ip(n[2]), _ip(n[3]), float(radius), int(o1), int(o2), int(n1), int(n2),
            _dp(grid.reshape(-1)), _dp(hab.reshape(-1)), _dp(pabc) if pabc is not None else null,
            _dp(forces.reshape(-1)) if forces is not None else null, null, null, null, null)

    @staticmethod
    def _ptr_array(mats):
        arr = (_dptr * len(mats))()
        for i, m in enumerate(mats):
            arr[i] = _dp(m.reshape(-1))
        return arr

    def collocate_pgf_products(self, *, orthorhombic, func, border_mask, la_max, la_min, lb_max, lb_min,
                               zeta, zetb, rscale, layout: "GridLayout", ra, rab, radius, o1, o2, pab,
                               grid: np.ndarray) -> None:
        """Batched form: n products for one grid in ONE device pass.  Per-product
        arguments are sequences of length n; ``pab`` is a list of [n2][n1] arrays."""
        pabs = [_f64(m) for m in pab]
        nprod = len(pabs)
        n1 = _i32([m.shape[1] for m in pabs])
        n2 = _i32([m.shape[0] for m in pabs])
        iv = [_i32(np.broadcast_to(v, (nprod,))) for v in (border_mask, la_max, la_min, lb_max, lb_min)]
        dv = [_f64(np.broadcast_to(v, (nprod,))) for v in (zeta, zetb, rscale)]
        ra, rab = _f64(ra).reshape(nprod, 3), _f64(rab).reshape(nprod, 3)
        rad = _f64(np.broadcast_to(radius, (nprod,)))
        o1, o2 = _i32(np.broadcast_to(o1, (nprod,))), _i32(np.broadcast_to(o2, (nprod,)))
        assert grid.dtype == np.float64 and grid.flags.c_contiguous
        a = [_f64(layout.dh).reshape(-1), _f64(layout.dh_inv).reshape(-1)]
        n = [_i32(layout.npts_global), _i32(layout.npts_local), _i32(layout.shift_local),
             _i32(layout.border_width)]
        self.lib.grid_b200_collocate_pgf_products(
            nprod, bool(orthorhombic), int(func), *[_ip(v) for v in iv], *[_dp(v) for v in dv],
            _dp(ra.reshape(-1)), _dp(rab.reshape(-1)), _dp(rad), _ip(o1), _ip(o2), _ip(n1), _ip(n2),
            self._ptr_array(pabs), _dp(a[0]), _dp(a[1]), _ip(n[0]), _ip(n[1]), _ip(n[2]), _ip(n[3]),
            _dp(grid.reshape(-1)))

    def integrate_pgf_products(self, *, orthorhombic, compute_tau, border_mask, la_max, la_min, lb_max,
                               lb_min, zeta, zetb, layout: "GridLayout", ra, rab, radius, o1, o2,
                               grid: np.ndarray, hab, pab=None, forces: Optional[np.ndarray] = None) -> None:
        """Batched form of integrate_pgf_product; ``hab`` is a list of float64
        [n2][n1] arrays (accumulated in place), ``forces`` is [n][2][3]."""
        nprod = len(hab)
        for m in hab:
            assert m.dtype == np.float64 and m.flags.c_contiguous and m.ndim == 2
        n1 = _i32([m.shape[1] for m in hab])
        n2 = _i32([m.shape[0] for m in hab])
        iv = [_i32(np.broadcast_to(v, (nprod,))) for v in (border_mask, la_max, la_min, lb_max, lb_min)]
        dv = [_f64(np.broadcast_to(v, (nprod,))) for v in (zeta, zetb)]
        ra, rab = _f64(ra).reshape(nprod, 3), _f64(rab).reshape(nprod, 3)
        rad = _f64(np.broadcast_to(radius, (nprod,)))
        o1, o2 = _i32(np.broadcast_to(o1, (nprod,))), _i32(np.broadcast_to(o2, (nprod,)))
        grid = _f64(grid)
        pabs = [_f64(m) for m in pab] if pab is not None else None
        if forces is not None:
            assert forces.dtype == np.float64 and forces.flags.c_contiguous and forces.size == 6 * nprod
            assert pabs is not None
        a = [_f64(layout.dh).reshape(-1), _f64(layout.dh_inv).reshape(-1)]
        n = [_i32(layout.npts_global), _i32(layout.npts_local), _i32(layout.shift_local),
             _i32(layout.border_width)]
        self.lib.grid_b200_integrate_pgf_products(
            nprod, bool(orthorhombic), bool(compute_tau), *[_ip(v) for v in iv], *[_dp(v) for v in dv],
            _dp(ra.reshape(-1)), _dp(rab.reshape(-1)), _dp(rad), _ip(o1), _ip(o2), _ip(n1), _ip(n2),
            _dp(a[0]), _dp(a[1]), _ip(n[0]), _ip(n[1]), _ip(n[2]), _ip(n[3]), _dp(grid.reshape(-1)),
            self._ptr_array(hab), self._ptr_array(pabs) if pabs is not None else C.cast(None, C.POINTER(_dptr)),
            _dp(forces.reshape(-1)) if forces is not None else C.cast(None, _dptr))

    def set_device(self, device: int) -> None:
        self.lib.grid_b200_set_device(int(device))

    def set_timing(self, flag: bool) -> None:
        self.lib.grid_b200_set_timing(bool(flag))

    def timings(self) -> dict:
        """Drain the device-side phase timers: {phase: (milliseconds, spans)}."""
        buf = np.zeros(14, dtype=np.float64)
        n = self.lib.grid_b200_get_timings(_dp(buf), 14)
        names = ["pab_to_coef", "collocate", "integrate", "coef_to_hab", "h2d", "d2h", "memset"]
        return {names[i]: (float(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n)}

    def set_device_resident(self, flag: bool) -> None:
        """When True, non-NULL ``device_buffer`` members are authoritative at
        the call boundary (no H2D/D2H inside the call) -- SURVEY.md 8(f) rank 1."""
        self.lib.grid_b200_set_device_resident(bool(flag))

    def set_stream(self, cuda_stream_ptr: int) -> None:
        self.lib.grid_b200_set_stream(C.c_void_p(cuda_stream_ptr))

    def launch_count(self) -> int:
        return int(self.lib.grid_b200_get_launch_count())

    def release_cache(self) -> None:
        """Return the device blocks freed task lists left in the library's arena to the driver."""
        self.lib.grid_b200_release_cache()

    def device_count(self) -> int:
        return int(self.lib.grid_b200_device_count())

    def set_kernel_variant(self, v: int) -> None:
        self.lib.grid_b200_set_kernel_variant(int(v))

    def stats(self, tl: TaskList) -> dict:
        buf = np.zeros(16, dtype=np.float64)
        n = self.lib.grid_b200_get_stats(tl.handle, _dp(buf), 16)
        keys = ["ntasks", "ntasks_fast", "ntasks_generic", "npairs", "npts_model", "flops_collocate",
                "flops_integrate", "max_lp", "max_cmax", "nlevels", "nblocks"]
        return {k: float(buf[i]) for i, k in enumerate(keys[:n])}
