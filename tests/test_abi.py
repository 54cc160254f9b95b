"""CPU-side checks of the drop-in boundary: the product library loads and
exports every symbol declared in include/grid_b200.h (no compute calls)."""
import ctypes
import os
import re

from cp2k_b200.grid_api import lib_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "grid_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(grid_b200_[a-z_]+)\s*\(", src)))


def test_header_declares_the_four_task_list_entry_points():
    syms = declared_symbols()
    for s in ("grid_b200_create_task_list", "grid_b200_free_task_list",
              "grid_b200_collocate_task_list", "grid_b200_integrate_task_list"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(lib_path()), "build the backend first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(lib_path())
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/grid_b200.h but not exported"


def test_struct_layouts_match_reference_structs():
    """grid_basis_set: 4 ints + 7 pointers; offload_buffer: size_t + 2 pointers
    (src/grid/common/grid_basis_set.h:14-26, src/offload/offload_buffer.h:16-20)."""
    from cp2k_b200.grid_api import _CBasisSet, _COffloadBuffer

    assert ctypes.sizeof(_CBasisSet) == 4 * 4 + 7 * 8
    assert ctypes.sizeof(_COffloadBuffer) == 3 * 8
    assert _CBasisSet.sphi.offset == 16 + 5 * 8


def test_product_does_not_import_oracle():
    """The product path must never route through the checker."""
    pkg = os.path.join(ROOT, "cp2k_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "grid_oracle" not in txt and "libgrid_ref" not in txt, f
