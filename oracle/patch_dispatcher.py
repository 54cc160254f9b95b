"""TEST INFRASTRUCTURE ONLY.  Applies INTEGRATION.md's dispatcher patch to a
scratch copy of the reference's two dispatcher files, so that the UNMODIFIED rest
of the reference grid library can be built with the B200 backend plugged in as
GRID_BACKEND_B200 = 15 (oracle/Makefile, target `ref_b200`).

Nothing of the reference is stored in this repository: the script reads
src/grid/grid_task_list.c and src/grid/grid_task_list_internal.h where they lie,
inserts OUR lines at anchored positions and writes the result to a temporary
build directory that the Makefile removes after the compile.  It fails loudly if an anchor is missing, i.e. if
the reference's dispatcher changed shape.

Usage: python patch_dispatcher.py <reference>/src/grid <build dir>
"""
import os
import sys

CREATE_ARM = """
  case GRID_BACKEND_B200:
    grid_b200_create_task_list(
        orthorhombic, ntasks, nlevels, natoms, nkinds, nblocks, block_offsets,
        &atom_positions[0][0], atom_kinds,
        (const grid_b200_basis_set **)basis_sets, level_list, iatom_list,
        jatom_list, iset_list, jset_list, ipgf_list, jpgf_list,
        border_mask_list, block_num_list, radius_list, &rab_list[0][0],
        &npts_global[0][0], &npts_local[0][0], &shift_local[0][0],
        &border_width[0][0], &dh[0][0][0], &dh_inv[0][0][0], &task_list->b200);
    break;
"""
FREE_ARM = """
  if (task_list->b200 != NULL) {
    grid_b200_free_task_list(task_list->b200);
    task_list->b200 = NULL;
  }
"""
COUNTERS = """
    { /* GRID STATISTICS, as gpu/grid_gpu_context.cu:538-552 (host thread) */
      int b200_ortho[20], b200_general[20];
      grid_b200_get_task_counts(task_list->b200, b200_ortho, b200_general);
      for (int lp = 0; lp < 20; lp++) {
        if (b200_ortho[lp] > 0)
          grid_library_counter_add(lp, GRID_BACKEND_B200, %s_ORTHO, b200_ortho[lp]);
        if (b200_general[lp] > 0)
          grid_library_counter_add(lp, GRID_BACKEND_B200, %s_GENERAL, b200_general[lp]);
      }
    }
"""
COLLOCATE_ARM = """
  case GRID_BACKEND_B200:
    grid_b200_collocate_task_list(task_list->b200, func, nlevels,
                                  (const grid_b200_buffer *)pab_blocks,
                                  (grid_b200_buffer **)grids);
""" + COUNTERS % ("GRID_COLLOCATE", "GRID_COLLOCATE") + """    break;
"""
INTEGRATE_ARM = """
  case GRID_BACKEND_B200:
    grid_b200_integrate_task_list(
        task_list->b200, compute_tau, natoms, nlevels,
        (const grid_b200_buffer *)pab_blocks, (const grid_b200_buffer **)grids,
        (grid_b200_buffer *)hab_blocks, forces ? &forces[0][0] : NULL,
        virial ? &virial[0][0] : NULL);
""" + COUNTERS % ("GRID_INTEGRATE", "GRID_INTEGRATE") + """    break;
"""
HEADER_LINES = """#include "grid_b200.h"
#ifndef GRID_BACKEND_B200
#define GRID_BACKEND_B200 15 /* the proper patch adds it to enum grid_backend */
#endif
"""


def insert_before(text: str, anchor: str, new: str, start: int = 0):
    pos = text.find(anchor, start)
    if pos < 0:
        raise SystemExit(f"patch_dispatcher: anchor not found: {anchor!r}")
    return text[:pos] + new + text[pos:], pos + len(new) + len(anchor)


def main(grid_dir: str, out_dir: str) -> None:
    os.makedirs(out_dir, exist_ok=True)
    # -- the internal handle struct: one more backend pointer
    hdr = open(os.path.join(grid_dir, "grid_task_list_internal.h")).read()
    hdr, _ = insert_before(hdr, "typedef struct {", HEADER_LINES)
    hdr, _ = insert_before(hdr, "  // more backends to be added here", "  grid_b200_task_list *b200;\n")
    open(os.path.join(out_dir, "grid_task_list_internal.h"), "w").write(hdr)
    # -- the dispatcher: one arm per entry point, each before the function's `default:`
    # (create, collocate, integrate in file order) and the free before free(npts_local)
    src = open(os.path.join(grid_dir, "grid_task_list.c")).read()
    src, at = insert_before(src, "\n  default:", CREATE_ARM)
    src, at = insert_before(src, "\n  free(task_list->npts_local);", FREE_ARM, at)
    src, at = insert_before(src, "\n  default:", COLLOCATE_ARM, at)
    src, at = insert_before(src, "\n  default:", INTEGRATE_ARM, at)
    if src.count("case GRID_BACKEND_B200") != 3 or src.count("grid_b200_free_task_list") != 1:
        raise SystemExit("patch_dispatcher: unexpected dispatcher layout")
    open(os.path.join(out_dir, "grid_task_list.c"), "w").write(src)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
