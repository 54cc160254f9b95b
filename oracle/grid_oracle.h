/*
 * TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE.
 *
 * Plain-C restatement of the arithmetic of CP2K's grid REF backend
 * (collocate / integrate of Gaussian products on real-space multigrids), used
 * as the parity oracle for the B200 backend.  Only tests/, bench.py's
 * cpu_baseline leg and __graft_entry__.smoke() may load this library.
 *
 * Pinning: tests/test_oracle_golden.py checks this restatement against the 13
 * golden `.task` vectors of the reference (src/grid/sample_tasks, converted to
 * tests/golden/*.npz) and against the unmodified reference REF backend built
 * into oracle/_ref/libgrid_ref.so on seeded multi-task lists.
 *
 * The entry points deliberately carry the per-backend signature of the
 * reference (src/grid/gpu/grid_gpu_task_list.h:25-60) so one harness drives
 * the reference, the oracle and the B200 backend alike.
 */
#ifndef GRID_ORACLE_H
#define GRID_ORACLE_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as grid_basis_set (src/grid/common/grid_basis_set.h:14-26). */
typedef struct {
  int nset, nsgf, maxco, maxpgf;
  int *lmin, *lmax, *npgf, *nsgf_set, *first_sgf;
  double *sphi, *zet;
} oracle_basis_set;

/* Same layout as offload_buffer (src/offload/offload_buffer.h:16-20). */
typedef struct {
  size_t size;
  double *host_buffer;
  double *device_buffer;
} oracle_buffer;

void grid_oracle_create_task_list(
    bool orthorhombic, int ntasks, int nlevels, int natoms, int nkinds,
    int nblocks, const int *block_offsets, const double *atom_positions,
    const int *atom_kinds, const oracle_basis_set **basis_sets,
    const int *level_list, const int *iatom_list, const int *jatom_list,
    const int *iset_list, const int *jset_list, const int *ipgf_list,
    const int *jpgf_list, const int *border_mask_list,
    const int *block_num_list, const double *radius_list,
    const double *rab_list, const int *npts_global, const int *npts_local,
    const int *shift_local, const int *border_width, const double *dh,
    const double *dh_inv, void **task_list);

void grid_oracle_free_task_list(void *task_list);

void grid_oracle_collocate_task_list(const void *task_list, int func,
                                     int nlevels,
                                     const oracle_buffer *pab_blocks,
                                     oracle_buffer **grids);

void grid_oracle_integrate_task_list(const void *task_list, bool compute_tau,
                                     int natoms, int nlevels,
                                     const oracle_buffer *pab_blocks,
                                     const oracle_buffer **grids,
                                     oracle_buffer *hab_blocks, double *forces,
                                     double *virial);

/* Single Gaussian product, cf. src/grid/ref/grid_ref_collocate.h and
 * src/grid/ref/grid_ref_integrate.h (used for the golden .task vectors). */
void grid_oracle_collocate_pgf_product(
    bool orthorhombic, int border_mask, int func, int la_max, int la_min,
    int lb_max, int lb_min, double zeta, double zetb, double rscale,
    const double *dh, const double *dh_inv, const double *ra,
    const double *rab, const int *npts_global, const int *npts_local,
    const int *shift_local, const int *border_width, double radius, int o1,
    int o2, int n1, int n2, const double *pab, double *grid);

void grid_oracle_integrate_pgf_product(
    bool orthorhombic, bool compute_tau, int border_mask, int la_max,
    int la_min, int lb_max, int lb_min, double zeta, double zetb,
    const double *dh, const double *dh_inv, const double *ra,
    const double *rab, const int *npts_global, const int *npts_local,
    const int *shift_local, const int *border_width, double radius, int o1,
    int o2, int n1, int n2, const double *grid, double *hab,
    const double *pab, double *forces /*[2][3] or NULL*/,
    double *virials /*[2][3][3] or NULL*/);

/* Work counters accumulated by every collocate/integrate since the last reset
 * (SURVEY.md 8(d) / Appendix A definition of the algorithmic flop count). */
typedef struct {
  double npts;      /* grid points visited                                  */
  double nrows;     /* (j,k) rows visited                                   */
  double nplanes;   /* k planes visited                                     */
  double ntasks;    /* Gaussian products mapped (not skipped by radius)     */
  double flops;     /* model FP64 flops (FMA = 2) of the REF loop nest      */
} oracle_counters;
void grid_oracle_reset_counters(void);
void grid_oracle_get_counters(oracle_counters *out);

#ifdef __cplusplus
}
#endif
#endif
