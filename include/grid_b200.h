/*
 * grid_b200.h -- C ABI of the B200-native backend for CP2K's grid library.
 *
 * Drop-in boundary.  The four task-list entry points have exactly the shape
 * every backend of the reference exports to its dispatcher, so that
 * src/grid/grid_task_list.c can select this backend with one more `case` per
 * switch (see INTEGRATION.md for the patch):
 *
 *   grid_b200_create_task_list     replaces  grid_gpu_create_task_list
 *                                  (src/grid/gpu/grid_gpu_task_list.h:25-38,
 *                                   called at src/grid/grid_task_list.c:116-124)
 *   grid_b200_free_task_list       replaces  grid_gpu_free_task_list   (:42, .c:159-164)
 *   grid_b200_collocate_task_list  replaces  grid_gpu_collocate_task_list
 *                                  (:48-52, called at .c:215-218)
 *   grid_b200_integrate_task_list  replaces  grid_gpu_integrate_task_list
 *                                  (:58-64, called at .c:327-331); `natoms` is
 *                                  passed like the REF/CPU backends do
 *                                  (src/grid/ref/grid_ref_task_list.h) so the
 *                                  forces array can be bounds-checked.
 *
 * Semantics are those of the public API (src/grid/grid_task_list.h:19-126):
 * task indices 1-based, block_offsets 0-based; collocate OVERWRITES every
 * level's grid; integrate OVERWRITES hab_blocks (all `size` bytes),
 * forces[natoms][3] and virial[3][3]; forces/virial may be NULL; a non-NULL
 * *task_list handle is reused.  All functions are void: failures print to
 * stderr and abort(), as in the reference.  Plain pointers and sizes only.
 */
#ifndef GRID_B200_H
#define GRID_B200_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-compatible with grid_basis_set (src/grid/common/grid_basis_set.h:14-26). */
typedef struct {
  int nset;
  int nsgf;
  int maxco;
  int maxpgf;
  int *lmin;
  int *lmax;
  int *npgf;
  int *nsgf_set;
  int *first_sgf;
  double *sphi;
  double *zet;
} grid_b200_basis_set;

/* Layout-compatible with offload_buffer (src/offload/offload_buffer.h:16-20).
 * host_buffer is the source/sink of truth at the call boundary unless
 * grid_b200_set_device_resident(true) was called AND device_buffer != NULL. */
typedef struct {
  size_t size; /* bytes */
  double *host_buffer;
  double *device_buffer;
} grid_b200_buffer;

typedef void grid_b200_task_list;

void grid_b200_create_task_list(
    const bool orthorhombic, const int ntasks, const int nlevels,
    const int natoms, const int nkinds, const int nblocks,
    const int *block_offsets, const double *atom_positions,
    const int *atom_kinds, const grid_b200_basis_set **basis_sets,
    const int *level_list, const int *iatom_list, const int *jatom_list,
    const int *iset_list, const int *jset_list, const int *ipgf_list,
    const int *jpgf_list, const int *border_mask_list,
    const int *block_num_list, const double *radius_list,
    const double *rab_list, const int *npts_global, const int *npts_local,
    const int *shift_local, const int *border_width, const double *dh,
    const double *dh_inv, grid_b200_task_list **task_list);

void grid_b200_free_task_list(grid_b200_task_list *task_list);

/* func: enum grid_func (src/grid/common/grid_constants.h:10-46). */
void grid_b200_collocate_task_list(const grid_b200_task_list *task_list,
                                   const int func, const int nlevels,
                                   const grid_b200_buffer *pab_blocks,
                                   grid_b200_buffer **grids);

void grid_b200_integrate_task_list(const grid_b200_task_list *task_list,
                                   const bool compute_tau, const int natoms,
                                   const int nlevels,
                                   const grid_b200_buffer *pab_blocks,
                                   const grid_b200_buffer **grids,
                                   grid_b200_buffer *hab_blocks, double *forces,
                                   double *virial);

/* ---- ad-hoc Gaussian products ---------------------
 *
 * grid_b200_collocate_pgf_product replaces grid_cpu_collocate_pgf_product
 * (src/grid/cpu/grid_cpu_collocate.h:49-58), the function behind module
 * grid_api's collocate_pgf_product (src/grid/grid_api.F:110-236), with the
 * same arguments in the same order (dh/dh_inv are [3][3], ra/rab [3], the
 * npts/shift/border arrays [3], pab is [n2][n1]); it ADDS the product
 * rscale * sum_ab pab[o2+b][o1+a] g_a g_b (after `func`'s transformation) to
 * `grid` (npts_local doubles, host memory).
 * grid_b200_integrate_pgf_product replaces grid_cpu_integrate_pgf_product
 * (src/grid/cpu/grid_cpu_integrate.h:51-62; grid_api.F:267-490): ADDS the
 * integrals to hab[o2+b][o1+a] and, if `forces` ([2][3]) is non-NULL, the
 * force contributions (needs pab).  `virials`, `hdab`, `hadb` and `a_hdab`
 * must be NULL: the call aborts otherwise (the summed virial is available
 * from grid_b200_integrate_task_list).
 *
 * The *_pgf_products forms take n products for ONE grid in one call (per-
 * product arrays of length n; ra/rab are [n][3]; pab/hab are arrays of n
 * pointers to [n2[p]][n1[p]] matrices; forces is [n][2][3]) -- one task-list
 * build, one upload and one kernel pass for the whole batch, which is how
 * callers that loop over products (src/qs_collocate_density.F:1808-2014)
 * should use a GPU. */
void grid_b200_collocate_pgf_product(
    const bool orthorhombic, const int border_mask, const int func,
    const int la_max, const int la_min, const int lb_max, const int lb_min,
    const double zeta, const double zetb, const double rscale,
    const double *dh, const double *dh_inv, const double *ra,
    const double *rab, const int *npts_global, const int *npts_local,
    const int *shift_local, const int *border_width, const double radius,
    const int o1, const int o2, const int n1, const int n2, const double *pab,
    double *grid);

void grid_b200_integrate_pgf_product(
    const bool orthorhombic, const bool compute_tau, const int border_mask,
    const int la_max, const int la_min, const int lb_max, const int lb_min,
    const double zeta, const double zetb, const double *dh,
    const double *dh_inv, const double *ra, const double *rab,
    const int *npts_global, const int *npts_local, const int *shift_local,
    const int *border_width, const double radius, const int o1, const int o2,
    const int n1, const int n2, const double *grid, double *hab,
    const double *pab, double *forces, double *virials, double *hdab,
    double *hadb, double *a_hdab);

void grid_b200_collocate_pgf_products(
    const int nproducts, const bool orthorhombic, const int func,
    const int *border_mask, const int *la_max, const int *la_min,
    const int *lb_max, const int *lb_min, const double *zeta,
    const double *zetb, const double *rscale, const double *ra,
    const double *rab, const double *radius, const int *o1, const int *o2,
    const int *n1, const int *n2, const double *const *pab, const double *dh,
    const double *dh_inv, const int *npts_global, const int *npts_local,
    const int *shift_local, const int *border_width, double *grid);

void grid_b200_integrate_pgf_products(
    const int nproducts, const bool orthorhombic, const bool compute_tau,
    const int *border_mask, const int *la_max, const int *la_min,
    const int *lb_max, const int *lb_min, const double *zeta,
    const double *zetb, const double *ra, const double *rab,
    const double *radius, const int *o1, const int *o2, const int *n1,
    const int *n2, const double *dh, const double *dh_inv,
    const int *npts_global, const int *npts_local, const int *shift_local,
    const int *border_width, const double *grid, double *const *hab,
    const double *const *pab, double *forces);

/* ---- z-slab distributed grids: halo exchange over NCCL ------------------------
 *
 * The rs_grid distribution restricted to 1-D z-slabs (src/pw/realspace_grid_types.F:413-415,
 * 514-519): rank r owns the global planes [owned_lo[r], owned_hi[r]) of a level and keeps
 * `border` halo planes on either side; its local grid is [owned + 2 border][ny][nx], the
 * `npts_local` / `shift_local` / `border_width` it hands to grid_b200_create_task_list.
 * grid_b200_halo_sum replaces the halo part of transfer_rs2pw_distributed
 * (realspace_grid_types.F:988-1204): after collocate every rank's halo planes are ADDED into
 * the planes of their owners (and zeroed).  grid_b200_halo_fill replaces that of
 * transfer_pw2rs_distributed (:1677-1893): before integrate the owners' planes are COPIED
 * into every halo.  Levels with distributed == false are replicated: the sum is an all-reduce
 * (:763-825), the fill does nothing.  `grid_dev` is device memory; the calls enqueue on the
 * communicator's stream and return (one grouped NCCL send/recv per level plus one add kernel).
 *
 * A communicator wraps ncclCommInitRank: one rank obtains the 128-byte unique id
 * (grid_b200_comm_unique_id) and hands it to the others -- in CP2K over the MPI communicator
 * of the rs_grid (mp_bcast), in the Python harness over torch.distributed. */
typedef struct grid_b200_comm grid_b200_comm;
typedef struct {
  int npts_global[3];
  int nranks, rank;
  int border;          /* halo planes on either side */
  bool distributed;    /* false: replicated level */
  const int *owned_lo; /* [nranks] */
  const int *owned_hi; /* [nranks] */
} grid_b200_slab;

void grid_b200_comm_unique_id(void *out128);
void grid_b200_comm_create(const int nranks, const int rank, const void *unique_id128, void *cuda_stream,
                           grid_b200_comm **comm_out);
void grid_b200_comm_destroy(grid_b200_comm *comm);

/* Sum of a replicated device buffer over the ranks of `comm`, enqueued on `cuda_stream`. */
void grid_b200_comm_allreduce(grid_b200_comm *comm, double *buf_dev, const size_t count, void *cuda_stream);
/* The same for n buffers (the levels of a call) as one grouped NCCL operation. */
void grid_b200_comm_allreduce_levels(grid_b200_comm *comm, const int n, double *const *bufs_dev, const size_t *counts,
                                     void *cuda_stream);

/* Replicated rs_grids (what CP2K uses for small systems and coarse levels,
 * src/pw/realspace_grid_types.F:763-825: every rank collocates its tasks onto full-size grids and
 * the grids are summed over the ranks): with a communicator set here,
 * grid_b200_collocate_task_list sums every level over the ranks itself, in device-resident mode
 * -- a level's sum is enqueued behind that level's kernels, on the level's stream.  NULL (default)
 * switches the sum off.  Measured on 8 B200: the grid kernels are persistent
 * and fill every SM, so an NCCL kernel enqueued this way does not start before the other levels
 * drain -- no faster than one grouped all-reduce after the call, which stays the default. */
void grid_b200_set_collocate_reduce(grid_b200_comm *comm);

/* Replicated grids in NVLink peer memory.  share_grids allocates this rank's grids of all levels
 * (npts[l] doubles each) in one slab that every rank of the node maps (CUDA IPC) and returns the
 * per-level device pointers; the caller uses them as the device_buffer of its grid buffers.
 * reduce_grid then sums such a grid over the ranks without a collective kernel: every rank pulls
 * its chunk of the peers' grids with the copy engines, adds them with one light kernel and pulls
 * the peers' reduced chunks, ordered by step counters in peer memory (cuStreamWaitValue32) --
 * all of it behind the work of `cuda_stream` and beside whatever runs on other streams.  Grids
 * that were not shared go through NCCL.  begin_grid must be enqueued before a shared grid is
 * overwritten again (grid_b200_collocate_task_list does both when set_collocate_reduce is on).
 * share_grids returns non-zero when peer memory is unavailable; collective over the ranks. */
int grid_b200_comm_share_grids(grid_b200_comm *comm, const int nlevels, const size_t *npts, double **grids_dev_out);
void grid_b200_comm_begin_grid(grid_b200_comm *comm, double *grid_dev, void *cuda_stream);
void grid_b200_comm_reduce_grid(grid_b200_comm *comm, double *grid_dev, const size_t count, void *cuda_stream);
void grid_b200_halo_sum(grid_b200_comm *comm, const grid_b200_slab *slab, double *grid_dev);
void grid_b200_halo_fill(grid_b200_comm *comm, const grid_b200_slab *slab, double *grid_dev);
/* All levels of a call in one grouped NCCL operation (what a level's exchange costs is the
 * launch latency of a group, not the NVLink transfer). */
void grid_b200_halo_sum_levels(grid_b200_comm *comm, const int nlevels, const grid_b200_slab *const *slabs,
                               double *const *grids_dev);
void grid_b200_halo_fill_levels(grid_b200_comm *comm, const int nlevels, const grid_b200_slab *const *slabs,
                                double *const *grids_dev);
/* The messages `slab->rank` takes part in, 11 ints each {src, dst, first, last+1 halo plane on
 * src, nruns, then per run: offset in the message, first owned local plane on dst, planes};
 * returns their number (needs no GPU: used to test the plan). */
int grid_b200_halo_plan(const grid_b200_slab *slab, int *out, const int max_msgs);

/* ---- backend controls (no counterpart in the reference) ------------------ */

/* Device selection follows offload_get_chosen_device()
 * (src/offload/offload_library.c); stand-alone it is set explicitly.
 * Negative = keep the CUDA current device. */
void grid_b200_set_device(const int device);
int grid_b200_device_count(void);

/* CUDA stream (cudaStream_t) all work of subsequent calls is enqueued on. */
void grid_b200_set_stream(void *cuda_stream);

/* true: a non-NULL device_buffer is authoritative on entry and exit (no
 * host<->device copies inside the call, the call returns after enqueueing).
 * false (default): host_buffer is authoritative; copies are part of the call. */
void grid_b200_set_device_resident(const bool flag);

/* Kernel family for orthorhombic tasks (the generic per-task kernels always cover what a
 * tiled family cannot): 0 = automatic, 1 = generic kernels everywhere, 2 = warp-tile kernels,
 * 3 = CTA-tile kernels (producer / consumer warps, bulk-async staging).  The family is fixed
 * when a task list is created; 1 can be switched per call.  Used by the parity tests to
 * cover all three. */
void grid_b200_set_kernel_variant(const int variant);

/* Tasks per lp (including the l growth of the list's last collocate / integrate call), for the
 * orthorhombic and the general path: what the dispatcher arm forwards to
 * grid_library_counter_add (src/grid/common/grid_library.c:135-147), like
 * gpu/grid_gpu_context.cu:538-552 does. */
void grid_b200_get_task_counts(const grid_b200_task_list *task_list, int ortho[20], int general[20]);

/* Task lists take their device memory from a caching arena (cudaMalloc / cudaFree of GB-sized
 * tables cost more than building them, and CP2K rebuilds a list of the same size every MD step):
 * freed blocks are kept, up to GRID_B200_CACHE_MB (environment, default 16384), and reused by the
 * next create.  This returns the cached blocks to the driver (the library does so itself when an
 * allocation fails).  No counterpart in the reference. */
void grid_b200_release_cache(void);

/* Number of CUDA kernels launched by this library since load. */
long long grid_b200_get_launch_count(void);

/* Device-side phase timing (CUDA events on the launching stream) for the
 * roofline report.  get_timings drains the recorded spans into
 * out[2*c] = milliseconds, out[2*c+1] = number of spans, for the classes
 * c = 0 pab->coef, 1 collocate grid kernels, 2 integrate grid kernels,
 * 3 coef->hab, 4 host->device copies, 5 device->host copies, 6 memsets;
 * returns the number of classes written (n/2 at most) and resets them. */
void grid_b200_set_timing(const bool flag);
int grid_b200_get_timings(double *out, const int n);

/* Workload statistics of a task list (model flop counts per SURVEY.md 8(d)).
 * Fills at most n doubles, returns how many were written:
 *  [0] ntasks [1] tasks on the tiled path [2] tasks on the generic path
 *  [3] (task,tile) pairs [4] grid points visited per pass (REF bounds)
 *  [5] model flops of collocate(AB) [6] model flops of integrate (no forces)
 *  [7] max lp [8] max cube half-width [9] nlevels [10] nblocks */
int grid_b200_get_stats(const grid_b200_task_list *task_list, double *out,
                        const int n);

#ifdef __cplusplus
}
#endif
#endif /* GRID_B200_H */
