// Ad-hoc Gaussian-product entry points of the B200 backend.
//
// The reference hard-wires the single-product API of module grid_api
// (collocate_pgf_product / integrate_pgf_product, src/grid/grid_api.F:110-236,
// 267-490) to its CPU backend (grid_cpu_collocate_pgf_product,
// src/grid/cpu/grid_cpu_collocate.h:49-58; grid_cpu_integrate_pgf_product,
// src/grid/cpu/grid_cpu_integrate.h:51-62).  Here the same calls -- and a
// BATCHED form that takes n products for one grid at once, which is how a GPU
// wants them -- run on the device by turning the products into an ad-hoc task
// list: every product side becomes one "atom" whose "basis set" is a single
// Cartesian set with an identity contraction (the construction the reference's
// replay harness uses, src/grid/grid_replay.c:114-214), every product one task
// with its own matrix block.  No CPU arithmetic path: the numbers come from the
// same kernels as grid_b200_collocate_task_list / grid_b200_integrate_task_list.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "../../include/grid_b200.h"

namespace {

inline int ncoset_h(const int l) { return (l < 0) ? 0 : (l + 1) * (l + 2) * (l + 3) / 6; }

[[noreturn]] void die(const char *msg) {
  fprintf(stderr, "grid_b200 (pgf products): %s\n", msg);
  abort();
}

// One Cartesian set (lmin..lmax, one primitive, identity sphi) per distinct
// (lmin, lmax, zet): the "kinds" of the ad-hoc list.
struct AdhocKinds {
  struct Storage {
    int lmin, lmax, npgf, nsgf_set, first_sgf;
    double zet;
    std::vector<double> sphi;
  };
  std::map<std::tuple<int, int, double>, int> index;
  std::vector<Storage *> store;
  std::vector<grid_b200_basis_set> sets;
  ~AdhocKinds() {
    for (Storage *s : store)
      delete s;
  }
  int get(const int lmin, const int lmax, const double zet) {
    const auto key = std::make_tuple(lmin, lmax, zet);
    const auto it = index.find(key);
    if (it != index.end())
      return it->second;
    Storage *s = new Storage;
    const int n = ncoset_h(lmax);
    s->lmin = lmin, s->lmax = lmax, s->npgf = 1, s->nsgf_set = n, s->first_sgf = 1, s->zet = zet;
    s->sphi.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++)
      s->sphi[(size_t)i * n + i] = 1.0;
    store.push_back(s);
    const int id = (int)store.size();  // 1-based kind
    index.emplace(key, id);
    return id;
  }
  void finish() {
    sets.resize(store.size());
    for (size_t k = 0; k < store.size(); k++) {
      Storage *s = store[k];
      grid_b200_basis_set &b = sets[k];
      b.nset = 1, b.nsgf = s->nsgf_set, b.maxco = s->nsgf_set, b.maxpgf = 1;
      b.lmin = &s->lmin, b.lmax = &s->lmax, b.npgf = &s->npgf, b.nsgf_set = &s->nsgf_set;
      b.first_sgf = &s->first_sgf, b.sphi = s->sphi.data(), b.zet = &s->zet;
    }
  }
};

struct AdhocList {
  grid_b200_task_list *list = nullptr;
  std::vector<int> block_offsets;  // per product
  std::vector<int> na, nb;         // ncoset(la_max), ncoset(lb_max) per product
  size_t nblock_doubles = 0;
  ~AdhocList() {
    if (list != nullptr)
      grid_b200_free_task_list(list);
  }
};

void build_adhoc(AdhocList &A, const bool orthorhombic, const int n, const int *border_mask,
                 const int *la_max, const int *la_min, const int *lb_max, const int *lb_min,
                 const double *zeta, const double *zetb, const double *ra, const double *rab,
                 const double *radius, const double *dh, const double *dh_inv, const int *npts_global,
                 const int *npts_local, const int *shift_local, const int *border_width) {
  AdhocKinds kinds;
  std::vector<int> atom_kinds(2 * (size_t)n), ones((size_t)n, 1), iatom((size_t)n), jatom((size_t)n),
      blocknum((size_t)n);
  std::vector<double> pos(6 * (size_t)n);
  A.block_offsets.resize(n), A.na.resize(n), A.nb.resize(n);
  size_t off = 0;
  for (int p = 0; p < n; p++) {
    if (la_min[p] < 0 || lb_min[p] < 0 || la_max[p] < la_min[p] || lb_max[p] < lb_min[p])
      die("invalid angular momentum range");
    atom_kinds[2 * p] = kinds.get(la_min[p], la_max[p], zeta[p]);
    atom_kinds[2 * p + 1] = kinds.get(lb_min[p], lb_max[p], zetb[p]);
    for (int d = 0; d < 3; d++) {
      pos[6 * p + d] = ra[3 * p + d];
      pos[6 * p + 3 + d] = ra[3 * p + d] + rab[3 * p + d];
    }
    iatom[p] = 2 * p + 1, jatom[p] = 2 * p + 2, blocknum[p] = p + 1;
    A.na[p] = ncoset_h(la_max[p]), A.nb[p] = ncoset_h(lb_max[p]);
    if (off > (size_t)0x7fffffff)
      die("too many products in one batch");
    A.block_offsets[p] = (int)off;
    off += (size_t)A.na[p] * A.nb[p];
  }
  A.nblock_doubles = off;
  kinds.finish();
  std::vector<const grid_b200_basis_set *> basis_ptrs(kinds.sets.size());
  for (size_t k = 0; k < kinds.sets.size(); k++)
    basis_ptrs[k] = &kinds.sets[k];
  // (the builder deep-copies every input, so the temporaries may go away)
  grid_b200_create_task_list(orthorhombic, n, 1, 2 * n, (int)basis_ptrs.size(), n, A.block_offsets.data(),
                             pos.data(), atom_kinds.data(), basis_ptrs.data(), ones.data(), iatom.data(),
                             jatom.data(), ones.data(), ones.data(), ones.data(), ones.data(), border_mask,
                             blocknum.data(), radius, rab, npts_global, npts_local, shift_local,
                             border_width, dh, dh_inv, &A.list);
}

size_t grid_len(const int *npts_local) {
  return (size_t)npts_local[0] * (size_t)npts_local[1] * (size_t)npts_local[2];
}

}  // namespace

extern "C" {

void grid_b200_collocate_pgf_products(
    const int nproducts, const bool orthorhombic, const int func, const int *border_mask, const int *la_max,
    const int *la_min, const int *lb_max, const int *lb_min, const double *zeta, const double *zetb,
    const double *rscale, const double *ra, const double *rab, const double *radius, const int *o1,
    const int *o2, const int *n1, const int *n2, const double *const *pab, const double *dh,
    const double *dh_inv, const int *npts_global, const int *npts_local, const int *shift_local,
    const int *border_width, double *grid) {
  if (nproducts <= 0)
    return;
  AdhocList A;
  build_adhoc(A, orthorhombic, nproducts, border_mask, la_max, la_min, lb_max, lb_min, zeta, zetb, ra, rab,
              radius, dh, dh_inv, npts_global, npts_local, shift_local, border_width);
  // blocks: the product's Cartesian sub-block, with the task list's factor
  // rscale = 2 for iatom != jatom (ref/grid_ref_task_list.c:369) divided out
  std::vector<double> blocks(A.nblock_doubles > 0 ? A.nblock_doubles : 1);
  for (int p = 0; p < nproducts; p++) {
    if (o1[p] < 0 || o2[p] < 0 || o1[p] + A.na[p] > n1[p] || o2[p] + A.nb[p] > n2[p])
      die("pab sub-block out of range");
    const double f = 0.5 * rscale[p];
    double *blk = blocks.data() + A.block_offsets[p];
    for (int j = 0; j < A.nb[p]; j++)
      for (int i = 0; i < A.na[p]; i++)
        blk[(size_t)j * A.na[p] + i] = f * pab[p][(size_t)(o2[p] + j) * n1[p] + o1[p] + i];
  }
  const size_t npts = grid_len(npts_local);
  std::vector<double> tmp(npts);
  grid_b200_buffer pab_buf{blocks.size() * sizeof(double), blocks.data(), nullptr};
  grid_b200_buffer grid_buf{npts * sizeof(double), tmp.data(), nullptr};
  grid_b200_buffer *grids[1] = {&grid_buf};
  grid_b200_collocate_task_list(A.list, func, 1, &pab_buf, grids);
  // the single-product API ACCUMULATES into the caller's grid
  for (size_t i = 0; i < npts; i++)
    grid[i] += tmp[i];
}

void grid_b200_integrate_pgf_products(
    const int nproducts, const bool orthorhombic, const bool compute_tau, const int *border_mask,
    const int *la_max, const int *la_min, const int *lb_max, const int *lb_min, const double *zeta,
    const double *zetb, const double *ra, const double *rab, const double *radius, const int *o1,
    const int *o2, const int *n1, const int *n2, const double *dh, const double *dh_inv,
    const int *npts_global, const int *npts_local, const int *shift_local, const int *border_width,
    const double *grid, double *const *hab, const double *const *pab, double *forces) {
  if (nproducts <= 0)
    return;
  AdhocList A;
  build_adhoc(A, orthorhombic, nproducts, border_mask, la_max, la_min, lb_max, lb_min, zeta, zetb, ra, rab,
              radius, dh, dh_inv, npts_global, npts_local, shift_local, border_width);
  const bool with_forces = (forces != nullptr);
  if (with_forces && pab == nullptr)
    die("forces need pab");
  const size_t nb = A.nblock_doubles > 0 ? A.nblock_doubles : 1;
  std::vector<double> pblocks(with_forces ? nb : 1), hblocks(nb);
  for (int p = 0; p < nproducts; p++) {
    if (o1[p] < 0 || o2[p] < 0 || o1[p] + A.na[p] > n1[p] || o2[p] + A.nb[p] > n2[p])
      die("hab sub-block out of range");
    if (with_forces) {
      double *blk = pblocks.data() + A.block_offsets[p];
      for (int j = 0; j < A.nb[p]; j++)
        for (int i = 0; i < A.na[p]; i++)
          blk[(size_t)j * A.na[p] + i] = pab[p][(size_t)(o2[p] + j) * n1[p] + o1[p] + i];
    }
  }
  const size_t npts = grid_len(npts_local);
  grid_b200_buffer pab_buf{pblocks.size() * sizeof(double), pblocks.data(), nullptr};
  grid_b200_buffer hab_buf{hblocks.size() * sizeof(double), hblocks.data(), nullptr};
  grid_b200_buffer grid_buf{npts * sizeof(double), const_cast<double *>(grid), nullptr};
  const grid_b200_buffer *grids[1] = {&grid_buf};
  std::vector<double> f(with_forces ? 6 * (size_t)nproducts : 0);
  grid_b200_integrate_task_list(A.list, compute_tau, 2 * nproducts, 1, with_forces ? &pab_buf : nullptr, grids,
                                &hab_buf, with_forces ? f.data() : nullptr, nullptr);
  // ACCUMULATE into the callers' hab / forces (cpu/grid_cpu_integrate.c semantics)
  for (int p = 0; p < nproducts; p++) {
    const double *blk = hblocks.data() + A.block_offsets[p];
    for (int j = 0; j < A.nb[p]; j++)
      for (int i = 0; i < A.na[p]; i++)
        hab[p][(size_t)(o2[p] + j) * n1[p] + o1[p] + i] += blk[(size_t)j * A.na[p] + i];
    if (with_forces)  // the list scales forces of iatom != jatom by 2 (ref/grid_ref_task_list.c:625-641)
      for (int i = 0; i < 6; i++)
        forces[6 * (size_t)p + i] += 0.5 * f[6 * (size_t)p + i];
  }
}

void grid_b200_collocate_pgf_product(
    const bool orthorhombic, const int border_mask, const int func, const int la_max, const int la_min,
    const int lb_max, const int lb_min, const double zeta, const double zetb, const double rscale,
    const double *dh, const double *dh_inv, const double *ra, const double *rab, const int *npts_global,
    const int *npts_local, const int *shift_local, const int *border_width, const double radius,
    const int o1, const int o2, const int n1, const int n2, const double *pab, double *grid) {
  grid_b200_collocate_pgf_products(1, orthorhombic, func, &border_mask, &la_max, &la_min, &lb_max, &lb_min,
                                   &zeta, &zetb, &rscale, ra, rab, &radius, &o1, &o2, &n1, &n2, &pab, dh,
                                   dh_inv, npts_global, npts_local, shift_local, border_width, grid);
}

void grid_b200_integrate_pgf_product(
    const bool orthorhombic, const bool compute_tau, const int border_mask, const int la_max,
    const int la_min, const int lb_max, const int lb_min, const double zeta, const double zetb,
    const double *dh, const double *dh_inv, const double *ra, const double *rab, const int *npts_global,
    const int *npts_local, const int *shift_local, const int *border_width, const double radius,
    const int o1, const int o2, const int n1, const int n2, const double *grid, double *hab,
    const double *pab, double *forces, double *virials, double *hdab, double *hadb, double *a_hdab) {
  if (virials != nullptr || hdab != nullptr || hadb != nullptr || a_hdab != nullptr)
    die("integrate_pgf_product: per-product virials / hdab / hadb / a_hdab are not provided by the "
        "B200 backend (use the task-list API for the virial)");
  grid_b200_integrate_pgf_products(1, orthorhombic, compute_tau, &border_mask, &la_max, &la_min, &lb_max,
                                   &lb_min, &zeta, &zetb, ra, rab, &radius, &o1, &o2, &n1, &n2, dh, dh_inv,
                                   npts_global, npts_local, shift_local, border_width, grid, &hab,
                                   (pab != nullptr) ? &pab : nullptr, forces);
}

}  // extern "C"
