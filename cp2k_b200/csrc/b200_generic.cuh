// Generic per-task collocate / integrate kernels (one CTA per Gaussian
// product).  They cover everything the tiled kernels do not: triclinic cells,
// border-masked ("generalised") tasks, very high lp and very large cubes, and
// serve as the in-backend cross-check of the tiled path
// (grid_b200_set_kernel_variant(1)).
//
// Reference semantics restated (paths relative to /root/reference/src/grid):
//   ortho bounds / tables / map   ref/grid_ref_collint.h:206-327, 30-200
//   general bounds / masks        ref/grid_ref_collint.h:582-691, 416-483, 333-389
// Design differences to the reference GPU backend
// (gpu/grid_gpu_collocate.cu:148-400, gpu/grid_gpu_integrate.cu:357-693):
//   * ortho: separable 1-D tables in shared memory, one exp per table entry
//     instead of one exp per grid point;
//   * the (y,z) contraction is done once per cube row by one thread, the row is
//     then swept by a half-warp with coalesced accesses along x;
//   * integrate walks the cube ONCE for all coefficients (no lbatch passes):
//     per-row partial sums are reduced by shuffles inside groups of eight lanes
//     (four rows per warp) and folded into the coefficients in two stages
//     (rows -> (lx, ly) sums per cube plane -> coefficients), by all threads.
#pragma once
#include "b200_coef.cuh"

namespace b200 {

constexpr int kGenThreads = 128;
constexpr int kGenRows = 128;  // cube rows per chunk (one per thread)
constexpr int kGenU = 1024;    // integrate: (plane, lx, ly) partial sums per fold batch
constexpr int kGenPairs = (kMaxLp + 1) * (kMaxLp + 2) / 2;

__device__ inline int pair_dist(const int k) { return (k <= 0) ? -k : k - 1; }

// ceil(-1e-8 - sqrt(max(0,rem)) * hinv) with the reference's operation order
// and without FMA contraction (the result decides which points are touched).
__device__ inline int sphere_start(const double rem, const double hinv) {
  return (int)ceil(__dsub_rn(-1e-8, __dmul_rn(sqrt(fmax(0.0, rem)), hinv)));
}

struct GenSmem {
  double *coef, *tab, *row, *aux;
  int *info, *map;
};

__host__ __device__ inline size_t gen_smem_bytes(const int max_lp, const int max_w) {
  const size_t nd = (size_t)ncoset(max_lp) + 3 * (size_t)(max_lp + 1) * max_w +
                    (size_t)kGenRows * (max_lp + 1) + 2 * (size_t)kGenRows;
  return nd * sizeof(double) + (4 * (size_t)kGenRows + (size_t)((max_w + 1) & ~1)) * sizeof(int) +
         kGenU * sizeof(double) + kGenPairs * sizeof(int);
}

template <bool COLLOCATE>
__global__ void __launch_bounds__(kGenThreads) generic_kernel(const GridLaunch L) {
  extern __shared__ double smem[];
  __shared__ unsigned s_orb[kOrbEntries];
  const int tid = threadIdx.x;
  if (!COLLOCATE) {
    stage_orb_table(s_orb, tid, kGenThreads);
    __syncthreads();
  }
  const int itask = L.task_ids[blockIdx.x];
  const TaskDev &T = L.tasks[itask];
  const int lp = T.la_max + T.lb_max + L.dl, lp1 = lp + 1, nc = ncoset(lp);
  double *coef_g = L.coef + L.coef_offsets[itask];
  if (T.skip) {
    if (!COLLOCATE)
      for (int c = tid; c < nc; c += kGenThreads)
        coef_g[c] = 0.0;
    return;
  }
  const int W = L.max_w;
  double *s_coef = smem;
  double *s_tab = s_coef + ncoset(L.max_lp);
  double *s_row = s_tab + 3 * (L.max_lp + 1) * W;
  double *s_aux = s_row + kGenRows * (L.max_lp + 1);
  int *s_info = (int *)(s_aux + 2 * kGenRows);
  int *s_map = s_info + 4 * kGenRows;
  // (the map is padded to an even number of ints so that the doubles behind it stay aligned)
  double *s_u = (double *)(s_map + ((W + 1) & ~1));
  int *s_pair = (int *)(s_u + kGenU);

  const LevelDev &G = L.level;
  const bool ortho = T.use_ortho;
  const double zetp = T.zetp;
  int lo[3], w[3];
  for (int d = 0; d < 3; d++) {
    lo[d] = ortho ? T.lb_cube[d] : T.index_min[d];
    w[d] = ortho ? 2 - 2 * T.lb_cube[d] : T.index_max[d] - T.index_min[d] + 1;
  }
  const double h[3] = {G.dh[0], G.dh[4], G.dh[8]};
  const double hinv[3] = {G.dh_inv[0], G.dh_inv[4], G.dh_inv[8]};

  // admissible local window of the general path (collint.h:591-609)
  int bnd[3][2];
  for (int d = 0; d < 3; d++) {
    bnd[d][0] = 0;
    bnd[d][1] = G.npts_local[d] - 1;
    if (T.border_mask & (1 << (2 * d)))
      bnd[d][0] += G.border_width[d];
    if (T.border_mask & (1 << (2 * d + 1)))
      bnd[d][1] -= G.border_width[d];
  }

  // ---- per-task tables ----------------------------------------------------
  if (COLLOCATE)
    for (int c = tid; c < nc; c += kGenThreads)
      s_coef[c] = coef_g[c];
  for (int q = tid; q < 3 * W; q += kGenThreads) {
    const int d = q / W, g = q % W;
    if (g >= w[d])
      continue;
    double *tb = s_tab + (size_t)d * lp1 * W + g;
    if (ortho) {
      const double x = (lo[d] + g) * h[d] - T.roffset[d];
      double v = exp(-zetp * x * x);
      for (int l = 0; l <= lp; l++, v *= x)
        tb[l * W] = v;
      if (d == 0)
        s_map[g] = pmod(T.cubecenter[0] + lo[0] + g - G.shift_local[0],
                        G.npts_global[0]);
    } else if (d > 0) {
      const double x = (lo[d] + g) - T.gp[d];
      double v = 1.0;
      for (int l = 0; l <= lp; l++, v *= x)
        tb[l * W] = v;
    }
  }
  const double *tabx = s_tab, *taby = s_tab + (size_t)lp1 * W,
               *tabz = s_tab + (size_t)2 * lp1 * W;

  // integrate: per-thread coefficient accumulators (coefficient tid + m * kGenThreads)
  constexpr int kCPT = 8;  // ncoset(16) = 969 <= 8 * 128
  double acc[kCPT];
  for (int m = 0; m < kCPT; m++)
    acc[m] = 0.0;
  // (lx, ly) pairs with lx + ly <= lp, pair index = ly * lp1 - ly (ly - 1) / 2 + lx
  const int npairs = lp1 * (lp1 + 1) / 2;
  if (!COLLOCATE)
    for (int q = tid; q < lp1 * lp1; q += kGenThreads) {
      const int ly = q / lp1, lx = q % lp1;
      if (lx + ly <= lp)
        s_pair[ly * lp1 - (ly * (ly - 1)) / 2 + lx] = lx | (ly << 8);
    }

  const int nrows = w[1] * w[2];
  // collocate: half-warps sweep a row; integrate: groups of eight lanes
  constexpr int kLanes = COLLOCATE ? 16 : 8;
  const int hw = tid / kLanes, sub = tid % kLanes;
  const int ny = G.npts_local[1], nx = G.npts_local[0];

  for (int r0 = 0; r0 < nrows; r0 += kGenRows) {
    __syncthreads();
    // ---- phase A: one thread per cube row: bounds (+ row coefficients) -----
    {
      const int row = r0 + tid;
      int base = -1, i0 = 0, i1 = -1;
      double aux_b = 0.0, aux_c = 0.0;
      int jj = 0, kk = 0;
      if (row < nrows) {
        kk = row / w[1], jj = row % w[1];
        const int k = lo[2] + kk, j = lo[1] + jj;
        if (ortho) {
          const double R = T.disr_radius;
          const double kr = pair_dist(k) * h[2];
          const double krem = __dsub_rn(__dmul_rn(R, R), __dmul_rn(kr, kr));
          const int jstart = sphere_start(krem, hinv[1]);
          if (jstart <= j && j <= 1 - jstart) {
            const double jr = pair_dist(j) * h[1];
            const double jrem = __dsub_rn(krem, __dmul_rn(jr, jr));
            const int istart = sphere_start(jrem, hinv[0]);
            i0 = istart, i1 = 1 - istart;
            const int kg = pmod(T.cubecenter[2] + k - G.shift_local[2], G.npts_global[2]);
            const int jg = pmod(T.cubecenter[1] + j - G.shift_local[1], G.npts_global[1]);
            base = (kg * ny + jg) * nx;
          }
        } else {
          const int kg = pmod(k - G.shift_local[2], G.npts_global[2]);
          const int jg = pmod(j - G.shift_local[1], G.npts_global[1]);
          if (bnd[2][0] <= kg && kg <= bnd[2][1] && bnd[1][0] <= jg && jg <= bnd[1][1]) {
            // |i*dh[0] + v|^2 = radius^2  (collint.h:450-463), no contraction
            double qa = 0.0, qb = 0.0, qc = 0.0;
            const double dj = j - T.gp[1], dk = k - T.gp[2];
            double vb = 0.0, vc = 0.0;  // same quadratic about di = i - gp[0]
            for (int c = 0; c < 3; c++) {
              const double h0 = G.dh[0 * 3 + c];
              const double v = __dadd_rn(
                  __dadd_rn(__dmul_rn(0.0 - T.gp[0], h0), __dmul_rn(dj, G.dh[1 * 3 + c])),
                  __dmul_rn(dk, G.dh[2 * 3 + c]));
              qa = __dadd_rn(qa, __dmul_rn(h0, h0));
              qb = __dadd_rn(qb, __dmul_rn(__dmul_rn(2.0, v), h0));
              qc = __dadd_rn(qc, __dmul_rn(v, v));
              const double v2 = dj * G.dh[1 * 3 + c] + dk * G.dh[2 * 3 + c];
              vb += 2.0 * v2 * h0;
              vc += v2 * v2;
            }
            const double disc = __dsub_rn(
                __dmul_rn(qb, qb),
                __dmul_rn(__dmul_rn(4.0, qa), __dsub_rn(qc, __dmul_rn(T.radius, T.radius))));
            if (0.0 < disc) {
              const double sq = sqrt(disc);
              const double inv2a = __ddiv_rn(1.0, __dmul_rn(2.0, qa));
              i0 = (int)ceil(__dmul_rn(__dsub_rn(-qb, sq), inv2a));
              i1 = (int)floor(__dmul_rn(__dadd_rn(-qb, sq), inv2a));
              base = (kg * ny + jg) * nx;
              aux_b = vb, aux_c = vc;
            }
          }
        }
      }
      s_info[4 * tid + 0] = base;
      s_info[4 * tid + 1] = i0;
      s_info[4 * tid + 2] = i1;
      s_info[4 * tid + 3] = (kk << 16) | jj;
      s_aux[2 * tid + 0] = aux_b;
      s_aux[2 * tid + 1] = aux_c;
      if (COLLOCATE && base >= 0) {
        double *cr = s_row + tid * lp1;
        for (int l = 0; l <= lp; l++)
          cr[l] = 0.0;
        for (int lz = 0; lz <= lp; lz++) {
          const double pz = tabz[lz * W + kk];
          for (int ly = 0; ly <= lp - lz; ly++) {
            const double pyz = taby[ly * W + jj] * pz;
            for (int lx = 0; lx <= lp - lz - ly; lx++)
              cr[lx] += s_coef[coset(lx, ly, lz)] * pyz;
          }
        }
      }
    }
    __syncthreads();

    // ---- phase B: half-warps sweep the rows along x --------------------------
    const int nr = min(kGenRows, nrows - r0);
    for (int r = hw; r < nr; r += kGenThreads / kLanes) {
      const int base = s_info[4 * r + 0];
      if (base < 0)
        continue;
      const int i0 = s_info[4 * r + 1], i1 = s_info[4 * r + 2];
      double *cr = s_row + r * lp1;
      if (COLLOCATE) {
        for (int i = i0 + sub; i <= i1; i += 16) {
          if (ortho) {
            const int g = i - lo[0];
            double v = 0.0;
            for (int l = 0; l <= lp; l++)
              v += cr[l] * tabx[l * W + g];
            atomicAdd(&L.grid[base + s_map[g]], v);
          } else {
            const int ig = pmod(i - G.shift_local[0], G.npts_global[0]);
            if (ig < bnd[0][0] || bnd[0][1] < ig)
              continue;
            const double di = i - T.gp[0];
            const double qa = G.dh[0] * G.dh[0] + G.dh[1] * G.dh[1] + G.dh[2] * G.dh[2];
            const double r2 = (qa * di + s_aux[2 * r]) * di + s_aux[2 * r + 1];
            double v = cr[lp];
            for (int l = lp - 1; l >= 0; l--)
              v = v * di + cr[l];
            atomicAdd(&L.grid[base + ig], v * exp(-zetp * r2));
          }
        }
      } else {
        // integrate: cr[l] = sum_i grid[i] * p_l(i), reduced over the group of eight lanes.
        // Only lane `sub == 0` of the group ever touches cr[] here.
        const unsigned hmask = 0xffu << (8 * ((tid >> 3) & 3));
        if (sub == 0)
          for (int l = 0; l <= lp; l++)
            cr[l] = 0.0;
        for (int ib = i0; ib <= i1; ib += 32) {
          double gv[4], xv[4];
          int gi[4];
#pragma unroll
          for (int m = 0; m < 4; m++) {
            const int i = ib + sub + 8 * m;
            gv[m] = 0.0, xv[m] = 0.0, gi[m] = 0;
            if (i <= i1) {
              if (ortho) {
                gi[m] = i - lo[0];
                gv[m] = L.grid[base + s_map[gi[m]]];
              } else {
                const int ig = pmod(i - G.shift_local[0], G.npts_global[0]);
                if (bnd[0][0] <= ig && ig <= bnd[0][1]) {
                  const double di = i - T.gp[0];
                  const double qa = G.dh[0] * G.dh[0] + G.dh[1] * G.dh[1] + G.dh[2] * G.dh[2];
                  const double r2 = (qa * di + s_aux[2 * r]) * di + s_aux[2 * r + 1];
                  gv[m] = L.grid[base + ig] * exp(-zetp * r2);
                  xv[m] = di;
                }
              }
            }
          }
          double pw[4] = {1.0, 1.0, 1.0, 1.0};
          for (int l = 0; l <= lp; l++) {
            double part = 0.0;
#pragma unroll
            for (int m = 0; m < 4; m++) {
              if (ortho) {
                part += gv[m] * tabx[l * W + gi[m]];
              } else {
                part += gv[m] * pw[m];
                pw[m] *= xv[m];
              }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1)
              part += __shfl_xor_sync(hmask, part, o, 8);
            if (sub == 0)
              cr[l] += part;
          }
        }
      }
    }

    if (!COLLOCATE) {
      __syncthreads();
      // ---- phase A': fold the row sums into the coefficients ----------------
      // Two stages per batch of cube planes kk: u[kk][lx,ly] = sum_jj row[kk,jj][lx] taby[ly][jj],
      // then coef[lx,ly,lz] += sum_kk u[kk][lx,ly] tabz[lz][kk]  (rows x pairs + planes x ncoset
      // operations instead of rows x ncoset).
      const int kk_first = r0 / w[1], kk_last = (r0 + nr - 1) / w[1];
      const int kbatch = max(1, kGenU / npairs);
      for (int kb0 = kk_first; kb0 <= kk_last; kb0 += kbatch) {
        const int nkb = min(kbatch, kk_last - kb0 + 1);
        for (int item = tid; item < nkb * npairs; item += kGenThreads) {
          const int kb = item / npairs, pr = item - kb * npairs;
          const int lx = s_pair[pr] & 0xff, ly = s_pair[pr] >> 8;
          const int kk = kb0 + kb;
          const int ra = max(kk * w[1] - r0, 0), rb = min((kk + 1) * w[1] - r0, nr);
          double u = 0.0;
          for (int r = ra; r < rb; r++)
            if (s_info[4 * r] >= 0)
              u += s_row[r * lp1 + lx] * taby[ly * W + (r0 + r - kk * w[1])];
          s_u[item] = u;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < kCPT; m++) {
          const int c = tid + m * kGenThreads;
          if (c < nc) {
            const Orb o = orb_of(s_orb, c);
            const int pr = o.l[1] * lp1 - (o.l[1] * (o.l[1] - 1)) / 2 + o.l[0];
            double a = acc[m];
            for (int kb = 0; kb < nkb; kb++)
              a += s_u[kb * npairs + pr] * tabz[o.l[2] * W + kb0 + kb];
            acc[m] = a;
          }
        }
        __syncthreads();
      }
    }
  }

  if (!COLLOCATE) {
#pragma unroll
    for (int m = 0; m < kCPT; m++) {
      const int c = tid + m * kGenThreads;
      if (c < nc)
        coef_g[c] = acc[m];
    }
  }
}

inline void launch_generic(const GridLaunch &L, const bool collocate) {
  if (L.ntasks == 0)
    return;
  const size_t bytes = gen_smem_bytes(L.max_lp, L.max_w);
  B200_ASSERT(bytes <= 220 * 1024, "generic kernel: cube too large for shared memory");
  if (collocate) {
    B200_CHECK(cudaFuncSetAttribute(generic_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    generic_kernel<true><<<L.ntasks, kGenThreads, bytes, L.stream>>>(L);
  } else {
    B200_CHECK(cudaFuncSetAttribute(generic_kernel<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    generic_kernel<false><<<L.ntasks, kGenThreads, bytes, L.stream>>>(L);
  }
  B200_CHECK(cudaGetLastError());
  count_launch();
}

}  // namespace b200
