// Tiled (grid-tile-centric) collocate / integrate path -- see DESIGN.md.
#pragma once
#include "b200_generic.cuh"

namespace b200 {

struct TiledLevel {
  long long npairs = 0;
  void release() {}
};

inline void build_tiled_level(TiledLevel &tl, const LevelDev &L, const std::vector<TaskDev> &tasks,
                              const int first, const int last, std::vector<int> &generic_ids,
                              cudaStream_t s) {
  (void)tl, (void)L, (void)s;
  for (int it = first; it < last; it++)
    generic_ids.push_back(it);
}
inline bool tiled_supports(const TiledLevel &tl, const int max_lp) {
  (void)tl, (void)max_lp;
  return false;
}
inline void launch_tiled_collocate(const TiledLevel &tl, const GridLaunch &L) { (void)tl, (void)L; }
inline void launch_tiled_integrate(const TiledLevel &tl, const GridLaunch &L) { (void)tl, (void)L; }

// ---------------------------------------------------------------------------
// Workload statistics: walks the reference's loop bounds for every task and
// accumulates the model flop count of SURVEY.md 8(d) / Appendix A.
// ---------------------------------------------------------------------------
struct StatsArgs {
  const TaskDev *tasks;
  int ntasks;
  const LevelDev *levels;
  double *out;  // [0] pts [1] flops collocate(AB) [2] flops integrate
};

__global__ void stats_kernel(const StatsArgs A) {
  double pts_sum = 0.0, fc_sum = 0.0, fi_sum = 0.0;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < A.ntasks; it += gridDim.x * blockDim.x) {
    const TaskDev &T = A.tasks[it];
    if (T.skip)
      continue;
    const LevelDev &G = A.levels[T.level];
    const int lp = T.la_max + T.lb_max;
    double pts = 0.0, rows = 0.0, planes = 0.0;
    if (T.use_ortho) {
      const double h[3] = {G.dh[0], G.dh[4], G.dh[8]};
      const double hinv[3] = {G.dh_inv[0], G.dh_inv[4], G.dh_inv[8]};
      const double R = T.disr_radius;
      for (int k = T.lb_cube[2]; k <= 1 - T.lb_cube[2]; k++) {
        planes += 1;
        const double kr = pair_dist(k) * h[2];
        const double krem = __dsub_rn(__dmul_rn(R, R), __dmul_rn(kr, kr));
        const int jstart = sphere_start(krem, hinv[1]);
        for (int j = jstart; j <= 1 - jstart; j++) {
          rows += 1;
          const double jr = pair_dist(j) * h[1];
          const double jrem = __dsub_rn(krem, __dmul_rn(jr, jr));
          pts += 2 - 2 * sphere_start(jrem, hinv[0]);
        }
      }
      const double t2 = 0.5 * (lp + 1) * (lp + 2);
      fc_sum += pts * (2.0 * (lp + 1) + 1) + rows * 2.0 * t2 + planes * 2.0 * ncoset(lp);
      fi_sum += pts * (2.0 * (lp + 1)) + rows * 2.0 * t2 + planes * 2.0 * ncoset(lp);
    } else {
      int bnd[3][2];
      for (int d = 0; d < 3; d++) {
        bnd[d][0] = 0, bnd[d][1] = G.npts_local[d] - 1;
        if (T.border_mask & (1 << (2 * d)))
          bnd[d][0] += G.border_width[d];
        if (T.border_mask & (1 << (2 * d + 1)))
          bnd[d][1] -= G.border_width[d];
      }
      for (int k = T.index_min[2]; k <= T.index_max[2]; k++) {
        const int kg = pmod(k - G.shift_local[2], G.npts_global[2]);
        if (kg < bnd[2][0] || bnd[2][1] < kg)
          continue;
        planes += 1;
        for (int j = T.index_min[1]; j <= T.index_max[1]; j++) {
          const int jg = pmod(j - G.shift_local[1], G.npts_global[1]);
          if (jg < bnd[1][0] || bnd[1][1] < jg)
            continue;
          double qa = 0.0, qb = 0.0, qc = 0.0;
          const double dj = j - T.gp[1], dk = k - T.gp[2];
          for (int c = 0; c < 3; c++) {
            const double h0 = G.dh[c];
            const double v = __dadd_rn(
                __dadd_rn(__dmul_rn(0.0 - T.gp[0], h0), __dmul_rn(dj, G.dh[3 + c])),
                __dmul_rn(dk, G.dh[6 + c]));
            qa = __dadd_rn(qa, __dmul_rn(h0, h0));
            qb = __dadd_rn(qb, __dmul_rn(__dmul_rn(2.0, v), h0));
            qc = __dadd_rn(qc, __dmul_rn(v, v));
          }
          const double disc = __dsub_rn(
              __dmul_rn(qb, qb),
              __dmul_rn(__dmul_rn(4.0, qa), __dsub_rn(qc, __dmul_rn(T.radius, T.radius))));
          if (!(0.0 < disc))
            continue;
          rows += 1;
          const double sq = sqrt(disc);
          const double inv2a = __ddiv_rn(1.0, __dmul_rn(2.0, qa));
          const int i0 = (int)ceil(__dmul_rn(__dsub_rn(-qb, sq), inv2a));
          const int i1 = (int)floor(__dmul_rn(__dadd_rn(-qb, sq), inv2a));
          for (int i = i0; i <= i1; i++) {
            const int ig = pmod(i - G.shift_local[0], G.npts_global[0]);
            if (bnd[0][0] <= ig && ig <= bnd[0][1])
              pts += 1;
          }
        }
      }
      const double t2 = 0.5 * (lp + 1) * (lp + 2);
      const double common = rows * (2.0 * t2 + (lp + 1) + 40.0) + planes * (2.0 * ncoset(lp) + (lp + 1));
      fc_sum += pts * (3.0 * (lp + 1) + 4) + common;
      fi_sum += pts * (3.0 * (lp + 1) + 3) + common;
    }
    pts_sum += pts;
  }
  atomicAdd(&A.out[0], pts_sum);
  atomicAdd(&A.out[1], fc_sum);
  atomicAdd(&A.out[2], fi_sum);
}

inline void compute_stats(const TaskDev *d_tasks, const int ntasks, const std::vector<LevelDev> &levels,
                          const std::vector<TaskDev> &h_tasks, double *stats, cudaStream_t s) {
  (void)h_tasks;
  LevelDev *d_levels = nullptr;
  double *d_out = nullptr;
  B200_CHECK(cudaMalloc((void **)&d_levels, levels.size() * sizeof(LevelDev)));
  B200_CHECK(cudaMalloc((void **)&d_out, 3 * sizeof(double)));
  B200_CHECK(cudaMemcpyAsync(d_levels, levels.data(), levels.size() * sizeof(LevelDev),
                             cudaMemcpyHostToDevice, s));
  B200_CHECK(cudaMemsetAsync(d_out, 0, 3 * sizeof(double), s));
  StatsArgs A{d_tasks, ntasks, d_levels, d_out};
  stats_kernel<<<std::min((ntasks + 127) / 128, 148 * 16), 128, 0, s>>>(A);
  B200_CHECK(cudaGetLastError());
  count_launch();
  double h[3];
  B200_CHECK(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, s));
  B200_CHECK(cudaStreamSynchronize(s));
  stats[4] = h[0], stats[5] = h[1], stats[6] = h[2];
  cudaFree(d_levels);
  cudaFree(d_out);
}

}  // namespace b200
