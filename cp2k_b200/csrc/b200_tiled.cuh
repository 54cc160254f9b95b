// Tiled (grid-tile-centric) collocate / integrate kernels for orthorhombic
// tasks -- the hot path of the backend.  See DESIGN.md section "Tiled kernels".
//
// Idea.  The reference scatters every Gaussian product into the grid
// (CPU: thread-private grids + reduction, ref/grid_ref_task_list.c:291-426;
// GPU: one FP64 atomicAdd per point, gpu/grid_gpu_collocate.cu:392-396).  Here
// the loop nest is inverted: a WARP owns an 8x8x16 block of the grid, every
// thread keeps its 2x16 grid points in REGISTERS, and the warp streams through
// the list of (task, block) pairs that overlap its block.  No atomics and no
// read-modify-write of the grid in the inner loop, no CTA-wide barrier either:
// warps are fully autonomous.  Each work item touches its block in memory
// exactly once (a RED flush for collocate, one load for integrate).
//
//   * pairs are generated once per task list ON THE GPU (count / scan / fill),
//     bucketed by (block, lp); periodic images become separate pairs; a pair
//     exists only if the task's sphere really meets the block, and carries the
//     warp-uniform part of the per-pair work precomputed (sphere-table index,
//     plane range);
//   * the separable Gaussian factors exp(-zetp (x-xp)^2) of every task are
//     tabulated once per task list (they depend on geometry only) -- there is
//     no exp() in the hot loop; a (pair, warp) step loads exactly one table
//     entry per lane (8 x-, 8 y-, 16 z-entries).  The rows are sized per task and
//     zero-padded by the block extent, the pair array is padded with copies of its
//     last record: the look-ahead loads and the table index need no clamping;
//   * which points of a cube are inside the (discretised-radius) sphere is two
//     table lookups per column: because the reference discretises the radius to
//     n*drmin (ref/grid_ref_collint.h:237-239) the admissible z-extent K of a
//     column depends only on (n, |j|, |i|); the table is built on the host with
//     the reference's own expressions, so the SET of touched points is
//     identical.  A second table turns (K, centre plane) into a 16-bit plane
//     mask; the masks drive PREDICATED DFMAs (no selects);
//   * planes no lane of the warp needs are skipped warp-uniformly;
//   * a work item's pairs are ordered by lp and each lp has its own fully
//     specialised pair loop (no switch inside the loop);
//   * the next pair's record, table entry and coefficients are prefetched
//     while the current pair is processed; the per-pair scratch in shared
//     memory is double buffered (one __syncwarp per pair).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "b200_generic.cuh"
#include "b200_pfma.cuh"

namespace b200 {

#ifndef B200_KBZ
#define B200_KBZ 16
#endif
#ifndef B200_CTAS_LO
#define B200_CTAS_LO 4
#endif
#ifndef B200_CTAS_HI
#define B200_CTAS_HI 3
#endif
#ifndef B200_LPHI_LO
#define B200_LPHI_LO 2   // kernels whose largest lp is <= this get B200_CTAS_LO CTAs per SM
#endif
constexpr int kBX = 8, kBY = 8, kBZ = B200_KBZ;  // block (warp footprint)
constexpr int kTiledThreads = 128;            // 4 autonomous warps
constexpr int kTiledWarps = kTiledThreads / 32;
constexpr int kTiledMaxLp = 6;                // max la_max + lb_max of a tiled task; beyond: generic kernels
constexpr int kTiledMaxLpCall = 7;            // max lp including the call's l growth (forces + virial: +3)
constexpr int kTiledMaxNb = 23;               // max cube half-width (sphere-table rows are 64 bytes)
constexpr int kTiledMaxN = 30;                // max discretised radius index
constexpr int kKPitch = 64;                   // sphere-table row pitch (bytes)
constexpr int kKPad = 8;                      // margin around the cube: block offsets need no bounds check
constexpr int kZmPitch = 66;                  // plane-mask table: [K+1][oz + kZmBias]; rows 33 banks apart
constexpr int kZmBias = 24;
constexpr int kZmRows = kTiledMaxNb + 2;
constexpr int kLpBuckets = kTiledMaxLp + 1;
constexpr int kItemPairs = 512;               // pairs per work item (upper bound)
constexpr int kPairPad = 9;                   // records after the last pair (copies of it)
// lp classes (by lp0 = la_max + lb_max): separate kernel instantiations keep the
// code of each kernel small (instruction cache!) and the registers low
constexpr int kNumClasses = 3;
__host__ __device__ inline int lp_class(const int lp0) { return lp0 <= 2 ? 0 : (lp0 <= 4 ? 1 : 2); }
constexpr int kClassLo[kNumClasses] = {0, 3, 5};
constexpr int kClassHi[kNumClasses] = {2, 4, 6};

struct alignas(16) TTask {  // static per-task data of the tiled path (80 bytes)
  double roff[3];
  double zl2;             // zetp * log2(e): exp(-zetp d^2) = 2^(-zl2 d^2)
  int cc[3];              // cubecenter - shift_local (not wrapped)
  int nb[3];              // -lb_cube per axis
  int n;                  // discretised radius index
  int lp0;                // la_max + lb_max
  int task;               // index into the TaskDev array
  int pad[3];
};

// exp(-zl2 * d * d) for zl2 = zetp * log2(e): 2^y with y = n / 64 + f, |f| <= 1/128,
// 2^(j/64) from a 64-entry table (shared memory) and a degree-5 polynomial for 2^f
// (truncation error 4e-17); relative error ~ 2 ulp plus |y| * 1.1e-16 from forming y.
// The binary exponent is clamped so that arguments far outside a task's cube (never
// used by anyone) still give finite numbers.  17 instructions: the kernels compute their
// 1-D Gaussian tables with it instead of reading per-task tables from HBM.
__device__ __forceinline__ double exp_neg_tab(const double zl2, const double d, const double *__restrict__ e2t) {
  const double y = -zl2 * (d * d);
  const double magic = 105553116266496.0;  // 1.5 * 2^46: one ulp is 1/64
  const double t = y + magic;
  const int n = __double2loint(t);
  const double f = y - (t - magic);
  double p = 1.3333558146428443e-03;           // ln2^5 / 120
  p = fma(p, f, 9.6181291076284772e-03);       // ln2^4 / 24
  p = fma(p, f, 5.5504108664821580e-02);       // ln2^3 / 6
  p = fma(p, f, 2.4022650695910071e-01);       // ln2^2 / 2
  p = fma(p, f, 6.9314718055994531e-01);       // ln2
  p = fma(p, f, 1.0);
  const double r = e2t[n & 63] * p;
  const int k = max(n >> 6, -1000);
  return __hiloint2double(__double2hiint(r) + (k << 20), __double2loint(r));
}

struct alignas(16) TPair {  // 16 bytes, read as one uint4 (warp-uniform)
  unsigned q;             // ttask index
  unsigned kbase;         // sphere-table index of block column (0,0)
  unsigned opk;           // bytes 0, 1, 3: cube centre (x, y, z) relative to the block origin (signed); byte 2: wlo | whi << 4
  unsigned spare;
};

struct alignas(16) TWork {  // 32 bytes
  int x0, y0, z0;         // block origin (local grid indices)
  int b[4];               // pair ranges per lp of the class: [b[i], b[i+1])
  int pad;
};

struct alignas(16) KTabHeader {  // per radius index n
  int offset;             // into the byte table, -1: not tabulated
  int nbx, nby, nbz;      // -lb per axis
};

struct TiledLevel {
  long long npairs = 0;
  int nwork = 0;
  int class_work_first[kNumClasses + 1] = {0, 0, 0, 0};  // work item ranges per lp class
  int class_ntasks[kNumClasses] = {0, 0, 0};
  int class_tt_first[kNumClasses + 1] = {0, 0, 0, 0};       // ttask ranges per lp class
  std::vector<int> h_tt_task;                                // TaskDev id per ttask (class-sorted)
  int coef_base[8][kNumClasses] = {};                        // per dl: start of the class's slots
  int max_nb = 0;
  int *d_class_task_ids[kNumClasses] = {nullptr, nullptr, nullptr};  // TaskDev ids per class
  int ntasks_tiled = 0;
  int max_lp0 = 0;
  int max_n = 0;
  TTask *d_ttasks = nullptr;
  TPair *d_pairs = nullptr;
  TWork *d_work = nullptr;
  int *d_counters = nullptr;         // one per (direction, class)
  KTabHeader *d_khead = nullptr;
  unsigned char *d_ktab = nullptr;   // K+1 per (n, dj, di), 0 outside the sphere
  unsigned short *d_zmask = nullptr; // [kZmRows][kZmPitch]
  void release() {
    dev_free(d_ttasks), dev_free(d_pairs), dev_free(d_work), dev_free(d_khead), dev_free(d_ktab);
    dev_free(d_zmask);
    dev_free(d_counters);
    d_counters = nullptr;
    for (auto &p : d_class_task_ids) {
      dev_free(p);
      p = nullptr;
    }
    d_ttasks = nullptr, d_pairs = nullptr, d_work = nullptr, d_khead = nullptr, d_ktab = nullptr;
    d_zmask = nullptr;
    npairs = 0, nwork = 0, ntasks_tiled = 0;
  }
};

// ---------------------------------------------------------------------------
// Host: sphere-extent tables.  For the discretised radius n*drmin the largest
// z pair-distance K such that the point (K, mj, mi) is visited by the
// reference's loop nest (ref/grid_ref_collint.h:144-147, 46-50, 237-254).
// Stored as K+1 (0 = column outside the sphere) by SIGNED cube offsets with a
// margin of kKPad on every side and a fixed row pitch:
//   tab[offset_n + (dj + nby + kKPad) * kKPitch + (di + nbx + kKPad)]
// so that a block column (li, lj) of a pair reads tab[kbase + lj*kKPitch + li]
// with kbase = offset_n + (nby + kKPad - oy) * kKPitch + (nbx + kKPad - ox).
// ---------------------------------------------------------------------------
inline void build_ktabs(const LevelDev &L, const int max_n, std::vector<KTabHeader> &heads,
                        std::vector<unsigned char> &tab, const int pitch = kKPitch, const int pad = kKPad) {
  const double h[3] = {L.dh[0], L.dh[4], L.dh[8]};
  const double hinv[3] = {L.dh_inv[0], L.dh_inv[4], L.dh_inv[8]};
  const double drmin = fmin(h[0], fmin(h[1], h[2]));
  heads.assign(max_n + 1, KTabHeader{-1, 0, 0, 0});
  tab.clear();
  for (int n = 1; n <= max_n; n++) {
    const double R = drmin * fmax(1.0, (double)n);
    int nb[3];
    for (int d = 0; d < 3; d++)
      nb[d] = -(int)ceil(-1e-8 - R * hinv[d]);
    KTabHeader H;
    H.offset = -1, H.nbx = nb[0], H.nby = nb[1], H.nbz = nb[2];
    heads[n] = H;
    if (nb[0] > kTiledMaxNb || nb[1] > kTiledMaxNb || nb[2] > kTiledMaxNb)
      continue;
    const int px = nb[0] + 1, py = nb[1] + 1;
    std::vector<int> K((size_t)px * py, -1);
    for (int kd = 0; kd <= nb[2]; kd++) {
      const double kr = kd * h[2];
      const double krem = R * R - kr * kr;
      const int jstart = (int)ceil(-1e-8 - sqrt(fmax(0.0, krem)) * hinv[1]);
      for (int jd = 0; jd <= -jstart && jd <= nb[1]; jd++) {
        const double jr = jd * h[1];
        const double jrem = krem - jr * jr;
        const int istart = (int)ceil(-1e-8 - sqrt(fmax(0.0, jrem)) * hinv[0]);
        for (int id = 0; id <= -istart && id <= nb[0]; id++)
          K[(size_t)jd * px + id] = std::max<int>(K[(size_t)jd * px + id], kd);
      }
    }
    H.offset = (int)tab.size();
    const int rows = 2 * nb[1] + 2 * pad + 1;
    tab.resize(tab.size() + (size_t)rows * pitch, (unsigned char)0);
    for (int dj = -nb[1]; dj <= nb[1] + 1; dj++)
      for (int di = -nb[0]; di <= nb[0] + 1; di++) {
        const int mj = (dj <= 0) ? -dj : dj - 1, mi = (di <= 0) ? -di : di - 1;
        tab[(size_t)H.offset + (size_t)(dj + nb[1] + pad) * pitch + (di + nb[0] + pad)] =
            (unsigned char)(K[(size_t)mj * px + mi] + 1);
      }
    heads[n] = H;
  }
  tab.resize((tab.size() + 15) / 16 * 16 + 16 * pitch, (unsigned char)0);  // slack for the +4 rows of column 1
}

// zmask[K1][oz + kZmBias]: bit p set <=> plane p of the block lies within
// pair-distance K = K1 - 1 of the cube centre plane oz:  oz - K <= p <= oz + K + 1
inline std::vector<unsigned short> build_zmask() {
  std::vector<unsigned short> zm((size_t)kZmRows * kZmPitch, 0);
  for (int k1 = 1; k1 < kZmRows; k1++)
    for (int ozb = 0; ozb < kZmPitch; ozb++) {
      const int K = k1 - 1, oz = ozb - kZmBias;
      unsigned m = 0;
      for (int p = 0; p < kBZ; p++)
        if (oz - K <= p && p <= oz + K + 1)
          m |= 1u << p;
      zm[(size_t)k1 * kZmPitch + ozb] = (unsigned short)m;
    }
  return zm;
}

// ---------------------------------------------------------------------------
// Device: pair generation.  One thread per tiled task; pass 0 counts the pairs
// per (block, lp) bucket, pass 1 writes them.
// ---------------------------------------------------------------------------
struct PairGenArgs {
  const TTask *ttasks;
  int nttasks;
  int nx, ny, nz, Nx, Ny, Nz;     // local / global grid size
  int nbx, nby, nbz;              // number of blocks per axis
  unsigned nblocks;
  const KTabHeader *khead;
  const unsigned char *ktab;
  unsigned int *bucket_count;     // pass 0
  const unsigned int *bucket_start;  // pass 1
  unsigned int *bucket_cursor;    // pass 1
  TPair *pairs;                   // pass 1
  unsigned long long *keys;       // pass 1: bucket << qbits | q, for the in-bucket ordering
  int qbits;
};

__device__ inline int floor_div(const int a, const int b) {  // b > 0
  return (a >= 0) ? a / b : -((-a + b - 1) / b);
}
__device__ inline int rel_dmin(const int a, const int b) {  // min pair distance over cube offsets [a,b]
  return (a <= 1 && b >= 0) ? 0 : ((a > 1) ? a - 1 : -b);
}

template <int PASS> __global__ void pairgen_kernel(const PairGenArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.nttasks)
    return;
  const TTask X = A.ttasks[q];
  const KTabHeader H = A.khead[X.n];
  const int nloc[3] = {A.nx, A.ny, A.nz}, N[3] = {A.Nx, A.Ny, A.Nz};
  const int B[3] = {kBX, kBY, kBZ}, nblk[3] = {A.nbx, A.nby, A.nbz};
  // per axis: image range
  int m_lo[3], m_hi[3];
  for (int d = 0; d < 3; d++) {
    const int lo = X.cc[d] - X.nb[d], hi = X.cc[d] + 1 + X.nb[d];
    // images m with [lo+mN, hi+mN] intersecting [0, nloc)
    // (a superset; images that miss the local grid are dropped below)
    m_lo[d] = floor_div(-hi, N[d]);
    m_hi[d] = floor_div(nloc[d] - 1 - lo, N[d]) + 1;
  }
  const unsigned char *ktab_n = A.ktab + H.offset;
  for (int mz = m_lo[2]; mz <= m_hi[2]; mz++) {
    const int cz = X.cc[2] + mz * N[2];
    const int az = max(cz - X.nb[2], 0), bz = min(cz + 1 + X.nb[2], nloc[2] - 1);
    if (az > bz)
      continue;
    for (int tz = az / B[2]; tz <= bz / B[2] && tz < nblk[2]; tz++) {
      const int oz = cz - tz * B[2];
      const int zlo = max(az, tz * B[2]) - tz * B[2], zhi = min(bz, tz * B[2] + B[2] - 1) - tz * B[2];
      const int mk = rel_dmin(zlo - oz, zhi - oz);
      for (int my = m_lo[1]; my <= m_hi[1]; my++) {
        const int cy = X.cc[1] + my * N[1];
        const int ay = max(cy - X.nb[1], 0), by = min(cy + 1 + X.nb[1], nloc[1] - 1);
        if (ay > by)
          continue;
        for (int ty = ay / B[1]; ty <= by / B[1] && ty < nblk[1]; ty++) {
          const int oy = cy - ty * B[1];
          const int mj = rel_dmin(max(ay, ty * B[1]) - cy, min(by, ty * B[1] + B[1] - 1) - cy);
          // the widest column of this row of blocks (mi = 0) must reach the block's planes
          if ((int)ktab_n[(X.nb[1] + kKPad - mj) * kKPitch + (X.nb[0] + kKPad)] - 1 < mk)
            continue;
          for (int mx = m_lo[0]; mx <= m_hi[0]; mx++) {
            const int cx = X.cc[0] + mx * N[0];
            const int ax = max(cx - X.nb[0], 0), bx = min(cx + 1 + X.nb[0], nloc[0] - 1);
            if (ax > bx)
              continue;
            for (int tx = ax / B[0]; tx <= bx / B[0] && tx < nblk[0]; tx++) {
              const int ox = cx - tx * B[0];
              const int mi = rel_dmin(max(ax, tx * B[0]) - cx, min(bx, tx * B[0] + B[0] - 1) - cx);
              // K of the block's column nearest to the centre bounds every other column's
              const int K = (int)ktab_n[(X.nb[1] + kKPad - mj) * kKPitch + (X.nb[0] + kKPad - mi)] - 1;
              const int wlo = max(zlo, oz - K), whi = min(zhi, oz + K + 1);
              if (K < 0 || wlo > whi)
                continue;  // the sphere misses this block
              const unsigned blk = (unsigned)((tz * nblk[1] + ty) * nblk[0] + tx);
              const unsigned bucket = ((unsigned)lp_class(X.lp0) * A.nblocks + blk) * kLpBuckets + X.lp0;
              if (PASS == 0) {
                atomicAdd(&A.bucket_count[bucket], 1u);
              } else {
                const unsigned pos = A.bucket_start[bucket] + atomicAdd(&A.bucket_cursor[bucket], 1u);
                TPair P;
                P.q = (unsigned)q;
                P.kbase = (unsigned)(H.offset + (X.nb[1] + kKPad - oy) * kKPitch + (X.nb[0] + kKPad - ox));
                P.opk = ((unsigned)ox & 0xffu) | (((unsigned)oy & 0xffu) << 8) | (((unsigned)oz & 0xffu) << 24) |
                        ((unsigned)wlo << 16) | ((unsigned)whi << 20);
                P.spare = 0u;
                A.pairs[pos] = P;
                A.keys[pos] = ((unsigned long long)bucket << A.qbits) | (unsigned long long)q;
              }
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Host: per-level build.
// ---------------------------------------------------------------------------
// GRID_B200_CREATE_TIMING=1: wall time of the builder's phases on stderr
struct CreateTimer {
  bool on;
  cudaStream_t s;
  std::chrono::steady_clock::time_point t;
  CreateTimer(cudaStream_t stream) : on(getenv("GRID_B200_CREATE_TIMING") != nullptr), s(stream) {
    t = std::chrono::steady_clock::now();
  }
  void tick(const char *what) {
    if (!on)
      return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "grid_b200 create:   %-26s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

// Device scratch of the per-level builders, kept across the levels of one list: GB-sized
// cudaMalloc / cudaFree pairs cost ~100 ms each at H2O-1024 size.
struct BuilderScratch {
  void *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[6] = {0, 0, 0, 0, 0, 0};
  void *get(const int slot, const size_t bytes) {
    if (bytes > cap[slot]) {
      dev_free(p[slot]);
      dev_alloc(&p[slot], std::max<size_t>(bytes, 16));
      cap[slot] = bytes;
    }
    return p[slot];
  }
  ~BuilderScratch() {
    for (auto q : p)
      dev_free(q);
  }
};

inline void build_tiled_level(TiledLevel &tl, const LevelDev &L, const TaskVec &tasks,
                              const TaskDev *d_tasks, const int first, const int last,
                              std::vector<int> &generic_ids, BuilderScratch &scratch, cudaStream_t s) {
  CreateTimer tm(s);
  tl.release();
  const double h[3] = {L.dh[0], L.dh[4], L.dh[8]};
  const double drmin = fmin(h[0], fmin(h[1], h[2]));
  const int nbx = (L.npts_local[0] + kBX - 1) / kBX, nby = (L.npts_local[1] + kBY - 1) / kBY,
            nbz = (L.npts_local[2] + kBZ - 1) / kBZ;
  const size_t nblocks = (size_t)nbx * nby * nbz;

  // Select the tasks of the tiled path and lay them out sorted by lp class (a task's coefficient
  // slot is then pure arithmetic on its index): two parallel passes over the level's records --
  // classify, then write each record to its final position -- with a serial prefix over one byte
  // per task in between; the result does not depend on the thread count.
  std::vector<TTask> tt;
  std::vector<KTabHeader> heads;
  std::vector<unsigned char> ktab;
  int max_n = 0, max_lp0 = 0, max_nb = 0;
  {
    const int nt = last - first;
    std::vector<int> nidx(std::max(nt, 1));            // radius index, 0 = generic path
    std::vector<unsigned char> cls(std::max(nt, 1));   // lp class of a tiled task
#pragma omp parallel for schedule(static) reduction(max : max_n, max_lp0, max_nb)
    for (int k = 0; k < nt; k++) {
      const TaskDev &T = tasks[first + k];
      bool ok = T.use_ortho && !T.skip;
      int n = 0;
      if (ok) {
        n = (int)llround(T.disr_radius / drmin);
        // the discretised radius must be exactly n*drmin for the tables to apply
        ok = (n >= 1 && n <= kTiledMaxN && T.disr_radius == drmin * fmax(1.0, (double)n));
        ok = ok && (T.la_max + T.lb_max <= kTiledMaxLp);
        for (int d = 0; d < 3; d++)
          ok = ok && (-T.lb_cube[d] <= kTiledMaxNb);
      }
      nidx[k] = ok ? n : 0;
      cls[k] = ok ? (unsigned char)lp_class(T.la_max + T.lb_max) : 0;
      if (ok) {
        max_n = std::max(max_n, n), max_lp0 = std::max(max_lp0, T.la_max + T.lb_max);
        for (int d = 0; d < 3; d++)
          max_nb = std::max(max_nb, -T.lb_cube[d]);
      }
    }
    for (int c = 0; c <= kNumClasses; c++)
      tl.class_tt_first[c] = 0;
    int n_generic = 0;
    for (int k = 0; k < nt; k++) {
      if (nidx[k])
        tl.class_tt_first[cls[k] + 1]++;
      else
        n_generic++;
    }
    for (int c = 0; c < kNumClasses; c++)
      tl.class_tt_first[c + 1] += tl.class_tt_first[c];
    std::vector<int> pos(std::max(nt, 1));  // final position among the tiled (or the generic) tasks
    {
      int next[kNumClasses], next_generic = 0;
      for (int c = 0; c < kNumClasses; c++)
        next[c] = tl.class_tt_first[c];
      for (int k = 0; k < nt; k++)
        pos[k] = nidx[k] ? next[cls[k]]++ : next_generic++;
    }
    const int n_tiled = tl.class_tt_first[kNumClasses];
    tl.ntasks_tiled = n_tiled;
    tl.max_lp0 = max_lp0;
    tl.max_n = max_n;
    tl.max_nb = max_nb;
    const size_t g0 = generic_ids.size();
    generic_ids.resize(g0 + n_generic);
    if (n_tiled > 0) {
      tt.resize(n_tiled);
      tl.h_tt_task.resize(n_tiled);
      build_ktabs(L, max_n, heads, ktab);
    }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nt; k++) {
      const int it = first + k;
      if (nidx[k] == 0) {
        generic_ids[g0 + pos[k]] = it;
        continue;
      }
      const TaskDev &T = tasks[it];
      TTask X;
      for (int d = 0; d < 3; d++) {
        X.roff[d] = T.roffset[d];
        X.cc[d] = T.cubecenter[d] - L.shift_local[d];
        X.nb[d] = -T.lb_cube[d];
      }
      X.n = nidx[k], X.lp0 = T.la_max + T.lb_max, X.task = it;
      X.zl2 = T.zetp * 1.4426950408889634074;
      X.pad[0] = X.pad[1] = X.pad[2] = 0;
      // the cube bounds stored with the task must agree with the table's
      B200_ASSERT(heads[X.n].offset >= 0 && X.nb[0] == heads[X.n].nbx && X.nb[1] == heads[X.n].nby &&
                      X.nb[2] == heads[X.n].nbz,
                  "cube bounds disagree with the sphere table");
      tt[pos[k]] = X;
      tl.h_tt_task[pos[k]] = it;
    }
  }
  tm.tick("level: select tasks, K tables");
  if (tt.empty())
    return;
  B200_ASSERT(nblocks * kLpBuckets * kNumClasses < (size_t)1 << 31, "too many grid blocks");
  auto up = [&](auto **dst, const auto &vec) {
    using T = typename std::remove_reference<decltype(vec)>::type::value_type;
    dev_alloc(dst, std::max<size_t>(vec.size(), 1) * sizeof(T));
    if (!vec.empty())
      B200_CHECK(cudaMemcpyAsync(*dst, vec.data(), vec.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  };
  up(&tl.d_ttasks, tt);
  up(&tl.d_khead, heads);
  up(&tl.d_ktab, ktab);
  const std::vector<unsigned short> zmask = build_zmask();
  up(&tl.d_zmask, zmask);
  tm.tick("level: uploads");

  // pairs: count, scan, fill
  const size_t nbuckets = nblocks * kLpBuckets * kNumClasses;
  unsigned int *d_count = nullptr, *d_start = nullptr;
  d_count = (unsigned int *)scratch.get(0, (nbuckets + 1) * sizeof(unsigned int));
  d_start = (unsigned int *)scratch.get(1, (nbuckets + 1) * sizeof(unsigned int));
  B200_CHECK(cudaMemsetAsync(d_count, 0, (nbuckets + 1) * sizeof(unsigned int), s));
  PairGenArgs PA;
  PA.ttasks = tl.d_ttasks, PA.nttasks = (int)tt.size();
  PA.nx = L.npts_local[0], PA.ny = L.npts_local[1], PA.nz = L.npts_local[2];
  PA.Nx = L.npts_global[0], PA.Ny = L.npts_global[1], PA.Nz = L.npts_global[2];
  PA.nbx = nbx, PA.nby = nby, PA.nbz = nbz, PA.nblocks = (unsigned)nblocks;
  PA.khead = tl.d_khead, PA.ktab = tl.d_ktab;
  PA.bucket_count = d_count, PA.bucket_start = d_start, PA.bucket_cursor = nullptr, PA.pairs = nullptr;
  PA.keys = nullptr, PA.qbits = 0;
  const int pg_blocks = ((int)tt.size() + 127) / 128;
  pairgen_kernel<0><<<pg_blocks, 128, 0, s>>>(PA);
  B200_CHECK(cudaGetLastError());
  void *d_temp = nullptr;
  size_t temp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, d_count, d_start, (int)(nbuckets + 1), s);
  d_temp = scratch.get(2, temp_bytes);
  cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_count, d_start, (int)(nbuckets + 1), s);
  std::vector<unsigned int> start(nbuckets + 1);
  B200_CHECK(cudaMemcpyAsync(start.data(), d_start, (nbuckets + 1) * sizeof(unsigned int),
                             cudaMemcpyDeviceToHost, s));
  B200_CHECK(cudaStreamSynchronize(s));
  const size_t npairs = start[nbuckets];
  tm.tick("level: count pairs, scan");
  B200_ASSERT(npairs < ((size_t)1 << 31), "too many (task, block) pairs on one level");
  tl.npairs = (long long)npairs;
  dev_alloc(&tl.d_pairs, (npairs + kPairPad) * sizeof(TPair));
  B200_CHECK(cudaMemsetAsync(d_count, 0, (nbuckets + 1) * sizeof(unsigned int), s));
  unsigned long long *d_keys[2] = {nullptr, nullptr};
  TPair *d_pairs_alt = nullptr;
  // both key buffers, the second pair buffer and the sort's temporary storage share slots 3..5
  d_keys[0] = (unsigned long long *)scratch.get(3, 2 * std::max<size_t>(npairs, 1) * sizeof(unsigned long long));
  d_keys[1] = d_keys[0] + std::max<size_t>(npairs, 1);
  int qbits = 1, bbits = 1;
  while (((size_t)1 << qbits) < tt.size())
    qbits++;
  while (((size_t)1 << bbits) < nbuckets)
    bbits++;
  PA.bucket_cursor = d_count, PA.pairs = tl.d_pairs, PA.keys = d_keys[0], PA.qbits = qbits;
  pairgen_kernel<1><<<pg_blocks, 128, 0, s>>>(PA);
  B200_CHECK(cudaGetLastError());
  count_launch(4);
  // Order every bucket by task: neighbouring blocks then walk (nearly) the same
  // tasks in the same order at the same time, so that the tasks' table rows and
  // coefficients are shared through L1/L2 instead of being re-read from HBM; it
  // also makes the accumulation order (and so the results) reproducible.
  if (npairs > 1) {
    d_pairs_alt = (TPair *)scratch.get(4, (npairs + kPairPad) * sizeof(TPair));
    cub::DoubleBuffer<unsigned long long> kb(d_keys[0], d_keys[1]);
    cub::DoubleBuffer<uint4> vb((uint4 *)tl.d_pairs, (uint4 *)d_pairs_alt);
    void *d_sort_temp = nullptr;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, kb, vb, (int)npairs, 0, qbits + bbits, s);
    d_sort_temp = scratch.get(5, std::max<size_t>(sort_bytes, 1));
    cub::DeviceRadixSort::SortPairs(d_sort_temp, sort_bytes, kb, vb, (int)npairs, 0, qbits + bbits, s);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(s));
    count_launch(6);
    if ((TPair *)vb.Current() != tl.d_pairs)  // the sorted pairs ended up in the scratch buffer
      B200_CHECK(cudaMemcpyAsync(tl.d_pairs, d_pairs_alt, npairs * sizeof(TPair), cudaMemcpyDeviceToDevice, s));
  }
  if (npairs > 0)
    for (int i = 0; i < kPairPad; i++)
      B200_CHECK(cudaMemcpyAsync(tl.d_pairs + npairs + i, tl.d_pairs + npairs - 1, sizeof(TPair),
                                 cudaMemcpyDeviceToDevice, s));

  tm.tick("level: fill + sort pairs");
  // work items per lp class: a block's pairs of that class (contiguous, ordered by lp) cut into chunks
  std::vector<TWork> work;
  const size_t target_items = (size_t)148 * 64;
  const int chunk = (int)std::min<size_t>(kItemPairs, std::max<size_t>(128, npairs / target_items + 1));
  for (int cls = 0; cls < kNumClasses; cls++) {
    tl.class_work_first[cls] = (int)work.size();
    for (size_t b = 0; b < nblocks; b++) {
      const size_t bb = ((size_t)cls * nblocks + b) * kLpBuckets;
      const unsigned f = start[bb], e = start[bb + kLpBuckets];
      if (e == f)
        continue;
      const int bx = (int)(b % nbx), by = (int)((b / nbx) % nby), bz = (int)(b / ((size_t)nbx * nby));
      const int cnt = (int)(e - f), nchunks = (cnt + chunk - 1) / chunk, per = (cnt + nchunks - 1) / nchunks;
      for (int c = 0; c < nchunks; c++) {
        TWork W;
        W.x0 = bx * kBX, W.y0 = by * kBY, W.z0 = bz * kBZ, W.pad = 0;
        const int lo = (int)f + c * per, hi = (int)f + std::min((c + 1) * per, cnt);
        W.b[0] = lo, W.b[3] = hi;
        for (int i = 1; i < 3; i++) {  // end of the class's i-th lp within this chunk
          const int lp_next = std::min(kClassLo[cls] + i, kLpBuckets);
          W.b[i] = std::min(std::max((int)start[bb + lp_next], lo), hi);
        }
        work.push_back(W);
      }
    }
    // Items stay in spatial (block) order: warps that run concurrently then work on
    // neighbouring blocks and share the tasks' table rows through L2.  They are
    // handed out dynamically (atomic counter), so no size sorting is needed.
  }
  tl.class_work_first[kNumClasses] = (int)work.size();
  // TaskDev ids per class (for calls whose l growth pushes a class out of the tiled range)
  // (the tiled tasks are class-sorted: a class's ids are a slice of h_tt_task)
  for (int cls = 0; cls < kNumClasses; cls++) {
    const int f = tl.class_tt_first[cls], n = tl.class_tt_first[cls + 1] - f;
    tl.class_ntasks[cls] = n;
    dev_alloc(&tl.d_class_task_ids[cls], std::max<size_t>(n, 1) * sizeof(int));
    if (n > 0)
      B200_CHECK(cudaMemcpyAsync(tl.d_class_task_ids[cls], tl.h_tt_task.data() + f, n * sizeof(int),
                                 cudaMemcpyHostToDevice, s));
  }
  tl.nwork = (int)work.size();
  up(&tl.d_work, work);
  dev_alloc(&tl.d_counters, 2 * kNumClasses * 4 * sizeof(int));
  B200_CHECK(cudaStreamSynchronize(s));
  tm.tick("level: work items");
}

inline bool tiled_supports(const TiledLevel &tl, const int max_lp) {
  (void)max_lp;
  return tl.ntasks_tiled > 0;
}

// ---------------------------------------------------------------------------
// Device: the hot kernels
// ---------------------------------------------------------------------------
struct TiledArgs {
  const TPair *pairs;
  const TWork *work;         // work items of ONE lp class
  int nwork;
  int *counter;              // dynamic work distribution (zeroed before the launch)
  const unsigned char *ktab;
  const unsigned short *zmask;
  const TTask *ttasks;       // (roff, zl2) per tiled task
  int max_nb;
  int tt_first;              // first ttask of this class
  int coef_base, coef_stride;  // slot of ttask q: coef_base + (q - tt_first) * coef_stride
  double *coef;
  double *grid;
  int nx, ny, nz;            // npts_local
  double hx, hy, hz;
};

// Transposing warp reduction: on return lane L holds in v[0] the warp-wide sum
// of element `idx` (returned); lanes whose idx >= N hold zeros.
template <int N> struct WarpVecReduce {
  template <int OFF, int M>
  static __device__ __forceinline__ void step(double (&v)[N], int &idx, const int lane) {
    if constexpr (OFF >= 1) {
      if constexpr (M > 1) {
        constexpr int H = (M + 1) / 2;
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int q = 0; q < H; q++) {
          const double hi = (H + q < M) ? v[H + q] : 0.0;
          const double send = up ? v[q] : hi;
          const double keep = up ? hi : v[q];
          v[q] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        if (up)
          idx += H;
        step<OFF / 2, H>(v, idx, lane);
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
        step<OFF / 2, 1>(v, idx, lane);
      }
    }
  }
  static __device__ __forceinline__ int run(double (&v)[N], const int lane) {
    int idx = 0;
    step<16, N>(v, idx, lane);
    return idx;
  }
};

// Number of lanes that end up with identical copies after WarpVecReduce<N>.
template <int N> struct DupLanes {
  static constexpr int splits = (N > 16) ? 5 : (N > 8 ? 4 : (N > 4 ? 3 : (N > 2 ? 2 : (N > 1 ? 1 : 0))));
  static constexpr int value = 32 >> splits;
};

// Lanes that hold a REAL element after WarpVecReduce<N> (bit L set <=> lane L's
// (idx, v[0]) is element idx of the vector, not the zero padding of an odd split).
// A padded lane can carry an idx < N that belongs to another lane's element: harmless
// for atomic adds of its zero, fatal for plain stores -- those must test this mask.
template <int N> struct ValidLanes {
  static constexpr unsigned compute() {
    unsigned mask = 0u;
    for (int lane = 0; lane < 32; lane++) {
      int m = N, cnt = N;
      for (int off = 16; off >= 1; off /= 2) {
        if (m > 1) {
          const int h = (m + 1) / 2;
          cnt = (lane & off) ? (cnt - h > 0 ? cnt - h : 0) : (cnt < h ? cnt : h);
          m = h;
        }
      }
      if (cnt >= 1)
        mask |= 1u << lane;
    }
    return mask;
  }
  static constexpr unsigned value = compute();
};

// Integrate epilogue of one (pair, warp): every thread holds the z-contracted
// sums S0/S1 of its two columns; coefficient (lx,ly,lz) receives
// X[lx] * (Y0[ly] S0[lz] + Y1[ly] S1[lz]) summed over the warp.
template <int LP> struct IntegrateReduce {
  // one slice lz at a time keeps the live register count at T2(LP-lz)
  template <int LZ>
  static __device__ __forceinline__ void slice(const double (&X)[LP + 1], const double (&Y0)[LP + 1],
                                               const double (&Y1)[LP + 1], const double (&S0)[LP + 1],
                                               const double (&S1)[LP + 1], double *__restrict__ gcoef,
                                               const int lane) {
    if constexpr (LZ <= LP) {
      constexpr int L2 = LP - LZ;
      constexpr int M = (L2 + 1) * (L2 + 2) / 2;
      double part[M];
      int q = 0;
#pragma unroll
      for (int ly = 0; ly <= L2; ly++) {
        const double w = fma(Y0[ly], S0[LZ], Y1[ly] * S1[LZ]);
#pragma unroll
        for (int lx = 0; lx <= L2 - ly; lx++)
          part[q++] = X[lx] * w;
      }
      // at most 32 values per transposing reduction (one result per lane)
      constexpr int M1 = (M > 32) ? 32 : M, M2 = M - M1;
      auto emit = [&](int idx, const double v) {
        int ly = 0;
        while (idx > L2 - ly) {
          idx -= L2 - ly + 1;
          ly++;
        }
        atomicAdd(&gcoef[coset(idx, ly, LZ)], v);
      };
      {
        double head[M1];
#pragma unroll
        for (int i = 0; i < M1; i++)
          head[i] = part[i];
        const int idx = WarpVecReduce<M1>::run(head, lane);
        if ((lane & (DupLanes<M1>::value - 1)) == 0 && idx < M1 && head[0] != 0.0)
          emit(idx, head[0]);
      }
      if constexpr (M2 > 0) {
        double rest[M2];
#pragma unroll
        for (int i = 0; i < M2; i++)
          rest[i] = part[M1 + i];
        const int idx = WarpVecReduce<M2>::run(rest, lane);
        if ((lane & (DupLanes<M2>::value - 1)) == 0 && idx < M2 && rest[0] != 0.0)
          emit(M1 + idx, rest[0]);
      }
      slice<LZ + 1>(X, Y0, Y1, S0, S1, gcoef, lane);
    }
  }
};

// Row pitch (doubles) of the per-pair scratch: even (16-byte rows) except lp = 0.
__host__ __device__ constexpr int row_pitch(const int lp) { return lp == 0 ? 1 : ((lp + 2) / 2) * 2; }
__host__ __device__ constexpr int stage_doubles(const int lp) {
  return 32 * row_pitch(lp) + ((ncoset(lp) + 1) / 2) * 2;
}

// Loads one scratch row (LP+1 doubles) with 16-byte accesses where possible.
template <int LP> __device__ __forceinline__ void load_row(const double *__restrict__ row, double (&z)[LP + 1]) {
  if constexpr (LP == 0) {
    z[0] = row[0];
  } else {
#pragma unroll
    for (int l = 0; l + 1 <= LP; l += 2) {
      const double2 v = *reinterpret_cast<const double2 *>(row + l);
      z[l] = v.x, z[l + 1] = v.y;
    }
    if constexpr ((LP & 1) == 0)
      z[LP] = row[LP];
  }
}

// The 16 planes of a block in pairs, entered at the first pair any lane needs
// and left after the last one (both warp-uniform, precomputed per pair).
#if B200_KBZ == 16
#define B200_PLANES(BODY)                                                      \
  switch (wlo >> 1) {                                                          \
  case 0: BODY(0) BODY(1) if (whi <= 1) break;                                 \
  case 1: BODY(2) BODY(3) if (whi <= 3) break;                                 \
  case 2: BODY(4) BODY(5) if (whi <= 5) break;                                 \
  case 3: BODY(6) BODY(7) if (whi <= 7) break;                                 \
  case 4: BODY(8) BODY(9) if (whi <= 9) break;                                 \
  case 5: BODY(10) BODY(11) if (whi <= 11) break;                              \
  case 6: BODY(12) BODY(13) if (whi <= 13) break;                              \
  default: BODY(14) BODY(15)                                                   \
  }
#else
#define B200_PLANES(BODY)                                                      \
  switch (wlo >> 1) {                                                          \
  case 0: BODY(0) BODY(1) if (whi <= 1) break;                                 \
  case 1: BODY(2) BODY(3) if (whi <= 3) break;                                 \
  case 2: BODY(4) BODY(5) if (whi <= 5) break;                                 \
  default: BODY(6) BODY(7)                                                     \
  }
#endif

// Sign-extended byte of `v` that a left shift by `sh` moves to the top.
__device__ __forceinline__ int top_byte(const unsigned v, const unsigned sh) { return (int)(v << sh) >> 24; }

// The kernels' dynamic shared memory, addressed through the symbol (an index, not a
// generic pointer: the compiler then folds the window base into the LDS/STS).
extern __shared__ double tiled_smem[];

// Per-lane constants of the pair loops.
struct LaneCtx {
  const uint4 *__restrict__ pairs;
  const double *__restrict__ ttasks;     // the level's TTask records, as doubles (10 per record)
  int e2t_index;                         // 2^(j/64) table in tiled_smem, in doubles
  double *__restrict__ coef0;            // coef + coef_base
  const unsigned char *__restrict__ ktab;
  int zm_index;                          // plane-mask table (biased) in tiled_smem, in 16-bit units
  int ws_index;                          // this warp's scratch (two stages) in tiled_smem, in doubles
  double my_h;
  int my_axis, my_t;
  int klane;                             // lj * kKPitch + li
  unsigned sel;                          // left shift that moves my axis' byte of opk to the top
  int tt_first, coef_stride;
  int lane, li, lj;
};

// All pairs [first, last) of ONE lp for this warp's block.
template <bool COLLOCATE, int LP, int STAGE>
__device__ __forceinline__ void run_pairs(const LaneCtx &c, const int first, const int last,
                                          double (&acc0)[kBZ], double (&acc1)[kBZ]) {
  if (first >= last)
    return;
  constexpr int PITCH = row_pitch(LP);
  constexpr int NC = (LP + 1) * (LP + 2) * (LP + 3) / 6;
  constexpr int NCL = (NC + 31) / 32;  // coefficient registers per lane
  const int lane = c.lane;
  __syncwarp();  // the previous run's scratch reads are complete

  // ---- software pipeline: the next pair's table entry and coefficients are fetched
  // while the current pair is processed.  The pipeline registers (e_n, roff_n, o_n,
  // c_n) are CONSUMED (scratch stores) before the next fetch overwrites them and the
  // pair records are re-read from L1 instead of being rotated through registers:
  // ptxas 12.9 (-O1 and up) miscompiles "copy, then overwrite the source" rotations
  // of loaded values in the large high-lp loop bodies -- later pairs of a run then
  // see stale data (tests/test_b200_parity.py::test_multi_pair_items pins this).
  double zl2_n, roff_n, c_n[NCL];
  int o_n;
#define B200_FETCH(R)                                                          \
  {                                                                            \
    o_n = top_byte(R.z, c.sel);                                                \
    const double *t_ = c.ttasks + (size_t)R.x * (sizeof(TTask) / sizeof(double)); \
    roff_n = __ldg(t_ + c.my_axis);                                            \
    zl2_n = __ldg(t_ + 3);                                                     \
    if (COLLOCATE) {                                                           \
      const double *c_ = c.coef0 + ((R.x - (unsigned)c.tt_first) * (unsigned)c.coef_stride + (unsigned)lane); \
      _Pragma("unroll") for (int k = 0; k < NCL; k++)                          \
        c_n[k] = (lane + 32 * k < NC) ? c_[32 * k] : 0.0;                      \
    }                                                                          \
  }
  const uint4 *pp = c.pairs + first;
  const uint4 *const pend = c.pairs + last;
  // (opaque: kept in a vector register pair; as a uniform register ptxas copies it
  // into fresh vector registers for every one of the loop's loads)
  asm volatile("" : "+l"(pp));
  {
    const uint4 Rf = __ldg(pp);
    B200_FETCH(Rf)
  }

  // The pair array is padded with kPairPad copies of its last record: the loop reads one
  // record ahead and prefetches eight ahead without clamping.
  int stage = 0;
  for (; pp < pend; pp++) {
    const uint4 R0 = __ldg(pp);                          // L1 hit (read as Rn one iteration ago)
    const uint4 Rn = __ldg(pp + 1);                      // same line 7 times out of 8
    asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + 8));

    // scratch: my table entry times the powers of (x - xp); the coefficients
    double *ws = tiled_smem + c.ws_index + stage;
    stage = STAGE - stage;
    {
      // my table entry exp(-zetp x^2), computed here: no exp tables in HBM.  (Computing it
      // one pair ahead, in the fetch stage, was measured slower: 9.0 / 11.6 ms against
      // 8.1 / 10.5 ms per H2O-256 collocate / integrate.)
      const double x = (double)(c.my_t - o_n) * c.my_h - roff_n;
      const double e_x = exp_neg_tab(zl2_n, x, tiled_smem + c.e2t_index);
      double *row = ws + lane * PITCH;
      if constexpr (LP == 0) {
        row[0] = e_x;
      } else {
        double v0 = e_x;
#pragma unroll
        for (int l = 0; l + 1 <= LP; l += 2) {
          const double v1 = v0 * x;
          *reinterpret_cast<double2 *>(row + l) = make_double2(v0, v1);
          v0 = v1 * x;
        }
        if constexpr ((LP & 1) == 0)
          row[LP] = v0;
      }
      if (COLLOCATE) {
#pragma unroll
        for (int k = 0; k < NCL; k++)
          if (lane + 32 * k < NC)
            ws[32 * PITCH + lane + 32 * k] = c_n[k];
      }
    }
    // (the tail fetches for the record after the run -- the next run's first pair or
    // the padding: harmless, and the loop stays branch-free)
    B200_FETCH(Rn)

    // sphere masks of my two columns: K+1 from the sphere table, then the plane mask
    const unsigned char *kp = c.ktab + (R0.y + (unsigned)c.klane);
    const unsigned k0 = __ldg(kp), k1 = __ldg(kp + 4 * kKPitch);
    const int ozb = (int)R0.z >> 24;  // oz; the table index is biased
    const unsigned short *s_zm = reinterpret_cast<const unsigned short *>(tiled_smem) + c.zm_index;
    const unsigned mask0 = s_zm[(int)(k0 * kZmPitch) + ozb];
    const unsigned mask1 = s_zm[(int)(k1 * kZmPitch) + ozb];
    const int wlo = (int)((R0.z >> 16) & 15u), whi = (int)((R0.z >> 20) & 15u);
    __syncwarp();

    const double *tZ = ws + 16 * PITCH;
    double X[LP + 1], Y0[LP + 1], Y1[LP + 1];
    load_row<LP>(ws + c.li * PITCH, X);
    load_row<LP>(ws + (8 + c.lj) * PITCH, Y0);
    load_row<LP>(ws + (12 + c.lj) * PITCH, Y1);

    if (COLLOCATE) {
      // E[ly][lz] = sum_lx C[lx,ly,lz] X[lx]  (shared by both columns: same x)
      // D_c[lz]   = sum_ly E[ly][lz] Y_c[ly]
      const double *C = ws + 32 * PITCH;
      double D0[LP + 1], D1[LP + 1];
#pragma unroll
      for (int lz = 0; lz <= LP; lz++)
        D0[lz] = 0.0, D1[lz] = 0.0;
#pragma unroll
      for (int ly = 0; ly <= LP; ly++) {
#pragma unroll
        for (int lz = 0; lz <= LP - ly; lz++) {
          double ev = C[coset(0, ly, lz)] * X[0];
#pragma unroll
          for (int lx = 1; lx <= LP - ly - lz; lx++)
            ev = fma(C[coset(lx, ly, lz)], X[lx], ev);
          D0[lz] = fma(ev, Y0[ly], D0[lz]);
          D1[lz] = fma(ev, Y1[ly], D1[lz]);
        }
      }
#define B200_BODY(p)                                                           \
  {                                                                            \
    double z[LP + 1];                                                          \
    load_row<LP>(tZ + (p)*PITCH, z);                                           \
    PFma<LP>::col(acc0[p], D0, z, mask0 & (1u << (p)));                        \
    PFma<LP>::col(acc1[p], D1, z, mask1 & (1u << (p)));                        \
  }
      B200_PLANES(B200_BODY)
#undef B200_BODY
    } else {
      double S0[LP + 1], S1[LP + 1];
#pragma unroll
      for (int l = 0; l <= LP; l++)
        S0[l] = 0.0, S1[l] = 0.0;
#define B200_BODY(p)                                                           \
  {                                                                            \
    double z[LP + 1];                                                          \
    load_row<LP>(tZ + (p)*PITCH, z);                                           \
    PFma<LP>::integ(S0, acc0[p], z, mask0 & (1u << (p)));                      \
    PFma<LP>::integ(S1, acc1[p], z, mask1 & (1u << (p)));                      \
  }
      B200_PLANES(B200_BODY)
#undef B200_BODY
      // table entries outside the cube are zero and inactive columns have S = 0:
      // nothing spurious enters the warp-wide sums
      double *__restrict__ gcoef = c.coef0 + (R0.x - (unsigned)c.tt_first) * (unsigned)c.coef_stride;
      if constexpr (LP <= 3) {
        // few coefficients: one transposing reduction over all of them
        double part[NC];
#pragma unroll
        for (int ly = 0; ly <= LP; ly++) {
#pragma unroll
          for (int lz = 0; lz <= LP - ly; lz++) {
            const double w = fma(Y0[ly], S0[lz], Y1[ly] * S1[lz]);
#pragma unroll
            for (int lx = 0; lx <= LP - ly - lz; lx++)
              part[coset(lx, ly, lz)] = X[lx] * w;
          }
        }
        const int idx = WarpVecReduce<NC>::run(part, lane);
        if ((lane & (DupLanes<NC>::value - 1)) == 0 && idx < NC && part[0] != 0.0)
          atomicAdd(&gcoef[idx], part[0]);
      } else {
        IntegrateReduce<LP>::template slice<0>(X, Y0, Y1, S0, S1, gcoef, lane);
      }
    }
  }
#undef B200_FETCH
}

template <bool COLLOCATE, int LPLO, int LPHI, int SUB = 0>
__global__ void __launch_bounds__(kTiledThreads, (LPHI <= B200_LPHI_LO) ? B200_CTAS_LO : B200_CTAS_HI) tiled_kernel(const TiledArgs A) {
  double *const smem = tiled_smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int STAGE = stage_doubles(LPHI);
  // shared memory: [per-warp scratch, two stages][2^(j/64), 64 doubles][plane-mask table]
  unsigned short *s_zm = (unsigned short *)(smem + (size_t)kTiledWarps * 2 * STAGE + 64);
  for (int q = tid; q < kZmRows * kZmPitch / 2; q += kTiledThreads)
    ((unsigned *)s_zm)[q] = ((const unsigned *)A.zmask)[q];
  if (tid < 64)
    smem[(size_t)kTiledWarps * 2 * STAGE + tid] = exp2((double)tid * (1.0 / 64.0));
  __syncthreads();  // the only CTA-wide barrier

  LaneCtx c;
  c.lane = lane, c.li = lane & 7, c.lj = lane >> 3;  // my columns: (li, lj) and (li, lj + 4)
  // which table entry this lane fetches: axis and block-local index
  const int my_axis = (lane < 8) ? 0 : ((lane < 16) ? 1 : 2);
  c.my_t = (lane < 8) ? lane : ((lane < 16) ? lane - 8 : lane - 16);
  c.my_h = (my_axis == 0) ? A.hx : ((my_axis == 1) ? A.hy : A.hz);
  c.sel = (my_axis == 0) ? 24u : ((my_axis == 1) ? 16u : 0u);
  c.my_axis = my_axis;
  c.ttasks = reinterpret_cast<const double *>(A.ttasks);
  c.e2t_index = kTiledWarps * 2 * STAGE;
  c.coef0 = A.coef + A.coef_base;
  c.tt_first = A.tt_first, c.coef_stride = A.coef_stride;
  c.pairs = (const uint4 *)A.pairs;
  c.ktab = A.ktab;
  c.klane = c.lj * kKPitch + c.li;
  // Opaque to the compiler from here on: it would otherwise re-derive these per-lane
  // constants from the thread index inside the pair loops.
  asm volatile("" : "+r"(c.my_t), "+r"(c.klane), "+r"(c.sel));
  c.zm_index = (kTiledWarps * 2 * STAGE + 64) * 4 + kZmBias;
  c.ws_index = warp * 2 * STAGE;

  // persistent warps: work items are handed out in spatial order
  for (;;) {
    int iw = 0;
    if (lane == 0)
      iw = atomicAdd(A.counter, 1);
    iw = __shfl_sync(0xffffffffu, iw, 0);
    if (iw >= A.nwork)
      break;
    const int4 W0 = ((const int4 *)A.work)[2 * iw], W1 = ((const int4 *)A.work)[2 * iw + 1];
    if constexpr (LPLO == LPHI) {  // experimental per-lp launches: only this lp's pairs of the item
      if (((SUB == 0) ? W0.w : ((SUB == 1) ? W1.x : W1.y)) >= ((SUB == 0) ? W1.x : ((SUB == 1) ? W1.y : W1.z)))
        continue;
    }
    const int x0 = W0.x, y0 = W0.y, z0 = W0.z;
    const int vx = min(kBX, A.nx - x0), vy = min(kBY, A.ny - y0), vz = min(kBZ, A.nz - z0);
    const size_t sy = A.nx, sz = (size_t)A.nx * A.ny;
    double *g0 = A.grid + (size_t)z0 * sz + (size_t)(y0 + c.lj) * sy + x0 + c.li;
    double *g1 = g0 + 4 * sy;
    const bool col0 = (c.li < vx && c.lj < vy), col1 = (c.li < vx && c.lj + 4 < vy);

    // Points outside the grid (partial edge blocks) need no masking in the pair
    // loops: they integrate zeros and their collocated values are never flushed.
    double acc0[kBZ], acc1[kBZ];
#pragma unroll
    for (int p = 0; p < kBZ; p++) {
      acc0[p] = 0.0, acc1[p] = 0.0;
      if (!COLLOCATE && p < vz) {
        if (col0)
          acc0[p] = g0[p * sz];
        if (col1)
          acc1[p] = g1[p * sz];
      }
    }

    if constexpr (LPLO == LPHI)
      run_pairs<COLLOCATE, LPLO, STAGE>(c, (SUB == 0) ? W0.w : ((SUB == 1) ? W1.x : W1.y),
                                        (SUB == 0) ? W1.x : ((SUB == 1) ? W1.y : W1.z), acc0, acc1);
    else
      run_pairs<COLLOCATE, LPLO, STAGE>(c, W0.w, W1.x, acc0, acc1);
    if constexpr (LPLO + 1 <= LPHI)
      run_pairs<COLLOCATE, LPLO + 1, STAGE>(c, W1.x, W1.y, acc0, acc1);
    if constexpr (LPLO + 2 <= LPHI)
      run_pairs<COLLOCATE, LPLO + 2, STAGE>(c, W1.y, W1.z, acc0, acc1);

    if (COLLOCATE) {
#pragma unroll
      for (int p = 0; p < kBZ; p++) {
        if (p < vz) {
          if (col0 && acc0[p] != 0.0)
            atomicAdd(&g0[p * sz], acc0[p]);
          if (col1 && acc1[p] != 0.0)
            atomicAdd(&g1[p * sz], acc1[p]);
        }
      }
    }
  }  // work items
}

inline size_t tiled_smem_bytes(const int lphi) {
  return (size_t)kTiledWarps * 2 * stage_doubles(lphi) * sizeof(double) + 64 * sizeof(double) +
         (size_t)kZmRows * kZmPitch * sizeof(unsigned short);
}

template <bool COLLOCATE, int LPLO, int LPHI, int SUB = 0>
inline void launch_tiled_class(const TiledArgs &A, const TiledLevel &tl, cudaStream_t s) {
  (void)tl;
  const size_t bytes = tiled_smem_bytes(LPHI);
  B200_ASSERT(bytes <= 200 * 1024, "tiled kernel: shared memory budget exceeded");
  B200_CHECK(cudaFuncSetAttribute(tiled_kernel<COLLOCATE, LPLO, LPHI, SUB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  int per_sm = 1;
  B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tiled_kernel<COLLOCATE, LPLO, LPHI, SUB>,
                                                          kTiledThreads, bytes));
  const int grid = std::min((A.nwork + kTiledWarps - 1) / kTiledWarps, 148 * std::max(per_sm, 1));
  tiled_kernel<COLLOCATE, LPLO, LPHI, SUB><<<grid, kTiledThreads, bytes, s>>>(A);
  B200_CHECK(cudaGetLastError());
  count_launch();
}

// Launches the tiled kernels of every lp class that stays within the tiled
// range for this call's l growth; returns a bit mask of the classes that must
// be handled by the generic kernel instead.
template <bool COLLOCATE> inline unsigned launch_tiled(TiledLevel &tl, const GridLaunch &L) {
  if (tl.ntasks_tiled == 0)
    return 0u;
  B200_ASSERT(L.dl >= 0 && L.dl < 8, "unexpected l growth");
  TiledArgs A;
  A.pairs = tl.d_pairs;
  A.ktab = tl.d_ktab, A.zmask = tl.d_zmask;
  A.ttasks = tl.d_ttasks, A.max_nb = tl.max_nb;
  A.coef = L.coef, A.grid = L.grid;
  A.nx = L.level.npts_local[0], A.ny = L.level.npts_local[1], A.nz = L.level.npts_local[2];
  A.hx = L.level.dh[0], A.hy = L.level.dh[4], A.hz = L.level.dh[8];
  unsigned leftover = 0u;
  // GRID_B200_PER_LP=1 (experimental, off by default): one kernel per lp instead of one per
  // lp class -- each lp loop then gets its own register allocation and a smaller scratch, at
  // the price of flushing / loading a block once per lp.
  static const bool per_lp = (getenv("GRID_B200_PER_LP") != nullptr && atoi(getenv("GRID_B200_PER_LP")) != 0);
  B200_CHECK(cudaMemsetAsync(tl.d_counters + (COLLOCATE ? 0 : kNumClasses * 4), 0, kNumClasses * 4 * sizeof(int),
                             L.stream));
  for (int cls = 0; cls < kNumClasses; cls++) {
    A.counter = tl.d_counters + (COLLOCATE ? 0 : kNumClasses * 4) + cls * 4;
    A.work = tl.d_work + tl.class_work_first[cls];
    A.nwork = tl.class_work_first[cls + 1] - tl.class_work_first[cls];
    if (A.nwork == 0)
      continue;
    const int lo = kClassLo[cls] + L.dl, hi = kClassHi[cls] + L.dl;
    if (hi > kTiledMaxLpCall) {
      leftover |= 1u << cls;
      continue;
    }
    A.tt_first = tl.class_tt_first[cls];
    A.coef_base = tl.coef_base[L.dl][cls];
    A.coef_stride = (ncoset(hi) + 1) / 2 * 2;  // slots are padded to even sizes (ensure_coef_offsets)
    cudaStream_t s = L.stream;
    if (per_lp && lo == 0) {  // the dominant class only (water: lp 0..2)
      int *const counter0 = A.counter;
      A.counter = counter0 + 0;
      launch_tiled_class<COLLOCATE, 0, 0, 0>(A, tl, s);
      A.counter = counter0 + 1;
      launch_tiled_class<COLLOCATE, 1, 1, 1>(A, tl, s);
      A.counter = counter0 + 2;
      launch_tiled_class<COLLOCATE, 2, 2, 2>(A, tl, s);
      A.counter = counter0;
      continue;
    }
    if (lo == 0) launch_tiled_class<COLLOCATE, 0, 2>(A, tl, s);
    else if (lo == 1) launch_tiled_class<COLLOCATE, 1, 3>(A, tl, s);
    else if (lo == 2) launch_tiled_class<COLLOCATE, 2, 4>(A, tl, s);
    else if (lo == 3 && hi == 5) launch_tiled_class<COLLOCATE, 3, 5>(A, tl, s);
    else if (lo == 3) launch_tiled_class<COLLOCATE, 3, 4>(A, tl, s);
    else if (lo == 4 && hi == 6) launch_tiled_class<COLLOCATE, 4, 6>(A, tl, s);
    else if (lo == 4) launch_tiled_class<COLLOCATE, 4, 5>(A, tl, s);
    else if (lo == 5 && hi == 7) launch_tiled_class<COLLOCATE, 5, 7>(A, tl, s);
    else if (lo == 5) launch_tiled_class<COLLOCATE, 5, 6>(A, tl, s);
    else if (lo == 6 && hi == 7) launch_tiled_class<COLLOCATE, 6, 7>(A, tl, s);
    else leftover |= 1u << cls;
  }
  return leftover;
}

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
                          const TaskVec &h_tasks, double *stats, cudaStream_t s) {
  (void)h_tasks;
  LevelDev *d_levels = nullptr;
  double *d_out = nullptr;
  dev_alloc(&d_levels, levels.size() * sizeof(LevelDev));
  dev_alloc(&d_out, 3 * sizeof(double));
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
  dev_free(d_levels);
  dev_free(d_out);
}

}  // namespace b200
