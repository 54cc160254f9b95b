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
//     bucketed by (block, lp); periodic images become separate pairs;
//   * the separable Gaussian factors exp(-zetp (x-xp)^2) of every task are
//     tabulated once per task list (they depend on geometry only) -- there is
//     no exp() in the hot loop; a (pair, warp) step loads exactly one table
//     entry per lane (8 x-, 8 y-, 16 z-entries);
//   * which points of a cube are inside the (discretised-radius) sphere is a
//     table lookup: because the reference discretises the radius to n*drmin
//     (ref/grid_ref_collint.h:237-239) the admissible z-extent of a column
//     depends only on (n, |j|, |i|); the table is built on the host with the
//     reference's own expressions, so the SET of touched points is identical;
//   * planes no lane of the warp needs are skipped warp-uniformly;
//   * the next pair's record, table entry and coefficients are prefetched
//     while the current pair is processed.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "b200_generic.cuh"

namespace b200 {

constexpr int kBX = 8, kBY = 8, kBZ = 16;     // block (warp footprint)
constexpr int kTiledThreads = 256;            // 8 autonomous warps
constexpr int kTiledWarps = kTiledThreads / 32;
constexpr int kTiledMaxLp = 6;                // beyond: generic kernels
constexpr int kTiledMaxN = 60;                // max discretised radius index
constexpr int kLpBuckets = kTiledMaxLp + 1;
constexpr int kItemPairs = 1024;              // pairs per work item (upper bound)

struct TTask {            // static per-task data of the tiled path
  double roff[3];
  int cc[3];              // cubecenter - shift_local (not wrapped)
  int nb[3];              // -lb_cube per axis
  int n;                  // discretised radius index
  int lp0;                // la_max + lb_max
  int task;               // index into the TaskDev array
};

struct TPair {            // 8 bytes
  int ttask;
  signed char o[3];       // cube centre relative to the block origin
  unsigned char lp0;
};

struct TWork {
  int x0, y0, z0;         // block origin (local grid indices)
  int first, last;        // pair range
};

struct KTabHeader {       // per radius index n
  int offset;             // into the byte table
  int nbx, nby, nbz;      // -lb per axis
};

struct TiledLevel {
  long long npairs = 0;
  int nwork = 0;
  int ntasks_tiled = 0;
  int max_lp0 = 0;
  int ktab_bytes = 0;
  int max_n = 0;
  int P = 0;              // exp-table pitch per axis: entries g = -P/2+1 .. P/2
  TTask *d_ttasks = nullptr;
  TPair *d_pairs = nullptr;
  TWork *d_work = nullptr;
  KTabHeader *d_khead = nullptr;
  signed char *d_ktab = nullptr;
  double *d_etab = nullptr;          // [ttask][3][P]
  int *d_tcoef[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void release() {
    cudaFree(d_ttasks), cudaFree(d_pairs), cudaFree(d_work), cudaFree(d_khead), cudaFree(d_ktab);
    cudaFree(d_etab);
    for (auto &p : d_tcoef) {
      cudaFree(p);
      p = nullptr;
    }
    d_ttasks = nullptr, d_pairs = nullptr, d_work = nullptr, d_khead = nullptr, d_ktab = nullptr;
    d_etab = nullptr;
    npairs = 0, nwork = 0, ntasks_tiled = 0;
  }
};

// ---------------------------------------------------------------------------
// Host: sphere-extent tables.  K[n][mj][mi] = largest z pair-distance kd such
// that the point (kd, mj, mi) is visited by the reference's loop nest
// (ref/grid_ref_collint.h:144-147, 46-50, 237-254), or -1.
// ---------------------------------------------------------------------------
inline void build_ktabs(const LevelDev &L, const int max_n, std::vector<KTabHeader> &heads,
                        std::vector<signed char> &tab) {
  const double h[3] = {L.dh[0], L.dh[4], L.dh[8]};
  const double hinv[3] = {L.dh_inv[0], L.dh_inv[4], L.dh_inv[8]};
  const double drmin = fmin(h[0], fmin(h[1], h[2]));
  heads.assign(max_n + 1, KTabHeader{0, 0, 0, 0});
  tab.clear();
  for (int n = 1; n <= max_n; n++) {
    const double R = drmin * fmax(1.0, (double)n);
    int nb[3];
    for (int d = 0; d < 3; d++)
      nb[d] = -(int)ceil(-1e-8 - R * hinv[d]);
    KTabHeader H;
    H.offset = (int)tab.size();
    H.nbx = nb[0], H.nby = nb[1], H.nbz = nb[2];
    const int px = nb[0] + 1, py = nb[1] + 1;
    std::vector<signed char> K((size_t)px * py, (signed char)-1);
    for (int kd = 0; kd <= nb[2]; kd++) {
      const double kr = kd * h[2];
      const double krem = R * R - kr * kr;
      const int jstart = (int)ceil(-1e-8 - sqrt(fmax(0.0, krem)) * hinv[1]);
      for (int jd = 0; jd <= -jstart && jd <= nb[1]; jd++) {
        const double jr = jd * h[1];
        const double jrem = krem - jr * jr;
        const int istart = (int)ceil(-1e-8 - sqrt(fmax(0.0, jrem)) * hinv[0]);
        for (int id = 0; id <= -istart && id <= nb[0]; id++)
          K[(size_t)jd * px + id] = (signed char)std::max<int>(K[(size_t)jd * px + id], kd);
      }
    }
    tab.insert(tab.end(), K.begin(), K.end());
    heads[n] = H;
  }
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
  double hx, hy, hz, drmin;
  unsigned int *bucket_count;     // pass 0
  const unsigned int *bucket_start;  // pass 1
  unsigned int *bucket_cursor;    // pass 1
  TPair *pairs;                   // pass 1
};

__device__ inline int floor_div(const int a, const int b) {  // b > 0
  return (a >= 0) ? a / b : -((-a + b - 1) / b);
}
__device__ inline int rel_dmin(const int a, const int b) {  // min pair distance over [a,b]
  return (a <= 1 && b >= 0) ? 0 : ((a > 1) ? a - 1 : -b);
}

template <int PASS> __global__ void pairgen_kernel(const PairGenArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.nttasks)
    return;
  const TTask X = A.ttasks[q];
  const double R = A.drmin * (double)X.n;
  const double R2 = R * R * (1.0 + 1e-9) + 1e-12;
  const int nloc[3] = {A.nx, A.ny, A.nz}, N[3] = {A.Nx, A.Ny, A.Nz};
  const int B[3] = {kBX, kBY, kBZ}, nblk[3] = {A.nbx, A.nby, A.nbz};
  const double h[3] = {A.hx, A.hy, A.hz};
  // per axis: image range
  int m_lo[3], m_hi[3];
  for (int d = 0; d < 3; d++) {
    const int lo = X.cc[d] - X.nb[d], hi = X.cc[d] + 1 + X.nb[d];
    // images m with [lo+mN, hi+mN] intersecting [0, nloc)
    // (a superset; images that miss the local grid are dropped below)
    m_lo[d] = floor_div(-hi, N[d]);
    m_hi[d] = floor_div(nloc[d] - 1 - lo, N[d]) + 1;
  }
  for (int mz = m_lo[2]; mz <= m_hi[2]; mz++) {
    const int cz = X.cc[2] + mz * N[2];
    const int az = max(cz - X.nb[2], 0), bz = min(cz + 1 + X.nb[2], nloc[2] - 1);
    if (az > bz)
      continue;
    for (int tz = az / B[2]; tz <= bz / B[2] && tz < nblk[2]; tz++) {
      const int oz = cz - tz * B[2];
      const double dz = rel_dmin(max(az, tz * B[2]) - cz, min(bz, tz * B[2] + B[2] - 1) - cz) * h[2];
      for (int my = m_lo[1]; my <= m_hi[1]; my++) {
        const int cy = X.cc[1] + my * N[1];
        const int ay = max(cy - X.nb[1], 0), by = min(cy + 1 + X.nb[1], nloc[1] - 1);
        if (ay > by)
          continue;
        for (int ty = ay / B[1]; ty <= by / B[1] && ty < nblk[1]; ty++) {
          const int oy = cy - ty * B[1];
          const double dy = rel_dmin(max(ay, ty * B[1]) - cy, min(by, ty * B[1] + B[1] - 1) - cy) * h[1];
          if (dz * dz + dy * dy > R2)
            continue;
          for (int mx = m_lo[0]; mx <= m_hi[0]; mx++) {
            const int cx = X.cc[0] + mx * N[0];
            const int ax = max(cx - X.nb[0], 0), bx = min(cx + 1 + X.nb[0], nloc[0] - 1);
            if (ax > bx)
              continue;
            for (int tx = ax / B[0]; tx <= bx / B[0] && tx < nblk[0]; tx++) {
              const int ox = cx - tx * B[0];
              const double dx = rel_dmin(max(ax, tx * B[0]) - cx, min(bx, tx * B[0] + B[0] - 1) - cx) * h[0];
              if (dx * dx + dy * dy + dz * dz > R2)
                continue;
              const unsigned bucket = ((unsigned)((tz * nblk[1] + ty) * nblk[0] + tx)) * kLpBuckets + X.lp0;
              if (PASS == 0) {
                atomicAdd(&A.bucket_count[bucket], 1u);
              } else {
                const unsigned pos = A.bucket_start[bucket] + atomicAdd(&A.bucket_cursor[bucket], 1u);
                TPair P;
                P.ttask = q;
                P.o[0] = (signed char)ox, P.o[1] = (signed char)oy, P.o[2] = (signed char)oz;
                P.lp0 = (unsigned char)X.lp0;
                A.pairs[pos] = P;
              }
            }
          }
        }
      }
    }
  }
}

// exp tables: etab[(q*3+d)*P + (g + P/2 - 1)] = exp(-zetp (g*h_d - roff_d)^2)
__global__ void etab_kernel(const TTask *ttasks, const TaskDev *tasks, const int nttasks, const int P,
                            const double hx, const double hy, const double hz, double *etab) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)nttasks * 3 * P;
  if (idx >= total)
    return;
  const int e = (int)(idx % P), d = (int)((idx / P) % 3);
  const int q = (int)(idx / ((size_t)3 * P));
  const TTask &X = ttasks[q];
  const int g = e - (P / 2 - 1);
  double v = 0.0;
  if (g >= -X.nb[d] && g <= X.nb[d] + 1) {
    const double h = (d == 0) ? hx : ((d == 1) ? hy : hz);
    const double x = g * h - X.roff[d];
    v = exp(-tasks[X.task].zetp * x * x);
  }
  etab[idx] = v;
}

__global__ void tcoef_kernel(const TTask *ttasks, const int nttasks, const int *coef_offsets, int *tcoef) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nttasks)
    tcoef[q] = coef_offsets[ttasks[q].task];
}

// ---------------------------------------------------------------------------
// Host: per-level build.
// ---------------------------------------------------------------------------
inline void build_tiled_level(TiledLevel &tl, const LevelDev &L, const std::vector<TaskDev> &tasks,
                              const TaskDev *d_tasks, const int first, const int last,
                              std::vector<int> &generic_ids, cudaStream_t s) {
  tl.release();
  const double h[3] = {L.dh[0], L.dh[4], L.dh[8]};
  const double drmin = fmin(h[0], fmin(h[1], h[2]));
  const int nbx = (L.npts_local[0] + kBX - 1) / kBX, nby = (L.npts_local[1] + kBY - 1) / kBY,
            nbz = (L.npts_local[2] + kBZ - 1) / kBZ;
  const size_t nblocks = (size_t)nbx * nby * nbz;

  std::vector<TTask> tt;
  int max_n = 0, max_lp0 = 0, max_nb = 0;
  for (int it = first; it < last; it++) {
    const TaskDev &T = tasks[it];
    bool ok = T.use_ortho && !T.skip;
    int n = 0;
    if (ok) {
      n = (int)llround(T.disr_radius / drmin);
      // the discretised radius must be exactly n*drmin for the tables to apply
      ok = (n >= 1 && n <= kTiledMaxN && T.disr_radius == drmin * fmax(1.0, (double)n));
      ok = ok && (T.la_max + T.lb_max <= kTiledMaxLp);
      for (int d = 0; d < 3; d++)
        ok = ok && (-T.lb_cube[d] <= 100);
    }
    if (!ok) {
      generic_ids.push_back(it);
      continue;
    }
    TTask X;
    for (int d = 0; d < 3; d++) {
      X.roff[d] = T.roffset[d];
      X.cc[d] = T.cubecenter[d] - L.shift_local[d];
      X.nb[d] = -T.lb_cube[d];
      max_nb = std::max(max_nb, X.nb[d]);
    }
    X.n = n, X.lp0 = T.la_max + T.lb_max, X.task = it;
    tt.push_back(X);
    max_n = std::max(max_n, n);
    max_lp0 = std::max(max_lp0, X.lp0);
  }
  tl.ntasks_tiled = (int)tt.size();
  tl.max_lp0 = max_lp0;
  tl.max_n = max_n;
  if (tt.empty())
    return;
  B200_ASSERT(nblocks * kLpBuckets < (size_t)1 << 31, "too many grid blocks");

  std::vector<KTabHeader> heads;
  std::vector<signed char> ktab;
  build_ktabs(L, max_n, heads, ktab);
  for (const TTask &X : tt)  // the cube bounds stored with the task must agree with the table's
    B200_ASSERT(X.nb[0] == heads[X.n].nbx && X.nb[1] == heads[X.n].nby && X.nb[2] == heads[X.n].nbz,
                "cube bounds disagree with the sphere table");
  tl.ktab_bytes = (int)ktab.size();
  tl.P = 2 * (max_nb + 1);

  auto up = [&](auto **dst, const auto &vec) {
    using T = typename std::remove_reference<decltype(vec)>::type::value_type;
    B200_CHECK(cudaMalloc((void **)dst, std::max<size_t>(vec.size(), 1) * sizeof(T)));
    if (!vec.empty())
      B200_CHECK(cudaMemcpyAsync(*dst, vec.data(), vec.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  };
  up(&tl.d_ttasks, tt);
  up(&tl.d_khead, heads);
  up(&tl.d_ktab, ktab);

  // exp tables
  const size_t etab_len = (size_t)tt.size() * 3 * tl.P;
  B200_CHECK(cudaMalloc((void **)&tl.d_etab, etab_len * sizeof(double)));
  etab_kernel<<<(unsigned)((etab_len + 255) / 256), 256, 0, s>>>(tl.d_ttasks, d_tasks, (int)tt.size(), tl.P,
                                                                h[0], h[1], h[2], tl.d_etab);
  B200_CHECK(cudaGetLastError());
  count_launch();

  // pairs: count, scan, fill
  const size_t nbuckets = nblocks * kLpBuckets;
  unsigned int *d_count = nullptr, *d_start = nullptr;
  B200_CHECK(cudaMalloc((void **)&d_count, (nbuckets + 1) * sizeof(unsigned int)));
  B200_CHECK(cudaMalloc((void **)&d_start, (nbuckets + 1) * sizeof(unsigned int)));
  B200_CHECK(cudaMemsetAsync(d_count, 0, (nbuckets + 1) * sizeof(unsigned int), s));
  PairGenArgs PA;
  PA.ttasks = tl.d_ttasks, PA.nttasks = (int)tt.size();
  PA.nx = L.npts_local[0], PA.ny = L.npts_local[1], PA.nz = L.npts_local[2];
  PA.Nx = L.npts_global[0], PA.Ny = L.npts_global[1], PA.Nz = L.npts_global[2];
  PA.nbx = nbx, PA.nby = nby, PA.nbz = nbz;
  PA.hx = h[0], PA.hy = h[1], PA.hz = h[2], PA.drmin = drmin;
  PA.bucket_count = d_count, PA.bucket_start = d_start, PA.bucket_cursor = nullptr, PA.pairs = nullptr;
  const int pg_blocks = ((int)tt.size() + 127) / 128;
  pairgen_kernel<0><<<pg_blocks, 128, 0, s>>>(PA);
  B200_CHECK(cudaGetLastError());
  void *d_temp = nullptr;
  size_t temp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, d_count, d_start, (int)(nbuckets + 1), s);
  B200_CHECK(cudaMalloc(&d_temp, temp_bytes));
  cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_count, d_start, (int)(nbuckets + 1), s);
  std::vector<unsigned int> start(nbuckets + 1);
  B200_CHECK(cudaMemcpyAsync(start.data(), d_start, (nbuckets + 1) * sizeof(unsigned int),
                             cudaMemcpyDeviceToHost, s));
  B200_CHECK(cudaStreamSynchronize(s));
  const size_t npairs = start[nbuckets];
  tl.npairs = (long long)npairs;
  B200_CHECK(cudaMalloc((void **)&tl.d_pairs, std::max<size_t>(npairs, 1) * sizeof(TPair)));
  B200_CHECK(cudaMemsetAsync(d_count, 0, (nbuckets + 1) * sizeof(unsigned int), s));
  PA.bucket_cursor = d_count, PA.pairs = tl.d_pairs;
  pairgen_kernel<1><<<pg_blocks, 128, 0, s>>>(PA);
  B200_CHECK(cudaGetLastError());
  count_launch(4);

  // work items: a block's pairs (all lp buckets, contiguous) cut into chunks
  std::vector<TWork> work;
  const size_t target_items = (size_t)148 * 2 * kTiledWarps * 6;
  const int chunk = (int)std::min<size_t>(kItemPairs, std::max<size_t>(128, npairs / target_items + 1));
  for (size_t b = 0; b < nblocks; b++) {
    const unsigned f = start[b * kLpBuckets], e = start[(b + 1) * kLpBuckets];
    if (e == f)
      continue;
    const int bx = (int)(b % nbx), by = (int)((b / nbx) % nby), bz = (int)(b / ((size_t)nbx * nby));
    const int cnt = (int)(e - f), nchunks = (cnt + chunk - 1) / chunk, per = (cnt + nchunks - 1) / nchunks;
    for (int c = 0; c < nchunks; c++) {
      TWork W;
      W.x0 = bx * kBX, W.y0 = by * kBY, W.z0 = bz * kBZ;
      W.first = (int)f + c * per;
      W.last = (int)f + std::min((c + 1) * per, cnt);
      work.push_back(W);
    }
  }
  // longest first; the 8 warps of a CTA then get items of similar length
  std::stable_sort(work.begin(), work.end(),
                   [](const TWork &a, const TWork &b) { return (a.last - a.first) > (b.last - b.first); });
  tl.nwork = (int)work.size();
  up(&tl.d_work, work);
  B200_CHECK(cudaStreamSynchronize(s));
  cudaFree(d_temp), cudaFree(d_count), cudaFree(d_start);
}

inline bool tiled_supports(const TiledLevel &tl, const int max_lp) {
  return tl.ntasks_tiled > 0 && max_lp <= kTiledMaxLp;
}

// ---------------------------------------------------------------------------
// Device: the hot kernels
// ---------------------------------------------------------------------------
struct TiledArgs {
  const TTask *ttasks;
  const TPair *pairs;
  const TWork *work;
  int nwork;
  const KTabHeader *khead;
  const signed char *ktab;
  int ktab_bytes, max_n;
  const double *etab;
  int P;
  const int *tcoef;          // coefficient offset per tiled task (for this call's dl)
  double *coef;
  double *grid;
  int nx, ny, nz;            // npts_local
  double hx, hy, hz;
  int dl;                    // lp growth of this call
  int lps;                   // table pitch (doubles) = max lp of the launch + 1
  int ncmax;                 // ncoset(lps-1)
};

template <int LP> struct NCo {
  static constexpr int value = (LP + 1) * (LP + 2) * (LP + 3) / 6;
};

// D[lz] = sum_{lx+ly <= LP-lz} C[coset(lx,ly,lz)] X[lx] Y[ly]
template <int LP>
__device__ __forceinline__ void column_coefs(const double *__restrict__ C, const double (&X)[LP + 1],
                                             const double (&Y)[LP + 1], double (&D)[LP + 1]) {
#pragma unroll
  for (int lz = 0; lz <= LP; lz++)
    D[lz] = 0.0;
#pragma unroll
  for (int ly = 0; ly <= LP; ly++) {
#pragma unroll
    for (int lx = 0; lx <= LP - ly; lx++) {
      const double xy = X[lx] * Y[ly];
#pragma unroll
      for (int lz = 0; lz <= LP - lx - ly; lz++)
        D[lz] = fma(C[coset(lx, ly, lz)], xy, D[lz]);
    }
  }
}

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
      int idx = WarpVecReduce<M>::run(part, lane);
      if ((lane & (DupLanes<M>::value - 1)) == 0 && idx < M && part[0] != 0.0) {
        int ly = 0;
        while (idx > L2 - ly) {
          idx -= L2 - ly + 1;
          ly++;
        }
        atomicAdd(&gcoef[coset(idx, ly, LZ)], part[0]);
      }
      slice<LZ + 1>(X, Y0, Y1, S0, S1, gcoef, lane);
    }
  }
  static __device__ __forceinline__ void run(const double (&X)[LP + 1], const double (&Y0)[LP + 1],
                                             const double (&Y1)[LP + 1], const double (&S0)[LP + 1],
                                             const double (&S1)[LP + 1], double *__restrict__ gcoef,
                                             const int lane) {
    slice<0>(X, Y0, Y1, S0, S1, gcoef, lane);
  }
};

// One (pair, warp) step.  `wtab` is this warp's table scratch
// [32 entries][lps]: entries 0..7 x, 8..15 y, 16..31 z.
template <bool COLLOCATE, int LP>
__device__ __forceinline__ void process_pair(const int lps, const double *__restrict__ wtab,
                                             const double *__restrict__ wC, double *__restrict__ gcoef,
                                             const int lo0, const int len0, const int lo1, const int len1,
                                             const bool on0, const bool on1, const int wlo, const int whi,
                                             const int li, const int lj, const int lane, double (&acc0)[kBZ],
                                             double (&acc1)[kBZ]) {
  double X[LP + 1], Y0[LP + 1], Y1[LP + 1];
#pragma unroll
  for (int l = 0; l <= LP; l++) {
    X[l] = wtab[li * lps + l];
    Y0[l] = wtab[(8 + lj) * lps + l];
    Y1[l] = wtab[(12 + lj) * lps + l];
  }
  const double *__restrict__ tZ = wtab + 16 * lps;

  if (COLLOCATE) {
    double D0[LP + 1], D1[LP + 1];
    column_coefs<LP>(wC, X, Y0, D0);
    column_coefs<LP>(wC, X, Y1, D1);
#define B200_PLANE(p)                                                          \
  {                                                                            \
    double z[LP + 1];                                                          \
    _Pragma("unroll") for (int l = 0; l <= LP; l++) z[l] = tZ[(p)*lps + l];    \
    if ((unsigned)((p)-lo0) <= (unsigned)len0) {                               \
      double v = acc0[p];                                                      \
      _Pragma("unroll") for (int l = 0; l <= LP; l++) v = fma(D0[l], z[l], v); \
      acc0[p] = v;                                                             \
    }                                                                          \
    if ((unsigned)((p)-lo1) <= (unsigned)len1) {                               \
      double v = acc1[p];                                                      \
      _Pragma("unroll") for (int l = 0; l <= LP; l++) v = fma(D1[l], z[l], v); \
      acc1[p] = v;                                                             \
    }                                                                          \
  }
#pragma unroll
    for (int p = 0; p < kBZ; p++) {
      if (p >= wlo && p <= whi)  // warp-uniform
        B200_PLANE(p)
    }
#undef B200_PLANE
  } else {
    double S0[LP + 1], S1[LP + 1];
#pragma unroll
    for (int l = 0; l <= LP; l++)
      S0[l] = 0.0, S1[l] = 0.0;
#pragma unroll
    for (int p = 0; p < kBZ; p++) {
      if (p >= wlo && p <= whi) {
        double z[LP + 1];
#pragma unroll
        for (int l = 0; l <= LP; l++)
          z[l] = tZ[p * lps + l];
        if ((unsigned)(p - lo0) <= (unsigned)len0) {
#pragma unroll
          for (int l = 0; l <= LP; l++)
            S0[l] = fma(acc0[p], z[l], S0[l]);
        }
        if ((unsigned)(p - lo1) <= (unsigned)len1) {
#pragma unroll
          for (int l = 0; l <= LP; l++)
            S1[l] = fma(acc1[p], z[l], S1[l]);
        }
      }
    }
    // table entries outside the cube are zero, inactive columns have S = 0:
    // nothing spurious enters the warp-wide sums
#pragma unroll
    for (int l = 0; l <= LP; l++) {
      if (!on0)
        Y0[l] = 0.0;
      if (!on1)
        Y1[l] = 0.0;
    }
    IntegrateReduce<LP>::run(X, Y0, Y1, S0, S1, gcoef, lane);
  }
}

template <bool COLLOCATE>
__global__ void __launch_bounds__(kTiledThreads, 2) tiled_kernel(const TiledArgs A) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lps = A.lps, ncmax = A.ncmax;
  // shared memory: per-warp table scratch + coefficients, then the K tables
  double *wtab = smem + (size_t)warp * (32 * lps + ncmax);
  double *wC = wtab + 32 * lps;
  KTabHeader *s_khead = (KTabHeader *)(smem + (size_t)kTiledWarps * (32 * lps + ncmax));
  signed char *s_ktab = (signed char *)(s_khead + (A.max_n + 1));
  for (int q = tid; q <= A.max_n; q += kTiledThreads)
    s_khead[q] = A.khead[q];
  for (int q = tid; q < A.ktab_bytes; q += kTiledThreads)
    s_ktab[q] = A.ktab[q];
  __syncthreads();  // the only CTA-wide barrier

  const int iw = blockIdx.x * kTiledWarps + warp;
  if (iw >= A.nwork)
    return;
  const TWork W = A.work[iw];
  const int vx = min(kBX, A.nx - W.x0), vy = min(kBY, A.ny - W.y0), vz = min(kBZ, A.nz - W.z0);
  const int li = lane & 7, lj = lane >> 3;  // my columns: (li, lj) and (li, lj + 4)
  const size_t sy = A.nx, sz = (size_t)A.nx * A.ny;
  double *g0 = A.grid + (size_t)W.z0 * sz + (size_t)(W.y0 + lj) * sy + W.x0 + li;
  double *g1 = g0 + 4 * sy;
  const bool col0 = (li < vx && lj < vy), col1 = (li < vx && lj + 4 < vy);

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

  // which table entry this lane fetches: axis and block-local index
  const int my_axis = (lane < 8) ? 0 : ((lane < 16) ? 1 : 2);
  const int my_t = (lane < 8) ? lane : ((lane < 16) ? lane - 8 : lane - 16);
  const double my_h = (my_axis == 0) ? A.hx : ((my_axis == 1) ? A.hy : A.hz);
  const int Phalf = A.P / 2 - 1;

  // ---- software pipeline: fetch(pair) -> process(pair) ----------------------
  TPair Pn = A.pairs[W.first];
  double e_n, roff_n, c_n = 0.0;
  int n_n, coff_n;
  {
    const TTask &X = A.ttasks[Pn.ttask];
    n_n = X.n;
    roff_n = X.roff[my_axis];
    coff_n = A.tcoef[Pn.ttask];
    const int g = my_t - Pn.o[my_axis];
    const int ge = min(max(g + Phalf, 0), A.P - 1);
    e_n = A.etab[((size_t)Pn.ttask * 3 + my_axis) * A.P + ge];
    if (g + Phalf != ge)
      e_n = 0.0;
    if (COLLOCATE && lane < ncoset(Pn.lp0 + A.dl))
      c_n = A.coef[coff_n + lane];
  }

  for (int ip = W.first; ip < W.last; ip++) {
    const TPair P = Pn;
    const double e = e_n, roff = roff_n;
    double c0 = c_n;
    const int n = n_n, coff = coff_n;
    const int lp = P.lp0 + A.dl;
    if (ip + 1 < W.last) {  // prefetch the next pair
      Pn = A.pairs[ip + 1];
      const TTask &X = A.ttasks[Pn.ttask];
      n_n = X.n;
      roff_n = X.roff[my_axis];
      coff_n = A.tcoef[Pn.ttask];
      const int g = my_t - Pn.o[my_axis];
      const int ge = min(max(g + Phalf, 0), A.P - 1);
      e_n = A.etab[((size_t)Pn.ttask * 3 + my_axis) * A.P + ge];
      if (g + Phalf != ge)
        e_n = 0.0;
      if (COLLOCATE && lane < ncoset(Pn.lp0 + A.dl))
        c_n = A.coef[coff_n + lane];
    }

    // admissible planes of my two columns
    const KTabHeader H = s_khead[n];
    const int mi = pair_dist(li - P.o[0]);
    const int px = H.nbx + 1;
    int lo[2], len[2], hi[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const int j = lj + 4 * c;
      const int mj = pair_dist(j - P.o[1]);
      int K = -1;
      if (mi <= H.nbx && mj <= H.nby && (c == 0 ? col0 : col1))
        K = s_ktab[H.offset + mj * px + mi];
      const int plo = max(P.o[2] - K, 0), phi = min(P.o[2] + K + 1, vz - 1);
      const bool on = (K >= 0 && plo <= phi);
      // inactive column: (unsigned)(p - lo) <= (unsigned)len is never true
      lo[c] = on ? plo : (1 << 20);
      len[c] = on ? phi - plo : 0;
      hi[c] = on ? phi : -1;
    }
    const int wlo = __reduce_min_sync(0xffffffffu, min(lo[0], lo[1]));
    const int whi = __reduce_max_sync(0xffffffffu, max(hi[0], hi[1]));
    if (wlo > whi)
      continue;  // warp-uniform: the sphere misses this block after all

    // table scratch: my entry times the powers of (x - xp)
    __syncwarp();
    {
      const double x = (my_t - P.o[my_axis]) * my_h - roff;
      double v = e;
      double *row = wtab + lane * lps;
      for (int l = 0; l <= lp; l++, v *= x)
        row[l] = v;
      if (COLLOCATE && lane < ncmax)
        wC[lane] = c0;
      if (COLLOCATE)
        for (int c = lane + 32; c < ncoset(lp); c += 32)
          wC[c] = A.coef[coff + c];
    }
    __syncwarp();

    double *gcoef = A.coef + coff;
    switch (lp) {
#define B200_CASE(LPV)                                                                             \
  case LPV:                                                                                        \
    process_pair<COLLOCATE, LPV>(lps, wtab, wC, gcoef, lo[0], len[0], lo[1], len[1], hi[0] >= 0,   \
                                 hi[1] >= 0, wlo, whi, li, lj, lane, acc0, acc1);                  \
    break;
      B200_CASE(0)
      B200_CASE(1)
      B200_CASE(2)
      B200_CASE(3)
      B200_CASE(4)
      B200_CASE(5)
      B200_CASE(6)
#undef B200_CASE
    default:
      break;
    }
  }

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
}

inline size_t tiled_smem_bytes(const int lps, const int max_n, const int ktab_bytes) {
  const size_t nd = (size_t)kTiledWarps * (32 * lps + ncoset(lps - 1));
  return nd * sizeof(double) + (max_n + 1) * sizeof(KTabHeader) + ktab_bytes + 16;
}

inline void launch_tiled(TiledLevel &tl, const GridLaunch &L, const bool collocate) {
  if (tl.nwork == 0)
    return;
  B200_ASSERT(L.dl >= 0 && L.dl < 8, "unexpected l growth");
  if (tl.d_tcoef[L.dl] == nullptr) {  // coefficient offsets per tiled task for this dl
    B200_CHECK(cudaMalloc((void **)&tl.d_tcoef[L.dl], tl.ntasks_tiled * sizeof(int)));
    tcoef_kernel<<<(tl.ntasks_tiled + 255) / 256, 256, 0, L.stream>>>(tl.d_ttasks, tl.ntasks_tiled,
                                                                      L.coef_offsets, tl.d_tcoef[L.dl]);
    B200_CHECK(cudaGetLastError());
    count_launch();
  }
  TiledArgs A;
  A.ttasks = tl.d_ttasks, A.pairs = tl.d_pairs, A.work = tl.d_work, A.nwork = tl.nwork;
  A.khead = tl.d_khead, A.ktab = tl.d_ktab, A.ktab_bytes = tl.ktab_bytes, A.max_n = tl.max_n;
  A.etab = tl.d_etab, A.P = tl.P, A.tcoef = tl.d_tcoef[L.dl];
  A.coef = L.coef, A.grid = L.grid;
  A.nx = L.level.npts_local[0], A.ny = L.level.npts_local[1], A.nz = L.level.npts_local[2];
  A.hx = L.level.dh[0], A.hy = L.level.dh[4], A.hz = L.level.dh[8];
  A.dl = L.dl;
  A.lps = tl.max_lp0 + L.dl + 1;
  A.ncmax = ncoset(A.lps - 1);
  const size_t bytes = tiled_smem_bytes(A.lps, tl.max_n, tl.ktab_bytes);
  B200_ASSERT(bytes <= 100 * 1024, "tiled kernel: shared memory budget exceeded");
  const int grid = (tl.nwork + kTiledWarps - 1) / kTiledWarps;
  if (collocate) {
    B200_CHECK(cudaFuncSetAttribute(tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    tiled_kernel<true><<<grid, kTiledThreads, bytes, L.stream>>>(A);
  } else {
    B200_CHECK(cudaFuncSetAttribute(tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    tiled_kernel<false><<<grid, kTiledThreads, bytes, L.stream>>>(A);
  }
  B200_CHECK(cudaGetLastError());
  count_launch();
}
inline void launch_tiled_collocate(TiledLevel &tl, const GridLaunch &L) { launch_tiled(tl, L, true); }
inline void launch_tiled_integrate(TiledLevel &tl, const GridLaunch &L) { launch_tiled(tl, L, false); }

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
