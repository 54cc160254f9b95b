// CTA-tile collocate / integrate kernels for orthorhombic tasks -- the hot path
// (round 2).  See DESIGN.md section "CTA-tile kernels".
//
// The round-1 kernels (b200_tiled.cuh) gave every WARP an 8x8x16 grid block and
// let it decode its own (task, block) pairs: 195 instructions per pair for 47
// DFMAs, each warp re-deriving table addresses, sphere masks and 1-D tables that
// its neighbours derive too.  Here a CTA of 8 warps owns a 16x16x32 TILE of the
// grid (2x2x2 warp blocks, every thread still keeps its 2x16 points in
// registers) and walks the list of VISITS (task, tile) of its tile together:
//
//   * per visit ONE warp (the visits' producers rotate over the 8 warps, so the
//     work is spread evenly over the four SM sub-partitions) prepares, in a ring
//     of shared-memory slots, everything the other warps need: the 1-D tables
//     exp(-zetp d^2) d^l of the tile's 16 + 16 + 32 coordinates (computed in the
//     kernel: no exp tables in HBM any more), the plane masks of the tile's 256
//     columns from the sphere-extent table, per warp block the activity flag and
//     plane range, and -- collocate -- the task's coefficients, fetched by a
//     bulk-async copy (cp.async.bulk, completion counted on the slot's mbarrier);
//   * slots are handed over with mbarriers (full: producer -> consumers, empty:
//     consumers -> next producer); producers run half a ring ahead, so consumers
//     only ever touch shared memory: no global-load latency in the FMA loops;
//   * a consumer whose warp block the sphere misses skips the visit after one
//     shared-memory load; otherwise it builds the per-column factors and runs the
//     predicated plane loop (b200_pfma.cuh) over its plane range;
//   * integrate: every warp reduces its contribution by warp shuffles and leaves
//     it in the slot; the producer that recycles the slot sums the (<= 8) partial
//     results in a fixed order and folds them into the task's coefficients with
//     ONE atomic per coefficient and visit (round 1: one per coefficient and
//     (task, warp block) pair).
//
// Reference semantics (which points a task touches, what it adds) are unchanged:
// ref/grid_ref_collint.h:206-327 (bounds, tables), :30-200 (loops).
#pragma once
#include "b200_tiled.cuh"

namespace b200 {

constexpr int kCtX = 16, kCtY = 16, kCtZ = 32;  // CTA tile
constexpr int kCtKPad = 16;    // sphere-table margin: tile offsets need no bounds check
constexpr int kCtKPitch = 96;  // sphere-table row pitch (bytes)
constexpr int kCtZmBias = 24, kCtZmPitch = 80, kCtZmRows = kTiledMaxNb + 2;
constexpr int kCtMaxLevels = 8;
constexpr int kCtItemVisits = 1536;  // visits per work item (upper bound)

struct alignas(16) CTask {  // what a producer needs of a task
  double roff[3];
  double zl2;  // zetp * log2(e)
};

struct alignas(16) CWork {  // 32 bytes
  int x0, y0, z0;           // tile origin (local grid indices)
  int level;
  int b[4];                 // visit ranges per lp of the class: [b[i], b[i+1])
};

struct alignas(64) CVisit {  // one (task, tile) visit, see visitgen_kernel
  uint4 a, b;
  CTask t;  // the task's data inlined: the producers stream the visit list, they never gather
};

struct CtileLevel {
  long long nvisits = 0;
  int ntasks_tiled = 0;
  int max_lp0 = 0, max_n = 0, max_nb = 0;
  int class_ntasks[kNumClasses] = {0, 0, 0};
  int class_tt_first[kNumClasses + 1] = {0, 0, 0, 0};
  std::vector<int> h_tt_task;  // TaskDev id per tiled task (class-sorted)
  int coef_base[8][kNumClasses] = {};
  int *d_class_task_ids[kNumClasses] = {nullptr, nullptr, nullptr};
  CTask *d_ctasks = nullptr;
  CVisit *d_visits = nullptr;
  KTabHeader *d_khead = nullptr;
  unsigned char *d_ktab = nullptr;
  std::vector<CWork> work[kNumClasses];  // host: merged over the levels by the list
  void release() {
    dev_free(d_ctasks), dev_free(d_visits), dev_free(d_khead), dev_free(d_ktab);
    d_ctasks = nullptr, d_visits = nullptr, d_khead = nullptr, d_ktab = nullptr;
    for (auto &p : d_class_task_ids) {
      dev_free(p);
      p = nullptr;
    }
    for (auto &w : work)
      w.clear();
    nvisits = 0, ntasks_tiled = 0;
  }
};

// zmask[K1][oz + kCtZmBias]: bit p set <=> plane p of the tile lies within
// pair-distance K = K1 - 1 of the cube centre plane oz:  oz - K <= p <= oz + K + 1
inline std::vector<unsigned> build_ct_zmask() {
  std::vector<unsigned> zm((size_t)kCtZmRows * kCtZmPitch, 0u);
  for (int k1 = 1; k1 < kCtZmRows; k1++)
    for (int ozb = 0; ozb < kCtZmPitch; ozb++) {
      const int K = k1 - 1, oz = ozb - kCtZmBias;
      unsigned m = 0;
      for (int p = 0; p < kCtZ; p++)
        if (oz - K <= p && p <= oz + K + 1)
          m |= 1u << p;
      zm[(size_t)k1 * kCtZmPitch + ozb] = m;
    }
  return zm;
}

// ---------------------------------------------------------------------------
// Device: visit generation.  One thread per tiled task; pass 0 counts the visits
// per (tile, lp) bucket, pass 1 writes them.  Periodic images are separate visits.
// A visit record is 32 bytes:
//   a.x  task index within the level (24 bits) | active warp blocks (8 bits)
//   a.y  sphere-table index of tile column (0, 0)
//   a.z  cube centre relative to the tile origin: x | y << 8 | z << 16 (signed bytes)
//   b.x, b.y  per warp block one byte: first | last << 4 plane any of its columns needs
//   t    the task's sub-grid offset and exponent (a copy: the record is self-contained)
// Which warp blocks a visit touches and their plane ranges are decided HERE, once per
// task list, with the same exact test the round-1 pair generation used (the sphere
// meets a block iff its column nearest to the centre reaches its nearest plane).
// ---------------------------------------------------------------------------
struct VisitGenArgs {
  const TTask *ttasks;
  const CTask *ctasks;
  int nttasks;
  int nx, ny, nz, Nx, Ny, Nz;  // local / global grid size
  int ntx, nty, ntz;           // tiles per axis
  unsigned ntiles;
  const KTabHeader *khead;
  const unsigned char *ktab;
  unsigned int *bucket_count;        // pass 0
  const unsigned int *bucket_start;  // pass 1
  unsigned int *bucket_cursor;       // pass 1
  CVisit *visits;                    // pass 1
  unsigned long long *keys;          // pass 1: bucket << qbits | q
  int qbits;
};

template <int PASS> __global__ void visitgen_kernel(const VisitGenArgs A) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.nttasks)
    return;
  const TTask X = A.ttasks[q];
  const KTabHeader H = A.khead[X.n];
  const int nloc[3] = {A.nx, A.ny, A.nz}, N[3] = {A.Nx, A.Ny, A.Nz};
  const int B[3] = {kCtX, kCtY, kCtZ}, nt[3] = {A.ntx, A.nty, A.ntz};
  int m_lo[3], m_hi[3];
  for (int d = 0; d < 3; d++) {
    const int lo = X.cc[d] - X.nb[d], hi = X.cc[d] + 1 + X.nb[d];
    m_lo[d] = floor_div(-hi, N[d]);
    m_hi[d] = floor_div(nloc[d] - 1 - lo, N[d]) + 1;
  }
  const unsigned char *ktab_n = A.ktab + H.offset;
  for (int mz = m_lo[2]; mz <= m_hi[2]; mz++) {
    const int cz = X.cc[2] + mz * N[2];
    const int az = max(cz - X.nb[2], 0), bz = min(cz + 1 + X.nb[2], nloc[2] - 1);
    if (az > bz)
      continue;
    for (int tz = az / B[2]; tz <= bz / B[2] && tz < nt[2]; tz++) {
      for (int my = m_lo[1]; my <= m_hi[1]; my++) {
        const int cy = X.cc[1] + my * N[1];
        const int ay = max(cy - X.nb[1], 0), by = min(cy + 1 + X.nb[1], nloc[1] - 1);
        if (ay > by)
          continue;
        for (int ty = ay / B[1]; ty <= by / B[1] && ty < nt[1]; ty++) {
          for (int mx = m_lo[0]; mx <= m_hi[0]; mx++) {
            const int cx = X.cc[0] + mx * N[0];
            const int ax = max(cx - X.nb[0], 0), bx = min(cx + 1 + X.nb[0], nloc[0] - 1);
            if (ax > bx)
              continue;
            for (int tx = ax / B[0]; tx <= bx / B[0] && tx < nt[0]; tx++) {
              // the 8 warp blocks (8 x 8 x 16) of this tile
              unsigned act = 0u, wr[2] = {0u, 0u};
              for (int w = 0; w < 8; w++) {
                const int x0 = tx * B[0] + 8 * (w & 1), y0 = ty * B[1] + 8 * ((w >> 1) & 1), z0 = tz * B[2] + 16 * (w >> 2);
                const int xl = max(ax, x0), xh = min(bx, x0 + 7), yl = max(ay, y0), yh = min(by, y0 + 7),
                          zl = max(az, z0), zh = min(bz, z0 + 15);
                if (xl > xh || yl > yh || zl > zh)
                  continue;
                const int mi = rel_dmin(xl - cx, xh - cx), mj = rel_dmin(yl - cy, yh - cy);
                // K of the block's column nearest to the centre bounds every other column's
                const int K = (int)ktab_n[(X.nb[1] + kCtKPad - mj) * kCtKPitch + (X.nb[0] + kCtKPad - mi)] - 1;
                const int ozb = cz - z0;
                const int wlo = max(zl - z0, ozb - K), whi = min(zh - z0, ozb + K + 1);
                if (K < 0 || wlo > whi)
                  continue;  // the sphere misses this warp block
                act |= 1u << w;
                wr[w >> 2] |= ((unsigned)wlo | ((unsigned)whi << 4)) << (8 * (w & 3));
              }
              if (act == 0u)
                continue;
              const unsigned tile = (unsigned)((tz * nt[1] + ty) * nt[0] + tx);
              const int ox = cx - tx * B[0], oy = cy - ty * B[1], oz = cz - tz * B[2];
              const unsigned bucket = ((unsigned)lp_class(X.lp0) * A.ntiles + tile) * kLpBuckets + X.lp0;
              if (PASS == 0) {
                atomicAdd(&A.bucket_count[bucket], 1u);
              } else {
                const unsigned pos = A.bucket_start[bucket] + atomicAdd(&A.bucket_cursor[bucket], 1u);
                CVisit V;
                V.a.x = (unsigned)q | (act << 24);
                V.a.y = (unsigned)(H.offset + (X.nb[1] + kCtKPad - oy) * kCtKPitch + (X.nb[0] + kCtKPad - ox));
                V.a.z = ((unsigned)ox & 0xffu) | (((unsigned)oy & 0xffu) << 8) | (((unsigned)oz & 0xffu) << 16);
                V.a.w = 0u;
                V.b = make_uint4(wr[0], wr[1], 0u, 0u);
                V.t = A.ctasks[q];
                A.visits[pos] = V;
                // Order inside a bucket: NOT by task.  Consecutive tasks (the pgf pairs of one atom
                // pair share their centre) hit the same few warp blocks, and the slot ring only
                // decouples the eight consumer warps by its depth: in task order a CTA runs on two
                // or three warps at a time (measured: 17.7 ms per H2O-256 collocate against 12.7 ms
                // with this order; dealing the visits round-robin by the warp block nearest to the
                // centre instead was worse, 14.8 ms).  A bijective hash of the task index spreads
                // every warp block's visits evenly over the list.  The order is still a pure
                // function of the task list: results stay reproducible.
                const unsigned long long hq = ((unsigned long long)q * 0x9E3779B1ull) & ((1ull << A.qbits) - 1ull);
                A.keys[pos] = ((unsigned long long)bucket << A.qbits) | hq;
              }
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Host: per-level build (task selection is the tiled path's: same criteria).
// ---------------------------------------------------------------------------
inline void build_ctile_level(CtileLevel &cl, const int level, const LevelDev &L, const TaskVec &tasks,
                              const int first, const int last, std::vector<int> &generic_ids, const int item_cap,
                              cudaStream_t s) {
  cl.release();
  const double h[3] = {L.dh[0], L.dh[4], L.dh[8]};
  const double drmin = fmin(h[0], fmin(h[1], h[2]));
  const int ntx = (L.npts_local[0] + kCtX - 1) / kCtX, nty = (L.npts_local[1] + kCtY - 1) / kCtY,
            ntz = (L.npts_local[2] + kCtZ - 1) / kCtZ;
  const size_t ntiles = (size_t)ntx * nty * ntz;

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
        ok = ok && (-T.lb_cube[d] <= kTiledMaxNb);
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
    X.zl2 = T.zetp * 1.4426950408889634074;
    X.pad[0] = X.pad[1] = X.pad[2] = 0;
    tt.push_back(X);
    max_n = std::max(max_n, n);
    max_lp0 = std::max(max_lp0, X.lp0);
  }
  cl.ntasks_tiled = (int)tt.size();
  cl.max_lp0 = max_lp0, cl.max_n = max_n, cl.max_nb = max_nb;
  if (tt.empty())
    return;
  {  // stable partition by lp class (three classes: one counting pass instead of a sort)
    std::vector<TTask> sorted(tt.size());
    size_t pos[kNumClasses + 1] = {0, 0, 0, 0};
    for (const TTask &X : tt)
      pos[lp_class(X.lp0) + 1]++;
    for (int c = 0; c < kNumClasses; c++)
      pos[c + 1] += pos[c];
    for (const TTask &X : tt)
      sorted[pos[lp_class(X.lp0)]++] = X;
    tt.swap(sorted);
  }
  cl.h_tt_task.resize(tt.size());
  for (int c = 0; c <= kNumClasses; c++)
    cl.class_tt_first[c] = 0;
  std::vector<CTask> ct(tt.size());
  for (size_t q = 0; q < tt.size(); q++) {
    cl.h_tt_task[q] = tt[q].task;
    cl.class_tt_first[lp_class(tt[q].lp0) + 1]++;
    for (int d = 0; d < 3; d++)
      ct[q].roff[d] = tt[q].roff[d];
    ct[q].zl2 = tt[q].zl2;
  }
  for (int c = 0; c < kNumClasses; c++)
    cl.class_tt_first[c + 1] += cl.class_tt_first[c];
  B200_ASSERT(ntiles * kLpBuckets * kNumClasses < (size_t)1 << 31, "too many grid tiles");

  std::vector<KTabHeader> heads;
  std::vector<unsigned char> ktab;
  build_ktabs(L, max_n, heads, ktab, kCtKPitch, kCtKPad);
  for (const TTask &X : tt)
    B200_ASSERT(heads[X.n].offset >= 0 && X.nb[0] == heads[X.n].nbx && X.nb[1] == heads[X.n].nby &&
                    X.nb[2] == heads[X.n].nbz,
                "cube bounds disagree with the sphere table");

  auto up = [&](auto **dst, const auto &vec) {
    using T = typename std::remove_reference<decltype(vec)>::type::value_type;
    dev_alloc(dst, std::max<size_t>(vec.size(), 1) * sizeof(T));
    if (!vec.empty())
      B200_CHECK(cudaMemcpyAsync(*dst, vec.data(), vec.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  };
  TTask *d_ttasks = nullptr;
  up(&d_ttasks, tt);
  up(&cl.d_ctasks, ct);
  up(&cl.d_khead, heads);
  up(&cl.d_ktab, ktab);

  // visits: count, scan, fill, sort by (bucket, hashed task index)
  const size_t nbuckets = ntiles * kLpBuckets * kNumClasses;
  unsigned int *d_count = nullptr, *d_start = nullptr;
  dev_alloc(&d_count, (nbuckets + 1) * sizeof(unsigned int));
  dev_alloc(&d_start, (nbuckets + 1) * sizeof(unsigned int));
  B200_CHECK(cudaMemsetAsync(d_count, 0, (nbuckets + 1) * sizeof(unsigned int), s));
  VisitGenArgs VA;
  VA.ttasks = d_ttasks, VA.ctasks = cl.d_ctasks, VA.nttasks = (int)tt.size();
  VA.nx = L.npts_local[0], VA.ny = L.npts_local[1], VA.nz = L.npts_local[2];
  VA.Nx = L.npts_global[0], VA.Ny = L.npts_global[1], VA.Nz = L.npts_global[2];
  VA.ntx = ntx, VA.nty = nty, VA.ntz = ntz, VA.ntiles = (unsigned)ntiles;
  VA.khead = cl.d_khead, VA.ktab = cl.d_ktab;
  VA.bucket_count = d_count, VA.bucket_start = d_start, VA.bucket_cursor = nullptr, VA.visits = nullptr;
  VA.keys = nullptr, VA.qbits = 0;
  const int vg_blocks = ((int)tt.size() + 127) / 128;
  visitgen_kernel<0><<<vg_blocks, 128, 0, s>>>(VA);
  B200_CHECK(cudaGetLastError());
  void *d_temp = nullptr;
  size_t temp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, d_count, d_start, (int)(nbuckets + 1), s);
  dev_alloc(&d_temp, temp_bytes);
  cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_count, d_start, (int)(nbuckets + 1), s);
  std::vector<unsigned int> start(nbuckets + 1);
  B200_CHECK(cudaMemcpyAsync(start.data(), d_start, (nbuckets + 1) * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
  B200_CHECK(cudaStreamSynchronize(s));
  const size_t nvis = start[nbuckets];
  B200_ASSERT(nvis < ((size_t)1 << 31), "too many (task, tile) visits on one level");
  cl.nvisits = (long long)nvis;
  B200_ASSERT(tt.size() < ((size_t)1 << 24), "too many tiled tasks on one level for the 24-bit visit field");
  dev_alloc(&cl.d_visits, (nvis + 40) * sizeof(CVisit));
  B200_CHECK(cudaMemsetAsync(cl.d_visits, 0, (nvis + 40) * sizeof(CVisit), s));  // (the kernels peek past the end)
  B200_CHECK(cudaMemsetAsync(d_count, 0, (nbuckets + 1) * sizeof(unsigned int), s));
  unsigned long long *d_keys[2] = {nullptr, nullptr};
  CVisit *d_visits_alt = nullptr;
  dev_alloc(&d_keys[0], std::max<size_t>(nvis, 1) * sizeof(unsigned long long));
  int qbits = 1, bbits = 1;
  while (((size_t)1 << qbits) < tt.size())
    qbits++;
  while (((size_t)1 << bbits) < nbuckets)
    bbits++;
  VA.bucket_cursor = d_count, VA.visits = cl.d_visits, VA.keys = d_keys[0], VA.qbits = qbits;
  visitgen_kernel<1><<<vg_blocks, 128, 0, s>>>(VA);
  B200_CHECK(cudaGetLastError());
  count_launch(4);
  if (nvis > 1) {
    dev_alloc(&d_keys[1], nvis * sizeof(unsigned long long));
    dev_alloc(&d_visits_alt, (nvis + 40) * sizeof(CVisit));
    B200_CHECK(cudaMemsetAsync(d_visits_alt, 0, (nvis + 40) * sizeof(CVisit), s));
    cub::DoubleBuffer<unsigned long long> kb(d_keys[0], d_keys[1]);
    cub::DoubleBuffer<CVisit> vb(cl.d_visits, d_visits_alt);
    void *d_sort_temp = nullptr;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, kb, vb, (int)nvis, 0, qbits + bbits, s);
    dev_alloc(&d_sort_temp, std::max<size_t>(sort_bytes, 1));
    cub::DeviceRadixSort::SortPairs(d_sort_temp, sort_bytes, kb, vb, (int)nvis, 0, qbits + bbits, s);
    B200_CHECK(cudaGetLastError());
    B200_CHECK(cudaStreamSynchronize(s));
    count_launch(6);
    if (vb.Current() != cl.d_visits)
      std::swap(cl.d_visits, d_visits_alt);
    dev_free(d_sort_temp);
  }
  dev_free(d_keys[0]), dev_free(d_keys[1]), dev_free(d_visits_alt);

  // work items per lp class: a tile's visits of that class (contiguous, ordered by lp) cut into chunks
  for (int cls = 0; cls < kNumClasses; cls++) {
    for (size_t b = 0; b < ntiles; b++) {
      const size_t bb = ((size_t)cls * ntiles + b) * kLpBuckets;
      const unsigned f = start[bb], e = start[bb + kLpBuckets];
      if (e == f)
        continue;
      const int tx = (int)(b % ntx), ty = (int)((b / ntx) % nty), tz = (int)(b / ((size_t)ntx * nty));
      const int cnt = (int)(e - f), nchunks = (cnt + item_cap - 1) / item_cap,
                per = (cnt + nchunks - 1) / nchunks;
      for (int c = 0; c < nchunks; c++) {
        CWork W;
        W.x0 = tx * kCtX, W.y0 = ty * kCtY, W.z0 = tz * kCtZ, W.level = level;
        const int lo = (int)f + c * per, hi = (int)f + std::min((c + 1) * per, cnt);
        W.b[0] = lo, W.b[3] = hi;
        for (int i = 1; i < 3; i++) {
          const int lp_next = std::min(kClassLo[cls] + i, kLpBuckets);
          W.b[i] = std::min(std::max((int)start[bb + lp_next], lo), hi);
        }
        cl.work[cls].push_back(W);
      }
    }
  }
  for (int cls = 0; cls < kNumClasses; cls++) {
    std::vector<int> ids;
    for (const TTask &X : tt)
      if (lp_class(X.lp0) == cls)
        ids.push_back(X.task);
    cl.class_ntasks[cls] = (int)ids.size();
    up(&cl.d_class_task_ids[cls], ids);
  }
  B200_CHECK(cudaStreamSynchronize(s));
  dev_free(d_temp), dev_free(d_count), dev_free(d_start), dev_free(d_ttasks);
}

// ---------------------------------------------------------------------------
// Device: mbarrier / bulk-copy primitives (PTX; sm_90+)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned ct_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(const unsigned bar, const unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(const unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(const unsigned bar, const unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(const unsigned bar, const unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE%=;\n\t"
      "bra WAIT%=;\n\t"
      "DONE%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// global -> shared bulk-async copy; completion (bytes) is counted on `bar`
__device__ __forceinline__ void bulk_g2s(const unsigned dst, const void *src, const unsigned bytes, const unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---------------------------------------------------------------------------
// Device: the hot kernels
// ---------------------------------------------------------------------------
struct CtLevelArgs {
  const CVisit *visits;
  const CTask *ctasks;
  const unsigned char *ktab;
  double *grid;
  int nx, ny, nz;  // npts_local
  int coef_base;   // slot of tiled task q: coef_base + (q - tt_first) * coef_stride
  int tt_first;
  int pad;
  double hx, hy, hz;
};

struct CtArgs {
  CtLevelArgs lev[kCtMaxLevels];
  const CWork *work;  // work items of ONE lp class, all levels
  int nwork;
  int *counter;  // dynamic work distribution (zeroed before the launch)
  const unsigned *zmask;
  double *coef;
  int coef_stride;
};

constexpr int kCtConsumers = 8, kCtProducers = 4;  // warps: two consumer warpgroups, one producer warpgroup
constexpr int kCtThreads = 32 * (kCtConsumers + kCtProducers);

// Shared-memory slot of one visit (byte offsets).
template <bool COLLOCATE, int LPHI> struct CtSlot {
  static constexpr int PITCH = (LPHI + 2) / 2 * 2;  // doubles per table row (even: 16-byte rows)
  static constexpr int NCP = (ncoset(LPHI) + 1) / 2 * 2;
  static constexpr int HDR = 0;  // u32[16]: [0..1] plane ranges of the 8 warp blocks, [2] coefficient slot,
                                 //          [3] lp, [4] active warp blocks, [5] the visit's sequence number,
                                 //          [6] plane-mask table column of the visit (oz + bias)
  static constexpr int ROWS = 64;                   // x rows | y rows | z rows
  static constexpr int KT = ROWS + 64 * PITCH * 8;  // u8 [16][16]: K + 1 of the tile's columns (0: outside)
  static constexpr int COEF = KT + 256;             // collocate: C_xyz; integrate: partial results [8][NCP]
  static constexpr int BYTES = COEF + (COLLOCATE ? NCP : 8 * NCP) * 8;
};

#ifndef B200_CT_CTAS
#define B200_CT_CTAS 2   // CTAs per SM of the lp <= 2 class (1: 216 registers per consumer, deeper ring)
#endif
#ifndef B200_CT_NS
#define B200_CT_NS 32    // ring depth of the lp <= 2 class
#endif
constexpr int kCtChunk = 16;   // visit records per staging chunk (one bulk copy)
constexpr int kCtChunks = 4;   // staging ring depth

template <bool COLLOCATE, int LPHI> struct CtConf {
  using Slot = CtSlot<COLLOCATE, LPHI>;
  // ring depth: as many slots as fit comfortably beside a second CTA (low classes)
  static constexpr int NS = (Slot::BYTES <= 3072) ? B200_CT_NS : ((Slot::BYTES <= 6144) ? 16 : ((Slot::BYTES <= 12288) ? 8 : 4));
  static constexpr int BAR = 0;                       // u64 full[NS], empty[NS], cfull[4], cempty[4]
  static constexpr int ITEM = 16 * NS + 16 * kCtChunks;  // int: the CTA's current work item
  static constexpr int E2T = ITEM + 16;               // double[64]: 2^(j/64)
  static constexpr int ZM = E2T + 512;                // u16 [2][kCtZmRows][kCtZmPitch]: plane masks per z half
  static constexpr int STAGE = ZM + 2 * kCtZmRows * kCtZmPitch * 2;  // CVisit [kCtChunks][kCtChunk]
  static constexpr int SLOTS = STAGE + kCtChunks * kCtChunk * 64;
  static constexpr int BYTES = SLOTS + NS * Slot::BYTES;
};

extern __shared__ __align__(128) unsigned char ct_smem[];

// Loads LP+1 doubles of a 16-byte aligned table row.
template <int LP> __device__ __forceinline__ void ct_load_row(const double *__restrict__ row, double (&z)[LP + 1]) {
#pragma unroll
  for (int l = 0; l + 1 <= LP; l += 2) {
    const double2 v = *reinterpret_cast<const double2 *>(row + l);
    z[l] = v.x, z[l + 1] = v.y;
  }
  if constexpr ((LP & 1) == 0)
    z[LP] = row[LP];
}

// Transposing reductions of the integrate epilogue with the result left in shared memory.
template <int LP> struct CtIntegrateReduce {
  template <int LZ>
  static __device__ __forceinline__ void slice(const double (&X)[LP + 1], const double (&Y0)[LP + 1],
                                               const double (&Y1)[LP + 1], const double (&S0)[LP + 1],
                                               const double (&S1)[LP + 1], double *__restrict__ part,
                                               const int lane) {
    if constexpr (LZ <= LP) {
      constexpr int L2 = LP - LZ;
      constexpr int M = (L2 + 1) * (L2 + 2) / 2;
      double v[M];
      int q = 0;
#pragma unroll
      for (int ly = 0; ly <= L2; ly++) {
        const double w = fma(Y0[ly], S0[LZ], Y1[ly] * S1[LZ]);
#pragma unroll
        for (int lx = 0; lx <= L2 - ly; lx++)
          v[q++] = X[lx] * w;
      }
      auto emit = [&](int idx, const double val) {
        int ly = 0;
        while (idx > L2 - ly) {
          idx -= L2 - ly + 1;
          ly++;
        }
        part[coset(idx, ly, LZ)] = val;
      };
      constexpr int M1 = (M > 32) ? 32 : M, M2 = M - M1;
      {
        double head[M1];
#pragma unroll
        for (int i = 0; i < M1; i++)
          head[i] = v[i];
        const int idx = WarpVecReduce<M1>::run(head, lane);
        if ((lane & (DupLanes<M1>::value - 1)) == 0 && ((ValidLanes<M1>::value >> lane) & 1u))
          emit(idx, head[0]);
      }
      if constexpr (M2 > 0) {
        double rest[M2];
#pragma unroll
        for (int i = 0; i < M2; i++)
          rest[i] = v[M1 + i];
        const int idx = WarpVecReduce<M2>::run(rest, lane);
        if ((lane & (DupLanes<M2>::value - 1)) == 0 && ((ValidLanes<M2>::value >> lane) & 1u))
          emit(M1 + idx, rest[0]);
      }
      slice<LZ + 1>(X, Y0, Y1, S0, S1, part, lane);
    }
  }
};

// Per-warp constants of a consumer.
struct CtWarp {
  int lane, warp;
  int xrow, yrow;  // table rows of my x and of my first y within the tile
  int zrow0;       // first z row of my warp block
  int kt_off;      // my first column within the slot's K tile
  const unsigned short *zm;  // plane-mask table of my z half
};

// ---- consumer: one visit, one warp (the visit is known to touch this warp block) ----
template <bool COLLOCATE, int LP, int LPHI>
__device__ __forceinline__ void ct_consume(const unsigned char *__restrict__ slot, const CtWarp &c,
                                           double (&acc0)[16], double (&acc1)[16]) {
  using S = CtSlot<COLLOCATE, LPHI>;
  constexpr int PITCH = S::PITCH;
  constexpr int NC = ncoset(LP);
  const unsigned wr = reinterpret_cast<const unsigned char *>(slot + S::HDR)[c.warp];
  const int wlo = (int)(wr & 15u), whi = (int)(wr >> 4);
  // sphere masks of my two columns: K + 1 from the staged tile of the sphere table, then the plane mask
  const unsigned char *kt = slot + S::KT + c.kt_off;
  const unsigned zcol = reinterpret_cast<const unsigned *>(slot + S::HDR)[6];
  const unsigned mask0 = c.zm[(unsigned)kt[0] * kCtZmPitch + zcol];
  const unsigned mask1 = c.zm[(unsigned)kt[4 * 16] * kCtZmPitch + zcol];
  const double *rows = reinterpret_cast<const double *>(slot + S::ROWS);
  double X[LP + 1], Y0[LP + 1], Y1[LP + 1];
  ct_load_row<LP>(rows + c.xrow * PITCH, X);
  ct_load_row<LP>(rows + c.yrow * PITCH, Y0);
  ct_load_row<LP>(rows + (c.yrow + 4) * PITCH, Y1);
  const double *tZ = rows + (32 + c.zrow0) * PITCH;

  if constexpr (COLLOCATE) {
    // E[ly][lz] = sum_lx C[lx,ly,lz] X[lx]  (shared by both columns: same x)
    // D_c[lz]   = sum_ly E[ly][lz] Y_c[ly]
    const double *C = reinterpret_cast<const double *>(slot + S::COEF);
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
    ct_load_row<LP>(tZ + (p)*PITCH, z);                                        \
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
    ct_load_row<LP>(tZ + (p)*PITCH, z);                                        \
    PFma<LP>::integ(S0, acc0[p], z, mask0 & (1u << (p)));                      \
    PFma<LP>::integ(S1, acc1[p], z, mask1 & (1u << (p)));                      \
  }
    B200_PLANES(B200_BODY)
#undef B200_BODY
    // inactive columns have S = 0: nothing spurious enters the warp-wide sums
    double *part = reinterpret_cast<double *>(const_cast<unsigned char *>(slot) + S::COEF) + c.warp * S::NCP;
    if constexpr (LP <= 3) {
      double v[NC];
#pragma unroll
      for (int ly = 0; ly <= LP; ly++) {
#pragma unroll
        for (int lz = 0; lz <= LP - ly; lz++) {
          const double w = fma(Y0[ly], S0[lz], Y1[ly] * S1[lz]);
#pragma unroll
          for (int lx = 0; lx <= LP - ly - lz; lx++)
            v[coset(lx, ly, lz)] = X[lx] * w;
        }
      }
      const int idx = WarpVecReduce<NC>::run(v, c.lane);
      if ((c.lane & (DupLanes<NC>::value - 1)) == 0 && ((ValidLanes<NC>::value >> c.lane) & 1u))
        part[idx] = v[0];
    } else {
      CtIntegrateReduce<LP>::template slice<0>(X, Y0, Y1, S0, S1, part, c.lane);
    }
  }
}

// ---- integrate: fold a finished visit's partial results into the coefficients
template <int LPHI>
__device__ __forceinline__ void ct_drain(const unsigned char *__restrict__ slot, double *__restrict__ coef,
                                         const int lane) {
  using S = CtSlot<false, LPHI>;
  const unsigned *hdr = reinterpret_cast<const unsigned *>(slot + S::HDR);
  const unsigned cs = hdr[2], lp = hdr[3], act = hdr[4];
  const int nc = ncoset((int)lp);
  const double *part = reinterpret_cast<const double *>(slot + S::COEF);
  for (int k = lane; k < nc; k += 32) {
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < kCtConsumers; w++)
      if (act & (1u << w))
        sum += part[w * S::NCP + k];
    atomicAdd(&coef[cs + k], sum);
  }
}

// ---- producer: prepare one visit in its slot (one warp) from its staged record ----
template <bool COLLOCATE, int LPHI>
__device__ __forceinline__ void ct_produce(unsigned char *__restrict__ slot, const unsigned full_bar,
                                           const unsigned empty_bar, const CtLevelArgs &L, const CtArgs &A,
                                           const double *__restrict__ s_e2t, const unsigned char *__restrict__ rec,
                                           const int lp, const unsigned seq, const int lane) {
  using S = CtSlot<COLLOCATE, LPHI>;
  constexpr int PITCH = S::PITCH;
  const uint4 ra = *reinterpret_cast<const uint4 *>(rec);
  const unsigned q = ra.x & 0xffffffu, act = ra.x >> 24;
  const int ox = (int)(signed char)(ra.z & 0xffu), oy = (int)(signed char)((ra.z >> 8) & 0xffu),
            oz = (int)(signed char)((ra.z >> 16) & 0xffu);
  const unsigned cs = (unsigned)(L.coef_base + (int)(q - (unsigned)L.tt_first) * A.coef_stride);
  if constexpr (COLLOCATE) {
    if (lane == 0)
      bulk_g2s(ct_smem_u32(slot + S::COEF), A.coef + cs, (unsigned)((ncoset(lp) + 1) / 2 * 16), full_bar);
  }
  // sphere extents of my 8 columns (row r = lane >> 1, columns 8 * (lane & 1) .. + 7): issued
  // first, needed last -- the table is small and mostly L1-resident
  const int r = lane >> 1, hf = lane & 1;
  const unsigned char *kp = L.ktab + (ra.y + (unsigned)(r * kCtKPitch + 8 * hf));
  const unsigned long long ka = (unsigned long long)kp;
  const uint2 *kap = reinterpret_cast<const uint2 *>(ka & ~7ull);
  const uint2 kw0 = __ldg(kap), kw1 = __ldg(kap + 1);
  // 1-D tables: lanes 0-15 x, 16-31 y; then 32 z entries (two independent chains)
  const double2 t0 = *reinterpret_cast<const double2 *>(rec + 32), t1 = *reinterpret_cast<const double2 *>(rec + 48);
  double *rows = reinterpret_cast<double *>(slot + S::ROWS);
  {
    const bool isy = lane >= 16;
    const double d1 = (double)((lane & 15) - (isy ? oy : ox)) * (isy ? L.hy : L.hx) - (isy ? t0.y : t0.x);
    const double d2 = (double)(lane - oz) * L.hz - t1.x;
    double e1 = exp_neg_tab(t1.y, d1, s_e2t), e2 = exp_neg_tab(t1.y, d2, s_e2t);
    double *row1 = rows + lane * PITCH, *row2 = rows + (32 + lane) * PITCH;
    row1[0] = e1, row2[0] = e2;
#pragma unroll
    for (int l = 1; l <= LPHI; l++)
      if (l <= lp) {
        e1 *= d1, e2 *= d2;
        row1[l] = e1, row2[l] = e2;
      }
  }
  {
    const unsigned sh = (unsigned)(ka & 7ull) * 8u;
    const unsigned long long lo = ((unsigned long long)kw0.y << 32) | kw0.x, hi = ((unsigned long long)kw1.y << 32) | kw1.x;
    const unsigned long long kbytes = (sh == 0u) ? lo : ((lo >> sh) | (hi << (64u - sh)));
    *reinterpret_cast<unsigned long long *>(slot + S::KT + r * 16 + 8 * hf) = kbytes;
  }
  if (lane == 0) {
    unsigned *hdr = reinterpret_cast<unsigned *>(slot + S::HDR);
    const uint2 rb = *reinterpret_cast<const uint2 *>(rec + 16);
    hdr[0] = rb.x, hdr[1] = rb.y;
    hdr[6] = (unsigned)(oz + kCtZmBias);
    if constexpr (!COLLOCATE)
      hdr[2] = cs, hdr[3] = (unsigned)lp, hdr[4] = act;
  }
  __syncwarp();
  if (lane == 0) {
    // Consumers skip the visits that miss their warp block, so one of them may ask for this
    // slot several ring turns early -- when a parity wait on `full` cannot tell the turns
    // apart.  The sequence number says which visit the slot holds (or is about to hold: the
    // full barrier's current phase is then this visit's, and the parity wait is exact).
    *reinterpret_cast<volatile unsigned *>(slot + S::HDR + 20) = seq;
    // the warp blocks this visit does not touch never look at the slot: release it on their behalf
    const int idle = kCtConsumers - __popc(act);
    if (idle > 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(empty_bar), "r"(idle) : "memory");
    if constexpr (COLLOCATE)
      mbar_arrive_expect_tx(full_bar, (unsigned)((ncoset(lp) + 1) / 2 * 16));
    else
      mbar_arrive(full_bar);
  }
}

// Ring bookkeeping of one CTA (uniform over its threads).
template <bool COLLOCATE, int LPHI> struct CtRing {
  using Conf = CtConf<COLLOCATE, LPHI>;
  using S = CtSlot<COLLOCATE, LPHI>;
  unsigned bar0;   // full[s] = bar0 + 8 s, empty[s] = bar0 + 8 (NS + s), then cfull[4], cempty[4]
  unsigned gbase;  // visits this CTA has been through before the current item (slot = g % NS, use = g / NS)
  unsigned cbase;  // staging chunks ... (buffer = k % kCtChunks, use = k / kCtChunks)
  int vbeg, vend;  // the item's visits (absolute indices into the level's visit array)
  int e0, e1;      // ends of the class's first and second lp within the item
  __device__ __forceinline__ unsigned char *slot(const unsigned s) const {
    return ct_smem + Conf::SLOTS + s * S::BYTES;
  }
  __device__ __forceinline__ unsigned full(const unsigned s) const { return bar0 + 8u * s; }
  __device__ __forceinline__ unsigned empty(const unsigned s) const { return bar0 + 8u * (Conf::NS + s); }
  __device__ __forceinline__ unsigned cfull(const unsigned b) const { return bar0 + 8u * (2 * Conf::NS + b); }
  __device__ __forceinline__ unsigned cempty(const unsigned b) const {
    return bar0 + 8u * (2 * Conf::NS + kCtChunks + b);
  }
};

// All visits [lo, hi) of ONE lp of the item, as seen by one consumer warp: it peeks at
// the activity bits of 32 visits at a time and only ever touches the slots of visits
// that reach its warp block.
template <bool COLLOCATE, int LP, int LPHI>
__device__ __forceinline__ void ct_consumer_run(const CtRing<COLLOCATE, LPHI> &R, const CtLevelArgs &L,
                                                const CtWarp &c, const int lo, const int hi, double (&acc0)[16],
                                                double (&acc1)[16]) {
  constexpr int NS = CtConf<COLLOCATE, LPHI>::NS;
  if (lo >= hi)
    return;
  const unsigned bit = 1u << (24 + c.warp);
  unsigned nxt = __ldg(&L.visits[lo + c.lane].a.x);  // (the visit array is padded: no bounds check)
  for (int base = lo; base < hi; base += 32) {
    const unsigned cur = nxt;
    nxt = __ldg(&L.visits[base + 32 + c.lane].a.x);
    unsigned m = __ballot_sync(0xffffffffu, (cur & bit) != 0u && base + c.lane < hi);
    while (m) {
      const int b = __ffs((int)m) - 1;
      m &= m - 1u;
      const unsigned g = R.gbase + (unsigned)(base + b - R.vbeg);
      const unsigned s = g % NS, use = g / NS;
      {
        const volatile unsigned *seq = reinterpret_cast<const volatile unsigned *>(R.slot(s) + CtSlot<COLLOCATE, LPHI>::HDR + 20);
        while (*seq != g + 1u)
          __nanosleep(20);
      }
      mbar_wait(R.full(s), use & 1u);
      ct_consume<COLLOCATE, LP, LPHI>(R.slot(s), c, acc0, acc1);
      __syncwarp();
      if (c.lane == 0)
        mbar_arrive(R.empty(s));
    }
  }
}

// The producer warps of a CTA: the item's visit records are streamed into a staging ring
// in chunks of kCtChunk (one bulk-async copy each, issued two chunks ahead by the first
// producer warp); inside a chunk the four warps take the visits round-robin.
template <bool COLLOCATE, int LPLO, int LPHI>
__device__ __forceinline__ void ct_producer_run(const CtRing<COLLOCATE, LPHI> &R, const CtLevelArgs &L,
                                                const CtArgs &A, const double *__restrict__ s_e2t, const int pw,
                                                const int lane) {
  using Conf = CtConf<COLLOCATE, LPHI>;
  constexpr int NS = Conf::NS;
  const int nvis = R.vend - R.vbeg;
  const int nch = (nvis + kCtChunk - 1) / kCtChunk;
  auto request = [&](const int c) {  // loader thread only
    const unsigned k = R.cbase + (unsigned)c, b = k % kCtChunks, cu = k / kCtChunks;
    mbar_wait(R.cempty(b), (cu & 1u) ^ 1u);
    const unsigned bytes = (unsigned)min(kCtChunk, nvis - c * kCtChunk) * 64u;
    mbar_arrive_expect_tx(R.cfull(b), bytes);
    bulk_g2s(ct_smem_u32(ct_smem + Conf::STAGE + b * (kCtChunk * 64)), L.visits + (R.vbeg + c * kCtChunk), bytes,
             R.cfull(b));
  };
  const bool loader = (pw == 0 && lane == 0);
  if (loader) {
    request(0);
    if (nch > 1)
      request(1);
  }
  for (int c = 0; c < nch; c++) {
    if (loader && c + 2 < nch)
      request(c + 2);
    const unsigned k = R.cbase + (unsigned)c, b = k % kCtChunks, cu = k / kCtChunks;
    mbar_wait(R.cfull(b), cu & 1u);
    const unsigned char *stage = ct_smem + Conf::STAGE + b * (kCtChunk * 64);
    const int clen = min(kCtChunk, nvis - c * kCtChunk);
    for (int j = pw; j < clen; j += kCtProducers) {
      const int ord = c * kCtChunk + j, v = R.vbeg + ord;
      const unsigned g = R.gbase + (unsigned)ord;
      const unsigned s = g % NS, use = g / NS;
      unsigned char *slot = R.slot(s);
      mbar_wait(R.empty(s), (use & 1u) ^ 1u);
      if constexpr (!COLLOCATE) {
        if (ord >= NS) {  // the slot's previous visit belongs to this item: fold its results
          ct_drain<LPHI>(slot, A.coef, lane);
          __syncwarp();
        }
      }
      const int lp = LPLO + (v >= R.e0 ? 1 : 0) + (v >= R.e1 ? 1 : 0);
      ct_produce<COLLOCATE, LPHI>(slot, R.full(s), R.empty(s), L, A, s_e2t, stage + j * 64, lp, g + 1u, lane);
    }
    __syncwarp();
    if (lane == 0)
      mbar_arrive(R.cempty(b));
  }
  if constexpr (!COLLOCATE) {
    // fold the results of the item's last visits (nobody recycles their slots within the item)
    for (int ord = max(0, nvis - NS); ord < nvis; ord++)
      if ((ord & (kCtProducers - 1)) == pw) {
        const unsigned g = R.gbase + (unsigned)ord;
        const unsigned s = g % NS, use = g / NS;
        mbar_wait(R.empty(s), use & 1u);
        ct_drain<LPHI>(R.slot(s), A.coef, lane);
      }
  }
}

// Register budgets of the two roles (setmaxnreg moves registers inside the CTA's own
// allocation: 8 RC + 4 RP <= 12 LAUNCH).
template <int LPHI> struct CtRegs {
  static constexpr int CTAS = (LPHI <= 2) ? 2 : 1;
  static constexpr int LAUNCH = (LPHI <= 2) ? 80 : 168;
  static constexpr int CONSUMER = (LPHI <= 2) ? 96 : 216;
  static constexpr int PRODUCER = (LPHI <= 2) ? 48 : 72;
};

template <bool COLLOCATE, int LPLO, int LPHI>
__global__ void __launch_bounds__(kCtThreads, CtRegs<LPHI>::CTAS) ctile_kernel(const __grid_constant__ CtArgs A) {
  using Conf = CtConf<COLLOCATE, LPHI>;
  constexpr int NS = Conf::NS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned short *s_zm = reinterpret_cast<unsigned short *>(ct_smem + Conf::ZM);
  double *s_e2t = reinterpret_cast<double *>(ct_smem + Conf::E2T);
  int *s_item = reinterpret_cast<int *>(ct_smem + Conf::ITEM);
  CtRing<COLLOCATE, LPHI> R;
  R.bar0 = ct_smem_u32(ct_smem + Conf::BAR);
  R.gbase = 0u, R.cbase = 0u;
  for (int q = tid; q < kCtZmRows * kCtZmPitch; q += kCtThreads) {  // split the plane masks per z half
    const unsigned m = A.zmask[q];
    s_zm[q] = (unsigned short)(m & 0xffffu);
    s_zm[kCtZmRows * kCtZmPitch + q] = (unsigned short)(m >> 16);
  }
  if (tid < 64)
    s_e2t[tid] = exp2((double)tid * (1.0 / 64.0));
  if (tid < NS)
    *reinterpret_cast<unsigned *>(R.slot(tid) + CtSlot<COLLOCATE, LPHI>::HDR + 20) = 0u;
  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(R.full(s), 1);
      mbar_init(R.empty(s), kCtConsumers);
    }
    for (int b = 0; b < kCtChunks; b++) {
      mbar_init(R.cfull(b), 1);
      mbar_init(R.cempty(b), kCtProducers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kCtConsumers) {
    // ===================== producer warpgroup =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CtRegs<LPHI>::PRODUCER));
    const int pw = warp - kCtConsumers;
    for (;;) {
      __syncthreads();
      if (tid == kCtConsumers * 32)
        *s_item = atomicAdd(A.counter, 1);
      __syncthreads();
      const int iw = *s_item;
      if (iw >= A.nwork)
        break;
      const int4 W0 = reinterpret_cast<const int4 *>(A.work)[2 * iw];      // x0, y0, z0, level
      const int4 W1 = reinterpret_cast<const int4 *>(A.work)[2 * iw + 1];  // b[0..3]
      const CtLevelArgs &L = A.lev[W0.w];
      R.vbeg = W1.x, R.vend = W1.w;
      R.e0 = (LPLO + 1 <= LPHI) ? W1.y : W1.w;
      R.e1 = (LPLO + 2 <= LPHI) ? W1.z : W1.w;
      ct_producer_run<COLLOCATE, LPLO, LPHI>(R, L, A, s_e2t, pw, lane);
      R.gbase += (unsigned)(R.vend - R.vbeg);
      R.cbase += (unsigned)((R.vend - R.vbeg + kCtChunk - 1) / kCtChunk);
    }
    return;
  }

  // ===================== consumer warpgroups =====================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CtRegs<LPHI>::CONSUMER));
  CtWarp c;
  c.lane = lane, c.warp = warp;
  const int li = lane & 7, lj = lane >> 3;  // my columns within the warp block: (li, lj) and (li, lj + 4)
  const int bx = warp & 1, by = (warp >> 1) & 1, bz = warp >> 2;
  c.xrow = 8 * bx + li;
  c.yrow = 16 + 8 * by + lj;
  c.zrow0 = 16 * bz;
  c.kt_off = (8 * by + lj) * 16 + 8 * bx + li;
  c.zm = s_zm + bz * (kCtZmRows * kCtZmPitch);

  for (;;) {
    __syncthreads();
    __syncthreads();  // (the producers' leader publishes the item between the two barriers)
    const int iw = *s_item;
    if (iw >= A.nwork)
      break;
    const int4 W0 = reinterpret_cast<const int4 *>(A.work)[2 * iw];      // x0, y0, z0, level
    const int4 W1 = reinterpret_cast<const int4 *>(A.work)[2 * iw + 1];  // b[0..3]
    const CtLevelArgs &L = A.lev[W0.w];
    const int x0 = W0.x + 8 * bx, y0 = W0.y + 8 * by, z0 = W0.z + 16 * bz;
    const int vx = L.nx - x0, vy = L.ny - y0, vz = L.nz - z0;  // valid extent of my warp block (may be <= 0)
    const size_t sy = L.nx, sz = (size_t)L.nx * L.ny;
    double *g0 = L.grid + (size_t)z0 * sz + (size_t)(y0 + lj) * sy + x0 + li;
    double *g1 = g0 + 4 * sy;
    const bool col0 = (li < vx && lj < vy), col1 = (li < vx && lj + 4 < vy);

    // Points outside the grid (partial edge tiles) need no masking in the visit loops:
    // they integrate zeros and their collocated values are never flushed.
    double acc0[16], acc1[16];
#pragma unroll
    for (int p = 0; p < 16; p++) {
      acc0[p] = 0.0, acc1[p] = 0.0;
      if (!COLLOCATE && p < vz) {
        if (col0)
          acc0[p] = g0[p * sz];
        if (col1)
          acc1[p] = g1[p * sz];
      }
    }

    R.vbeg = W1.x, R.vend = W1.w;
    R.e0 = (LPLO + 1 <= LPHI) ? W1.y : W1.w;
    R.e1 = (LPLO + 2 <= LPHI) ? W1.z : W1.w;
    ct_consumer_run<COLLOCATE, LPLO, LPHI>(R, L, c, R.vbeg, R.e0, acc0, acc1);
    if constexpr (LPLO + 1 <= LPHI)
      ct_consumer_run<COLLOCATE, LPLO + 1, LPHI>(R, L, c, R.e0, R.e1, acc0, acc1);
    if constexpr (LPLO + 2 <= LPHI)
      ct_consumer_run<COLLOCATE, LPLO + 2, LPHI>(R, L, c, R.e1, R.vend, acc0, acc1);
    R.gbase += (unsigned)(R.vend - R.vbeg);

    if (COLLOCATE) {
#pragma unroll
      for (int p = 0; p < 16; p++) {
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

// ---------------------------------------------------------------------------
// Host: launch
// ---------------------------------------------------------------------------
struct CtileList {  // list-wide data of the CTA-tile path
  CWork *d_work[kNumClasses] = {nullptr, nullptr, nullptr};
  std::vector<int> level_first[kNumClasses];  // [nlevels + 1] ranges into d_work[cls], by level
  int *d_counters = nullptr;                  // [2 directions][kNumClasses][8]
  unsigned *d_zmask = nullptr;
  void release() {
    for (auto &p : d_work) {
      dev_free(p);
      p = nullptr;
    }
    dev_free(d_counters), dev_free(d_zmask);
    d_counters = nullptr, d_zmask = nullptr;
    for (auto &v : level_first)
      v.clear();
  }
};

inline void finish_ctile_list(CtileList &cl, std::vector<CtileLevel *> &levels, cudaStream_t s) {
  cl.release();
  const std::vector<unsigned> zm = build_ct_zmask();
  dev_alloc(&cl.d_zmask, zm.size() * sizeof(unsigned));
  B200_CHECK(cudaMemcpyAsync(cl.d_zmask, zm.data(), zm.size() * sizeof(unsigned), cudaMemcpyHostToDevice, s));
  dev_alloc(&cl.d_counters, 2 * kNumClasses * 8 * sizeof(int));
  for (int cls = 0; cls < kNumClasses; cls++) {
    std::vector<CWork> all;
    cl.level_first[cls].assign(levels.size() + 1, 0);
    for (size_t l = 0; l < levels.size(); l++) {
      cl.level_first[cls][l] = (int)all.size();
      all.insert(all.end(), levels[l]->work[cls].begin(), levels[l]->work[cls].end());
      levels[l]->work[cls].clear();
      levels[l]->work[cls].shrink_to_fit();
    }
    cl.level_first[cls][levels.size()] = (int)all.size();
    dev_alloc(&cl.d_work[cls], std::max<size_t>(all.size(), 1) * sizeof(CWork));
    if (!all.empty())
      B200_CHECK(cudaMemcpyAsync(cl.d_work[cls], all.data(), all.size() * sizeof(CWork), cudaMemcpyHostToDevice, s));
    B200_CHECK(cudaStreamSynchronize(s));  // `all` is a temporary
  }
}

template <bool COLLOCATE, int LPLO, int LPHI> inline void launch_ctile_class(const CtArgs &A, cudaStream_t s) {
  constexpr int bytes = CtConf<COLLOCATE, LPHI>::BYTES;
  static_assert(bytes <= 200 * 1024, "ctile kernel: shared memory budget exceeded");
  static bool configured = false;
  static int per_sm = 1;
  if (!configured) {
    B200_CHECK(cudaFuncSetAttribute(ctile_kernel<COLLOCATE, LPLO, LPHI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    bytes));
    B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctile_kernel<COLLOCATE, LPLO, LPHI>,
                                                            kCtThreads, bytes));
    configured = true;
  }
  int nsm = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = std::min(A.nwork, nsm * std::max(per_sm, 1));
  ctile_kernel<COLLOCATE, LPLO, LPHI><<<grid, kCtThreads, bytes, s>>>(A);
  B200_CHECK(cudaGetLastError());
  count_launch();
}

// What a call hands to launch_ctile: the levels [l0, l1) of the list.
struct CtileCall {
  const CtileList *list;
  CtileLevel *const *levels;       // [nlevels]
  const LevelDev *level_dev;       // [nlevels]
  double *const *grids;            // [nlevels] device pointers
  int l0, l1;
  int dl;
  double *coef;
  cudaStream_t stream;
};

// Launches the CTA-tile kernels of every lp class that stays within the tiled range for
// this call's l growth; returns a bit mask of the classes that must be handled by the
// generic kernel instead.
template <bool COLLOCATE> inline unsigned launch_ctile(const CtileCall &C) {
  B200_ASSERT(C.dl >= 0 && C.dl < 8, "unexpected l growth");
  B200_ASSERT(C.l1 <= kCtMaxLevels, "more grid levels than the CTA-tile kernels support");
  unsigned leftover = 0u;
  int *counters = C.list->d_counters + (COLLOCATE ? 0 : kNumClasses * 8);
  B200_CHECK(cudaMemsetAsync(counters, 0, kNumClasses * 8 * sizeof(int), C.stream));
  for (int cls = 0; cls < kNumClasses; cls++) {
    const int w0 = C.list->level_first[cls][C.l0], w1 = C.list->level_first[cls][C.l1];
    if (w1 == w0)
      continue;
    const int lo = kClassLo[cls] + C.dl, hi = kClassHi[cls] + C.dl;
    if (hi > kTiledMaxLpCall) {
      leftover |= 1u << cls;
      continue;
    }
    CtArgs A;
    memset(&A, 0, sizeof(A));
    for (int l = C.l0; l < C.l1; l++) {
      const CtileLevel &T = *C.levels[l];
      const LevelDev &D = C.level_dev[l];
      CtLevelArgs &LA = A.lev[l];
      LA.visits = T.d_visits, LA.ctasks = T.d_ctasks, LA.ktab = T.d_ktab, LA.grid = C.grids[l];
      LA.nx = D.npts_local[0], LA.ny = D.npts_local[1], LA.nz = D.npts_local[2];
      LA.coef_base = T.coef_base[C.dl][cls], LA.tt_first = T.class_tt_first[cls];
      LA.hx = D.dh[0], LA.hy = D.dh[4], LA.hz = D.dh[8];
    }
    A.work = C.list->d_work[cls] + w0, A.nwork = w1 - w0;
    A.counter = counters + cls * 8;
    A.zmask = C.list->d_zmask;
    A.coef = C.coef;
    A.coef_stride = (ncoset(hi) + 1) / 2 * 2;
    cudaStream_t s = C.stream;
    if (lo == 0) launch_ctile_class<COLLOCATE, 0, 2>(A, s);
    else if (lo == 1) launch_ctile_class<COLLOCATE, 1, 3>(A, s);
    else if (lo == 2) launch_ctile_class<COLLOCATE, 2, 4>(A, s);
    else if (lo == 3 && hi == 5) launch_ctile_class<COLLOCATE, 3, 5>(A, s);
    else if (lo == 3) launch_ctile_class<COLLOCATE, 3, 4>(A, s);
    else if (lo == 4 && hi == 6) launch_ctile_class<COLLOCATE, 4, 6>(A, s);
    else if (lo == 4) launch_ctile_class<COLLOCATE, 4, 5>(A, s);
    else if (lo == 5 && hi == 7) launch_ctile_class<COLLOCATE, 5, 7>(A, s);
    else if (lo == 5) launch_ctile_class<COLLOCATE, 5, 6>(A, s);
    else if (lo == 6 && hi == 7) launch_ctile_class<COLLOCATE, 6, 7>(A, s);
    else leftover |= 1u << cls;
  }
  return leftover;
}

}  // namespace b200
