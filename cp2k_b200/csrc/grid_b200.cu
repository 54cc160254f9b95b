// B200-native backend of CP2K's grid library: task-list builder, device
// residency and the C ABI declared in include/grid_b200.h.
//
// Reference behaviour mirrored here (paths relative to /root/reference/src/grid):
//   dispatcher contract            grid_task_list.c:23-135 (create, handle reuse,
//                                  empty lists), :175-270, :277-436
//   task sort + per-level ranges   ref/grid_ref_task_list.c:26-37, 140-163
//   per-task geometry              ref/grid_ref_collint.h:222-254, 611-640, 929-947
//   GPU-backend call protocol      gpu/grid_gpu_context.cu:479-556, 562-655
// Single translation unit: the kernel headers are included below so that the
// __constant__ tables are shared.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>
#include <omp.h>
#include <parallel/algorithm>
#include <vector>

#include "b200_internal.cuh"
#include "b200_coef.cuh"
#include "b200_generic.cuh"
#include "b200_tiled.cuh"
#include "b200_ctile.cuh"

namespace b200 {

// ---------------------------------------------------------------------------
// library-wide state
// ---------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
static int g_device = -1;
static cudaStream_t g_stream = nullptr;
static bool g_device_resident = false;
// kernel family: 0 = automatic (kDefaultPath), 1 = generic kernels only, 2 = warp-tile kernels
// (b200_tiled.cuh), 3 = CTA-tile kernels (b200_ctile.cuh)
static int g_variant = 0;
constexpr int kDefaultPath = 2;  // what "automatic" resolves to: the faster family on the H2O benchmarks
static inline int variant_path() {
  const int v = (g_variant == 0) ? kDefaultPath : g_variant;
  return v == 3 ? 0 : 2;  // TaskList::path: 0 = CTA-tile data, 2 = warp-tile data
}
static bool g_tables_uploaded[64] = {false};

void count_launch(int n) { g_launches += n; }

// Optional device-side timing of the library's own phases (CUDA events on the
// launching stream), used by bench.py for the per-kernel roofline numbers.
enum TimedClass { T_PAB2COEF = 0, T_COLLOCATE, T_INTEGRATE, T_COEF2HAB, T_H2D, T_D2H, T_MEMSET, T_NCLASS };
struct TimedSpan {
  cudaEvent_t a, b;
  int cls;
};
static bool g_timing = false;
static std::vector<TimedSpan> g_spans;
static double g_ms[T_NCLASS] = {0};
static double g_nspans[T_NCLASS] = {0};
struct ScopedTimer {
  TimedSpan sp;
  cudaStream_t s;
  bool on;
  ScopedTimer(int cls, cudaStream_t stream) : s(stream), on(g_timing) {
    if (!on)
      return;
    sp.cls = cls;
    B200_CHECK(cudaEventCreate(&sp.a));
    B200_CHECK(cudaEventCreate(&sp.b));
    B200_CHECK(cudaEventRecord(sp.a, s));
  }
  ~ScopedTimer() {
    if (!on)
      return;
    B200_CHECK(cudaEventRecord(sp.b, s));
    if (g_spans.size() >= 65536) {  // nobody drains the timers: drop the oldest span
      cudaEventDestroy(g_spans.front().a);
      cudaEventDestroy(g_spans.front().b);
      g_spans.erase(g_spans.begin());
    }
    g_spans.push_back(sp);
  }
};

static void activate_device() {
  if (g_device >= 0)
    B200_CHECK(cudaSetDevice(g_device));
  int dev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  B200_ASSERT(dev >= 0 && dev < 64, "device index beyond the constant-table bookkeeping (64 devices)");
  if (!g_tables_uploaded[dev]) {
    OrbTable tab;
    memset(&tab, 0, sizeof(tab));
    for (int lx = 0; lx <= 15; lx++)
      for (int ly = 0; lx + ly <= 15; ly++)
        for (int lz = 0; lx + ly + lz <= 15; lz++) {
          const int c = coset(lx, ly, lz);
          tab.l[c][0] = lx, tab.l[c][1] = ly, tab.l[c][2] = lz;
        }
    B200_CHECK(cudaMemcpyToSymbol(c_orb, &tab, sizeof(tab)));
    double binom[kMaxLSide + 1][kMaxLSide + 1];
    for (int n = 0; n <= kMaxLSide; n++)
      for (int k = 0; k <= kMaxLSide; k++) {
        double r = (k <= n) ? 1.0 : 0.0;
        for (int i = 1; i <= k && k <= n; i++)
          r = r * (double)(n - k + i) / (double)i;
        binom[n][k] = r;
      }
    B200_CHECK(cudaMemcpyToSymbol(c_binom, binom, sizeof(binom)));
    g_tables_uploaded[dev] = true;
  }
}

template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  void ensure(size_t n) {
    if (n <= cap)
      return;
    if (p)
      dev_free(p);
    dev_alloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    cap = n;
  }
  template <class Vec> void upload(const Vec &v, cudaStream_t s) {
    ensure(v.size());
    if (!v.empty()) {
      B200_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
      B200_CHECK(cudaStreamSynchronize(s));  // v may be a temporary
    }
  }
  void release() {
    if (p)
      dev_free(p);
    p = nullptr, cap = 0;
  }
};

struct Kind {
  int nset, nsgf, maxco, maxpgf;
  std::vector<int> lmin, lmax, npgf, nsgf_set, first_sgf;
  std::vector<double> zet;
  int sphi_off;
};

struct LevelInfo {
  int first = 0, last = 0;           // task range [first,last) in sorted order
  int n_generic = 0;                 // tasks that must use the generic kernel
  int max_lp0 = 0;                   // max la_max+lb_max over the level
  int max_w = 1;                     // max cube / index-box edge
  int max_lp0_general = -1;          // over non-ortho tasks
  TiledLevel tiled;                  // warp-tile path data (b200_tiled.cuh), path == 2
  CtileLevel ctile;                  // CTA-tile path data (b200_ctile.cuh), path == 0
};

constexpr int kNumSizeClasses = 3;
inline int size_class(const TaskDev &T) {
  const int lp0 = T.la_max + T.lb_max;
  return lp0 <= 1 ? 0 : (lp0 == 2 ? 1 : 2);
}

struct TaskList {
  bool empty = true;
  bool ortho = false;
  int ntasks = 0, nlevels = 0, natoms = 0, nkinds = 0, nblocks = 0;
  std::vector<LevelDev> levels;
  std::vector<LevelInfo> linfo;
  TaskVec h_tasks;
  DevBuf<TaskDev> d_tasks;
  DevBuf<double> d_sphi;
  DevBuf<int> d_iota, d_generic_ids, d_block_task_ids, d_block_first;
  std::vector<int> h_generic_ids;  // per level, concatenated; offsets below
  std::vector<int> generic_first;
  DevBuf<int> d_coef_off[8];
  size_t coef_total[8] = {0};
  bool coef_ready[8] = {false};
  DevBuf<double> d_coef, d_pab, d_hab, d_fv;
  std::vector<DevBuf<double>> d_grids;
  std::map<int, DevBuf<double>> Tmats;  // key = level*(kMaxLp+1)+lp
  std::vector<const double *> h_Tptrs;
  DevBuf<const double *> d_Tptrs;
  // index lists of the cab <-> cxyz transform per (la, lb) after ldiffs
  struct GatherBufs {
    DevBuf<unsigned long long> by_k, by_ab;
    DevBuf<int> kstart, kperm, abstart;
  };
  std::map<int, GatherBufs> gather_bufs;    // key = la * (kMaxLSide+1) + lb
  std::vector<GatherList> h_glists;
  DevBuf<GatherList> d_glists;
  std::vector<std::pair<int, int>> l_combos;  // (la_max, lb_max) present in the list
  int max_nsgf_set = 1, max_ncoset_raw = 1, max_la = 0, max_lb = 0, maxco = 1;
  int max_block_size = 1;
  size_t pab_len = 0;
  // host<->device copies of the P/H blocks are pipelined against the coefficient
  // kernels in chunks of consecutive matrix blocks (needs monotonic offsets)
  struct Chunk {
    int t0, t1;        // range in the by-block task order
    size_t off0;       // first double of the chunk in the block buffers
    int c0[kNumSizeClasses + 1];  // the chunk's tasks per size class: ranges in d_class_ids_chunked
  };
  // The coefficient kernels size their per-task scratch in shared memory for the largest
  // task of a launch; tasks are therefore launched per SIZE CLASS (by la_max + lb_max), so
  // that the many small products are not throttled to the occupancy of the few large ones.
  struct SizeClass {
    int max_nsgf_set = 1, max_ncoset_raw = 1, max_la = 0, max_lb = 0;
  };
  SizeClass sclass[kNumSizeClasses];
  DevBuf<int> d_class_ids_chunked;  // per chunk: class 0 | class 1 | class 2, block order inside
  DevBuf<int> d_class_ids_global;   // whole list: class 0 | class 1 | class 2, block order inside
  int g0[kNumSizeClasses + 1] = {0};
  std::vector<Chunk> chunks;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> ev_chunk;
  cudaEvent_t ev_copy_done = nullptr;
  std::vector<cudaStream_t> level_streams;   // the levels' grid kernels run concurrently
  cudaEvent_t ev_fork = nullptr;
  std::vector<cudaEvent_t> ev_join;
  double stats[16] = {0};
  bool stats_ready = false;
  int last_dl = 0;                   // l growth of the last collocate / integrate call (statistics)
  int path = 0;                      // which tiled family the list was built for (0 / 2)
  CtileList ct;

  void release() {
    for (auto st : level_streams)
      cudaStreamDestroy(st);
    level_streams.clear();
    for (auto ev : ev_join)
      cudaEventDestroy(ev);
    ev_join.clear();
    if (ev_fork)
      cudaEventDestroy(ev_fork);
    ev_fork = nullptr;
    for (auto ev : ev_chunk)
      cudaEventDestroy(ev);
    ev_chunk.clear();
    if (ev_copy_done)
      cudaEventDestroy(ev_copy_done);
    ev_copy_done = nullptr;
    if (copy_stream)
      cudaStreamDestroy(copy_stream);
    copy_stream = nullptr;
    chunks.clear();
    d_tasks.release(), d_sphi.release(), d_iota.release(), d_generic_ids.release();
    d_block_task_ids.release(), d_block_first.release();
    d_class_ids_chunked.release(), d_class_ids_global.release();
    for (auto &c : sclass)
      c = SizeClass();
    for (auto &b : d_coef_off)
      b.release();
    d_coef.release(), d_pab.release(), d_hab.release(), d_fv.release();
    for (auto &g : d_grids)
      g.release();
    d_grids.clear();
    for (auto &kv : Tmats)
      kv.second.release();
    Tmats.clear();
    d_Tptrs.release();
    for (auto &kv : gather_bufs) {
      kv.second.by_k.release(), kv.second.by_ab.release();
      kv.second.kstart.release(), kv.second.kperm.release(), kv.second.abstart.release();
    }
    gather_bufs.clear();
    h_glists.clear();
    d_glists.release();
    for (auto &li : linfo)
      li.tiled.release(), li.ctile.release();
    ct.release();
    for (int i = 0; i < 8; i++)
      coef_ready[i] = false, coef_total[i] = 0;
    stats_ready = false;
  }
};

// ---------------------------------------------------------------------------
// Cartesian -> lattice polynomial basis for triclinic cells
// (ref/grid_ref_collint.h:697-762), built by polynomial multiplication.
// T[q][c]: coefficient of di^il dj^jl dk^kl (q = coset(il,jl,kl)) in
// x^lx y^ly z^lz (c = coset(lx,ly,lz)).
// ---------------------------------------------------------------------------
static void poly3_mul(int lp, const std::vector<double> &a, const std::vector<double> &b,
                      std::vector<double> &c) {
  const int n = lp + 1;
  std::fill(c.begin(), c.end(), 0.0);
  for (int ak = 0; ak < n; ak++)
    for (int aj = 0; aj < n - ak; aj++)
      for (int ai = 0; ai < n - ak - aj; ai++) {
        const double av = a[(ak * n + aj) * n + ai];
        if (av == 0.0)
          continue;
        for (int bk = 0; ak + bk < n; bk++)
          for (int bj = 0; ak + bk + aj + bj < n; bj++)
            for (int bi = 0; ak + bk + aj + bj + ai + bi < n; bi++)
              c[((ak + bk) * n + aj + bj) * n + ai + bi] += av * b[(bk * n + bj) * n + bi];
      }
}

static std::vector<double> build_cijk_transform(int lp, const double *dh) {
  const int n = lp + 1, n3 = n * n * n, nc = ncoset(lp);
  std::vector<std::vector<double>> pw(3 * n, std::vector<double>(n3, 0.0));
  std::vector<double> lin(n3), t1(n3), t2(n3);
  for (int c = 0; c < 3; c++) {
    std::fill(lin.begin(), lin.end(), 0.0);
    if (lp >= 1) {
      lin[(0 * n + 0) * n + 1] = dh[0 * 3 + c];
      lin[(0 * n + 1) * n + 0] = dh[1 * 3 + c];
      lin[(1 * n + 0) * n + 0] = dh[2 * 3 + c];
    }
    pw[c * n + 0][0] = 1.0;
    for (int p = 1; p <= lp; p++)
      poly3_mul(lp, pw[c * n + p - 1], lin, pw[c * n + p]);
  }
  std::vector<double> T((size_t)nc * nc, 0.0);
  for (int lz = 0; lz <= lp; lz++)
    for (int ly = 0; ly <= lp - lz; ly++)
      for (int lx = 0; lx <= lp - lz - ly; lx++) {
        poly3_mul(lp, pw[0 * n + lx], pw[1 * n + ly], t1);
        poly3_mul(lp, t1, pw[2 * n + lz], t2);
        const int c = coset(lx, ly, lz);
        for (int kl = 0; kl <= lp; kl++)
          for (int jl = 0; jl <= lp - kl; jl++)
            for (int il = 0; il <= lp - kl - jl; il++)
              T[(size_t)coset(il, jl, kl) * nc + c] = t2[(kl * n + jl) * n + il];
      }
  return T;
}

static void ensure_transforms(TaskList &tl, int dl, cudaStream_t s) {
  bool changed = false;
  for (int lev = 0; lev < tl.nlevels; lev++) {
    const int top = tl.linfo[lev].max_lp0_general;
    if (top < 0)
      continue;
    for (int lp = 0; lp <= top + dl; lp++) {
      const int key = lev * (kMaxLp + 1) + lp;
      if (tl.Tmats.count(key))
        continue;
      B200_ASSERT(lp <= kMaxLp, "lp too large");
      std::vector<double> T = build_cijk_transform(lp, tl.levels[lev].dh);
      tl.Tmats[key].upload(T, s);
      tl.h_Tptrs[key] = tl.Tmats[key].p;
      changed = true;
    }
  }
  if (changed)
    tl.d_Tptrs.upload(tl.h_Tptrs, s);
}

// Index lists of the Cartesian <-> polynomial transform (GatherList,
// b200_internal.cuh) for every (la_max + dla, lb_max + dlb) this call meets.
static void ensure_gather_lists(TaskList &tl, const int dla, const int dlb, cudaStream_t s) {
  constexpr int W = kMaxLSide + 1;
  if (tl.h_glists.empty())
    tl.h_glists.assign((size_t)W * W, GatherList{nullptr, nullptr, nullptr, nullptr, nullptr});
  bool changed = false;
  for (const auto &lc : tl.l_combos) {
    const int la = lc.first + dla, lb = lc.second + dlb;
    B200_ASSERT(la <= kMaxLSide && lb <= kMaxLSide, "angular momentum too large");
    const int key = la * W + lb;
    if (tl.gather_bufs.count(key))
      continue;
    const int n1 = ncoset(la), n2 = ncoset(lb), lp = la + lb, lp1 = lp + 1, nc = ncoset(lp);
    std::vector<int> lx(std::max(n1, n2)), ly(lx.size()), lz(lx.size());
    for (int x = 0; x <= std::max(la, lb); x++)
      for (int y = 0; x + y <= std::max(la, lb); y++)
        for (int z = 0; x + y + z <= std::max(la, lb); z++) {
          const int c = coset(x, y, z);
          lx[c] = x, ly[c] = y, lz[c] = z;
        }
    auto AL = [&](int d, int a, int b, int k) { return (((d * (la + 1) + a) * (lb + 1) + b) * lp1) + k; };
    B200_ASSERT(n1 * n2 < 65536 && AL(2, la, lb, lp) < 65536 && nc < 65536, "gather list index overflow");
    std::vector<unsigned long long> by_ab;
    std::vector<int> abstart(n1 * n2 + 1, 0);
    std::vector<std::vector<unsigned long long>> per_k(nc);
    for (int ib = 0; ib < n2; ib++)
      for (int ia = 0; ia < n1; ia++) {
        abstart[ib * n1 + ia] = (int)by_ab.size();
        for (int kz = 0; kz <= lz[ia] + lz[ib]; kz++)
          for (int ky = 0; ky <= ly[ia] + ly[ib]; ky++)
            for (int kx = 0; kx <= lx[ia] + lx[ib]; kx++) {
              const unsigned long long fac = ((unsigned long long)AL(0, lx[ia], lx[ib], kx) << 16) |
                                             ((unsigned long long)AL(1, ly[ia], ly[ib], ky) << 32) |
                                             ((unsigned long long)AL(2, lz[ia], lz[ib], kz) << 48);
              const int k = coset(kx, ky, kz);
              by_ab.push_back(fac | (unsigned long long)k);
              per_k[k].push_back(fac | (unsigned long long)(ib * n1 + ia));
            }
      }
    abstart[n1 * n2] = (int)by_ab.size();
    std::vector<unsigned long long> by_k;
    std::vector<int> kstart(nc + 1, 0), kperm(nc);
    for (int k = 0; k < nc; k++) {
      kstart[k] = (int)by_k.size();
      by_k.insert(by_k.end(), per_k[k].begin(), per_k[k].end());
      kperm[k] = k;
    }
    kstart[nc] = (int)by_k.size();
    std::stable_sort(kperm.begin(), kperm.end(),
                     [&](int a, int b) { return per_k[a].size() > per_k[b].size(); });
    TaskList::GatherBufs &G = tl.gather_bufs[key];
    G.by_k.upload(by_k, s), G.by_ab.upload(by_ab, s);
    G.kstart.upload(kstart, s), G.kperm.upload(kperm, s), G.abstart.upload(abstart, s);
    tl.h_glists[key] = GatherList{G.by_k.p, G.kstart.p, G.kperm.p, G.by_ab.p, G.abstart.p};
    changed = true;
  }
  if (changed || tl.d_glists.p == nullptr)
    tl.d_glists.upload(tl.h_glists, s);
}

// Coefficient buffer layout for one l growth `dl`.  Tasks on the tiled path
// get fixed-size slots per (level, lp class) so that the hot kernels compute a
// task's slot from its index without a load; all other tasks are packed behind.
static void ensure_coef_offsets(TaskList &tl, int dl, cudaStream_t s) {
  B200_ASSERT(dl >= 0 && dl < 8, "unexpected l growth");
  if (tl.coef_ready[dl])
    return;
  std::vector<int> off(tl.ntasks, -1);
  size_t total = 0;
  // (slots start at even offsets and have even strides: 16-byte aligned for the bulk copies)
  // The class kernels' look-ahead reads the slot of one pair past their range with THEIR class's
  // base and stride; that address stays inside the level's run of slots (and the buffer) only
  // because a level's classes are laid out back to back in ascending order with non-decreasing
  // strides, and the buffer ends with slack:
  static_assert(kClassHi[0] < kClassHi[1] && kClassHi[1] < kClassHi[2] && kNumClasses == 3,
                "coefficient slots: class strides must ascend (look-ahead of the pair loops)");
  for (int lev = 0; lev < tl.nlevels; lev++) {
    LevelInfo &li = tl.linfo[lev];
    const bool ct = (tl.path == 0);
    const int ntiled = ct ? li.ctile.ntasks_tiled : li.tiled.ntasks_tiled;
    const int *tt_first = ct ? li.ctile.class_tt_first : li.tiled.class_tt_first;
    const std::vector<int> &tt_task = ct ? li.ctile.h_tt_task : li.tiled.h_tt_task;
    for (int cls = 0; cls < kNumClasses; cls++) {
      const int stride = (ncoset(kClassHi[cls] + dl) + 1) / 2 * 2;
      B200_ASSERT(total < (size_t)INT_MAX, "coefficient buffer exceeds 2^31 entries");
      (ct ? li.ctile.coef_base : li.tiled.coef_base)[dl][cls] = (int)total;
      if (ntiled == 0)
        continue;
      for (int q = tt_first[cls]; q < tt_first[cls + 1]; q++) {
        off[tt_task[q]] = (int)total;
        total += stride;
      }
    }
  }
  for (int i = 0; i < tl.ntasks; i++)
    if (off[i] < 0) {
      off[i] = (int)total;
      total += (ncoset(tl.h_tasks[i].la_max + tl.h_tasks[i].lb_max + dl) + 1) / 2 * 2;
    }
  total += 256;  // slack for whole-slot prefetches
  B200_ASSERT(total < (size_t)INT_MAX, "coefficient buffer exceeds 2^31 entries");
  tl.d_coef_off[dl].upload(off, s);
  tl.coef_total[dl] = total;
  tl.coef_ready[dl] = true;
}

// ---------------------------------------------------------------------------
// create
// ---------------------------------------------------------------------------
static void build_task_list(
    TaskList &tl, const bool orthorhombic, const int ntasks, const int nlevels, const int natoms,
    const int nkinds, const int nblocks, const int *block_offsets, const double *atom_positions,
    const int *atom_kinds, const grid_b200_basis_set **basis_sets, const int *level_list,
    const int *iatom_list, const int *jatom_list, const int *iset_list, const int *jset_list,
    const int *ipgf_list, const int *jpgf_list, const int *border_mask_list,
    const int *block_num_list, const double *radius_list, const double *rab_list,
    const int *npts_global, const int *npts_local, const int *shift_local,
    const int *border_width, const double *dh, const double *dh_inv) {
  cudaStream_t s = g_stream;
  // GRID_B200_CREATE_TIMING=1: wall time of the builder's phases on stderr
  static const bool timing = (getenv("GRID_B200_CREATE_TIMING") != nullptr);
  auto t_last = std::chrono::steady_clock::now();
  auto tick = [&](const char *what) {
    if (!timing)
      return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "grid_b200 create: %-28s %8.1f ms\n", what,
            std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  tl.path = (nlevels > kCtMaxLevels) ? 2 : variant_path();
  tl.ortho = orthorhombic;
  tl.ntasks = ntasks, tl.nlevels = nlevels, tl.natoms = natoms;
  tl.nkinds = nkinds, tl.nblocks = nblocks;

  // basis sets: deep copy of what the kernels need, sphi into one pool
  std::vector<Kind> kinds(nkinds);
  std::vector<double> sphi_pool;
  for (int k = 0; k < nkinds; k++) {
    const grid_b200_basis_set *b = basis_sets[k];
    Kind &K = kinds[k];
    K.nset = b->nset, K.nsgf = b->nsgf, K.maxco = b->maxco, K.maxpgf = b->maxpgf;
    K.lmin.assign(b->lmin, b->lmin + b->nset);
    K.lmax.assign(b->lmax, b->lmax + b->nset);
    K.npgf.assign(b->npgf, b->npgf + b->nset);
    K.nsgf_set.assign(b->nsgf_set, b->nsgf_set + b->nset);
    K.first_sgf.assign(b->first_sgf, b->first_sgf + b->nset);
    K.zet.assign(b->zet, b->zet + (size_t)b->nset * b->maxpgf);
    K.sphi_off = (int)sphi_pool.size();
    sphi_pool.insert(sphi_pool.end(), b->sphi, b->sphi + (size_t)b->nsgf * b->maxco);
    tl.maxco = std::max(tl.maxco, b->maxco);
  }
  tl.d_sphi.upload(sphi_pool, s);

  tl.levels.resize(nlevels);
  tl.linfo.assign(nlevels, LevelInfo());
  std::vector<double> dh_max(nlevels, 0.0);
  for (int l = 0; l < nlevels; l++) {
    LevelDev &L = tl.levels[l];
    for (int i = 0; i < 3; i++) {
      L.npts_global[i] = npts_global[3 * l + i];
      L.npts_local[i] = npts_local[3 * l + i];
      L.shift_local[i] = shift_local[3 * l + i];
      L.border_width[i] = border_width[3 * l + i];
    }
    for (int i = 0; i < 9; i++) {
      L.dh[i] = dh[9 * l + i];
      L.dh_inv[i] = dh_inv[9 * l + i];
      dh_max[l] = fmax(dh_max[l], fabs(L.dh[i]));
    }
    const size_t npts = (size_t)L.npts_local[0] * L.npts_local[1] * L.npts_local[2];
    B200_ASSERT(npts < (size_t)INT_MAX, "local grid exceeds 2^31 points");
  }

  // sort like the reference: (level, block, iset, jset), stable
  std::vector<int> order(ntasks);
  std::iota(order.begin(), order.end(), 0);
  // (multi-threaded: the list has 10^6 - 10^7 tasks and is rebuilt every MD step)
  __gnu_parallel::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (level_list[a] != level_list[b])
      return level_list[a] < level_list[b];
    if (block_num_list[a] != block_num_list[b])
      return block_num_list[a] < block_num_list[b];
    if (iset_list[a] != iset_list[b])
      return iset_list[a] < iset_list[b];
    return jset_list[a] < jset_list[b];
  });

  tick("basis sets, task sort");
  tl.h_tasks.resize(ntasks);
#pragma omp parallel for schedule(static)
  for (int it = 0; it < ntasks; it++) {
    const int i = order[it];
    TaskDev &T = tl.h_tasks[it];
    memset(&T, 0, sizeof(T));
    T.level = level_list[i] - 1;
    T.iatom = iatom_list[i] - 1, T.jatom = jatom_list[i] - 1;
    T.iset = iset_list[i] - 1, T.jset = jset_list[i] - 1;
    const int ipgf = ipgf_list[i] - 1, jpgf = jpgf_list[i] - 1;
    T.border_mask = border_mask_list[i];
    T.block_num = block_num_list[i] - 1;
    B200_ASSERT(T.level >= 0 && T.level < nlevels, "task level out of range");
    B200_ASSERT(T.block_num >= 0 && T.block_num < nblocks, "task block out of range");
    B200_ASSERT(T.iatom >= 0 && T.iatom < natoms && T.jatom >= 0 && T.jatom < natoms,
                "task atom out of range");
    T.block_offset = block_offsets[T.block_num];
    T.radius = radius_list[i];
    const Kind &Ka = kinds[atom_kinds[T.iatom] - 1], &Kb = kinds[atom_kinds[T.jatom] - 1];
    T.zeta = Ka.zet[(size_t)T.iset * Ka.maxpgf + ipgf];
    T.zetb = Kb.zet[(size_t)T.jset * Kb.maxpgf + jpgf];
    T.la_max = Ka.lmax[T.iset], T.la_min = Ka.lmin[T.iset];
    T.lb_max = Kb.lmax[T.jset], T.lb_min = Kb.lmin[T.jset];
    T.ncoseta = ncoset(T.la_max), T.ncosetb = ncoset(T.lb_max);
    T.ncoa = Ka.npgf[T.iset] * T.ncoseta, T.ncob = Kb.npgf[T.jset] * T.ncosetb;
    T.sgfa = Ka.first_sgf[T.iset] - 1, T.sgfb = Kb.first_sgf[T.jset] - 1;
    T.nsgf_seta = Ka.nsgf_set[T.iset], T.nsgf_setb = Kb.nsgf_set[T.jset];
    T.nsgfa = Ka.nsgf, T.nsgfb = Kb.nsgf;
    T.o1 = ipgf * T.ncoseta, T.o2 = jpgf * T.ncosetb;
    T.sphi_a = Ka.sphi_off, T.sphi_b = Kb.sphi_off;
    T.maxcoa = Ka.maxco, T.maxcob = Kb.maxco;
    T.transpose = (T.iatom <= T.jatom);
    T.use_ortho = (orthorhombic && T.border_mask == 0);
    const LevelDev &L = tl.levels[T.level];

    // product centre and prefactor (collint.h:939-947)
    T.zetp = T.zeta + T.zetb;
    const double f = T.zetb / T.zetp;
    double rab2 = 0.0;
    for (int d = 0; d < 3; d++) {
      T.ra[d] = atom_positions[3 * T.iatom + d];
      T.rab[d] = rab_list[3 * i + d];
      rab2 += T.rab[d] * T.rab[d];
    }
    T.prefactor = exp(-T.zeta * f * rab2);
    for (int d = 0; d < 3; d++)
      T.rp[d] = T.ra[d] + f * T.rab[d];
    T.skip = (2.0 * T.radius < dh_max[T.level]);

    if (T.use_ortho) {  // collint.h:222-254
      const double h[3] = {L.dh[0], L.dh[4], L.dh[8]};
      const double drmin = fmin(h[0], fmin(h[1], h[2]));
      T.disr_radius = drmin * fmax(1.0, ceil(T.radius / drmin));
      for (int d = 0; d < 3; d++) {
        double sacc = 0.0;
        for (int j = 0; j < 3; j++)
          sacc += L.dh_inv[j * 3 + d] * T.rp[j];
        T.cubecenter[d] = (int)floor(sacc);
        T.roffset[d] = T.rp[d] - ((double)T.cubecenter[d]) * h[d];
        T.lb_cube[d] = (int)ceil(-1e-8 - T.disr_radius * L.dh_inv[d * 3 + d]);
        if (!T.skip && L.npts_global[d] != L.npts_local[d]) {
          const int ub = 1 - T.lb_cube[d];
          const int off =
              pmod(T.cubecenter[d] + T.lb_cube[d] - L.shift_local[d], L.npts_global[d]) -
              T.lb_cube[d];
          B200_ASSERT(off + ub < L.npts_local[d] && off + T.lb_cube[d] >= 0,
                      "cube does not fit the non-periodic local grid (collint.h:247-253)");
        }
      }
    } else {  // collint.h:611-640
      for (int d = 0; d < 3; d++) {
        T.gp[d] = 0.0;
        for (int j = 0; j < 3; j++)
          T.gp[d] += L.dh_inv[j * 3 + d] * T.rp[j];
        T.index_min[d] = INT_MAX, T.index_max[d] = INT_MIN;
      }
      for (int a = -1; a <= 1; a++)
        for (int b = -1; b <= 1; b++)
          for (int c = -1; c <= 1; c++) {
            const double x = T.rp[0] + a * T.radius, y = T.rp[1] + b * T.radius,
                         z = T.rp[2] + c * T.radius;
            for (int d = 0; d < 3; d++) {
              const double resc =
                  L.dh_inv[0 * 3 + d] * x + L.dh_inv[1 * 3 + d] * y + L.dh_inv[2 * 3 + d] * z;
              T.index_min[d] = std::min(T.index_min[d], (int)floor(resc));
              T.index_max[d] = std::max(T.index_max[d], (int)ceil(resc));
            }
          }
    }
  }

  tick("task records (host)");
  // per-level ranges and maxima (threads reduce into private copies, merged at the end)
  for (int l = 0; l < nlevels; l++)
    tl.linfo[l].first = tl.linfo[l].last = 0;
  {
    struct Acc {
      std::vector<LevelInfo> li;
      TaskList::SizeClass sc[kNumSizeClasses];
      int max_nsgf_set = 1, max_ncoset_raw = 1, max_la = 0, max_lb = 0, max_block_size = 1;
      bool combo[kMaxLSide + 1][kMaxLSide + 1] = {};
    };
    std::vector<Acc> accs;
#pragma omp parallel
    {
#pragma omp single
      accs.resize(omp_get_num_threads());
      Acc &A = accs[omp_get_thread_num()];
      A.li.assign(nlevels, LevelInfo());
      for (auto &x : A.li)
        x.first = INT_MAX, x.last = 0;
#pragma omp for schedule(static)
      for (int it = 0; it < ntasks; it++) {
        const TaskDev &T = tl.h_tasks[it];
        LevelInfo &li = A.li[T.level];
        li.first = std::min(li.first, it);
        li.last = std::max(li.last, it + 1);
        const int lp0 = T.la_max + T.lb_max;
        li.max_lp0 = std::max(li.max_lp0, lp0);
        if (!T.skip)
          for (int d = 0; d < 3; d++) {
            const int w = T.use_ortho ? 2 - 2 * T.lb_cube[d] : T.index_max[d] - T.index_min[d] + 1;
            li.max_w = std::max(li.max_w, w);
          }
        if (!T.use_ortho)
          li.max_lp0_general = std::max(li.max_lp0_general, lp0);
        A.max_nsgf_set = std::max({A.max_nsgf_set, T.nsgf_seta, T.nsgf_setb});
        A.max_ncoset_raw = std::max({A.max_ncoset_raw, T.ncoseta, T.ncosetb});
        TaskList::SizeClass &C = A.sc[size_class(T)];
        C.max_nsgf_set = std::max({C.max_nsgf_set, T.nsgf_seta, T.nsgf_setb});
        C.max_ncoset_raw = std::max({C.max_ncoset_raw, T.ncoseta, T.ncosetb});
        C.max_la = std::max(C.max_la, T.la_max), C.max_lb = std::max(C.max_lb, T.lb_max);
        A.max_la = std::max(A.max_la, T.la_max), A.max_lb = std::max(A.max_lb, T.lb_max);
        B200_ASSERT(T.la_max <= kMaxLSide && T.lb_max <= kMaxLSide, "angular momentum beyond kMaxLSide");
        A.combo[T.la_max][T.lb_max] = true;
        A.max_block_size = std::max(A.max_block_size, T.nsgfa * T.nsgfb);
      }
    }
    bool combo[kMaxLSide + 1][kMaxLSide + 1] = {};
    for (const Acc &A : accs) {
      for (int l = 0; l < nlevels; l++) {
        LevelInfo &li = tl.linfo[l];
        const LevelInfo &x = A.li[l];
        if (x.last > 0) {
          li.first = (li.last == 0) ? x.first : std::min(li.first, x.first);
          li.last = std::max(li.last, x.last);
        }
        li.max_lp0 = std::max(li.max_lp0, x.max_lp0), li.max_w = std::max(li.max_w, x.max_w);
        li.max_lp0_general = std::max(li.max_lp0_general, x.max_lp0_general);
      }
      for (int k = 0; k < kNumSizeClasses; k++) {
        TaskList::SizeClass &C = tl.sclass[k];
        C.max_nsgf_set = std::max(C.max_nsgf_set, A.sc[k].max_nsgf_set);
        C.max_ncoset_raw = std::max(C.max_ncoset_raw, A.sc[k].max_ncoset_raw);
        C.max_la = std::max(C.max_la, A.sc[k].max_la), C.max_lb = std::max(C.max_lb, A.sc[k].max_lb);
      }
      tl.max_nsgf_set = std::max(tl.max_nsgf_set, A.max_nsgf_set);
      tl.max_ncoset_raw = std::max(tl.max_ncoset_raw, A.max_ncoset_raw);
      tl.max_la = std::max(tl.max_la, A.max_la), tl.max_lb = std::max(tl.max_lb, A.max_lb);
      tl.max_block_size = std::max(tl.max_block_size, A.max_block_size);
      for (int a = 0; a <= kMaxLSide; a++)
        for (int b2 = 0; b2 <= kMaxLSide; b2++)
          combo[a][b2] = combo[a][b2] || A.combo[a][b2];
    }
    for (int a = 0; a <= kMaxLSide; a++)
      for (int b2 = 0; b2 <= kMaxLSide; b2++)
        if (combo[a][b2])
          tl.l_combos.push_back(std::make_pair(a, b2));
  }
  B200_ASSERT(tl.max_la + 3 <= kMaxLSide && tl.max_lb + 3 <= kMaxLSide,
              "angular momentum beyond what this build supports (kMaxLSide)");
  for (int l = 0; l < nlevels; l++)
    if (tl.linfo[l].last == 0)
      tl.linfo[l].first = 0;

  tick("level ranges, maxima");
  tl.d_tasks.upload(tl.h_tasks, s);
  tick("task records upload");
  std::vector<int> iota(ntasks);
  std::iota(iota.begin(), iota.end(), 0);
  tl.d_iota.upload(iota, s);

  // tasks grouped by matrix block for the hab/forces kernel: sorted by a packed key
  // (block, iset, jset) read from compact arrays, not from the 360-byte records
  std::vector<int> by_block(ntasks);
  std::vector<unsigned long long> bkey(ntasks);
  std::vector<unsigned char> tcls(ntasks);
#pragma omp parallel for schedule(static)
  for (int it = 0; it < ntasks; it++) {
    const TaskDev &T = tl.h_tasks[it];
    B200_ASSERT(T.iset < 1024 && T.jset < 1024, "more than 1024 sets per kind");
    bkey[it] = ((unsigned long long)T.block_num << 20) | ((unsigned long long)T.iset << 10) | (unsigned long long)T.jset;
    tcls[it] = (unsigned char)size_class(T);
    by_block[it] = it;
  }
  __gnu_parallel::stable_sort(by_block.begin(), by_block.end(), [&](int a, int b) { return bkey[a] < bkey[b]; });
  std::vector<int> block_first(nblocks + 1, 0);
  for (int it = 0; it < ntasks; it++)
    block_first[(size_t)(bkey[it] >> 20) + 1]++;
  for (int b = 0; b < nblocks; b++)
    block_first[b + 1] += block_first[b];
  tl.d_block_task_ids.upload(by_block, s);
  tl.d_block_first.upload(block_first, s);
  {  // copy/compute pipeline chunks
    bool monotonic = true;
    for (int b = 1; b < nblocks; b++)
      monotonic = monotonic && (block_offsets[b] > block_offsets[b - 1]);
    const int nchunks = (monotonic && ntasks > 20000) ? 6 : 1;
    tl.chunks.clear();
    for (int c = 0; c < nchunks; c++) {
      const int b0 = (int)((long long)nblocks * c / nchunks);
      const int b1 = (int)((long long)nblocks * (c + 1) / nchunks);
      if (b1 > b0)
        tl.chunks.push_back(TaskList::Chunk{block_first[b0], block_first[b1],
                                            (c == 0) ? (size_t)0 : (size_t)block_offsets[b0], {0, 0, 0, 0}});
    }
    // size-class lists (block order inside a class)
    std::vector<int> chunked, global;
    chunked.reserve(ntasks), global.reserve(ntasks);
    for (auto &ch : tl.chunks)
      for (int k = 0; k < kNumSizeClasses; k++) {
        ch.c0[k] = (int)chunked.size();
        for (int it = ch.t0; it < ch.t1; it++)
          if (tcls[by_block[it]] == k)
            chunked.push_back(by_block[it]);
        ch.c0[k + 1] = (int)chunked.size();
      }
    for (int k = 0; k < kNumSizeClasses; k++) {
      tl.g0[k] = (int)global.size();
      for (int it = 0; it < ntasks; it++)
        if (tcls[by_block[it]] == k)
          global.push_back(by_block[it]);
      tl.g0[k + 1] = (int)global.size();
    }
    B200_ASSERT((int)chunked.size() == ntasks && (int)global.size() == ntasks, "size-class lists incomplete");
    tl.d_class_ids_chunked.upload(chunked, s);
    tl.d_class_ids_global.upload(global, s);
    B200_CHECK(cudaStreamCreateWithFlags(&tl.copy_stream, cudaStreamNonBlocking));
    B200_CHECK(cudaEventCreateWithFlags(&tl.ev_copy_done, cudaEventDisableTiming));
    tl.ev_chunk.resize(tl.chunks.size());
    for (auto &ev : tl.ev_chunk)
      B200_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }

  tick("block order, size classes");
  tl.h_Tptrs.assign((size_t)nlevels * (kMaxLp + 1), nullptr);
  tl.d_Tptrs.upload(tl.h_Tptrs, s);
  tl.d_grids.resize(nlevels);

  // split every level into tiled-path and generic-path tasks
  tl.h_generic_ids.clear();
  tl.generic_first.assign(nlevels + 1, 0);
  BuilderScratch tiled_scratch;
  for (int l = 0; l < nlevels; l++) {
    LevelInfo &li = tl.linfo[l];
    std::vector<int> generic_ids;
    if (tl.path == 0) {
      // work items of at most item_cap visits: about a dozen items per resident CTA (a task makes
      // about 8 visits), capped so that a tile's accumulators are not flushed too often
      const int item_cap = (int)std::min<double>(kCtItemVisits, std::max(256.0, 8.5 * ntasks / (148.0 * 2 * 12)));
      build_ctile_level(li.ctile, l, tl.levels[l], tl.h_tasks, li.first, li.last, generic_ids, item_cap, s);
    }
    else
      build_tiled_level(li.tiled, tl.levels[l], tl.h_tasks, tl.d_tasks.p, li.first, li.last, generic_ids,
                        tiled_scratch, s);
    li.n_generic = (int)generic_ids.size();
    tl.generic_first[l] = (int)tl.h_generic_ids.size();
    tl.h_generic_ids.insert(tl.h_generic_ids.end(), generic_ids.begin(), generic_ids.end());
  }
  tick("tiled levels (pairs)");
  tl.generic_first[nlevels] = (int)tl.h_generic_ids.size();
  tl.d_generic_ids.upload(tl.h_generic_ids, s);
  if (tl.path == 0) {
    std::vector<CtileLevel *> cls;
    for (int l = 0; l < nlevels; l++)
      cls.push_back(&tl.linfo[l].ctile);
    finish_ctile_list(tl.ct, cls, s);
  }
  tl.level_streams.resize(nlevels);
  tl.ev_join.resize(nlevels);
  B200_CHECK(cudaEventCreateWithFlags(&tl.ev_fork, cudaEventDisableTiming));
  for (int l = 0; l < nlevels; l++) {
    B200_CHECK(cudaStreamCreateWithFlags(&tl.level_streams[l], cudaStreamNonBlocking));
    B200_CHECK(cudaEventCreateWithFlags(&tl.ev_join[l], cudaEventDisableTiming));
  }
}

// ---------------------------------------------------------------------------
// buffers at the call boundary
// ---------------------------------------------------------------------------
// A caller's device_buffer is used only if it really is device (or managed) memory: a
// reference build without __OFFLOAD fills offload_buffer.device_buffer from its memory pool
// with a plain host allocation (src/offload/offload_buffer.c:78-83 with
// OFFLOAD_BUFFER_MEMPOOL, src/offload/offload_mempool.c:69-106) -- non-NULL, but not a device
// pointer.  Found by running the reference's own replay harness against this backend.
static inline bool use_caller_device(const grid_b200_buffer *b) {
  if (b == nullptr || b->device_buffer == nullptr)
    return false;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, b->device_buffer) != cudaSuccess) {
    cudaGetLastError();  // clear the sticky-free error of an unknown pointer
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

}  // namespace b200

using namespace b200;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

void grid_b200_set_device(const int device) { g_device = device; }

int grid_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;
  return n;
}

void grid_b200_set_stream(void *cuda_stream) { g_stream = (cudaStream_t)cuda_stream; }
void grid_b200_set_device_resident(const bool flag) { g_device_resident = flag; }
void grid_b200_set_kernel_variant(const int variant) { g_variant = variant; }
long long grid_b200_get_launch_count(void) { return g_launches.load(); }

void grid_b200_set_timing(const bool flag) { g_timing = flag; }

int grid_b200_get_timings(double *out, const int n) {
  for (auto &sp : g_spans) {
    B200_CHECK(cudaEventSynchronize(sp.b));
    float ms = 0.f;
    B200_CHECK(cudaEventElapsedTime(&ms, sp.a, sp.b));
    g_ms[sp.cls] += ms;
    g_nspans[sp.cls] += 1;
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  g_spans.clear();
  const int m = std::min(n / 2, (int)T_NCLASS);
  for (int i = 0; i < m; i++) {
    out[2 * i] = g_ms[i];
    out[2 * i + 1] = g_nspans[i];
    g_ms[i] = 0, g_nspans[i] = 0;
  }
  return m;
}

void grid_b200_create_task_list(
    const bool orthorhombic, const int ntasks, const int nlevels, const int natoms,
    const int nkinds, const int nblocks, const int *block_offsets, const double *atom_positions,
    const int *atom_kinds, const grid_b200_basis_set **basis_sets, const int *level_list,
    const int *iatom_list, const int *jatom_list, const int *iset_list, const int *jset_list,
    const int *ipgf_list, const int *jpgf_list, const int *border_mask_list,
    const int *block_num_list, const double *radius_list, const double *rab_list,
    const int *npts_global, const int *npts_local, const int *shift_local,
    const int *border_width, const double *dh, const double *dh_inv,
    grid_b200_task_list **task_list_out) {
  activate_device();
  TaskList *tl = nullptr;
  if (*task_list_out == nullptr) {
    tl = new TaskList();
    *task_list_out = tl;
  } else {  // reuse the handle, rebuild the content (grid_task_list.c:42-63)
    tl = (TaskList *)*task_list_out;
    tl->release();
    TaskList fresh;
    std::swap(*tl, fresh);
  }
  if (ntasks == 0 || nblocks == 0 || nlevels == 0) {
    tl->empty = true;
    tl->nlevels = nlevels;
    tl->natoms = natoms;
    return;
  }
  tl->empty = false;
  build_task_list(*tl, orthorhombic, ntasks, nlevels, natoms, nkinds, nblocks, block_offsets,
                  atom_positions, atom_kinds, basis_sets, level_list, iatom_list, jatom_list,
                  iset_list, jset_list, ipgf_list, jpgf_list, border_mask_list, block_num_list,
                  radius_list, rab_list, npts_global, npts_local, shift_local, border_width, dh,
                  dh_inv);
}

void grid_b200_release_cache(void) { dev_arena().release(); }

static grid_b200_comm *g_collocate_comm = nullptr;
void grid_b200_set_collocate_reduce(grid_b200_comm *comm) { g_collocate_comm = comm; }

void grid_b200_free_task_list(grid_b200_task_list *ptr) {
  if (ptr == nullptr)
    return;
  activate_device();
  TaskList *tl = (TaskList *)ptr;
  tl->release();
  delete tl;
}

void grid_b200_collocate_task_list(const grid_b200_task_list *ptr, const int func,
                                   const int nlevels, const grid_b200_buffer *pab_blocks,
                                   grid_b200_buffer **grids) {
  if (ptr == nullptr)
    return;
  activate_device();
  TaskList &tl = *(TaskList *)ptr;
  cudaStream_t s = g_stream;
  FuncDesc F;
  if (!describe_func(func, F)) {
    fprintf(stderr, "grid_b200: unknown grid_func %d\n", func);
    abort();
  }
  if (tl.empty) {  // grid_task_list.c:185-189
    for (int l = 0; l < nlevels; l++) {
      if (g_collocate_comm != nullptr && g_device_resident && use_caller_device(grids[l]))
        grid_b200_comm_begin_grid(g_collocate_comm, grids[l]->device_buffer, (void *)s);
      if (g_device_resident && use_caller_device(grids[l]))
        B200_CHECK(cudaMemsetAsync(grids[l]->device_buffer, 0, grids[l]->size, s));
      else
        memset(grids[l]->host_buffer, 0, grids[l]->size);
      if (g_collocate_comm != nullptr) {  // the other ranks' lists need not be empty
        B200_ASSERT(g_device_resident && use_caller_device(grids[l]), "the collocate reduction needs device-resident grids");
        grid_b200_comm_reduce_grid(g_collocate_comm, grids[l]->device_buffer, grids[l]->size / sizeof(double), (void *)s);
      }
    }
    return;
  }
  B200_ASSERT(tl.nlevels == nlevels, "nlevels differs from the task list");
  const int dl = F.dla_max + F.dlb_max;
  tl.last_dl = dl;
  ensure_coef_offsets(tl, dl, s);
  ensure_transforms(tl, dl, s);
  ensure_gather_lists(tl, F.dla_max, F.dlb_max, s);
  tl.d_coef.ensure(tl.coef_total[dl]);

  // density blocks -> coefficients.  With host-authoritative buffers the upload
  // is pipelined chunk-wise against the coefficient kernel.
  CoefLaunch CL;
  CL.tasks = tl.d_tasks.p, CL.task_ids = nullptr, CL.ntasks = tl.ntasks;
  CL.sphi_pool = tl.d_sphi.p, CL.coef_offsets = tl.d_coef_off[dl].p, CL.coef = tl.d_coef.p;
  CL.cijk_T = tl.d_Tptrs.p, CL.glists = tl.d_glists.p, CL.stream = s;
  const double *d_pab = nullptr;
  const bool pab_from_host = !(g_device_resident && use_caller_device(pab_blocks));
  if (!pab_from_host) {
    d_pab = pab_blocks->device_buffer;
    ScopedTimer tm(T_PAB2COEF, s);
    for (int k = 0; k < kNumSizeClasses; k++) {
      const TaskList::SizeClass &C = tl.sclass[k];
      CL.task_ids = tl.d_class_ids_global.p + tl.g0[k];
      CL.ntasks = tl.g0[k + 1] - tl.g0[k];
      launch_pab_to_coef(CL, func, d_pab, C.max_nsgf_set, C.max_ncoset_raw, C.max_la + F.dla_max,
                         C.max_lb + F.dlb_max);
    }
  } else {
    double *dst = use_caller_device(pab_blocks) ? pab_blocks->device_buffer : nullptr;
    if (dst == nullptr) {
      tl.d_pab.ensure(pab_blocks->size / sizeof(double));
      dst = tl.d_pab.p;
    }
    d_pab = dst;
    const size_t total = pab_blocks->size / sizeof(double);
    ScopedTimer tm(T_PAB2COEF, s);  // includes the exposed part of the upload
    B200_CHECK(cudaEventRecord(tl.ev_fork, s));
    B200_CHECK(cudaStreamWaitEvent(tl.copy_stream, tl.ev_fork, 0));
    for (size_t c = 0; c < tl.chunks.size(); c++) {
      const size_t o0 = tl.chunks[c].off0;
      const size_t o1 = (c + 1 < tl.chunks.size()) ? tl.chunks[c + 1].off0 : total;
      if (o1 > o0)
        B200_CHECK(cudaMemcpyAsync(dst + o0, pab_blocks->host_buffer + o0, (o1 - o0) * sizeof(double),
                                   cudaMemcpyHostToDevice, tl.copy_stream));
      B200_CHECK(cudaEventRecord(tl.ev_chunk[c], tl.copy_stream));
      B200_CHECK(cudaStreamWaitEvent(s, tl.ev_chunk[c], 0));
      for (int k = 0; k < kNumSizeClasses; k++) {
        const TaskList::SizeClass &C = tl.sclass[k];
        CL.task_ids = tl.d_class_ids_chunked.p + tl.chunks[c].c0[k];
        CL.ntasks = tl.chunks[c].c0[k + 1] - tl.chunks[c].c0[k];
        launch_pab_to_coef(CL, func, d_pab, C.max_nsgf_set, C.max_ncoset_raw, C.max_la + F.dla_max,
                           C.max_lb + F.dlb_max);
      }
    }
  }

  if (tl.path == 0 && g_variant != 1) {
    // CTA-tile path: one launch per lp class covers all levels (the work items of the
    // small levels fill the tail of the big one); generic leftovers follow per level.
    ScopedTimer *tm_grid = new ScopedTimer(T_COLLOCATE, s);
    std::vector<double *> dg(nlevels);
    std::vector<CtileLevel *> cl(nlevels);
    bool any_host = false;
    for (int l = 0; l < nlevels; l++) {
      const LevelDev &L = tl.levels[l];
      const size_t npts = (size_t)L.npts_local[0] * L.npts_local[1] * L.npts_local[2];
      B200_ASSERT(grids[l]->size >= npts * sizeof(double), "grid buffer smaller than npts_local");
      double *d_grid = use_caller_device(grids[l]) ? grids[l]->device_buffer : nullptr;
      if (d_grid == nullptr) {
        tl.d_grids[l].ensure(npts);
        d_grid = tl.d_grids[l].p;
      }
      dg[l] = d_grid, cl[l] = &tl.linfo[l].ctile;
      if (g_collocate_comm != nullptr)
        grid_b200_comm_begin_grid(g_collocate_comm, d_grid, (void *)s);
      B200_CHECK(cudaMemsetAsync(d_grid, 0, npts * sizeof(double), s));
    }
    CtileCall C;
    C.list = &tl.ct, C.levels = cl.data(), C.level_dev = tl.levels.data(), C.grids = dg.data();
    C.l0 = 0, C.l1 = nlevels, C.dl = dl, C.coef = tl.d_coef.p, C.stream = s;
    const unsigned leftover = launch_ctile<true>(C);
    for (int l = 0; l < nlevels; l++) {
      LevelInfo &li = tl.linfo[l];
      GridLaunch GL;
      GL.tasks = tl.d_tasks.p, GL.level = tl.levels[l], GL.dl = dl;
      GL.coef_offsets = tl.d_coef_off[dl].p, GL.coef = tl.d_coef.p, GL.grid = dg[l];
      GL.max_lp = li.max_lp0 + dl, GL.max_w = li.max_w, GL.stream = s;
      GL.task_ids = tl.d_generic_ids.p + tl.generic_first[l], GL.ntasks = li.n_generic;
      launch_generic(GL, true);
      for (int cls = 0; cls < kNumClasses; cls++)
        if (leftover & (1u << cls)) {
          GL.task_ids = li.ctile.d_class_task_ids[cls], GL.ntasks = li.ctile.class_ntasks[cls];
          launch_generic(GL, true);
        }
    }
    if (g_collocate_comm != nullptr)
      for (int l = 0; l < nlevels; l++) {
        const LevelDev &L = tl.levels[l];
        B200_ASSERT(g_device_resident && use_caller_device(grids[l]), "the collocate reduction needs device-resident grids");
        grid_b200_comm_reduce_grid(g_collocate_comm, dg[l], (size_t)L.npts_local[0] * L.npts_local[1] * L.npts_local[2],
                                   (void *)s);
      }
    delete tm_grid;
    for (int l = 0; l < nlevels; l++) {
      const bool resident = g_device_resident && use_caller_device(grids[l]);
      if (!resident) {
        const LevelDev &L = tl.levels[l];
        const size_t npts = (size_t)L.npts_local[0] * L.npts_local[1] * L.npts_local[2];
        ScopedTimer tm(T_D2H, s);
        B200_CHECK(cudaMemcpyAsync(grids[l]->host_buffer, dg[l], npts * sizeof(double), cudaMemcpyDeviceToHost, s));
        any_host = true;
      }
    }
    if (!g_device_resident || any_host)
      B200_CHECK(cudaStreamSynchronize(s));
    else if (pab_from_host && !tl.chunks.empty())  // the caller may change host P once the call has returned
      B200_CHECK(cudaEventSynchronize(tl.ev_chunk[tl.chunks.size() - 1]));
    return;
  }

  // The levels are independent: fork one stream per level, join afterwards.
  // T_COLLOCATE spans fork..join, i.e. all grid kernels of this call.
  ScopedTimer *tm_grid = new ScopedTimer(T_COLLOCATE, s);
  bool any_host_copy = false;
  B200_CHECK(cudaEventRecord(tl.ev_fork, s));
  for (int l = 0; l < nlevels; l++) {
    const LevelDev &L = tl.levels[l];
    LevelInfo &li = tl.linfo[l];
    const size_t npts = (size_t)L.npts_local[0] * L.npts_local[1] * L.npts_local[2];
    B200_ASSERT(grids[l]->size >= npts * sizeof(double), "grid buffer smaller than npts_local");
    const bool resident = g_device_resident && use_caller_device(grids[l]);
    double *d_grid = use_caller_device(grids[l]) ? grids[l]->device_buffer : nullptr;
    if (d_grid == nullptr) {
      tl.d_grids[l].ensure(npts);
      d_grid = tl.d_grids[l].p;
    }
    cudaStream_t ls = tl.level_streams[l];
    B200_CHECK(cudaStreamWaitEvent(ls, tl.ev_fork, 0));
    if (g_collocate_comm != nullptr)
      grid_b200_comm_begin_grid(g_collocate_comm, d_grid, (void *)ls);
    B200_CHECK(cudaMemsetAsync(d_grid, 0, npts * sizeof(double), ls));
    const bool force_generic = (g_variant == 1);
    GridLaunch GL;
    GL.tasks = tl.d_tasks.p, GL.level = L, GL.dl = dl;
    GL.coef_offsets = tl.d_coef_off[dl].p, GL.coef = tl.d_coef.p, GL.grid = d_grid;
    GL.max_lp = li.max_lp0 + dl, GL.max_w = li.max_w, GL.stream = ls;
    {
      if (force_generic || !tiled_supports(li.tiled, li.max_lp0 + dl)) {
        GL.task_ids = tl.d_iota.p + li.first, GL.ntasks = li.last - li.first;
        launch_generic(GL, true);
      } else {
        const unsigned leftover = launch_tiled<true>(li.tiled, GL);
        GL.task_ids = tl.d_generic_ids.p + tl.generic_first[l], GL.ntasks = li.n_generic;
        launch_generic(GL, true);
        for (int cls = 0; cls < kNumClasses; cls++)
          if (leftover & (1u << cls)) {
            GL.task_ids = li.tiled.d_class_task_ids[cls], GL.ntasks = li.tiled.class_ntasks[cls];
            launch_generic(GL, true);
          }
      }
    }
    if (g_collocate_comm != nullptr) {
      B200_ASSERT(resident, "the collocate reduction needs device-resident grids");
      grid_b200_comm_reduce_grid(g_collocate_comm, d_grid, npts, (void *)ls);
    }
    if (!resident) {
      B200_CHECK(cudaMemcpyAsync(grids[l]->host_buffer, d_grid, npts * sizeof(double),
                                 cudaMemcpyDeviceToHost, ls));
      any_host_copy = true;
    }
    B200_CHECK(cudaEventRecord(tl.ev_join[l], ls));
  }
  for (int l = 0; l < nlevels; l++)
    B200_CHECK(cudaStreamWaitEvent(s, tl.ev_join[l], 0));
  delete tm_grid;
  // host_buffer is the source of truth for every buffer that was not resident: the
  // copies back must have landed when the call returns
  if (!g_device_resident || any_host_copy)
    B200_CHECK(cudaStreamSynchronize(s));
  else if (pab_from_host && !tl.chunks.empty())  // the caller may change host P once the call has returned
    B200_CHECK(cudaEventSynchronize(tl.ev_chunk[tl.chunks.size() - 1]));
}

void grid_b200_integrate_task_list(const grid_b200_task_list *ptr, const bool compute_tau,
                                   const int natoms, const int nlevels,
                                   const grid_b200_buffer *pab_blocks,
                                   const grid_b200_buffer **grids, grid_b200_buffer *hab_blocks,
                                   double *forces, double *virial) {
  if (ptr == nullptr)
    return;
  activate_device();
  TaskList &tl = *(TaskList *)ptr;
  cudaStream_t s = g_stream;
  const bool do_f = (forces != nullptr), do_v = (virial != nullptr);
  if (tl.empty) {  // grid_task_list.c:290-311
    if (g_device_resident && use_caller_device(hab_blocks))
      B200_CHECK(cudaMemsetAsync(hab_blocks->device_buffer, 0, hab_blocks->size, s));
    else
      memset(hab_blocks->host_buffer, 0, hab_blocks->size);
    if (do_f)
      memset(forces, 0, sizeof(double) * 3 * natoms);
    if (do_v)
      memset(virial, 0, sizeof(double) * 9);
    return;
  }
  B200_ASSERT(tl.nlevels == nlevels, "nlevels differs from the task list");
  B200_ASSERT(tl.natoms == natoms, "natoms differs from the task list");
  B200_ASSERT(!do_v || do_f, "virial requires forces (ref/grid_ref_integrate.c:58)");
  B200_ASSERT(!(do_f || do_v) || pab_blocks != nullptr,
              "forces/virial require pab_blocks (grid_task_list.c:321-322)");

  // l growth, common/grid_process_vab.h:222-251
  int dla_max = 0, dla_min = 0, dlb_max = 0, dlb_min = 0;
  if (do_f || do_v)
    dla_max += 1, dla_min -= 1, dlb_min -= 1;
  if (do_v)
    dla_max += 1, dlb_max += 1;
  if (compute_tau)
    dla_max += 1, dlb_max += 1, dla_min -= 1, dlb_min -= 1;
  const int dl = dla_max + dlb_max;
  tl.last_dl = dl;
  ensure_coef_offsets(tl, dl, s);
  ensure_transforms(tl, dl, s);
  ensure_gather_lists(tl, dla_max, dlb_max, s);
  tl.d_coef.ensure(tl.coef_total[dl]);
  {  // the tiled integrate kernel accumulates into the coefficients
    ScopedTimer tm(T_MEMSET, s);
    B200_CHECK(cudaMemsetAsync(tl.d_coef.p, 0, tl.coef_total[dl] * sizeof(double), s));
  }

  if (tl.path == 0 && g_variant != 1) {
    std::vector<double *> dg(nlevels);
    std::vector<CtileLevel *> cl(nlevels);
    for (int l = 0; l < nlevels; l++) {
      const LevelDev &L = tl.levels[l];
      const size_t npts = (size_t)L.npts_local[0] * L.npts_local[1] * L.npts_local[2];
      B200_ASSERT(grids[l]->size >= npts * sizeof(double), "grid buffer smaller than npts_local");
      double *d_grid = use_caller_device(grids[l]) ? grids[l]->device_buffer : nullptr;
      if (!(g_device_resident && d_grid != nullptr)) {
        if (d_grid == nullptr) {
          tl.d_grids[l].ensure(npts);
          d_grid = tl.d_grids[l].p;
        }
        ScopedTimer tm(T_H2D, s);
        B200_CHECK(cudaMemcpyAsync(d_grid, grids[l]->host_buffer, npts * sizeof(double), cudaMemcpyHostToDevice, s));
      }
      dg[l] = d_grid, cl[l] = &tl.linfo[l].ctile;
    }
    ScopedTimer tm_g(T_INTEGRATE, s);
    CtileCall C;
    C.list = &tl.ct, C.levels = cl.data(), C.level_dev = tl.levels.data(), C.grids = dg.data();
    C.l0 = 0, C.l1 = nlevels, C.dl = dl, C.coef = tl.d_coef.p, C.stream = s;
    const unsigned leftover = launch_ctile<false>(C);
    for (int l = 0; l < nlevels; l++) {
      LevelInfo &li = tl.linfo[l];
      GridLaunch GL;
      GL.tasks = tl.d_tasks.p, GL.level = tl.levels[l], GL.dl = dl;
      GL.coef_offsets = tl.d_coef_off[dl].p, GL.coef = tl.d_coef.p, GL.grid = dg[l];
      GL.max_lp = li.max_lp0 + dl, GL.max_w = li.max_w, GL.stream = s;
      GL.task_ids = tl.d_generic_ids.p + tl.generic_first[l], GL.ntasks = li.n_generic;
      launch_generic(GL, false);
      for (int cls = 0; cls < kNumClasses; cls++)
        if (leftover & (1u << cls)) {
          GL.task_ids = li.ctile.d_class_task_ids[cls], GL.ntasks = li.ctile.class_ntasks[cls];
          launch_generic(GL, false);
        }
    }
  } else {
  ScopedTimer *tm_grid = new ScopedTimer(T_INTEGRATE, s);
  B200_CHECK(cudaEventRecord(tl.ev_fork, s));
  // Host-authoritative grids are uploaded coarsest level first: its kernels start after a
  // few microseconds of copy and the big uploads overlap with compute.  (Resident grids:
  // finest level first, so that the short kernels fill the tail.)
  bool any_upload = false;
  for (int l = 0; l < nlevels; l++)
    any_upload = any_upload || !(g_device_resident && use_caller_device(grids[l]));
  for (int k = 0; k < nlevels; k++) {
    const int l = any_upload ? nlevels - 1 - k : k;
    const LevelDev &L = tl.levels[l];
    LevelInfo &li = tl.linfo[l];
    const size_t npts = (size_t)L.npts_local[0] * L.npts_local[1] * L.npts_local[2];
    B200_ASSERT(grids[l]->size >= npts * sizeof(double), "grid buffer smaller than npts_local");
    cudaStream_t ls = tl.level_streams[l];
    B200_CHECK(cudaStreamWaitEvent(ls, tl.ev_fork, 0));
    double *d_grid = use_caller_device(grids[l]) ? grids[l]->device_buffer : nullptr;
    if (!(g_device_resident && d_grid != nullptr)) {
      if (d_grid == nullptr) {
        tl.d_grids[l].ensure(npts);
        d_grid = tl.d_grids[l].p;
      }
      B200_CHECK(cudaMemcpyAsync(d_grid, grids[l]->host_buffer, npts * sizeof(double),
                                 cudaMemcpyHostToDevice, ls));
    }
    const bool force_generic = (g_variant == 1);
    GridLaunch GL;
    GL.tasks = tl.d_tasks.p, GL.level = L, GL.dl = dl;
    GL.coef_offsets = tl.d_coef_off[dl].p, GL.coef = tl.d_coef.p, GL.grid = d_grid;
    GL.max_lp = li.max_lp0 + dl, GL.max_w = li.max_w, GL.stream = ls;
    if (force_generic || !tiled_supports(li.tiled, li.max_lp0 + dl)) {
      GL.task_ids = tl.d_iota.p + li.first, GL.ntasks = li.last - li.first;
      launch_generic(GL, false);
    } else {
      const unsigned leftover = launch_tiled<false>(li.tiled, GL);
      GL.task_ids = tl.d_generic_ids.p + tl.generic_first[l], GL.ntasks = li.n_generic;
      launch_generic(GL, false);
      for (int cls = 0; cls < kNumClasses; cls++)
        if (leftover & (1u << cls)) {
          GL.task_ids = li.tiled.d_class_task_ids[cls], GL.ntasks = li.tiled.class_ntasks[cls];
          launch_generic(GL, false);
        }
    }
    B200_CHECK(cudaEventRecord(tl.ev_join[l], ls));
  }
  for (int l = 0; l < nlevels; l++)
    B200_CHECK(cudaStreamWaitEvent(s, tl.ev_join[l], 0));
  delete tm_grid;
  }

  const double *d_pab = nullptr;
  if (do_f) {
    if (g_device_resident && use_caller_device(pab_blocks)) {
      d_pab = pab_blocks->device_buffer;
    } else {
      double *dst = use_caller_device(pab_blocks) ? pab_blocks->device_buffer : nullptr;
      if (dst == nullptr) {
        tl.d_pab.ensure(pab_blocks->size / sizeof(double));
        dst = tl.d_pab.p;
      }
      ScopedTimer tm(T_H2D, s);
      B200_CHECK(cudaMemcpyAsync(dst, pab_blocks->host_buffer, pab_blocks->size,
                                 cudaMemcpyHostToDevice, s));
      d_pab = dst;
    }
  }
  const bool hab_resident = g_device_resident && use_caller_device(hab_blocks);
  double *d_hab = use_caller_device(hab_blocks) ? hab_blocks->device_buffer : nullptr;
  if (d_hab == nullptr) {
    tl.d_hab.ensure(hab_blocks->size / sizeof(double));
    d_hab = tl.d_hab.p;
  }
  tl.d_fv.ensure((size_t)3 * natoms + 9);
  {
    ScopedTimer tm(T_MEMSET, s);
    B200_CHECK(cudaMemsetAsync(d_hab, 0, hab_blocks->size, s));
    B200_CHECK(cudaMemsetAsync(tl.d_fv.p, 0, ((size_t)3 * natoms + 9) * sizeof(double), s));
  }

  HabLaunch HL;
  HL.tasks = tl.d_tasks.p, HL.block_task_ids = tl.d_block_task_ids.p;
  HL.block_first = tl.d_block_first.p, HL.nblocks = tl.nblocks, HL.sphi_pool = tl.d_sphi.p;
  HL.coef_offsets = tl.d_coef_off[dl].p, HL.coef = tl.d_coef.p, HL.cijk_T = tl.d_Tptrs.p;
  HL.glists = tl.d_glists.p;
  HL.pab = d_pab, HL.hab = d_hab;
  HL.forces = do_f ? tl.d_fv.p : nullptr;
  HL.virial = do_v ? tl.d_fv.p + (size_t)3 * natoms : nullptr;
  HL.compute_tau = compute_tau, HL.maxco = tl.maxco, HL.max_nsgf_set = tl.max_nsgf_set;
  HL.max_la_l = tl.max_la + dla_max, HL.max_lb_l = tl.max_lb + dlb_max, HL.stream = s;
  if (hab_resident) {
    ScopedTimer tm(T_COEF2HAB, s);
    for (int k = 0; k < kNumSizeClasses; k++) {
      const TaskList::SizeClass &C = tl.sclass[k];
      HL.block_task_ids = tl.d_class_ids_global.p + tl.g0[k];
      HL.max_nsgf_set = C.max_nsgf_set, HL.max_la_l = C.max_la + dla_max, HL.max_lb_l = C.max_lb + dlb_max;
      launch_coef_to_hab(HL, tl.g0[k + 1] - tl.g0[k], C.max_ncoset_raw, dla_max, dla_min, dlb_max, dlb_min);
    }
  } else {
    // chunk-wise: the download of finished blocks overlaps the remaining kernels
    const size_t total = hab_blocks->size / sizeof(double);
    ScopedTimer tm(T_COEF2HAB, s);
    for (size_t c = 0; c < tl.chunks.size(); c++) {
      for (int k = 0; k < kNumSizeClasses; k++) {
        const TaskList::SizeClass &C = tl.sclass[k];
        HL.block_task_ids = tl.d_class_ids_chunked.p + tl.chunks[c].c0[k];
        HL.max_nsgf_set = C.max_nsgf_set, HL.max_la_l = C.max_la + dla_max, HL.max_lb_l = C.max_lb + dlb_max;
        launch_coef_to_hab(HL, tl.chunks[c].c0[k + 1] - tl.chunks[c].c0[k], C.max_ncoset_raw, dla_max, dla_min,
                           dlb_max, dlb_min);
      }
      B200_CHECK(cudaEventRecord(tl.ev_chunk[c], s));
      B200_CHECK(cudaStreamWaitEvent(tl.copy_stream, tl.ev_chunk[c], 0));
      const size_t o0 = tl.chunks[c].off0;
      const size_t o1 = (c + 1 < tl.chunks.size()) ? tl.chunks[c + 1].off0 : total;
      if (o1 > o0)
        B200_CHECK(cudaMemcpyAsync(hab_blocks->host_buffer + o0, d_hab + o0, (o1 - o0) * sizeof(double),
                                   cudaMemcpyDeviceToHost, tl.copy_stream));
    }
    B200_CHECK(cudaEventRecord(tl.ev_copy_done, tl.copy_stream));
    B200_CHECK(cudaStreamWaitEvent(s, tl.ev_copy_done, 0));
  }
  if (do_f)
    B200_CHECK(cudaMemcpyAsync(forces, tl.d_fv.p, sizeof(double) * 3 * natoms,
                               cudaMemcpyDeviceToHost, s));
  if (do_v)
    B200_CHECK(cudaMemcpyAsync(virial, tl.d_fv.p + (size_t)3 * natoms, sizeof(double) * 9,
                               cudaMemcpyDeviceToHost, s));
  // (hab not resident: its copy back to host_buffer must have landed on return; grids taken from
  // host_buffer: the caller may change them once the call has returned)
  bool grids_from_host = false;
  for (int l = 0; l < nlevels; l++)
    grids_from_host = grids_from_host || !(g_device_resident && use_caller_device(grids[l]));
  if (!g_device_resident || !hab_resident || grids_from_host || do_f || do_v)
    B200_CHECK(cudaStreamSynchronize(s));
}

// Tasks per lp (la_max + lb_max plus the l growth of the list's last call), orthorhombic and
// general path separately -- what gpu/grid_gpu_context.cu:538-552 forwards to
// grid_library_counter_add after a call (the GRID STATISTICS table of a CP2K run).
void grid_b200_get_task_counts(const grid_b200_task_list *ptr, int ortho[20], int general[20]) {
  for (int i = 0; i < 20; i++)
    ortho[i] = 0, general[i] = 0;
  if (ptr == nullptr)
    return;
  const TaskList &tl = *(const TaskList *)ptr;
  for (const TaskDev &T : tl.h_tasks) {
    if (T.skip)
      continue;
    const int lp = std::min(T.la_max + T.lb_max + tl.last_dl, 19);
    (T.use_ortho ? ortho : general)[lp]++;
  }
}

int grid_b200_get_stats(const grid_b200_task_list *ptr, double *out, const int n) {
  if (ptr == nullptr)
    return 0;
  activate_device();
  TaskList &tl = *(TaskList *)ptr;
  if (!tl.stats_ready && !tl.empty) {
    compute_stats(tl.d_tasks.p, tl.ntasks, tl.levels, tl.h_tasks, tl.stats, g_stream);
    int max_lp = 0, max_w = 0, n_generic = 0;
    double npairs = 0;
    for (int l = 0; l < tl.nlevels; l++) {
      max_lp = std::max(max_lp, tl.linfo[l].max_lp0);
      max_w = std::max(max_w, tl.linfo[l].max_w);
      n_generic += tl.linfo[l].n_generic;
      npairs += (double)(tl.path == 0 ? tl.linfo[l].ctile.nvisits : tl.linfo[l].tiled.npairs);
    }
    tl.stats[0] = tl.ntasks, tl.stats[1] = tl.ntasks - n_generic, tl.stats[2] = n_generic;
    tl.stats[3] = npairs, tl.stats[7] = max_lp, tl.stats[8] = max_w / 2;
    tl.stats[9] = tl.nlevels, tl.stats[10] = tl.nblocks;
    tl.stats_ready = true;
  }
  const int m = std::min(n, 11);
  for (int i = 0; i < m; i++)
    out[i] = tl.stats[i];
  return m;
}

}  // extern "C"
