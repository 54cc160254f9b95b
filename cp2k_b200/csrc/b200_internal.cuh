// Internal declarations of the B200 grid backend (sm_100a only).
//
// Data layout in HBM:
//   TaskDev tasks[ntasks]       one record per Gaussian product, sorted by
//                               (level, block, iset, jset) like the reference
//                               (src/grid/ref/grid_ref_task_list.c:26-37,140)
//   double  coef[sum ncoset(lp)] per-task polynomial coefficients C_xyz in coset
//                               order (compact: only lx+ly+lz <= lp)
//   double  sphi_pool[]         all kinds' sphi matrices back to back
//   grids / pab / hab           caller buffers or backend-owned mirrors
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <algorithm>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>
#include <cstdio>
#include <cstdlib>

#include "../../include/grid_b200.h"

// Failures are loud, as in the reference (assert/abort): message to stderr and,
// because test runners capture stderr and lose it on abort(), to a log file.
inline void b200_fatal(const char *what, const char *detail, const char *file, const int line) {
  fprintf(stderr, "grid_b200: %s (%s) at %s:%d\n", what, detail, file, line);
  fflush(stderr);
  const char *path = getenv("GRID_B200_ABORT_LOG");
  FILE *f = fopen(path ? path : "/tmp/grid_b200_abort.log", "a");
  if (f) {
    fprintf(f, "grid_b200: %s (%s) at %s:%d\n", what, detail, file, line);
    fclose(f);
  }
  abort();
}

#define B200_CHECK(cmd)                                                        \
  do {                                                                         \
    cudaError_t e_ = (cmd);                                                    \
    if (e_ != cudaSuccess)                                                     \
      b200_fatal("CUDA error", cudaGetErrorString(e_), __FILE__, __LINE__);    \
  } while (0)

#define B200_ASSERT(cond, msg)                                                 \
  do {                                                                         \
    if (!(cond))                                                               \
      b200_fatal(msg, #cond, __FILE__, __LINE__);                              \
  } while (0)

namespace b200 {

// Device memory of the task lists goes through a small caching arena: cudaMalloc / cudaFree of
// GB-sized buffers cost 50-500 ms each on the boxes we measured (and vary by 5x from call to
// call), and CP2K rebuilds a list of the same size every MD step.  Freed blocks are kept per
// device and handed out again to requests of (nearly) the same size; the cache is dropped when
// an allocation fails, by grid_b200_release_cache(), or beyond GRID_B200_CACHE_MB (default 16384).
struct DevArena {
  struct Block {
    void *p;
    size_t bytes;
    int dev;
  };
  std::mutex mu;
  std::vector<Block> free_blocks;
  std::unordered_map<void *, std::pair<size_t, int>> live;
  size_t cached_bytes = 0, cap_bytes = 0;
  bool cap_read = false;

  static size_t round_up(const size_t bytes) {
    const size_t q = (bytes >= ((size_t)1 << 20)) ? ((size_t)2 << 20) : 512;
    return (std::max<size_t>(bytes, 1) + q - 1) / q * q;
  }
  void drop_cache_locked(const int dev) {
    size_t k = 0;
    for (size_t i = 0; i < free_blocks.size(); i++) {
      if (dev < 0 || free_blocks[i].dev == dev) {
        cudaFree(free_blocks[i].p);
        cached_bytes -= free_blocks[i].bytes;
      } else {
        free_blocks[k++] = free_blocks[i];
      }
    }
    free_blocks.resize(k);
  }
  void *alloc(const size_t want) {
    const size_t bytes = round_up(want);
    int dev = 0;
    B200_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int best = -1;
    for (int i = 0; i < (int)free_blocks.size(); i++) {
      const Block &b = free_blocks[i];
      if (b.dev == dev && b.bytes >= bytes && b.bytes <= bytes + bytes / 8 &&
          (best < 0 || b.bytes < free_blocks[best].bytes))
        best = i;
    }
    void *p = nullptr;
    size_t got = bytes;
    if (best >= 0) {
      p = free_blocks[best].p, got = free_blocks[best].bytes;
      cached_bytes -= got;
      free_blocks[best] = free_blocks.back();
      free_blocks.pop_back();
    } else if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      drop_cache_locked(dev);
      B200_CHECK(cudaMalloc(&p, bytes));
    }
    live[p] = {got, dev};
    return p;
  }
  void free(void *p) {
    if (p == nullptr)
      return;
    // what cudaFree does implicitly: nothing in flight may still use the block when it is reused
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lock(mu);
    if (!cap_read) {
      const char *env = getenv("GRID_B200_CACHE_MB");
      cap_bytes = (size_t)(env ? atoll(env) : 16384) << 20;
      cap_read = true;
    }
    auto it = live.find(p);
    if (it == live.end()) {  // not ours (never happens inside the library)
      cudaFree(p);
      return;
    }
    const Block b{p, it->second.first, it->second.second};
    live.erase(it);
    if (cached_bytes + b.bytes > cap_bytes) {
      cudaFree(p);
    } else {
      free_blocks.push_back(b);
      cached_bytes += b.bytes;
    }
  }
  void release() {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lock(mu);
    drop_cache_locked(-1);
  }
};
inline DevArena &dev_arena() {
  static DevArena *a = new DevArena();  // leaked on purpose: no CUDA calls at process exit
  return *a;
}
template <typename T> inline void dev_alloc(T **p, const size_t bytes) { *p = (T *)dev_arena().alloc(bytes); }
inline void dev_free(void *p) { dev_arena().free(p); }

constexpr int kMaxLSide = 8;   // max angular momentum per side after ldiffs
constexpr int kMaxLp = 16;     // max la+lb after ldiffs
constexpr int kNumOrb = 165;   // ncoset(kMaxLSide)

__host__ __device__ constexpr int ncoset(const int l) {
  return (l < 0) ? 0 : ((l + 1) * (l + 2) * (l + 3)) / 6;
}
__host__ __device__ constexpr int coset(const int lx, const int ly, const int lz) {
  const int l = lx + ly + lz;
  return ncoset(l - 1) + ((l - lx) * (l - lx + 1)) / 2 + lz;
}
__host__ __device__ inline int pmod(const int a, const int m) {
  return ((a % m) + m) % m;
}

// Per grid level (src/grid/ref/grid_ref_task_list_internal.h:38-46).
struct LevelDev {
  int npts_global[3];
  int npts_local[3];
  int shift_local[3];
  int border_width[3];
  double dh[9];      // dh[i*3+j]: j-th Cartesian component of lattice step i
  double dh_inv[9];
};

// One Gaussian product.  Everything that does not depend on `func` is
// precomputed on the host once per task list.
struct TaskDev {
  double ra[3], rab[3], rp[3];
  double zeta, zetb, zetp;
  double prefactor;  // exp(-zeta*zetb/zetp*|rab|^2), without rscale
  double radius;
  // orthorhombic geometry (src/grid/ref/grid_ref_collint.h:222-254)
  double roffset[3];
  double disr_radius;
  int cubecenter[3];
  int lb_cube[3];
  // general geometry (src/grid/ref/grid_ref_collint.h:611-640)
  double gp[3];
  int index_min[3], index_max[3];
  int level, border_mask;
  int la_max, la_min, lb_max, lb_min;
  int iatom, jatom;
  int block_num, block_offset;
  int iset, jset;
  int sgfa, sgfb, nsgf_seta, nsgf_setb, nsgfa, nsgfb;
  int ncoseta, ncosetb;        // ncoset(la_max), ncoset(lb_max)
  int ncoa, ncob;              // npgf * ncoset: Cartesian size of the set
  int o1, o2;                  // (ipgf-1)*ncoseta, (jpgf-1)*ncosetb
  int sphi_a, sphi_b;          // offsets of the kinds' sphi in sphi_pool
  int maxcoa, maxcob;
  int transpose;               // iatom <= jatom
  int use_ortho;               // orthorhombic && border_mask == 0
  int skip;                    // 2*radius < max|dh|  (collint.h:929-937)
};

// Host copy of the task records: a vector that does NOT zero its elements on resize (the
// builder fills 10^6 - 10^7 records of 360 bytes in parallel; a value-initialising resize
// would first touch and zero all of them on one thread).
template <class T> struct DefaultInitAlloc : std::allocator<T> {
  template <class U> struct rebind {
    using other = DefaultInitAlloc<U>;
  };
  template <class U, class... A> void construct(U *p, A &&...a) {
    if constexpr (sizeof...(A) == 0)
      ::new ((void *)p) U;
    else
      ::new ((void *)p) U(std::forward<A>(a)...);
  }
};
using TaskVec = std::vector<TaskDev, DefaultInitAlloc<TaskDev>>;

// (lx,ly,lz) of coset index c, for c < ncoset(kMaxLp) -- filled at load time.
struct OrbTable {
  unsigned char l[816][3];  // ncoset(15) = 816
};

// Index lists of the Cartesian <-> polynomial transform for one (la, lb) after
// ldiffs (ref/grid_ref_collint.h:827-911): every (a, b, k) with k_d <= a_d + b_d,
// once ordered by the polynomial index k (collocate: gather into C_xyz) and once
// by the pair index ib * ncoset(la) + ia (integrate: gather into cab).  An entry
// packs four 16-bit indices: [0] the other side's index (pair index resp. k),
// [1..3] the positions of the three binomial factors in the task's alpha table.
struct GatherList {
  const unsigned long long *by_k;
  const int *kstart;   // [ncoset(lp) + 1]
  const int *kperm;    // outputs by decreasing list length
  const unsigned long long *by_ab;
  const int *abstart;  // [ncoset(la) * ncoset(lb) + 1]
};

// ---- launch bookkeeping ---------------------------------------------------
void count_launch(int n = 1);

// ---- kernels: coefficients (b200_coef.cu) ----------------------------------
struct CoefLaunch {
  const TaskDev *tasks;
  const int *task_ids;     // nullptr = identity
  int ntasks;
  const double *sphi_pool;
  const int *coef_offsets; // per task (indexed by task id)
  double *coef;
  const double *const *cijk_T;  // per (level*(kMaxLp+1)+lp) transform or null
  const GatherList *glists;     // [(kMaxLSide+1)^2], indexed la * (kMaxLSide+1) + lb
  cudaStream_t stream;
};

struct HabLaunch {
  const TaskDev *tasks;
  const int *block_task_ids;    // tasks sorted by (block, iset, jset)
  const int *block_first;       // [nblocks+1] ranges into block_task_ids
  int nblocks;
  const double *sphi_pool;
  const int *coef_offsets;
  const double *coef;
  const double *const *cijk_T;
  const GatherList *glists;
  const double *pab;            // may be null
  double *hab;
  double *forces;               // device [natoms][3] or null
  double *virial;               // device [9] or null
  bool compute_tau;
  int maxco;
  int max_nsgf_set;
  int max_la_l, max_lb_l;       // after process ldiffs
  cudaStream_t stream;
};

// ---- kernels: generic per-task collocate / integrate (b200_generic.cu) ----
struct GridLaunch {
  const TaskDev *tasks;
  const int *task_ids;     // tasks of this level handled by this kernel
  int ntasks;
  LevelDev level;
  int dl;                  // lp growth for this call
  const int *coef_offsets;
  double *coef;            // read (collocate) or written (integrate)
  double *grid;
  int max_lp;              // over these tasks, including dl
  int max_w;               // max cube/box edge over these tasks
  cudaStream_t stream;
};

}  // namespace b200
