// Halo exchange of z-slab distributed real-space grids over NCCL -- the exchange step
// either side of the hot path when the grids are sharded over the GPUs of a node
// (SURVEY.md 8(e)).  C-callable, so that the Fortran host that owns
// transfer_rs2pw_distributed / transfer_pw2rs_distributed can use it on device buffers.
//
// Reference behaviour mirrored (paths relative to /root/reference/src/pw):
//   slab descriptor (owned planes + 2 * border)   realspace_grid_types.F:413-415, 514-519
//   halo SUM after collocate                      realspace_grid_types.F:988-1204
//   halo FILL before integrate                    realspace_grid_types.F:1677-1893
//   replicated levels: sum over all ranks         realspace_grid_types.F:763-825
// The reference shifts the halo through the ring of neighbours with MPI sendrecv, n_shifts
// rounds when the halo is wider than a slab.  Here every rank sends each contiguous range
// of halo planes straight to the rank that owns it (z is the slowest index: ranges are
// contiguous), all messages of a level in ONE grouped NCCL call over NVLink, and a single
// kernel adds what arrived into the owned planes.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, reusing the copy the process has
// already loaded -- e.g. PyTorch's): the library has no link-time dependency on it and
// loads on machines without NCCL.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "b200_internal.cuh"

namespace b200 {

// ---- the few NCCL entry points used, resolved lazily ------------------------
struct NcclUniqueId {
  char internal[128];
};
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclDouble = 8, kNcclSum = 0;  // ncclFloat64, ncclSum (nccl.h)

static NcclApi &nccl() {
  static NcclApi api;
  if (api.ok)
    return api;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy already in the process
  if (h == nullptr)
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr)
    h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  B200_ASSERT(h != nullptr, "libnccl.so.2 not found (needed for the halo exchange)");
  auto sym = [&](const char *name) {
    void *p = dlsym(h, name);
    if (p == nullptr)
      b200_fatal("NCCL symbol missing", name, __FILE__, __LINE__);
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.ok = true;
  return api;
}

#define B200_NCCL(cmd)                                                         \
  do {                                                                         \
    const int r_ = (cmd);                                                      \
    if (r_ != 0)                                                               \
      b200_fatal("NCCL error", nccl().GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

// ---- replicated grids in NVLink peer memory ------------------------------------------
// The sum of a replicated level over the ranks WITHOUT a collective kernel: the grid kernels are
// persistent and fill every SM, so an NCCL all-reduce enqueued behind one level cannot start
// before the other levels' kernels have drained (measured: no overlap).  Here every rank's
// grids live in one cudaMalloc'd slab that all ranks of the node map (CUDA IPC); rank r owns
// the r-th chunk of each level and
//   1. pulls that chunk of every peer's grid into a staging buffer   (copy engines, P2P reads)
//   2. adds the staged chunks into its own                           (one light kernel)
//   3. pulls every peer's reduced chunk into its own grid            (copy engines)
// The steps are ordered across ranks by 32-bit step counters in a page of host memory that all
// ranks map (POSIX shared memory, registered with CUDA): a rank publishes "my level is
// collocated / my chunk is reduced / I have gathered" with ONE cuStreamWriteValue32, a stream
// waits for a peer's counter with cuStreamWaitValue32 -- no kernels, no per-peer messages (a
// first version signalled with 4-byte peer copies: 140 tiny copies per call on the copy engines
// cost more than the transfers).  Only step 2 needs an SM slot.
enum PeerFlag : int { PF_DONE = 0, PF_RED = 1, PF_GATH = 2, PF_KINDS = 3 };
typedef int (*StreamValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
struct PeerGrids {
  int nlevels = 0;
  std::vector<size_t> npts, off;       // per level: doubles, offset in the slab (doubles)
  double *base = nullptr;              // this rank's slab: the levels, then the flag block
  std::vector<double *> peer_base;     // [rank] (peer_base[rank] == base)
  unsigned *hflags = nullptr;          // host page shared by the ranks: [rank][kind][level]
  unsigned *dflags = nullptr;          // its device address
  size_t hflags_bytes = 0;
  std::vector<double *> stage;         // per level: (nranks - 1) chunks
  std::vector<cudaStream_t> cs;
  std::vector<cudaEvent_t> ev_in, ev_out;
  // one more stream per (level, peer): the pulls from the peers run side by side on the copy
  // engines (in one stream they are serialised at the rate of a single engine: measured 0.56 ms
  // exposed per call at 8 GPUs)
  std::vector<cudaStream_t> pull;      // [level * (nranks - 1) + k]
  std::vector<cudaEvent_t> ev_pull;    // [(level * (nranks - 1) + k) * 2 + phase]
  std::vector<unsigned> step;          // per level: exchanges enqueued so far
  StreamValue32Fn wait_value = nullptr, write_value = nullptr;
};

struct HaloComm {
  void *comm = nullptr;
  int nranks = 1, rank = 0;
  double *tmp = nullptr;  // receive buffer of the halo sum
  size_t tmp_cap = 0;     // doubles
  cudaStream_t stream = nullptr;
  PeerGrids *peer = nullptr;
  unsigned long long tag = 0;  // hash of the unique id: names the shared flag page
};

// ---- the exchange plan of one level ------------------------------------------
// A message = a contiguous range of the SENDER's halo planes owned by the receiver; on the
// receiver it lands on one or two runs of owned planes (two when the owned range is met
// across the periodic boundary).
struct HaloRun {
  int k, d, m;  // planes [k, k + m) of the message <-> local planes [d, d + m) of the owner
};
struct HaloMsg {
  int src, dst;
  int a, b;  // local halo planes [a, b) on src
  std::vector<HaloRun> runs;
};

static std::vector<HaloMsg> halo_plan(const grid_b200_slab &S) {
  std::vector<HaloMsg> plan;
  const int nz = S.npts_global[2], B = S.border;
  for (int src = 0; src < S.nranks; src++) {
    const int lo_s = S.owned_lo[src], nown = S.owned_hi[src] - lo_s, nloc = nown + 2 * B;
    for (int dst = 0; dst < S.nranks; dst++) {
      if (dst == src)
        continue;
      const int lo_d = S.owned_lo[dst], hi_d = S.owned_hi[dst];
      HaloMsg cur;
      int prev_src = -2, prev_dst = -2;
      bool open = false;
      auto close = [&]() {
        if (open)
          plan.push_back(cur);
        open = false;
      };
      for (int p = 0; p < nloc; p++) {
        if (p >= B && p < B + nown)
          continue;  // owned plane of src
        const int g = ((lo_s - B + p) % nz + nz) % nz;
        if (g < lo_d || g >= hi_d) {
          close();
          prev_src = -2;
          continue;
        }
        const int dl = g - lo_d + B;  // local plane on dst
        if (!open || p != prev_src + 1) {
          close();
          cur = HaloMsg{src, dst, p, p + 1, {HaloRun{0, dl, 1}}};
          open = true;
        } else {
          cur.b = p + 1;
          if (dl == prev_dst + 1)
            cur.runs.back().m++;
          else
            cur.runs.push_back(HaloRun{p - cur.a, dl, 1});
        }
        prev_src = p, prev_dst = dl;
      }
      close();
    }
  }
  return plan;
}

__global__ void halo_add_kernel(double *__restrict__ grid, const double *__restrict__ buf, const size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    grid[i] += buf[i];
}

// 64-thread CTAs with at most 32 registers per thread: 2048 registers, what the persistent grid
// kernels (4 CTAs x 128 threads x 124 registers) leave free on an SM -- the sum does not have to
// wait for one of them to exit.
constexpr int kPeerSumThreads = 64;
__global__ void __launch_bounds__(kPeerSumThreads, 32)
peer_sum_kernel(double *__restrict__ mine, const double *__restrict__ stage, const unsigned chunk_stride, const int npeers,
                const unsigned n) {
  for (unsigned i = blockIdx.x * kPeerSumThreads + threadIdx.x; i < n; i += gridDim.x * kPeerSumThreads) {
    double v = mine[i];
    const double *p = stage + i;
#pragma unroll 1
    for (int k = 0; k < npeers; k++, p += chunk_stride)
      v += *p;
    mine[i] = v;
  }
}

// chunk of level l owned by rank r: [lo, hi) in doubles, boundaries on multiples of 32
static void peer_chunk(const PeerGrids &P, const int nranks, const int l, const int r, size_t &lo, size_t &hi) {
  const size_t n = P.npts[l], per = ((n + nranks - 1) / nranks + 31) / 32 * 32;
  lo = std::min(n, per * (size_t)r), hi = std::min(n, per * (size_t)(r + 1));
}

static int peer_level_of(const HaloComm &H, const double *grid_dev) {
  if (H.peer == nullptr)
    return -1;
  for (int l = 0; l < H.peer->nlevels; l++)
    if (grid_dev == H.peer->base + H.peer->off[l])
      return l;
  return -1;
}

static void check_slab(const HaloComm &H, const grid_b200_slab &S) {
  B200_ASSERT(S.nranks == H.nranks && S.rank == H.rank, "slab descriptor and communicator disagree");
  B200_ASSERT(S.border >= 0 && S.owned_lo != nullptr && S.owned_hi != nullptr, "incomplete slab descriptor");
}

}  // namespace b200

using namespace b200;

extern "C" {

void grid_b200_comm_unique_id(void *out128) {
  NcclUniqueId id;
  B200_NCCL(nccl().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
}

void grid_b200_comm_create(const int nranks, const int rank, const void *unique_id128, void *cuda_stream,
                           grid_b200_comm **comm_out) {
  HaloComm *H = new HaloComm();
  H->nranks = nranks, H->rank = rank, H->stream = (cudaStream_t)cuda_stream;
  NcclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  H->tag = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof(id); i++)
    H->tag = (H->tag ^ (unsigned char)id.internal[i]) * 1099511628211ull;
  B200_NCCL(nccl().CommInitRank(&H->comm, nranks, id, rank));
  *comm_out = (grid_b200_comm *)H;
}

// One slab per rank holding all levels, mapped by every rank of the node.  Returns 0 and this
// rank's per-level device pointers, or non-zero when peer memory is not available (the caller
// then keeps its own buffers and the sums go through NCCL).
int grid_b200_comm_share_grids(grid_b200_comm *comm, const int nlevels, const size_t *npts, double **grids_dev_out) {
  HaloComm &H = *(HaloComm *)comm;
  B200_ASSERT(H.peer == nullptr, "grids are already shared on this communicator");
  int dev = 0, ndev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  B200_CHECK(cudaGetDeviceCount(&ndev));
  PeerGrids *P = new PeerGrids();
  cudaDriverEntryPointQueryResult q1, q2;
  void *f1 = nullptr, *f2 = nullptr;
  const bool have_ops = cudaGetDriverEntryPoint("cuStreamWaitValue32", &f1, cudaEnableDefault, &q1) == cudaSuccess &&
                        cudaGetDriverEntryPoint("cuStreamWriteValue32", &f2, cudaEnableDefault, &q2) == cudaSuccess &&
                        f1 != nullptr && f2 != nullptr;
  // every rank must come to the same decision: ok = min over ranks
  int ok = (have_ops && H.nranks <= ndev) ? 1 : 0;
  P->wait_value = (StreamValue32Fn)f1, P->write_value = (StreamValue32Fn)f2;
  P->nlevels = nlevels;
  size_t total = 0;
  for (int l = 0; l < nlevels; l++) {
    P->npts.push_back(npts[l]);
    P->off.push_back(total);
    total += (npts[l] + 63) / 64 * 64;
  }
  const size_t nflags = (size_t)H.nranks * PF_KINDS * nlevels;
  const size_t slab_bytes = total * sizeof(double);
  // the flag page: rank 0 creates it before the first exchange below, the others open it after
  char shm_name[64];
  snprintf(shm_name, sizeof(shm_name), "/grid_b200_%016llx", H.tag);
  P->hflags_bytes = (nflags * sizeof(unsigned) + 4095) / 4096 * 4096;
  int shm_fd = -1;
  if (H.rank == 0) {
    shm_unlink(shm_name);
    shm_fd = shm_open(shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (shm_fd < 0 || ftruncate(shm_fd, (off_t)P->hflags_bytes) != 0)
      ok = 0;
  }
  if (ok && cudaMalloc((void **)&P->base, slab_bytes) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine, P->base) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  // all-gather {ok, handle} over the communicator
  struct Rec {
    int ok, pad[15];
    cudaIpcMemHandle_t h;
  };
  static_assert(sizeof(Rec) == 128, "IPC record");
  Rec rec;
  memset(&rec, 0, sizeof(rec));
  rec.ok = ok, rec.h = mine;
  Rec *d_recs = nullptr;
  std::vector<Rec> recs(H.nranks);
  B200_CHECK(cudaMalloc((void **)&d_recs, sizeof(Rec) * H.nranks));
  B200_CHECK(cudaMemcpy(d_recs + H.rank, &rec, sizeof(Rec), cudaMemcpyHostToDevice));
  B200_NCCL(nccl().AllGather(d_recs + H.rank, d_recs, sizeof(Rec), /*ncclInt8*/ 0, H.comm, H.stream));
  B200_CHECK(cudaStreamSynchronize(H.stream));
  B200_CHECK(cudaMemcpy(recs.data(), d_recs, sizeof(Rec) * H.nranks, cudaMemcpyDeviceToHost));
  for (const Rec &r : recs)
    ok = std::min(ok, r.ok);
  P->peer_base.assign(H.nranks, nullptr);
  if (ok) {
    if (H.rank != 0)
      shm_fd = shm_open(shm_name, O_RDWR, 0600);
    void *m = (shm_fd >= 0) ? mmap(nullptr, P->hflags_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, shm_fd, 0) : MAP_FAILED;
    if (m == MAP_FAILED) {
      ok = 0;
    } else {
      P->hflags = (unsigned *)m;
      if (H.rank == 0)
        memset(m, 0, P->hflags_bytes);
      if (cudaHostRegister(m, P->hflags_bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess ||
          cudaHostGetDevicePointer((void **)&P->dflags, m, 0) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      }
    }
  }
  if (shm_fd >= 0)
    close(shm_fd);
  if (ok) {
    for (int r = 0; r < H.nranks && ok; r++) {
      if (r == H.rank) {
        P->peer_base[r] = P->base;
      } else if (cudaIpcOpenMemHandle((void **)&P->peer_base[r], recs[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      }
    }
  }
  // the mapping can fail on one rank only: agree once more
  int *d_ok = (int *)d_recs;
  std::vector<int> oks(H.nranks * 32);
  B200_CHECK(cudaMemcpy(d_ok + 32 * H.rank, &ok, sizeof(int), cudaMemcpyHostToDevice));
  B200_NCCL(nccl().AllGather(d_ok + 32 * H.rank, d_ok, 128, 0, H.comm, H.stream));
  B200_CHECK(cudaStreamSynchronize(H.stream));
  B200_CHECK(cudaMemcpy(oks.data(), d_ok, sizeof(int) * 32 * H.nranks, cudaMemcpyDeviceToHost));
  for (int r = 0; r < H.nranks; r++)
    ok = std::min(ok, oks[32 * r]);
  cudaFree(d_recs);
  if (H.rank == 0)
    shm_unlink(shm_name);  // everybody has opened it (or given up)
  if (!ok) {
    if (P->hflags != nullptr) {
      cudaHostUnregister(P->hflags);
      munmap(P->hflags, P->hflags_bytes);
    }
    for (int r = 0; r < H.nranks; r++)
      if (r != H.rank && P->peer_base[r] != nullptr)
        cudaIpcCloseMemHandle(P->peer_base[r]);
    cudaFree(P->base);
    delete P;
    return 1;
  }
  B200_CHECK(cudaMemset(P->base, 0, slab_bytes));
  int prio_least = 0, prio_greatest = 0;
  B200_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  for (int l = 0; l < nlevels; l++) {
    size_t lo, hi;
    peer_chunk(*P, H.nranks, l, 0, lo, hi);
    double *st = nullptr;
    B200_CHECK(cudaMalloc((void **)&st, std::max<size_t>((hi - lo) * (H.nranks - 1), 1) * sizeof(double)));
    P->stage.push_back(st);
    cudaStream_t cs;
    cudaEvent_t a, b;
    B200_CHECK(cudaStreamCreateWithPriority(&cs, cudaStreamNonBlocking, prio_greatest));
    B200_CHECK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    B200_CHECK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    P->cs.push_back(cs), P->ev_in.push_back(a), P->ev_out.push_back(b);
    P->step.push_back(0);
    for (int k = 0; k < H.nranks - 1; k++) {
      cudaStream_t q;
      B200_CHECK(cudaStreamCreateWithPriority(&q, cudaStreamNonBlocking, prio_greatest));
      P->pull.push_back(q);
      for (int ph = 0; ph < 2; ph++) {
        cudaEvent_t e;
        B200_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        P->ev_pull.push_back(e);
      }
    }
    grids_dev_out[l] = P->base + P->off[l];
  }
  B200_CHECK(cudaDeviceSynchronize());
  // nobody may signal into a flag block that is not zeroed yet
  B200_CHECK(cudaMalloc((void **)&d_ok, 128));
  B200_NCCL(nccl().AllReduce(d_ok, d_ok, 1, /*ncclInt32*/ 2, kNcclSum, H.comm, H.stream));
  B200_CHECK(cudaStreamSynchronize(H.stream));
  cudaFree(d_ok);
  H.peer = P;
  return 0;
}

#define B200_DRV(cmd)                                                          \
  do {                                                                         \
    const int r_ = (cmd);                                                      \
    if (r_ != 0)                                                               \
      b200_fatal("CUDA driver error in a stream memory operation", #cmd, __FILE__, __LINE__); \
  } while (0)

// Called before a shared grid is overwritten (the collocate's memset): the peers must have
// pulled this rank's reduced chunk of the previous exchange.
void grid_b200_comm_begin_grid(grid_b200_comm *comm, double *grid_dev, void *cuda_stream) {
  HaloComm &H = *(HaloComm *)comm;
  const int l = peer_level_of(H, grid_dev);
  if (l < 0)
    return;
  PeerGrids &P = *H.peer;
  if (P.step[l] == 0)
    return;
  for (int p = 0; p < H.nranks; p++)
    if (p != H.rank)
      B200_DRV(P.wait_value((cudaStream_t)cuda_stream,
                            (unsigned long long)(P.dflags + ((size_t)p * PF_KINDS + PF_GATH) * P.nlevels + l), P.step[l],
                            /*CU_STREAM_WAIT_VALUE_GEQ*/ 0));
}

// Sum of a replicated grid over the ranks, enqueued behind the work of `cuda_stream`; the stream
// continues when this rank holds the sum.  Shared grids go through peer memory, others through NCCL.
void grid_b200_comm_reduce_grid(grid_b200_comm *comm, double *grid_dev, const size_t count, void *cuda_stream) {
  HaloComm &H = *(HaloComm *)comm;
  const int l = peer_level_of(H, grid_dev);
  if (l < 0 || H.nranks == 1) {
    grid_b200_comm_allreduce(comm, grid_dev, count, cuda_stream);
    return;
  }
  PeerGrids &P = *H.peer;
  B200_ASSERT(count == P.npts[l], "shared grid: size differs from the one it was created with");
  cudaStream_t ps = (cudaStream_t)cuda_stream, cs = P.cs[l];
  const unsigned step = ++P.step[l];
  const int me = H.rank, n = H.nranks;
  auto flag_of = [&](const int src, const int kind) { return P.dflags + ((size_t)src * PF_KINDS + kind) * P.nlevels + l; };
  auto signal = [&](const int kind, cudaStream_t st) {
    B200_DRV(P.write_value(st, (unsigned long long)flag_of(me, kind), step, 0));
  };
  auto wait_for = [&](cudaStream_t st, const int p, const int kind) {
    B200_DRV(P.wait_value(st, (unsigned long long)flag_of(p, kind), step, 0));
  };
  // my level is complete
  signal(PF_DONE, ps);
  B200_CHECK(cudaEventRecord(P.ev_in[l], ps));
  B200_CHECK(cudaStreamWaitEvent(cs, P.ev_in[l], 0));
  // 1. the peers' contributions to my chunk
  size_t lo, hi;
  peer_chunk(P, n, l, me, lo, hi);
  size_t lo0, hi0;
  peer_chunk(P, n, l, 0, lo0, hi0);
  const size_t stride = hi0 - lo0;
  for (int d = 1; d < n; d++) {  // start with the next rank: spreads the reads over the peers
    const int p = (me + d) % n, k = d - 1;
    cudaStream_t q = P.pull[(size_t)l * (n - 1) + k];
    cudaEvent_t e = P.ev_pull[((size_t)l * (n - 1) + k) * 2 + 0];
    B200_CHECK(cudaStreamWaitEvent(q, P.ev_in[l], 0));  // (the staging buffer is free: cs has passed the last sum)
    wait_for(q, p, PF_DONE);
    if (hi > lo)
      B200_CHECK(cudaMemcpyAsync(P.stage[l] + (size_t)k * stride, P.peer_base[p] + P.off[l] + lo, (hi - lo) * sizeof(double),
                                 cudaMemcpyDefault, q));
    B200_CHECK(cudaEventRecord(e, q));
    B200_CHECK(cudaStreamWaitEvent(cs, e, 0));
  }
  // 2. add them
  if (hi > lo) {
    const size_t cnt = hi - lo;
    B200_ASSERT(cnt < ((size_t)1 << 31) && stride < ((size_t)1 << 31), "shared grid too large for the sum kernel");
    peer_sum_kernel<<<(unsigned)std::min<size_t>((cnt + kPeerSumThreads - 1) / kPeerSumThreads, 148 * 8), kPeerSumThreads, 0,
                      cs>>>(P.base + P.off[l] + lo, P.stage[l], (unsigned)stride, n - 1, (unsigned)cnt);
    B200_CHECK(cudaGetLastError());
    count_launch();
  }
  signal(PF_RED, cs);
  // 3. the other chunks, reduced by their owners (a peer's RED also says that it has read my
  //    contribution to its chunk, which the pull overwrites)
  for (int d = 1; d < n; d++) {
    const int p = (me + d) % n, k = d - 1;
    cudaStream_t q = P.pull[(size_t)l * (n - 1) + k];
    cudaEvent_t e = P.ev_pull[((size_t)l * (n - 1) + k) * 2 + 1];
    size_t plo, phi;
    peer_chunk(P, n, l, p, plo, phi);
    wait_for(q, p, PF_RED);
    if (phi > plo)
      B200_CHECK(cudaMemcpyAsync(P.base + P.off[l] + plo, P.peer_base[p] + P.off[l] + plo, (phi - plo) * sizeof(double),
                                 cudaMemcpyDefault, q));
    B200_CHECK(cudaEventRecord(e, q));
    B200_CHECK(cudaStreamWaitEvent(cs, e, 0));
  }
  signal(PF_GATH, cs);
  B200_CHECK(cudaEventRecord(P.ev_out[l], cs));
  B200_CHECK(cudaStreamWaitEvent(ps, P.ev_out[l], 0));
}

void grid_b200_comm_destroy(grid_b200_comm *comm) {
  if (comm == nullptr)
    return;
  HaloComm *H = (HaloComm *)comm;
  if (H->peer != nullptr) {
    PeerGrids *P = H->peer;
    cudaDeviceSynchronize();
    for (int r = 0; r < H->nranks; r++)
      if (r != H->rank)
        cudaIpcCloseMemHandle(P->peer_base[r]);
    for (cudaStream_t q : P->pull)
      cudaStreamDestroy(q);
    for (cudaEvent_t e : P->ev_pull)
      cudaEventDestroy(e);
    for (size_t l = 0; l < P->cs.size(); l++) {
      cudaStreamDestroy(P->cs[l]);
      cudaEventDestroy(P->ev_in[l]), cudaEventDestroy(P->ev_out[l]);
      cudaFree(P->stage[l]);
    }
    cudaHostUnregister(P->hflags);
    munmap(P->hflags, P->hflags_bytes);
    cudaFree(P->base);
    delete P;
  }
  if (H->comm)
    nccl().CommDestroy(H->comm);
  cudaFree(H->tmp);
  delete H;
}

// Sum of a replicated buffer over the ranks, enqueued on the given stream (NCCL orders the
// operations of a communicator in issue order, whichever streams they are issued on).
void grid_b200_comm_allreduce(grid_b200_comm *comm, double *buf_dev, const size_t count, void *cuda_stream) {
  HaloComm &H = *(HaloComm *)comm;
  if (H.nranks > 1 && count > 0)
    B200_NCCL(nccl().AllReduce(buf_dev, buf_dev, count, kNcclDouble, kNcclSum, H.comm, (cudaStream_t)cuda_stream));
}

// The same for several buffers (the levels of a call) as ONE grouped NCCL operation: one kernel
// launch instead of one per level.
void grid_b200_comm_allreduce_levels(grid_b200_comm *comm, const int n, double *const *bufs_dev, const size_t *counts,
                                     void *cuda_stream) {
  HaloComm &H = *(HaloComm *)comm;
  if (H.nranks <= 1)
    return;
  B200_NCCL(nccl().GroupStart());
  for (int i = 0; i < n; i++)
    if (counts[i] > 0)
      B200_NCCL(nccl().AllReduce(bufs_dev[i], bufs_dev[i], counts[i], kNcclDouble, kNcclSum, H.comm, (cudaStream_t)cuda_stream));
  B200_NCCL(nccl().GroupEnd());
}

// Messages this rank takes part in, in plan order; for tests of the plan on a CPU.
// out[i] = {src, dst, a, b, nruns, k0, d0, m0, k1, d1, m1} (11 ints per message).
int grid_b200_halo_plan(const grid_b200_slab *slab, int *out, const int max_msgs) {
  const std::vector<HaloMsg> plan = halo_plan(*slab);
  int n = 0;
  for (const HaloMsg &M : plan) {
    if (M.src != slab->rank && M.dst != slab->rank)
      continue;
    B200_ASSERT(M.runs.size() <= 2, "a halo message meets more than two owned runs");
    if (n < max_msgs) {
      int *o = out + 11 * n;
      o[0] = M.src, o[1] = M.dst, o[2] = M.a, o[3] = M.b, o[4] = (int)M.runs.size();
      for (int r = 0; r < 2; r++) {
        const bool have = r < (int)M.runs.size();
        o[5 + 3 * r] = have ? M.runs[r].k : 0, o[6 + 3 * r] = have ? M.runs[r].d : 0, o[7 + 3 * r] = have ? M.runs[r].m : 0;
      }
    }
    n++;
  }
  return n;
}

// All levels of a call in ONE grouped NCCL operation (the launch latency of a group, not the
// NVLink transfer, is what a level's exchange costs): halo planes to their owners for the
// distributed levels, an all-reduce for the replicated ones; then the adds and the halo reset.
void grid_b200_halo_sum_levels(grid_b200_comm *comm, const int nlevels, const grid_b200_slab *const *slabs,
                               double *const *grids_dev) {
  HaloComm &H = *(HaloComm *)comm;
  cudaStream_t s = H.stream;
  std::vector<std::vector<HaloMsg>> plans(nlevels);
  size_t need = 0;
  for (int l = 0; l < nlevels; l++) {
    const grid_b200_slab &S = *slabs[l];
    check_slab(H, S);
    if (!S.distributed)
      continue;
    plans[l] = halo_plan(S);
    const size_t plane = (size_t)S.npts_global[0] * S.npts_global[1];
    for (const HaloMsg &M : plans[l])
      if (M.dst == S.rank)
        need += (size_t)(M.b - M.a) * plane;
  }
  if (need > H.tmp_cap) {
    cudaFree(H.tmp);
    B200_CHECK(cudaMalloc((void **)&H.tmp, need * sizeof(double)));
    H.tmp_cap = need;
  }
  B200_NCCL(nccl().GroupStart());
  size_t off = 0;
  for (int l = 0; l < nlevels; l++) {
    const grid_b200_slab &S = *slabs[l];
    const size_t plane = (size_t)S.npts_global[0] * S.npts_global[1];
    if (!S.distributed) {  // replicated level: every rank holds the whole grid
      if (H.nranks > 1)
        B200_NCCL(nccl().AllReduce(grids_dev[l], grids_dev[l], plane * S.npts_global[2], kNcclDouble, kNcclSum,
                                   H.comm, s));
      continue;
    }
    for (const HaloMsg &M : plans[l]) {
      const size_t cnt = (size_t)(M.b - M.a) * plane;
      if (M.src == S.rank)
        B200_NCCL(nccl().Send(grids_dev[l] + (size_t)M.a * plane, cnt, kNcclDouble, M.dst, H.comm, s));
      if (M.dst == S.rank) {
        B200_NCCL(nccl().Recv(H.tmp + off, cnt, kNcclDouble, M.src, H.comm, s));
        off += cnt;
      }
    }
  }
  B200_NCCL(nccl().GroupEnd());
  off = 0;
  for (int l = 0; l < nlevels; l++) {
    const grid_b200_slab &S = *slabs[l];
    if (!S.distributed)
      continue;
    const size_t plane = (size_t)S.npts_global[0] * S.npts_global[1];
    for (const HaloMsg &M : plans[l]) {
      if (M.dst != S.rank)
        continue;
      for (const HaloRun &R : M.runs) {
        const size_t n = (size_t)R.m * plane;
        halo_add_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(
            grids_dev[l] + (size_t)R.d * plane, H.tmp + off + (size_t)R.k * plane, n);
        count_launch();
      }
      off += (size_t)(M.b - M.a) * plane;
    }
    // the halo has been handed over: zero it, so that a later sum is idempotent
    const int nown = S.owned_hi[S.rank] - S.owned_lo[S.rank];
    B200_CHECK(cudaMemsetAsync(grids_dev[l], 0, (size_t)S.border * plane * sizeof(double), s));
    B200_CHECK(cudaMemsetAsync(grids_dev[l] + (size_t)(S.border + nown) * plane, 0,
                               (size_t)S.border * plane * sizeof(double), s));
  }
  B200_CHECK(cudaGetLastError());
}

void grid_b200_halo_sum(grid_b200_comm *comm, const grid_b200_slab *slab, double *grid_dev) {
  grid_b200_halo_sum_levels(comm, 1, &slab, &grid_dev);
}

// The halo sum's plan run backwards: owners send their runs, halo holders receive them straight
// into the halo planes -- no staging buffer, no kernel; all levels in one grouped operation.
void grid_b200_halo_fill_levels(grid_b200_comm *comm, const int nlevels, const grid_b200_slab *const *slabs,
                                double *const *grids_dev) {
  HaloComm &H = *(HaloComm *)comm;
  cudaStream_t s = H.stream;
  bool any = false;
  for (int l = 0; l < nlevels; l++) {
    check_slab(H, *slabs[l]);
    any = any || slabs[l]->distributed;
  }
  if (!any)
    return;
  B200_NCCL(nccl().GroupStart());
  for (int l = 0; l < nlevels; l++) {
    const grid_b200_slab &S = *slabs[l];
    if (!S.distributed)
      continue;
    const size_t plane = (size_t)S.npts_global[0] * S.npts_global[1];
    for (const HaloMsg &M : halo_plan(S))
      for (const HaloRun &R : M.runs) {
        const size_t cnt = (size_t)R.m * plane;
        if (M.dst == S.rank)
          B200_NCCL(nccl().Send(grids_dev[l] + (size_t)R.d * plane, cnt, kNcclDouble, M.src, H.comm, s));
        if (M.src == S.rank)
          B200_NCCL(nccl().Recv(grids_dev[l] + (size_t)(M.a + R.k) * plane, cnt, kNcclDouble, M.dst, H.comm, s));
      }
  }
  B200_NCCL(nccl().GroupEnd());
}

void grid_b200_halo_fill(grid_b200_comm *comm, const grid_b200_slab *slab, double *grid_dev) {
  grid_b200_halo_fill_levels(comm, 1, &slab, &grid_dev);
}

}  // extern "C"
