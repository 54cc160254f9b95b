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

struct HaloComm {
  void *comm = nullptr;
  int nranks = 1, rank = 0;
  double *tmp = nullptr;  // receive buffer of the halo sum
  size_t tmp_cap = 0;     // doubles
  cudaStream_t stream = nullptr;
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
  B200_NCCL(nccl().CommInitRank(&H->comm, nranks, id, rank));
  *comm_out = (grid_b200_comm *)H;
}

void grid_b200_comm_destroy(grid_b200_comm *comm) {
  if (comm == nullptr)
    return;
  HaloComm *H = (HaloComm *)comm;
  if (H->comm)
    nccl().CommDestroy(H->comm);
  cudaFree(H->tmp);
  delete H;
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
