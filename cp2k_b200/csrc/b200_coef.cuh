// Coefficient kernels of the B200 grid backend:
//   pab_to_coef : density block -> per-task polynomial coefficients  (collocate)
//   coef_to_hab : per-task coefficients -> Hamiltonian block, forces, virial
//
// Reference semantics restated here (paths relative to /root/reference/src/grid):
//   load_pab / store_hab            ref/grid_ref_task_list.c:237-271, 462-499
//   prepare_pab (35 grid_func's)    common/grid_prepare_pab.h:33-430, 447-517
//   cab <-> cxyz re-centring        ref/grid_ref_collint.h:827-911
//   cxyz <-> cijk (general cells)   ref/grid_ref_collint.h:697-762
//   hab / forces / virial           common/grid_process_vab.h:29-251,
//                                   ref/grid_ref_integrate.c:44-152
// Unlike the reference GPU backend (gpu/grid_gpu_collocate.cu:34-118) the
// decontraction is two-staged through shared memory, the density transforms
// are generated from a small operator algebra instead of one routine per
// grid_func, and a whole atom-pair block is accumulated in shared memory and
// flushed once.
#pragma once
#include "b200_internal.cuh"

namespace b200 {

__constant__ OrbTable c_orb;
__constant__ double c_binom[kMaxLSide + 1][kMaxLSide + 1];

struct Orb {
  int l[3];
};
// The orbital table is staged into shared memory by every kernel that decodes
// coset indices with lane-varying arguments: divergent __constant__ reads are
// serialised, shared-memory reads are not.
constexpr int kOrbEntries = 816;  // ncoset(15)
__device__ inline void stage_orb_table(unsigned *s_orb, const int tid, const int nthr) {
  for (int c = tid; c < kOrbEntries; c += nthr)
    s_orb[c] = (unsigned)c_orb.l[c][0] | ((unsigned)c_orb.l[c][1] << 8) | ((unsigned)c_orb.l[c][2] << 16);
}
__device__ inline Orb orb_of(const unsigned *__restrict__ s_orb, const int c) {
  const unsigned v = s_orb[c];
  return Orb{{(int)(v & 255u), (int)((v >> 8) & 255u), (int)(v >> 16)}};
}
__device__ inline int oidx(const Orb &a) { return coset(a.l[0], a.l[1], a.l[2]); }
__device__ inline Orb oup(const int i, Orb a) {
  a.l[i] += 1;
  return a;
}
__device__ inline Orb odn(const int i, Orb a) {
  a.l[i] = max(0, a.l[i] - 1);
  return a;
}

// ---------------------------------------------------------------------------
// Operator algebra for the density transforms.  A primitive Cartesian Gaussian
// g(l) with exponent z obeys  d/dx_i g(l) = l_i g(l-e_i) - 2 z g(l+e_i)  and
// x_j g(l) = g(l+e_j); each grid_func is a short sum of
// (operator on a) x (operator on b).
// ---------------------------------------------------------------------------
enum OpKind : int { OP_ID = 0, OP_D, OP_N, OP_DD, OP_RD, OP_R, OP_CORE };
struct Op {
  int k, i, j;
};
struct FuncTerm {
  double c;
  Op a, b;
};
struct FuncDesc {
  int nterms;
  FuncTerm t[3];
  int dla_max, dla_min, dlb_max, dlb_min;
};

// Host+device so the host can size buffers from the same table.
__host__ __device__ inline bool describe_func(const int func, FuncDesc &F) {
  const Op ID = {OP_ID, 0, 0};
  F.nterms = 0;
  F.dla_max = +1, F.dla_min = -1, F.dlb_max = +1, F.dlb_min = -1;
  auto add = [&](double c, Op a, Op b) { F.t[F.nterms++] = FuncTerm{c, a, b}; };
  if (func == 100) {
    F.dla_max = F.dla_min = F.dlb_max = F.dlb_min = 0;
    add(1.0, ID, ID);
  } else if (func == 200) {
    for (int i = 0; i < 3; i++)
      add(0.5, Op{OP_D, i, 0}, Op{OP_D, i, 0});
  } else if (func >= 301 && func <= 303) {
    const int i = func - 301;
    add(+1.0, ID, Op{OP_D, i, 0});
    add(-1.0, Op{OP_D, i, 0}, ID);
  } else if (func >= 411 && func <= 433) {
    const int i = (func - 400) / 10 - 1, j = (func - 400) % 10 - 1;
    if (i < 0 || i > 2 || j < 0 || j > 2)
      return false;
    F.dlb_max = +2;
    add(+1.0, ID, Op{OP_RD, i, j});
    add(-1.0, Op{OP_D, i, 0}, Op{OP_R, 0, j});
  } else if (func >= 501 && func <= 503) {
    const int i = func - 501;
    add(1.0, ID, Op{OP_D, i, 0});
    add(1.0, Op{OP_D, i, 0}, ID);
  } else if (func >= 601 && func <= 603) {
    const int i = func - 601;
    add(1.0, Op{OP_D, i, 0}, Op{OP_D, i, 0});
  } else if (func >= 701 && func <= 703) {
    const int i = func - 701, j = (i + 1) % 3;
    F.dla_max = +2, F.dla_min = -2, F.dlb_max = +2, F.dlb_min = -2;
    add(1.0, Op{OP_DD, i, j}, Op{OP_DD, i, j});
  } else if (func >= 801 && func <= 803) {
    const int i = func - 801;
    F.dla_max = +2, F.dla_min = -2, F.dlb_max = +2, F.dlb_min = -2;
    add(1.0, Op{OP_DD, i, i}, Op{OP_DD, i, i});
  } else if (func >= 901 && func <= 903) {
    add(1.0, Op{OP_N, func - 901, 0}, ID);
  } else if (func >= 904 && func <= 906) {
    add(1.0, ID, Op{OP_N, func - 904, 0});
  } else if (func >= 1001 && func <= 1003) {
    add(1.0, Op{OP_CORE, func - 1001, 0}, ID);
  } else {
    return false;
  }
  return true;
}

struct OpTerm {
  Orb o;
  double c;
};
__device__ inline int op_expand(const Op op, const Orb l, const double z,
                                OpTerm out[4]) {
  const int i = op.i, j = op.j;
  switch (op.k) {
  case OP_ID:
    out[0] = {l, 1.0};
    return 1;
  case OP_D:
    out[0] = {odn(i, l), (double)l.l[i]};
    out[1] = {oup(i, l), -2.0 * z};
    return 2;
  case OP_N:
    out[0] = {odn(i, l), -(double)l.l[i]};
    out[1] = {oup(i, l), 2.0 * z};
    return 2;
  case OP_DD:
    if (i != j) {
      out[0] = {odn(i, odn(j, l)), (double)(l.l[i] * l.l[j])};
      out[1] = {oup(i, odn(j, l)), -2.0 * z * l.l[j]};
      out[2] = {odn(i, oup(j, l)), -2.0 * z * l.l[i]};
      out[3] = {oup(i, oup(j, l)), 4.0 * z * z};
      return 4;
    }
    out[0] = {odn(i, odn(i, l)), (double)(l.l[i] * (l.l[i] - 1))};
    out[1] = {l, -2.0 * z * (2 * l.l[i] + 1)};
    out[2] = {oup(i, oup(i, l)), 4.0 * z * z};
    return 3;
  case OP_RD:  // x_j d/dx_i : raise j first, then lower i (clamped)
    out[0] = {odn(i, oup(j, l)), (double)l.l[i]};
    out[1] = {oup(i, oup(j, l)), -2.0 * z};
    return 2;
  case OP_R:
    out[0] = {oup(j, l), 1.0};
    return 1;
  default:  // OP_CORE
    out[0] = {oup(i, l), 2.0 * z};
    return 1;
  }
}

// ---------------------------------------------------------------------------
// Shared building blocks (executed by `nthr` cooperating threads, rank `t`).
// ---------------------------------------------------------------------------
// The task record (360 bytes, read all over the kernels) of the NEXT task is
// fetched into registers while the current task is processed and published
// to shared memory at the top of the next iteration: the HBM latency of the
// record (and of the task-id indirection) is off the critical path.
constexpr int kTaskWords = (int)(sizeof(TaskDev) / 4);
constexpr int kTaskDoubles = (int)((sizeof(TaskDev) + 7) / 8);
static_assert(sizeof(TaskDev) % 4 == 0, "TaskDev is copied word-wise");
template <int G> struct TaskPrefetch {
  static constexpr int NW = (kTaskWords + G - 1) / G;
  unsigned w[NW];
  __device__ __forceinline__ void issue(const TaskDev *src, const int lane) {
    const unsigned *p = reinterpret_cast<const unsigned *>(src);
#pragma unroll
    for (int k = 0; k < NW; k++)
      w[k] = (lane + k * G < kTaskWords) ? __ldg(p + lane + k * G) : 0u;
  }
  __device__ __forceinline__ void commit(TaskDev *dst, const int lane) const {
    unsigned *p = reinterpret_cast<unsigned *>(dst);
#pragma unroll
    for (int k = 0; k < NW; k++)
      if (lane + k * G < kTaskWords)
        p[lane + k * G] = w[k];
  }
};
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ inline double block_elem(const TaskDev &T, const double *block,
                                    const int sa, const int sb) {
  return T.transpose ? block[(T.sgfb + sb) * T.nsgfa + T.sgfa + sa]
                     : block[(T.sgfa + sa) * T.nsgfb + T.sgfb + sb];
}
__device__ inline int block_index(const TaskDev &T, const int sa, const int sb) {
  return T.transpose ? (T.sgfb + sb) * T.nsgfa + T.sgfa + sa
                     : (T.sgfa + sa) * T.nsgfb + T.sgfb + sb;
}

template <typename SyncF>
__device__ inline void decontract_task(const TaskDev &T, const double *block,
                                       const double *sphi_pool, double *s_work,
                                       double *s_raw, const int t,
                                       const int nthr, SyncF sync) {
  // raw[jco][ico] = sum_{sa,sb} sphi_b[sb][o2+jco] P(sa,sb) sphi_a[sa][o1+ico]
  const int na = T.ncoseta, nb = T.ncosetb;
  const int nsa = T.nsgf_seta, nsb = T.nsgf_setb, maxcoa = T.maxcoa, maxcob = T.maxcob;
  const double *sphi_a = sphi_pool + T.sphi_a + T.sgfa * maxcoa + T.o1;
  const double *sphi_b = sphi_pool + T.sphi_b + T.sgfb * maxcob + T.o2;
  // element (sa, sb) of the sub-block = blk0[sa * str_a + sb * str_b]
  const int str_a = T.transpose ? 1 : T.nsgfb, str_b = T.transpose ? T.nsgfa : 1;
  const double *blk0 = block + (T.transpose ? T.sgfb * T.nsgfa + T.sgfa : T.sgfa * T.nsgfb + T.sgfb);
  for (int q = t; q < nsb * na; q += nthr) {
    const int sb = q / na, ico = q - sb * na;
    const double *pb = blk0 + sb * str_b, *ps = sphi_a + ico;
    double acc = 0.0;
#pragma unroll 4
    for (int sa = 0; sa < nsa; sa++, pb += str_a, ps += maxcoa)
      acc += __ldg(pb) * __ldg(ps);
    s_work[q] = acc;
  }
  sync();
  for (int q = t; q < nb * na; q += nthr) {
    const int jco = q / na, ico = q - jco * na;
    const double *ps = sphi_b + jco, *pw = s_work + ico;
    double acc = 0.0;
#pragma unroll 4
    for (int sb = 0; sb < nsb; sb++, ps += maxcob, pw += na)
      acc += __ldg(ps) * *pw;
    s_raw[q] = acc;
  }
  sync();
}

// alpha[d][la][lb][k] for (x-a)^la (x-b)^lb = sum_k alpha_k (x-p)^k
template <typename SyncF>
__device__ inline void make_alpha(const TaskDev &T, const int la_c,
                                  const int lb_c, double *s_alpha, const int t,
                                  const int nthr, SyncF sync) {
  // One lane per (d, lb): the rows la = 0..la_c follow from each other by one
  // multiplication with (x - a) = (x - p) + pa.
  const int lp1 = la_c + lb_c + 1;
  const int n = 3 * (lb_c + 1);
  for (int q = t; q < n; q += nthr) {
    const int lb = q % (lb_c + 1), d = q / (lb_c + 1);
    const double pa = T.rp[d] - T.ra[d];
    const double pb = T.rp[d] - (T.ra[d] + T.rab[d]);
    // la = 0: (x - b)^lb = sum_k binom(lb, k) pb^(lb-k) (x-p)^k, by repeated multiplication
    double *al = s_alpha + ((d * (la_c + 1) + 0) * (lb_c + 1) + lb) * lp1;
    al[0] = 1.0;
    for (int k = 1; k < lp1; k++)
      al[k] = 0.0;
    for (int m = 1; m <= lb; m++)
      for (int k = m; k >= 0; k--)
        al[k] = ((k > 0) ? al[k - 1] : 0.0) + pb * al[k];
    for (int la = 1; la <= la_c; la++) {
      double *nx = s_alpha + ((d * (la_c + 1) + la) * (lb_c + 1) + lb) * lp1;
      const int top = la + lb;
      for (int k = lp1 - 1; k > top; k--)
        nx[k] = 0.0;
      for (int k = top; k >= 0; k--)
        nx[k] = ((k > 0) ? al[k - 1] : 0.0) + pa * al[k];
      al = nx;
    }
  }
  sync();
}
#define B200_AL(d, la, lb, k)                                                  \
  s_alpha[((((d) * (la_c + 1) + (la)) * (lb_c + 1) + (lb)) * lp1) + (k)]

// ---------------------------------------------------------------------------
// pab_to_coef: one warp per task.
// ---------------------------------------------------------------------------
struct CoefDims {
  int work, raw, cab, alpha, cxyz;  // doubles per warp
  __host__ __device__ int total() const { return work + raw + cab + alpha + cxyz + kTaskDoubles; }
};

// Tasks are tiny: a group of kCoefGroup lanes (half a warp) handles one task, so
// a warp works on two tasks at once and keeps its lanes busy.
template <int G>
__global__ void __launch_bounds__(128)
pab_to_coef_kernel(const CoefLaunch L, const int func, const double *pab,
                   const CoefDims D) {
  extern __shared__ double smem[];
  __shared__ unsigned s_orb[kOrbEntries];
  stage_orb_table(s_orb, threadIdx.x, blockDim.x);
  __syncthreads();
  const int lane = threadIdx.x & (G - 1);            // rank within the group
  const int warp = threadIdx.x / G;                   // group index within the CTA
  const int wpc = blockDim.x / G;                     // groups per CTA
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
  double *s_work = smem + (size_t)warp * D.total();
  double *s_raw = s_work + D.work;
  double *s_cab = s_raw + D.raw;
  double *s_alpha = s_cab + D.cab;
  double *s_cxyz = s_alpha + D.alpha;
  TaskDev *s_task = reinterpret_cast<TaskDev *>(s_cxyz + D.cxyz);
  auto sync = [gmask] { __syncwarp(gmask); };

  FuncDesc F;
  describe_func(func, F);

  TaskPrefetch<G> pf;
  const int it0 = blockIdx.x * wpc + warp, stride = gridDim.x * wpc;
  int id_next = 0;
  if (it0 < L.ntasks) {
    id_next = L.task_ids ? L.task_ids[it0] : it0;
    pf.issue(L.tasks + id_next, lane);
  }
  for (int it = it0; it < L.ntasks; it += stride) {
    const int itask = id_next;
    __syncwarp(gmask);  // the previous task's reads of the record are complete
    pf.commit(s_task, lane);
    __syncwarp(gmask);
    if (it + stride < L.ntasks) {
      id_next = L.task_ids ? L.task_ids[it + stride] : it + stride;
      pf.issue(L.tasks + id_next, lane);
    }
    const TaskDev &T = *s_task;
    const int la_c = T.la_max + F.dla_max, lb_c = T.lb_max + F.dlb_max;
    const int la_min_c = max(T.la_min + F.dla_min, 0);
    const int lb_min_c = max(T.lb_min + F.dlb_min, 0);
    const int lp = la_c + lb_c, nc = ncoset(lp);
    double *out = L.coef + L.coef_offsets[itask];
    if (T.skip) {
      for (int c = lane; c < nc; c += G)
        out[c] = 0.0;
      continue;
    }
    decontract_task(T, pab + T.block_offset, L.sphi_pool, s_work, s_raw, lane,
                    G, sync);

    // prepare: cab[idx(b')][idx(a')] += coef * raw[idx(b)][idx(a)]
    const int n1c = ncoset(la_c), n2c = ncoset(lb_c);
    for (int q = lane; q < n1c * n2c; q += G)
      s_cab[q] = 0.0;
    __syncwarp(gmask);
    const int a_lo = ncoset(T.la_min - 1), a_hi = ncoset(T.la_max);
    const int b_lo = ncoset(T.lb_min - 1), b_hi = ncoset(T.lb_max);
    const int npa = a_hi - a_lo, npb = b_hi - b_lo;
    for (int q = lane; q < npa * npb; q += G) {
      const int ia = a_lo + q % npa, ib = b_lo + q / npa;
      const double p = s_raw[ib * T.ncoseta + ia];
      if (func == 100) {
        s_cab[ib * n1c + ia] = p;  // identity: distinct targets
      } else {
        const Orb a = orb_of(s_orb, ia), b = orb_of(s_orb, ib);
        for (int tt = 0; tt < F.nterms; tt++) {
          OpTerm ta[4], tb[4];
          const int na_t = op_expand(F.t[tt].a, a, T.zeta, ta);
          const int nb_t = op_expand(F.t[tt].b, b, T.zetb, tb);
          for (int x = 0; x < na_t; x++)
            for (int y = 0; y < nb_t; y++)
              atomicAdd(&s_cab[oidx(tb[y].o) * n1c + oidx(ta[x].o)],
                        F.t[tt].c * ta[x].c * tb[y].c * p);
        }
      }
    }
    __syncwarp(gmask);

    make_alpha(T, la_c, lb_c, s_alpha, lane, G, sync);

    // gather: cxyz[k] = prefactor * sum_{a,b} cab[b][a] ax ay az
    const double rscale = (T.iatom == T.jatom) ? 1.0 : 2.0;
    const double pref = rscale * T.prefactor;
    const int ca_lo = ncoset(la_min_c - 1), cb_lo = ncoset(lb_min_c - 1);
    const bool to_cijk = !T.use_ortho;
    // Every output coefficient k walks the precomputed list of (a, b) pairs that
    // reach it (b200_internal.cuh: GatherList) -- no tests, no index arithmetic.
    // Outputs are dealt to the lanes longest list first.
    {
      const GatherList GLs = L.glists[la_c * (kMaxLSide + 1) + lb_c];
      (void)ca_lo, (void)cb_lo;  // entries below the minimum l are zero in s_cab
      for (int cc = lane; cc < nc; cc += G) {
        const int c = GLs.kperm[cc];
        double acc = 0.0;
        const int e1 = GLs.kstart[c + 1];
        for (int e = GLs.kstart[c]; e < e1; e++) {
          const unsigned long long w = GLs.by_k[e];
          acc += s_cab[(unsigned)(w & 0xffffu)] *
                 (s_alpha[(unsigned)((w >> 16) & 0xffffu)] * s_alpha[(unsigned)((w >> 32) & 0xffffu)] *
                  s_alpha[(unsigned)(w >> 48)]);
        }
        acc *= pref;
        if (to_cijk)
          s_cxyz[c] = acc;
        else
          out[c] = acc;
      }
    }
    if (to_cijk) {  // lattice-polynomial basis for the general path
      __syncwarp(gmask);
      const double *Tm = L.cijk_T[T.level * (kMaxLp + 1) + lp];
      for (int q = lane; q < nc; q += G) {
        double acc = 0.0;
        for (int c = 0; c < nc; c++)
          acc += __ldg(&Tm[q * nc + c]) * s_cxyz[c];
        out[q] = acc;
      }
    }
    __syncwarp(gmask);
  }
}

// ---------------------------------------------------------------------------
// coef_to_hab: one CTA per atom-pair block.
// ---------------------------------------------------------------------------
struct PCtx {
  const double *cab;
  int m1;
  double zeta, zetb;
  double rab[3];
};
__device__ inline double cabv(const PCtx &p, const Orb &a, const Orb &b) {
  return p.cab[oidx(b) * p.m1 + oidx(a)];
}
// what: 0 hab, 1 force a (i), 2 force b (i), 3 virial a (i,j), 4 virial b (i,j)
__device__ inline double vab_plain(const PCtx &p, const int what, const int i,
                                   const int j, const Orb &a, const Orb &b) {
  switch (what) {
  case 0:
    return cabv(p, a, b);
  case 1:
    return 2.0 * p.zeta * cabv(p, oup(i, a), b) - a.l[i] * cabv(p, odn(i, a), b);
  case 2:
    return 2.0 * p.zetb * (cabv(p, oup(i, a), b) - p.rab[i] * cabv(p, a, b)) -
           b.l[i] * cabv(p, a, odn(i, b));
  case 3:
    return 2.0 * p.zeta * cabv(p, oup(i, oup(j, a)), b) -
           a.l[j] * cabv(p, oup(i, odn(j, a)), b);
  default:
    return 2.0 * p.zetb *
               (cabv(p, oup(i, oup(j, a)), b) -
                cabv(p, oup(i, a), b) * p.rab[j] -
                cabv(p, oup(j, a), b) * p.rab[i] +
                cabv(p, a, b) * p.rab[j] * p.rab[i]) -
           b.l[j] * cabv(p, a, oup(i, odn(j, b)));
  }
}
__device__ inline double vab_elem(const PCtx &p, const bool tau, const int what,
                                  const int i, const int j, const Orb &a,
                                  const Orb &b) {
  if (!tau)
    return vab_plain(p, what, i, j, a, b);
  double s = 0.0;
  for (int k = 0; k < 3; k++) {
    s += 0.5 * a.l[k] * b.l[k] * vab_plain(p, what, i, j, odn(k, a), odn(k, b));
    s -= p.zeta * b.l[k] * vab_plain(p, what, i, j, oup(k, a), odn(k, b));
    s -= a.l[k] * p.zetb * vab_plain(p, what, i, j, odn(k, a), oup(k, b));
    s += 2.0 * p.zeta * p.zetb * vab_plain(p, what, i, j, oup(k, a), oup(k, b));
  }
  return s;
}

struct HabDims {
  int work, raw, cab, alpha, cxyz, h, g;  // doubles per warp
  __host__ __device__ int total() const { return work + raw + cab + alpha + 2 * cxyz + h + g + kTaskDoubles; }
};

// Forces and virial are sums over matrix elements of  P(a,b) * d(vab)(a,b).  With tau, vab is
// itself a 12-term stencil over the plain elements (common/grid_process_vab.h:186-251): instead
// of expanding the stencil inside each of the 24 derivatives (~950 cab look-ups per element) the
// density is pushed through the ADJOINT stencil once (the integrate-side twin of prepare_pab's
// tau transform, common/grid_prepare_pab.h:76-130), after which forces and virial are plain sums
// over the (l+1)-grown element range: ~80 look-ups per element.
__device__ inline void accumulate_fv(const PCtx &P, const double pv, const Orb &a, const Orb &b,
                                     const bool do_v, double facc[15]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    facc[i] += pv * vab_plain(P, 1, i, 0, a, b);
    facc[3 + i] += pv * vab_plain(P, 2, i, 0, a, b);
  }
  if (do_v)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        facc[6 + 3 * i + j] += pv * (vab_plain(P, 3, i, j, a, b) + vab_plain(P, 4, i, j, a, b));
}

// One warp per task (tasks visited in block order for locality).  The spherical
// block receives the task's contribution through FP64 atomics: tasks of one
// block are few (~8 for water) and the atomics are spread over the whole block.
template <int G>
__global__ void __launch_bounds__(128, 4)
coef_to_hab_kernel(const HabLaunch L, const HabDims D, const int ntasks, const int dla_max,
                   const int dla_min, const int dlb_max, const int dlb_min) {
  extern __shared__ double smem[];
  __shared__ unsigned s_orb[kOrbEntries];
  stage_orb_table(s_orb, threadIdx.x, blockDim.x);
  __syncthreads();
  const int lane = threadIdx.x & (G - 1), warp = threadIdx.x / G, wpc = blockDim.x / G;
  // G = 16, 32: a (half-)warp per task; G = 128: the whole CTA on one task (large l: the
  // per-task scratch allows few tasks per SM, so more threads must share one)
  const unsigned gmask = (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
  double *s_work = smem + (size_t)warp * D.total();
  double *s_raw = s_work + D.work;
  double *s_cab = s_raw + D.raw;
  double *s_alpha = s_cab + D.cab;
  double *s_cxyz = s_alpha + D.alpha;
  double *s_cijk = s_cxyz + D.cxyz;
  double *s_h = s_cijk + D.cxyz;
  double *s_g = s_h + D.h;
  TaskDev *s_task = reinterpret_cast<TaskDev *>(s_g + D.g);
  auto sync = [gmask] {
    if (G > 32)
      __syncthreads();
    else
      __syncwarp(gmask);
  };
  const bool do_f = (L.forces != nullptr), do_v = (L.virial != nullptr);

  TaskPrefetch<G> pf;
  const int it0 = blockIdx.x * wpc + warp, stride = gridDim.x * wpc;
  int id_next = 0;
  if (it0 < ntasks) {
    id_next = L.block_task_ids[it0];
    pf.issue(L.tasks + id_next, lane);
  }
  for (int it = it0; it < ntasks; it += stride) {
    const int itask = id_next;
    sync();  // the previous task's reads of the record are complete
    pf.commit(s_task, lane);
    sync();
    if (it + stride < ntasks) {
      id_next = L.block_task_ids[it + stride];
      pf.issue(L.tasks + id_next, lane);
      // its coefficients: up to G cache lines from the start of its slot
      if (lane < 32)
        prefetch_l2(L.coef + L.coef_offsets[id_next] + 16 * lane);
    }
    const TaskDev &T = *s_task;
    if (T.skip)
      continue;
    const int la_c = T.la_max + dla_max, lb_c = T.lb_max + dlb_max;
    const int la_min_c = max(T.la_min + dla_min, 0);
    const int lb_min_c = max(T.lb_min + dlb_min, 0);
    const int lp = la_c + lb_c, nc = ncoset(lp);
    const double *in = L.coef + L.coef_offsets[itask];

    // (1) coefficients, back to the Cartesian polynomial basis if needed
    if (T.use_ortho) {
      for (int c = lane; c < nc; c += G)
        s_cxyz[c] = in[c];
    } else {
      for (int c = lane; c < nc; c += G)
        s_cijk[c] = in[c];
      sync();
      const double *Tm = L.cijk_T[T.level * (kMaxLp + 1) + lp];
      for (int c = lane; c < nc; c += G) {
        double acc = 0.0;
        for (int q = 0; q < nc; q++)
          acc += __ldg(&Tm[q * nc + c]) * s_cijk[q];
        s_cxyz[c] = acc;
      }
    }
    make_alpha(T, la_c, lb_c, s_alpha, lane, G, sync);  // syncs

    // (2) cab[b][a] = prefactor * sum_k cxyz[k] ax ay az
    const int n1c = ncoset(la_c), n2c = ncoset(lb_c);
    const int ca_lo = ncoset(la_min_c - 1), cb_lo = ncoset(lb_min_c - 1);
    {
      const GatherList GLs = L.glists[la_c * (kMaxLSide + 1) + lb_c];
      for (int q = lane; q < n1c * n2c; q += G) {
        const int ia = q % n1c, ib = q / n1c;
        double acc = 0.0;
        if (ia >= ca_lo && ib >= cb_lo) {
          const int e1 = GLs.abstart[q + 1];
          for (int e = GLs.abstart[q]; e < e1; e++) {
            const unsigned long long w = GLs.by_ab[e];
            acc += s_cxyz[(unsigned)(w & 0xffffu)] *
                   (s_alpha[(unsigned)((w >> 16) & 0xffffu)] * s_alpha[(unsigned)((w >> 32) & 0xffffu)] *
                    s_alpha[(unsigned)(w >> 48)]);
          }
          acc *= T.prefactor;
        }
        s_cab[q] = acc;
      }
    }
    // (3) density sub-block for forces / virial
    if (do_f || do_v)
      decontract_task(T, L.pab + T.block_offset, L.sphi_pool, s_work, s_raw, lane, G, sync);
    sync();

    // (4) matrix elements for the original l-range
    PCtx P;
    P.cab = s_cab, P.m1 = n1c, P.zeta = T.zeta, P.zetb = T.zetb;
    P.rab[0] = T.rab[0], P.rab[1] = T.rab[1], P.rab[2] = T.rab[2];
    const int na = T.ncoseta, nb = T.ncosetb;
    const int a_lo = ncoset(T.la_min - 1), b_lo = ncoset(T.lb_min - 1);
    double facc[15];  // force a (3), force b (3), virial a+b (9)
#pragma unroll
    for (int i = 0; i < 15; i++)
      facc[i] = 0.0;
    const bool fv_by_adjoint = do_f && L.compute_tau;
    if (fv_by_adjoint) {
      const int na1 = ncoset(T.la_max + 1), nb1 = ncoset(T.lb_max + 1);
      for (int q = lane; q < na1 * nb1; q += G) {
        const int ia = q % na1, ib = q / na1;
        const Orb a = orb_of(s_orb, ia), b = orb_of(s_orb, ib);
        double g = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const int ua = oidx(oup(k, a)), ub = oidx(oup(k, b));
          const int da = oidx(odn(k, a)), db = oidx(odn(k, b));
          const bool ok_ua = (ua >= a_lo && ua < na), ok_ub = (ub >= b_lo && ub < nb);
          const bool ok_da = (a.l[k] >= 1 && da >= a_lo && da < na);
          const bool ok_db = (b.l[k] >= 1 && db >= b_lo && db < nb);
          if (ok_ua && ok_ub)
            g += 0.5 * (a.l[k] + 1) * (b.l[k] + 1) * s_raw[ub * na + ua];
          if (ok_da && ok_ub)
            g -= P.zeta * (b.l[k] + 1) * s_raw[ub * na + da];
          if (ok_ua && ok_db)
            g -= (a.l[k] + 1) * P.zetb * s_raw[db * na + ua];
          if (ok_da && ok_db)
            g += 2.0 * P.zeta * P.zetb * s_raw[db * na + da];
        }
        s_g[q] = g;
      }
      sync();
      for (int q = lane; q < na1 * nb1; q += G) {
        const double g = s_g[q];
        if (g != 0.0)
          accumulate_fv(P, g, orb_of(s_orb, q % na1), orb_of(s_orb, q / na1), do_v, facc);
      }
    }
    for (int q = lane; q < na * nb; q += G) {
      const int ia = q % na, ib = q / na;
      double hval = 0.0;
      if (ia >= a_lo && ib >= b_lo) {
        const Orb a = orb_of(s_orb, ia), b = orb_of(s_orb, ib);
        hval = vab_elem(P, L.compute_tau, 0, 0, 0, a, b);
        if (do_f && !fv_by_adjoint)
          accumulate_fv(P, s_raw[ib * na + ia], a, b, do_v, facc);
      }
      s_h[q] = hval;
    }
    sync();

    // (5) contract into the spherical block: block += sphi_a h sphi_b^T
    const double *sphi_a = L.sphi_pool + T.sphi_a + T.sgfa * T.maxcoa + T.o1;
    const double *sphi_b = L.sphi_pool + T.sphi_b + T.sgfb * T.maxcob + T.o2;
    for (int q = lane; q < T.nsgf_setb * na; q += G) {
      const int sb = q / na, ico = q - sb * na;
      const double *ps = sphi_b + sb * T.maxcob, *ph = s_h + ico;
      double acc = 0.0;
#pragma unroll 4
      for (int jco = 0; jco < nb; jco++, ps++, ph += na)
        acc += __ldg(ps) * *ph;
      s_work[q] = acc;
    }
    sync();
    double *g_block = L.hab + T.block_offset;
    for (int q = lane; q < T.nsgf_seta * T.nsgf_setb; q += G) {
      const int sb = q / T.nsgf_seta, sa = q - sb * T.nsgf_seta;
      const double *pw = s_work + sb * na, *ps = sphi_a + sa * T.maxcoa;
      double acc = 0.0;
#pragma unroll 4
      for (int ico = 0; ico < na; ico++)
        acc += pw[ico] * __ldg(ps + ico);
      if (acc != 0.0)
        atomicAdd(&g_block[block_index(T, sa, sb)], acc);
    }

    // (6) forces / virial of this task (src/grid/ref/grid_ref_task_list.c:625-641)
    if (do_f) {
      const int nred = do_v ? 15 : 6;
      const double scale = (T.iatom == T.jatom) ? 1.0 : 2.0;
#pragma unroll
      for (int i = 0; i < 15; i++) {
        if (i >= nred)
          break;
        double v = facc[i];
#pragma unroll
        for (int o = ((G > 32) ? 32 : G) / 2; o > 0; o >>= 1)
          v += __shfl_xor_sync(gmask, v, o);
        if ((lane & 31) == 0 && v != 0.0) {
          if (i < 3)
            atomicAdd(&L.forces[3 * T.iatom + i], scale * v);
          else if (i < 6)
            atomicAdd(&L.forces[3 * T.jatom + i - 3], scale * v);
          else
            atomicAdd(&L.virial[i - 6], scale * v);
        }
      }
    }
    sync();
  }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
constexpr size_t kSmemBudget = 200 * 1024;

inline void launch_pab_to_coef(const CoefLaunch &L, const int func, const double *pab,
                               const int max_nsgf_set, const int max_ncoset_raw,
                               const int max_la_c, const int max_lb_c) {
  if (L.ntasks == 0)
    return;
  CoefDims D;
  D.work = max_nsgf_set * max_ncoset_raw;
  D.raw = max_ncoset_raw * max_ncoset_raw;
  D.cab = ncoset(max_la_c) * ncoset(max_lb_c);
  D.alpha = 3 * (max_la_c + 1) * (max_lb_c + 1) * (max_la_c + max_lb_c + 1);
  D.cxyz = ncoset(max_la_c + max_lb_c);
  const size_t per_group = (size_t)D.total() * sizeof(double);
  B200_ASSERT(per_group <= kSmemBudget, "basis too large for the coefficient kernel");
  // half-warp groups when two of them fit (small tasks), else one warp per task
  const int gpw = (8 * per_group <= kSmemBudget) ? 2 : 1;
  const int wpc = (int)std::min<size_t>(4, kSmemBudget / (per_group * gpw));
  const size_t bytes = per_group * gpw * wpc;
  // persistent CTAs (one wave): the per-CTA set-up (orbital table) is paid once per SM slot
  auto grid_for = [&](auto kernel) {
    B200_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int per_sm = 1;
    B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * wpc, bytes));
    return std::min((L.ntasks + gpw * wpc - 1) / (gpw * wpc), 148 * std::max(per_sm, 1));
  };
  if (gpw == 2) {
    const int grid = grid_for(pab_to_coef_kernel<16>);
    pab_to_coef_kernel<16><<<grid, 32 * wpc, bytes, L.stream>>>(L, func, pab, D);
  } else {
    const int grid = grid_for(pab_to_coef_kernel<32>);
    pab_to_coef_kernel<32><<<grid, 32 * wpc, bytes, L.stream>>>(L, func, pab, D);
  }
  B200_CHECK(cudaGetLastError());
  count_launch();
}

inline void launch_coef_to_hab(const HabLaunch &L, const int ntasks, const int max_ncoset_raw,
                               const int dla_max, const int dla_min, const int dlb_max,
                               const int dlb_min) {
  if (ntasks == 0)
    return;
  HabDims D;
  D.work = L.max_nsgf_set * max_ncoset_raw;
  D.raw = max_ncoset_raw * max_ncoset_raw;
  D.cab = ncoset(L.max_la_l) * ncoset(L.max_lb_l);
  D.alpha = 3 * (L.max_la_l + 1) * (L.max_lb_l + 1) * (L.max_la_l + L.max_lb_l + 1);
  D.cxyz = ncoset(L.max_la_l + L.max_lb_l);
  D.h = max_ncoset_raw * max_ncoset_raw;
  D.g = (L.compute_tau && L.forces != nullptr)
            ? ncoset(L.max_la_l - dla_max + 1) * ncoset(L.max_lb_l - dlb_max + 1)
            : 0;
  const size_t per_group = (size_t)D.total() * sizeof(double);
  B200_ASSERT(per_group <= kSmemBudget, "basis too large for the hab kernel");
  const int gpw = (8 * per_group <= kSmemBudget) ? 2 : 1;
  if (16 * per_group > kSmemBudget) {
    // fewer than 16 tasks fit an SM: one CTA of four warps per task instead of one warp
    B200_CHECK(cudaFuncSetAttribute(coef_to_hab_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)per_group));
    int per_sm = 1;
    B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coef_to_hab_kernel<128>, 128, per_group));
    const int grid = std::min(ntasks, 148 * std::max(per_sm, 1));
    coef_to_hab_kernel<128><<<grid, 128, per_group, L.stream>>>(L, D, ntasks, dla_max, dla_min, dlb_max, dlb_min);
    B200_CHECK(cudaGetLastError());
    count_launch();
    return;
  }
  const int wpc = (int)std::min<size_t>(4, kSmemBudget / (per_group * gpw));
  const size_t bytes = per_group * gpw * wpc;
  auto grid_for = [&](auto kernel) {
    B200_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int per_sm = 1;
    B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * wpc, bytes));
    return std::min((ntasks + gpw * wpc - 1) / (gpw * wpc), 148 * std::max(per_sm, 1));
  };
  if (gpw == 2) {
    const int grid = grid_for(coef_to_hab_kernel<16>);
    coef_to_hab_kernel<16><<<grid, 32 * wpc, bytes, L.stream>>>(L, D, ntasks, dla_max, dla_min, dlb_max,
                                                               dlb_min);
  } else {
    const int grid = grid_for(coef_to_hab_kernel<32>);
    coef_to_hab_kernel<32><<<grid, 32 * wpc, bytes, L.stream>>>(L, D, ntasks, dla_max, dla_min, dlb_max,
                                                               dlb_min);
  }
  B200_CHECK(cudaGetLastError());
  count_launch();
}

}  // namespace b200
