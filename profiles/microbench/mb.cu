// Micro-benchmarks that size the design of the grid kernels (not product code):
//  1. FP64 FMA peak (the roofline denominator MEASURED_PEAKS.json lacks)
//  2. global FP64 atomic-add (RED.E.ADD.F64) throughput, coalesced rows, vs
//     grid size (L2-resident) and row length
//  3. shared-memory FP64 read-modify-write throughput
//  4. cp.reduce.async.bulk .add.f64 (TMA bulk reduce smem->global) throughput
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)

__global__ void k_dfma(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// each warp adds rows of `rowlen` consecutive doubles at pseudo-random row starts
__global__ void k_red(double *grid, size_t npts, int rowlen, int rows_per_warp, unsigned seed) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  unsigned s = seed + warp * 2654435761u;
  for (int r = 0; r < rows_per_warp; r++) {
    s = s * 1664525u + 1013904223u;
    size_t base = ((size_t)s * 4u) % (npts - 64);
    for (int i = lane; i < rowlen; i += 32) atomicAdd(&grid[base + i], 1.0 + i);
  }
}

__global__ void k_smem_rmw(double *out, int iters, int rowlen) {
  extern __shared__ double tile[];
  const int n = 24 * 1024;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned s = 12345u + warp * 2654435761u + blockIdx.x;
  for (int it = 0; it < iters; it++) {
    s = s * 1664525u + 1013904223u;
    // each warp owns a private slice (no races): slice = warp * (n/nw)
    int base = warp * (n / nw) + (s >> 8) % (n / nw - 64);
    if (lane < rowlen) tile[base + lane] += 1.0 + lane;
  }
  __syncthreads();
  double acc = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += tile[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// TMA bulk reduce: each CTA repeatedly reduces a smem row block into global
__global__ void k_bulk_red(double *grid, size_t npts, int bytes, int iters) {
  extern __shared__ __align__(128) double stile[];
  for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) stile[i] = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = 777u + blockIdx.x * 2654435761u;
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stile);
    for (int it = 0; it < iters; it++) {
      s = s * 1664525u + 1013904223u;
      size_t base = (((size_t)s * 4u) % (npts - bytes / 8 - 2)) & ~(size_t)1;  // 16B aligned
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                   :: "l"(grid + base), "r"(saddr), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if ((it & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double *out; CK(cudaMalloc(&out, 148 * 8 * 1024 * 8));
  { // 1. DFMA
    for (int tpb : {128, 256, 512, 1024}) {
      int blocks = 148 * (2048 / tpb); int iters = 20000;
      float ms = timeit([&] { k_dfma<<<blocks, tpb>>>(out, iters); });
      double fl = 2.0 * 8 * iters * (double)blocks * tpb;
      printf("DFMA tpb=%4d blocks=%d : %.3f ms  %.2f TFLOP/s\n", tpb, blocks, ms, fl / ms * 1e-9);
    }
  }
  { // 2. RED.F64
    for (size_t npts : {(size_t)64000, (size_t)343000, (size_t)1728000, (size_t)8000000, (size_t)64000000}) {
      double *g; CK(cudaMalloc(&g, npts * 8)); CK(cudaMemset(g, 0, npts * 8));
      for (int rowlen : {8, 16, 32}) {
        int blocks = 148 * 8, tpb = 256, rows = 2000;
        float ms = timeit([&] { k_red<<<blocks, tpb>>>(g, npts, rowlen, rows, 1u); });
        double n = (double)blocks * (tpb / 32) * rows * rowlen;
        printf("RED npts=%9zu rowlen=%2d : %.3f ms  %.1f Gatom/s\n", npts, rowlen, ms, n / ms * 1e-6);
      }
      CK(cudaFree(g));
    }
  }
  { // 3. smem RMW
    CK(cudaFuncSetAttribute(k_smem_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024 * 8));
    for (int rowlen : {16, 32}) for (int tpb : {256, 512, 1024}) {
      int iters = 20000;
      float ms = timeit([&] { k_smem_rmw<<<148, tpb, 24 * 1024 * 8>>>(out, iters, rowlen); });
      double n = 148.0 * (tpb / 32) * iters * rowlen;
      printf("SMEM-RMW tpb=%4d rowlen=%2d : %.3f ms  %.1f Gupd/s  (%.2f upd/clk/SM @1.9GHz)\n", tpb, rowlen, ms, n / ms * 1e-6, n / ms * 1e-6 / 148 / 1.9);
    }
  }
  { // 4. bulk reduce
    CK(cudaFuncSetAttribute(k_bulk_red, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (size_t npts : {(size_t)343000, (size_t)8000000}) {
      double *g; CK(cudaMalloc(&g, npts * 8)); CK(cudaMemset(g, 0, npts * 8));
      for (int bytes : {128, 256, 1024, 8192}) {
        int iters = 4000;
        float ms = timeit([&] { k_bulk_red<<<148 * 4, 32, 64 * 1024>>>(g, npts, bytes, iters); });
        double n = 148.0 * 4 * iters * (bytes / 8);
        printf("BULK-RED npts=%9zu bytes=%5d : %.3f ms  %.1f Gatom/s\n", npts, bytes, ms, n / ms * 1e-6);
      }
      CK(cudaFree(g));
    }
  }
  return 0;
}
