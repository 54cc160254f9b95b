/*
 * TEST INFRASTRUCTURE ONLY -- link-time shims for building the UNMODIFIED
 * reference grid library (REF + CPU backends) out of /root/reference into
 * oracle/_ref/libgrid_ref.so.  Nothing here is product code.
 *
 *  - dgemm_: the REF/CPU backends call BLAS only for the tiny (<= ~50x50)
 *    (de)contraction products in load_pab / store_hab
 *    (src/grid/ref/grid_ref_task_list.c:216-231, 237-271, 462-499).  The image
 *    has no system BLAS, so a plain column-major implementation with unit-stride
 *    inner loops is supplied; the time spent in it is counted per thread so that
 *    the benchmark can state its share of the CPU baseline.
 *  - grid_dgemm_*: the dispatcher (src/grid/grid_task_list.c:105,156,209,334)
 *    references the DGEMM backend, which is out of scope; the
 *    stubs abort loudly if it is ever selected.
 */
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>

#include <omp.h>

/* Per-thread time spent in the shim, so that bench.py can state its share of the
 * CPU baseline (ref_shims_dgemm_stats). */
#define SHIM_MAX_THREADS 1024
static double shim_seconds[SHIM_MAX_THREADS][8]; /* padded against false sharing */
static long long shim_calls[SHIM_MAX_THREADS][8];

void ref_shims_dgemm_stats(double *seconds_sum, long long *calls_sum, const int reset) {
  double s = 0.0;
  long long c = 0;
  for (int t = 0; t < SHIM_MAX_THREADS; t++) {
    s += shim_seconds[t][0];
    c += shim_calls[t][0];
    if (reset) {
      shim_seconds[t][0] = 0.0;
      shim_calls[t][0] = 0;
    }
  }
  *seconds_sum = s;
  *calls_sum = c;
}

void dgemm_(const char *transa, const char *transb, const int *m, const int *n,
            const int *k, const double *alpha, const double *a, const int *lda,
            const double *b, const int *ldb, const double *beta, double *c,
            const int *ldc) {
  const double t0 = omp_get_wtime();
  const bool ta = (*transa == 'T' || *transa == 't');
  const bool tb = (*transb == 'T' || *transb == 't');
  const int M = *m, N = *n, K = *k;
  const size_t LDA = (size_t)*lda, LDB = (size_t)*ldb, LDC = (size_t)*ldc;
  const double al = *alpha, be = *beta;
  /* C = beta C first, then rank-1 / dot updates with unit-stride inner loops */
  for (int j = 0; j < N; j++)
    for (int i = 0; i < M; i++)
      c[i + j * LDC] = (be == 0.0) ? 0.0 : be * c[i + j * LDC];
  if (!ta) { /* columns of A are contiguous: axpy form */
    for (int j = 0; j < N; j++)
      for (int l = 0; l < K; l++) {
        const double bv = al * (tb ? b[j + l * LDB] : b[l + j * LDB]);
        const double *restrict ac = a + l * LDA;
        double *restrict cc = c + j * LDC;
        for (int i = 0; i < M; i++)
          cc[i] += ac[i] * bv;
      }
  } else if (!tb) { /* A^T B: dot products over contiguous columns */
    for (int j = 0; j < N; j++)
      for (int i = 0; i < M; i++) {
        const double *restrict ac = a + i * LDA;
        const double *restrict bc = b + j * LDB;
        double acc = 0.0;
        for (int l = 0; l < K; l++)
          acc += ac[l] * bc[l];
        c[i + j * LDC] += al * acc;
      }
  } else { /* A^T B^T */
    for (int j = 0; j < N; j++)
      for (int i = 0; i < M; i++) {
        const double *restrict ac = a + i * LDA;
        double acc = 0.0;
        for (int l = 0; l < K; l++)
          acc += ac[l] * b[j + l * LDB];
        c[i + j * LDC] += al * acc;
      }
  }
  const int t = omp_get_thread_num();
  if (t < SHIM_MAX_THREADS) {
    shim_seconds[t][0] += omp_get_wtime() - t0;
    shim_calls[t][0] += 1;
  }
}

static void no_dgemm_backend(void) {
  fprintf(stderr, "oracle/_ref: the DGEMM grid backend is not built (out of "
                  "scope); select REF (11) or CPU (12).\n");
  abort();
}

void grid_dgemm_create_task_list(void) { no_dgemm_backend(); }
void grid_dgemm_free_task_list(void *p) { (void)p; }
void grid_dgemm_collocate_task_list(void) { no_dgemm_backend(); }
void grid_dgemm_integrate_task_list(void) { no_dgemm_backend(); }
