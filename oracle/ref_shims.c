/*
 * TEST INFRASTRUCTURE ONLY -- link-time shims for building the UNMODIFIED
 * reference grid library (REF + CPU backends) out of /root/reference into
 * oracle/_ref/libgrid_ref.so.  Nothing here is product code.
 *
 *  - dgemm_: the REF/CPU backends call BLAS only for the tiny (<= ~50x50)
 *    (de)contraction products in load_pab / store_hab
 *    (src/grid/ref/grid_ref_task_list.c:216-231, 237-271, 462-499).  The image
 *    has no system BLAS, so a plain column-major triple loop is supplied.
 *  - grid_dgemm_*: the dispatcher (src/grid/grid_task_list.c:105,156,209,334)
 *    references the DGEMM backend, which is out of scope; the
 *    stubs abort loudly if it is ever selected.
 */
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>

void dgemm_(const char *transa, const char *transb, const int *m, const int *n,
            const int *k, const double *alpha, const double *a, const int *lda,
            const double *b, const int *ldb, const double *beta, double *c,
            const int *ldc) {
  const bool ta = (*transa == 'T' || *transa == 't');
  const bool tb = (*transb == 'T' || *transb == 't');
  for (int j = 0; j < *n; j++) {
    for (int i = 0; i < *m; i++) {
      double acc = 0.0;
      for (int l = 0; l < *k; l++) {
        const double av = ta ? a[l + (size_t)i * *lda] : a[i + (size_t)l * *lda];
        const double bv = tb ? b[j + (size_t)l * *ldb] : b[l + (size_t)j * *ldb];
        acc += av * bv;
      }
      double *cij = &c[i + (size_t)j * *ldc];
      *cij = (*beta == 0.0) ? (*alpha) * acc : (*alpha) * acc + (*beta) * (*cij);
    }
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
