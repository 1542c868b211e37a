/* TEST / BASELINE INFRASTRUCTURE -- not part of the product (the product path is hubbardtn_b200/libhtn.so).
 *
 * CPU execution of the staged H_AC apply of oracle/heff.py:HeffACPlan (the restatement of what MPSKit's `∂AC` does for
 * /root/reference/src/HubbardFunctions.jl:1012,1017,1027) with native threads: the same GEMM list (stage L: T = GL.x,
 * stage R: y += U.GR) through single-threaded OpenBLAS `cblas_dgemm` calls, one worker thread per host core over the
 * blocks -- the reference's own CPU policy (BLAS threads = 1, HF:29; all Julia threads over sector blocks, HF:37) --
 * without Python's interpreter lock between the calls.  The recoupling mix (stage W) is plain C loops.
 *
 * Built by oracle/native/Makefile (gcc), loaded with ctypes by oracle/native/__init__.py; the BLAS entry point is handed
 * in as a function pointer (numpy's bundled ILP64 OpenBLAS, symbol `scipy_cblas_dgemm64_`).  */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef void (*dgemm_fn)(int order, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const double* a,
                         int64_t lda, const double* b, int64_t ldb, double beta, double* c, int64_t ldc);

typedef struct {
  int64_t a_off, b_off, c_off; /* element offsets into the GL / x|U / T|y arrays */
  int32_t m, n, k, pad;
} htn_cpu_gemm;

typedef struct {
  int64_t dst_off;
  int32_t dst_space; /* 2 = U, 3 = y */
  int32_t nelem;
  int32_t src_begin, src_end;
} htn_cpu_mix;

typedef struct {
  int64_t off;
  int32_t space; /* 0 = x, 1 = T */
  int32_t pad;
  double coef;
} htn_cpu_src;

typedef struct {
  dgemm_fn dgemm;
  const double *GL, *GR, *x;
  double *y, *T, *U;
  const htn_cpu_gemm* gl; /* stage L: T[c_off] = GL[a_off] (m x k) . x[b_off] (k x n) */
  int32_t ngl;
  const htn_cpu_mix* mix;
  const htn_cpu_src* src;
  int32_t nmix;
  const htn_cpu_gemm* gr;    /* stage R, grouped by output block: y[c_off] += U[a_off] (m x k) . GR[b_off] (k x n) */
  const int32_t* gr_group;   /* ngroups + 1 offsets into gr: one group = one y block, processed by one thread */
  int32_t ngroups;
  int32_t nthreads;
  atomic_int next[3];
  pthread_barrier_t bar;
} job_t;

static void* worker(void* arg) {
  job_t* J = (job_t*)arg;
  /* stage L */
  for (;;) {
    const int i = atomic_fetch_add(&J->next[0], 1);
    if (i >= J->ngl) break;
    const htn_cpu_gemm* g = &J->gl[i];
    J->dgemm(101, 111, 111, g->m, g->n, g->k, 1.0, J->GL + g->a_off, g->k, J->x + g->b_off, g->n, 0.0, J->T + g->c_off, g->n);
  }
  pthread_barrier_wait(&J->bar);
  /* stage W: dst = sum coef * src */
  for (;;) {
    const int i = atomic_fetch_add(&J->next[1], 1);
    if (i >= J->nmix) break;
    const htn_cpu_mix* t = &J->mix[i];
    double* dst = (t->dst_space == 2 ? J->U : J->y) + t->dst_off;
    const int n = t->nelem;
    if (t->dst_space == 2) memset(dst, 0, (size_t)n * sizeof(double)); /* y was cleared before the launch */
    for (int s = t->src_begin; s < t->src_end; ++s) {
      const double* sp = (J->src[s].space == 0 ? J->x : J->T) + J->src[s].off;
      const double cf = J->src[s].coef;
      for (int e = 0; e < n; ++e) dst[e] += cf * sp[e];
    }
  }
  pthread_barrier_wait(&J->bar);
  /* stage R */
  for (;;) {
    const int gi = atomic_fetch_add(&J->next[2], 1);
    if (gi >= J->ngroups) break;
    for (int i = J->gr_group[gi]; i < J->gr_group[gi + 1]; ++i) {
      const htn_cpu_gemm* g = &J->gr[i];
      J->dgemm(101, 111, 111, g->m, g->n, g->k, 1.0, J->U + g->a_off, g->k, J->GR + g->b_off, g->n, 1.0, J->y + g->c_off, g->n);
    }
  }
  return 0;
}

/* one H_AC apply; y (ny elements) is cleared first.  Returns 0, or -1 if threads could not be started. */
int htn_cpu_heff_apply(void* dgemm, const double* GL, const double* GR, const double* x, double* y, int64_t ny, double* T,
                       double* U, const htn_cpu_gemm* gl, int32_t ngl, const htn_cpu_mix* mix, const htn_cpu_src* src,
                       int32_t nmix, const htn_cpu_gemm* gr, const int32_t* gr_group, int32_t ngroups, int32_t nthreads) {
  job_t J;
  memset(&J, 0, sizeof(J));
  J.dgemm = (dgemm_fn)dgemm;
  J.GL = GL; J.GR = GR; J.x = x; J.y = y; J.T = T; J.U = U;
  J.gl = gl; J.ngl = ngl; J.mix = mix; J.src = src; J.nmix = nmix; J.gr = gr; J.gr_group = gr_group; J.ngroups = ngroups;
  J.nthreads = nthreads < 1 ? 1 : nthreads;
  memset(y, 0, (size_t)ny * sizeof(double));
  for (int i = 0; i < 3; ++i) atomic_init(&J.next[i], 0);
  if (pthread_barrier_init(&J.bar, 0, (unsigned)J.nthreads)) return -1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)J.nthreads);
  int started = 0;
  for (int i = 1; i < J.nthreads; ++i) {
    if (pthread_create(&th[i], 0, worker, &J)) break;
    ++started;
  }
  if (started != J.nthreads - 1) { /* cannot run with a partial barrier */
    /* threads already started are blocked on the barrier: give up hard rather than deadlock */
    abort();
  }
  worker(&J);
  for (int i = 1; i < J.nthreads; ++i) pthread_join(th[i], 0);
  free(th);
  pthread_barrier_destroy(&J.bar);
  return 0;
}
