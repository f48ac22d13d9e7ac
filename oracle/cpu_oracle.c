/*
 * cpu_oracle.c — plain-C restatement of the reference's CPU contraction path.
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs as the checker / timed baseline; never linked into or called by the product library.
 *
 * What it restates (file:line under /root/reference):
 *   - tile permutation, image convention: src/TiledArray/permutation.h:69-79,
 *     src/TiledArray/tensor/permute.h:118-209                          -> oracle_permute_f64
 *   - row-major tile GEMM with op flags: src/TiledArray/tensor/kernels.h:92-231,
 *     src/TiledArray/math/blas.h:171-177 (col-major BLAS call with swapped operands)
 *                                                                      -> oracle_gemm_naive_f64 (triple loop,
 *                                                                         as tests/math_blas.cpp:156-250)
 *                                                                      -> oracle_gemm_blas_f64 (vendor DGEMM)
 *   - SparseShape::gemm arithmetic in the oracle's fixed order: src/TiledArray/sparse_shape.h:1589-1663
 *                                                                      -> oracle_shape_gemm_f32
 *   - SUMMA on a 1x1 grid with TiledArray's CPU execution model: one single-threaded DGEMM per
 *     tile pair (TiledArray forces BLAS to 1 thread, src/TiledArray/tiledarray.cpp:112) spread
 *     over a pool of task threads (MAD_NUM_THREADS analogue); pair list per step as
 *     Summa::contract, src/TiledArray/dist_eval/contraction_eval.h:1311-1384; beta = 0 on the
 *     first pair of a result tile then 1, src/TiledArray/tensor/tensor.h:3134-3140
 *                                                                      -> oracle_cpu_contract_f64
 * The FP64 arithmetic itself is the third-party BLAS the reference calls through BLAS++ (absent
 * from the reference tree); here it is OpenBLAS 0.3.30 from the numpy wheel, dlopen'ed at run time
 * (ILP64 symbol scipy_cblas_dgemm64_).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -shared -fPIC ... -lpthread -ldl)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- permutation ------------------------------------------------------------------------- */
/* out[perm(idx)] = in[idx]; out.extent[perm[i]] = extent[i] */
int oracle_permute_f64(int rank, const int64_t* extent, const int32_t* perm, const double* in, double* out) {
  int64_t out_ext[16], out_stride[16], idx[16];
  if (rank < 0 || rank > 16) return 1;
  int64_t total = 1;
  for (int i = 0; i < rank; ++i) { out_ext[perm[i]] = extent[i]; total *= extent[i]; }
  int64_t st = 1;
  for (int j = rank - 1; j >= 0; --j) { out_stride[j] = st; st *= out_ext[j]; }
  for (int i = 0; i < rank; ++i) idx[i] = 0;
  for (int64_t o = 0; o < total; ++o) {
    int64_t dst = 0;
    for (int i = 0; i < rank; ++i) dst += idx[i] * out_stride[perm[i]];
    out[dst] = in[o];
    for (int i = rank - 1; i >= 0; --i) {
      if (++idx[i] < extent[i]) break;
      idx[i] = 0;
    }
  }
  return 0;
}

/* ---- tile GEMM --------------------------------------------------------------------------- */
/* C[m x n] = alpha*op(A)*op(B) + beta*C, all row-major with natural leading dimensions */
void oracle_gemm_naive_f64(int opA, int opB, int m, int n, int k, double alpha, const double* A, const double* B,
                           double beta, double* C) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int x = 0; x < k; ++x) {
        const double a = opA ? A[(size_t)x * m + i] : A[(size_t)i * k + x];
        const double b = opB ? B[(size_t)j * k + x] : B[(size_t)x * n + j];
        s += a * b;
      }
      C[(size_t)i * n + j] = alpha * s + (beta == 0.0 ? 0.0 : beta * C[(size_t)i * n + j]);
    }
}

typedef void (*dgemm64_fn)(int order, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                           const double* a, int64_t lda, const double* b, int64_t ldb, double beta, double* c,
                           int64_t ldc);
typedef void (*set_threads_fn)(int);
static dgemm64_fn g_dgemm = NULL;
static set_threads_fn g_set_threads = NULL;

int oracle_load_blas(const char* path) {
  if (g_dgemm) return 0;
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "oracle: dlopen(%s) failed: %s\n", path, dlerror()); return 1; }
  g_dgemm = (dgemm64_fn)dlsym(h, "scipy_cblas_dgemm64_");
  g_set_threads = (set_threads_fn)dlsym(h, "scipy_openblas_set_num_threads64_");
  if (!g_dgemm) { fprintf(stderr, "oracle: scipy_cblas_dgemm64_ not found in %s\n", path); return 2; }
  if (g_set_threads) g_set_threads(1); /* tiledarray.cpp:112 */
  return 0;
}

/* math/blas.h:171-177: row-major product realised as a column-major call with swapped operands */
int oracle_gemm_blas_f64(int opA, int opB, int m, int n, int k, double alpha, const double* A, const double* B,
                         double beta, double* C) {
  if (!g_dgemm) return 1;
  const int64_t lda = opA ? m : k, ldb = opB ? k : n, ldc = n;
  /* CblasColMajor = 102, CblasNoTrans = 111, CblasTrans = 112 */
  g_dgemm(102, opB ? 112 : 111, opA ? 112 : 111, n, m, k, alpha, B, ldb, A, lda, beta, C, ldc);
  return 0;
}

/* ---- SparseShape::gemm arithmetic (fixed order, no FMA contraction) ------------------------ */
int64_t oracle_shape_gemm_f32(int M, int N, int K, const float* a, const float* b, const float* ksz, float abs_factor,
                              float thr, float* out) {
  int64_t nzero = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      volatile float acc = 0.0f;
      for (int k = 0; k < K; ++k) {
        volatile float la = a[(size_t)m * K + k] * ksz[k];
        volatile float rb = b[(size_t)k * N + n] * ksz[k];
        volatile float p = la * rb;
        acc = acc + p;
      }
      float v = abs_factor * acc;
      if (v < thr) { v = 0.0f; ++nzero; }
      out[(size_t)m * N + n] = v;
    }
  return nzero;
}

/* ---- SUMMA on a 1x1 grid, TiledArray CPU execution model ----------------------------------- */
typedef struct {
  int Mt, Nt, Kt;
  const int64_t *m_ext, *n_ext, *k_ext;
  int opA, opB;
  double alpha;
  const double* const* a_tiles; /* [Mt*Kt] host pointers, NULL = zero tile */
  const double* const* b_tiles; /* [Kt*Nt] */
  double* const* c_tiles;       /* [Mt*Nt], NULL = zero result tile */
  volatile int64_t next;        /* work counter over result tiles */
  int64_t npairs;
  pthread_mutex_t mu;
} contract_job;

static void* contract_worker(void* arg) {
  contract_job* J = (contract_job*)arg;
  int64_t pairs = 0;
  for (;;) {
    const int64_t w = __sync_fetch_and_add(&J->next, 1);
    if (w >= (int64_t)J->Mt * J->Nt) break;
    const int i = (int)(w / J->Nt), j = (int)(w % J->Nt);
    double* C = J->c_tiles[w];
    if (!C) continue; /* zero result tile: skipped (contraction_eval.h:1370) */
    int first = 1;
    for (int k = 0; k < J->Kt; ++k) { /* SUMMA steps in order; one reduce task per result tile */
      const double* A = J->a_tiles[(size_t)i * J->Kt + k];
      const double* B = J->b_tiles[(size_t)k * J->Nt + j];
      if (!A || !B) continue;
      oracle_gemm_blas_f64(J->opA, J->opB, (int)J->m_ext[i], (int)J->n_ext[j], (int)J->k_ext[k], J->alpha, A, B,
                           first ? 0.0 : 1.0, C);
      first = 0;
      ++pairs;
    }
    if (first) memset(C, 0, sizeof(double) * (size_t)J->m_ext[i] * J->n_ext[j]);
  }
  pthread_mutex_lock(&J->mu);
  J->npairs += pairs;
  pthread_mutex_unlock(&J->mu);
  return NULL;
}

/* returns wall seconds (clock_gettime MONOTONIC), or < 0 on error */
double oracle_cpu_contract_f64(int Mt, int Nt, int Kt, const int64_t* m_ext, const int64_t* n_ext, const int64_t* k_ext,
                               int opA, int opB, double alpha, const double* const* a_tiles,
                               const double* const* b_tiles, double* const* c_tiles, int nthreads, int64_t* npairs) {
  if (!g_dgemm) return -1.0;
  if (nthreads < 1) nthreads = 1;
  contract_job J;
  memset(&J, 0, sizeof(J));
  J.Mt = Mt; J.Nt = Nt; J.Kt = Kt; J.m_ext = m_ext; J.n_ext = n_ext; J.k_ext = k_ext;
  J.opA = opA; J.opB = opB; J.alpha = alpha; J.a_tiles = a_tiles; J.b_tiles = b_tiles; J.c_tiles = c_tiles;
  pthread_mutex_init(&J.mu, NULL);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, contract_worker, &J);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  if (npairs) *npairs = J.npairs;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
