/* ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into or called by the product (dawn_b200/).
 *
 * Plain-C restatements of what the reference's cxxnaive backend executes for the headline
 * stencils, in exactly the loop shape `dawn/src/dawn/CodeGen/CXXNaive/CXXNaiveCodeGen.cpp:380-495`
 * emits (multistage -> k interval -> stage -> i outer -> j inner, stage extents widening the ij
 * loops, temporaries as full 3-D storages, expressions fully parenthesised as
 * `dawn/src/dawn/CodeGen/ASTCodeGenCXX.cpp:109-153` prints them).  Compile with
 * -ffp-contract=off so every node is one IEEE operation.
 *
 * No pre-generated naive text exists in the reference for these stencils and Dawn itself cannot be
 * built in this image (needs Clang>=9 dev libs + protobuf, dawn/CMakeLists.txt:139-146), so these
 * follow the DSL sources cited per function; the general IIR interpreter oracle/naive_interp.py
 * (pinned against the reference's own generated naive code, oracle/_ref) cross-checks every one of
 * them bit-for-bit in tests/test_oracle.py.
 *
 * Buffers: dense doubles, i fastest; element (i,j,k) of an ijk storage at i + j*sj + k*sk.
 * `ni,nj,nk` are the domain sizes INCLUDING halos (driver-includes/domain.hpp:42-58); `h` is the
 * horizontal halo (GRIDTOOLS_DAWN_HALO_EXTENT, default 3).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j, k) ((size_t)(i) + (size_t)(j) * sj + (size_t)(k) * sk)

/* driver-includes/verify.hpp:48-64 (fillMath) + :216-233 (for_each over total lengths).
 * mask bit0/1/2 = storage has an i/j/k dimension; masked dimensions have length 1. */
void oracle_fillmath(double a, double b, double c, double d, double e, double f, int ni, int nj,
                     int nk, int mask, double* out) {
  const int d1 = (mask & 1) ? ni : 1, d2 = (mask & 2) ? nj : 1, d3 = (mask & 4) ? nk : 1;
  const double pi = atan(1.) * 4.;
  for(int i = 0; i < d1; ++i)
    for(int j = 0; j < d2; ++j)
      for(int k = 0; k < d3; ++k) {
        double x = i / (double)d1;
        double y = j / (double)d2;
        out[(size_t)i + (size_t)j * d1 + (size_t)k * d1 * d2] =
            k * 10e-3 + a * (b + cos(pi * (x + c * y)) + sin(d * pi * (x + e * y))) / f;
      }
}

/* ---------------------------------------------------------------------------------------------
 * hori_diff_type2_stencil: gtclang/test/integration-test/CodeGen/hori_diff_type2_stencil.cpp:5-55
 * API order run(out, in, crlato, crlatu, hdmask) (…_benchmark.cpp:69); crlato/crlatu are storage_j.
 * Stage 1 (lap) runs on the interior widened by (-1,+1,-1,+1) (read extents of lap in stage 2),
 * stage 2 on the interior.  `lap` is a caller-provided scratch of the same shape as `in`.
 * k range [k0,k1] so that a threaded driver can split levels (Parallel loop order).
 * -------------------------------------------------------------------------------------------*/
void oracle_hori_diff_type2_krange(int ni, int nj, int nk, int h, size_t sj, size_t sk, double* out,
                                   const double* in, const double* crlato, const double* crlatu,
                                   const double* hdmask, double* lap, int k0, int k1) {
  const int iMin = h, iMax = ni - h - 1, jMin = h, jMax = nj - h - 1;
  (void)nk;
  for(int k = k0; k <= k1; ++k) {
    for(int i = iMin + -1; i <= iMax + 1; ++i) {
      for(int j = jMin + -1; j <= jMax + 1; ++j) {
        lap[IDX(i, j, k)] =
            ((((in[IDX(i + 1, j, k)] + in[IDX(i - 1, j, k)]) - (2.0 * in[IDX(i, j, k)])) +
              (crlato[j] * (in[IDX(i, j + 1, k)] - in[IDX(i, j, k)]))) +
             (crlatu[j] * (in[IDX(i, j - 1, k)] - in[IDX(i, j, k)])));
      }
    }
    for(int i = iMin; i <= iMax; ++i) {
      for(int j = jMin; j <= jMax; ++j) {
        const double flx_a = (lap[IDX(i + 1, j, k)] - lap[IDX(i, j, k)]);
        const double fx_a =
            ((flx_a * (in[IDX(i + 1, j, k)] - in[IDX(i, j, k)])) > 0.0) ? 0.0 : flx_a;
        const double flx_b = (lap[IDX(i, j, k)] - lap[IDX(i - 1, j, k)]);
        const double fx_b =
            ((flx_b * (in[IDX(i, j, k)] - in[IDX(i - 1, j, k)])) > 0.0) ? 0.0 : flx_b;
        const double delta_flux_x = (fx_a - fx_b);
        const double fly_a = (crlato[j] * (lap[IDX(i, j + 1, k)] - lap[IDX(i, j, k)]));
        const double fy_a =
            ((fly_a * (in[IDX(i, j + 1, k)] - in[IDX(i, j, k)])) > 0.0) ? 0.0 : fly_a;
        const double fly_b = (crlato[j - 1] * (lap[IDX(i, j, k)] - lap[IDX(i, j - 1, k)]));
        const double fy_b =
            ((fly_b * (in[IDX(i, j, k)] - in[IDX(i, j - 1, k)])) > 0.0) ? 0.0 : fly_b;
        const double delta_flux_y = (fy_a - fy_b);
        out[IDX(i, j, k)] =
            (in[IDX(i, j, k)] - (hdmask[IDX(i, j, k)] * (delta_flux_x + delta_flux_y)));
      }
    }
  }
}

void oracle_hori_diff_type2(int ni, int nj, int nk, int h, size_t sj, size_t sk, double* out,
                            const double* in, const double* crlato, const double* crlatu,
                            const double* hdmask, double* lap) {
  oracle_hori_diff_type2_krange(ni, nj, nk, h, sj, sk, out, in, crlato, crlatu, hdmask, lap, 0,
                                nk - 1);
}

/* hori_diff_stencil_01: gtclang/test/integration-test/CodeGen/hori_diff_stencil_01.cpp:21-31
 *   lap = u(i+1)+u(i-1)+u(j+1)+u(j-1)-4.0*u ; out = lap(i+1)+lap(i-1)+lap(j+1)+lap(j-1)-4.0*lap */
void oracle_hori_diff_01_krange(int ni, int nj, int nk, int h, size_t sj, size_t sk, const double* u,
                                double* out, double* lap, int k0, int k1) {
  const int iMin = h, iMax = ni - h - 1, jMin = h, jMax = nj - h - 1;
  (void)nk;
  for(int k = k0; k <= k1; ++k) {
    for(int i = iMin + -1; i <= iMax + 1; ++i)
      for(int j = jMin + -1; j <= jMax + 1; ++j)
        lap[IDX(i, j, k)] = ((((u[IDX(i + 1, j, k)] + u[IDX(i - 1, j, k)]) + u[IDX(i, j + 1, k)]) +
                              u[IDX(i, j - 1, k)]) -
                             (4.0 * u[IDX(i, j, k)]));
    for(int i = iMin; i <= iMax; ++i)
      for(int j = jMin; j <= jMax; ++j)
        out[IDX(i, j, k)] =
            ((((lap[IDX(i + 1, j, k)] + lap[IDX(i - 1, j, k)]) + lap[IDX(i, j + 1, k)]) +
              lap[IDX(i, j - 1, k)]) -
             (4.0 * lap[IDX(i, j, k)]));
  }
}

/* dawn4py hori_diff (in, out, coeff): expression trees pinned by the golden CUDA text
 * dawn/test/integration-test/dawn4py-tests/data/hori_diff_stencil_ref.cpp:105-119 */
void oracle_hori_diff_dawn4py_krange(int ni, int nj, int nk, int h, size_t sj, size_t sk,
                                     const double* in, double* out, const double* coeff, double* lap,
                                     int k0, int k1) {
  const int iMin = h, iMax = ni - h - 1, jMin = h, jMax = nj - h - 1;
  (void)nk;
  for(int k = k0; k <= k1; ++k) {
    for(int i = iMin + -1; i <= iMax + 1; ++i)
      for(int j = jMin + -1; j <= jMax + 1; ++j)
        lap[IDX(i, j, k)] =
            ((-4.0 * in[IDX(i, j, k)]) +
             (coeff[IDX(i, j, k)] *
              (in[IDX(i + 1, j, k)] +
               (in[IDX(i - 1, j, k)] + (in[IDX(i, j + 1, k)] + in[IDX(i, j - 1, k)])))));
    for(int i = iMin; i <= iMax; ++i)
      for(int j = jMin; j <= jMax; ++j)
        out[IDX(i, j, k)] =
            ((-4.0 * lap[IDX(i, j, k)]) +
             (coeff[IDX(i, j, k)] *
              (lap[IDX(i + 1, j, k)] +
               (lap[IDX(i - 1, j, k)] + (lap[IDX(i, j + 1, k)] + lap[IDX(i, j - 1, k)])))));
  }
}

/* tridiagonal_solve: gtclang/test/integration-test/CodeGen/tridiagonal_solve.cpp:21-38
 * API order run(d, a, b, c).  Forward multistage (intervals [0,0] and [1,end]), then Backward
 * multistage [end-1 .. 0]; ij loops inside each k (CXXNaiveCodeGen.cpp:422-489).
 * i range [i0,i1] so a threaded driver can split columns. */
void oracle_tridiagonal_solve_irange(int ni, int nj, int nk, int h, size_t sj, size_t sk, double* d,
                                     const double* a, const double* b, double* c, int i0, int i1) {
  const int jMin = h, jMax = nj - h - 1, kMin = 0, kMax = nk - 1;
  (void)ni;
  for(int k = kMin + 0; k <= kMin + 0; ++k)
    for(int i = i0; i <= i1; ++i)
      for(int j = jMin; j <= jMax; ++j) {
        c[IDX(i, j, k)] = (c[IDX(i, j, k)] / b[IDX(i, j, k)]);
        d[IDX(i, j, k)] = (d[IDX(i, j, k)] / b[IDX(i, j, k)]);
      }
  for(int k = kMin + 1; k <= kMax + 0; ++k)
    for(int i = i0; i <= i1; ++i)
      for(int j = jMin; j <= jMax; ++j) {
        double m = (1.0 / (b[IDX(i, j, k)] - (a[IDX(i, j, k)] * c[IDX(i, j, k - 1)])));
        c[IDX(i, j, k)] = (c[IDX(i, j, k)] * m);
        d[IDX(i, j, k)] = ((d[IDX(i, j, k)] - (a[IDX(i, j, k)] * d[IDX(i, j, k - 1)])) * m);
      }
  for(int k = kMax + -1; k >= kMin + 0; --k)
    for(int i = i0; i <= i1; ++i)
      for(int j = jMin; j <= jMax; ++j)
        d[IDX(i, j, k)] -= (c[IDX(i, j, k)] * d[IDX(i, j, k + 1)]);
}

void oracle_tridiagonal_solve(int ni, int nj, int nk, int h, size_t sj, size_t sk, double* d,
                              const double* a, const double* b, double* c) {
  oracle_tridiagonal_solve_irange(ni, nj, nk, h, sj, sk, d, a, b, c, h, ni - h - 1);
}

/* tutorial laplacian: dawn/examples/tutorial/laplacian_stencil.cpp:1-13 (storage_ij fields,
 * global dx):  out = (-4*in + in(i+1) + in(i-1) + in(j-1) + in(j+1)) / (dx*dx), for every k. */
void oracle_tutorial_laplacian(int ni, int nj, int nk, int h, size_t sj, double dx, double* out,
                               const double* in) {
  const int iMin = h, iMax = ni - h - 1, jMin = h, jMax = nj - h - 1;
  const size_t sk = 0;
  for(int k = 0; k <= nk - 1; ++k)
    for(int i = iMin; i <= iMax; ++i)
      for(int j = jMin; j <= jMax; ++j)
        out[IDX(i, j, k)] =
            ((((((-4 * in[IDX(i, j, k)]) + in[IDX(i + 1, j, k)]) + in[IDX(i - 1, j, k)]) +
               in[IDX(i, j - 1, k)]) +
              in[IDX(i, j + 1, k)]) /
             (dx * dx));
}

/* ---------------------------------------------------------------------------------------------
 * N-thread drivers for the CPU baseline: the naive backend itself is serial
 * (README.md:66); these split k (Parallel multistages) or i columns (solver) over pthreads without
 * changing any per-point arithmetic.
 * -------------------------------------------------------------------------------------------*/
typedef struct {
  int kind, ni, nj, nk, h, lo, hi;
  size_t sj, sk;
  double* p[6];
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  if(j->hi < j->lo)
    return 0;
  switch(j->kind) {
  case 0:
    oracle_hori_diff_type2_krange(j->ni, j->nj, j->nk, j->h, j->sj, j->sk, j->p[0], j->p[1], j->p[2],
                                  j->p[3], j->p[4], j->p[5], j->lo, j->hi);
    break;
  case 1:
    oracle_tridiagonal_solve_irange(j->ni, j->nj, j->nk, j->h, j->sj, j->sk, j->p[0], j->p[1],
                                    j->p[2], j->p[3], j->lo, j->hi);
    break;
  case 2:
    oracle_hori_diff_01_krange(j->ni, j->nj, j->nk, j->h, j->sj, j->sk, j->p[0], j->p[1], j->p[2],
                               j->lo, j->hi);
    break;
  }
  return 0;
}

/* kind 0: hori_diff_type2 p={out,in,crlato,crlatu,hdmask,lap}; 1: tridiagonal p={d,a,b,c};
 * 2: hori_diff_01 p={u,out,lap}.  Returns 0 on success. */
int oracle_run_threaded(int kind, int nthreads, int ni, int nj, int nk, int h, size_t sj, size_t sk,
                        double** p, int np) {
  if(nthreads < 1)
    nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthreads);
  int lo = (kind == 1) ? h : 0, hi = (kind == 1) ? ni - h - 1 : nk - 1;
  int n = hi - lo + 1;
  for(int t = 0; t < nthreads; ++t) {
    job_t* j = &jobs[t];
    j->kind = kind, j->ni = ni, j->nj = nj, j->nk = nk, j->h = h, j->sj = sj, j->sk = sk;
    j->lo = lo + (int)((long)n * t / nthreads);
    j->hi = lo + (int)((long)n * (t + 1) / nthreads) - 1;
    for(int q = 0; q < 6; ++q)
      j->p[q] = q < np ? p[q] : 0;
    pthread_create(&th[t], 0, worker, j);
  }
  for(int t = 0; t < nthreads; ++t)
    pthread_join(th[t], 0);
  free(th);
  free(jobs);
  return 0;
}
