// TEST / BENCH INFRASTRUCTURE ONLY (oracle/_ref build).  The reference's own pre-generated CUDA code,
// compiled UNMODIFIED for sm_100a from where it lies under /root/reference:
//   dawn/test/integration-test/dawn4py-tests/data/hori_diff_stencil_ref.cpp          (4th-order hori-diff)
//   dawn/test/integration-test/dawn4py-tests/data/tridiagonal_solve_stencil_ref.cpp  (Thomas solver)
// behind a C ABI over caller-owned device buffers: what the reference's CUDA backend would run on a B200
// if it were simply recompiled - the baseline the b200 backend is measured against, and a second parity
// check (same inputs, bit-identical outputs; both are compiled with -fmad=false).
#include REF_CUDA_HORI_DIFF
#include REF_CUDA_TRIDIAGONAL

namespace {
gridtools::dawn::domain make_domain(int ni, int nj, int nk, int halo) {
  gridtools::dawn::domain dom(ni, nj, nk);
  dom.set_halos(halo, halo, halo, halo, 0, 0);
  return dom;
}
} // namespace

extern "C" {
// fields: element (0,0,0) of storages of ni x nj x (>= nk) elements incl. halos, i-stride 1.
// Runs the generated wrapper's run() `reps` times; *seconds = what its own timer_cuda accumulated.
int refcuda_hori_diff(int ni, int nj, int nk, int halo, int stride_j, int stride_k, double* in, double* out,
                      double* coeff, int reps, double* seconds) {
  using namespace gridtools::dawn;
  dawn_generated::cuda::hori_diff_stencil st(make_domain(ni, nj, nk, halo));
  storage_ijk_t s_in(in, stride_j, stride_k), s_out(out, stride_j, stride_k), s_coeff(coeff, stride_j, stride_k);
  for(int r = 0; r < reps; ++r)
    st.run(s_in, s_out, s_coeff);
  if(seconds)
    *seconds = st.get_total_time();
  return (int)cudaDeviceSynchronize() + (int)cudaGetLastError();
}

int refcuda_tridiagonal_solve(int ni, int nj, int nk, int halo, int stride_j, int stride_k, double* a,
                              double* b, double* c, double* d, int reps, double* seconds) {
  using namespace gridtools::dawn;
  dawn_generated::cuda::tridiagonal_solve_stencil st(make_domain(ni, nj, nk, halo));
  storage_ijk_t s_a(a, stride_j, stride_k), s_b(b, stride_j, stride_k), s_c(c, stride_j, stride_k),
      s_d(d, stride_j, stride_k);
  for(int r = 0; r < reps; ++r)
    st.run(s_a, s_b, s_c, s_d);
  if(seconds)
    *seconds = st.get_total_time();
  return (int)cudaDeviceSynchronize() + (int)cudaGetLastError();
}
}
