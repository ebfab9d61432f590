// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  One of the reference's small pre-generated CUDA texts
// (two ijk fields: in, out), compiled UNMODIFIED for sm_100a; built once per text (the texts reuse class
// names, so each gets its own shared library):
//   dawn/test/unit-test/dawn/CodeGen/reference/{conditional_stencil,global_indexing,nonoverlapping_stencil}.cu
//   dawn/test/integration-test/dawn4py-tests/data/copy_stencil_ref.cpp
// Macros: REF_TEXT (file), REF_CLASS (wrapper class in dawn_generated::cuda), REF_HAS_VAR1 (global var1).
#include REF_TEXT

extern "C" int refcuda_run(int ni, int nj, int nk, int halo, int stride_j, int stride_k, double* in,
                           double* out, int var1) {
  using namespace gridtools::dawn;
  domain dom(ni, nj, nk);
  dom.set_halos(halo, halo, halo, halo, 0, 0);
  dawn_generated::cuda::REF_CLASS st(dom);
#ifdef REF_HAS_VAR1
  st.set_var1(var1);
#else
  (void)var1;
#endif
  storage_ijk_t s_in(in, stride_j, stride_k), s_out(out, stride_j, stride_k);
  st.run(s_in, s_out);
  return (int)cudaDeviceSynchronize() + (int)cudaGetLastError();
}
