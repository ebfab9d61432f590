// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  The reference's pre-generated CUDA text of its largest
// code-generation fixture, dawn/test/unit-test/dawn/CodeGen/reference/update_dz_c.cu (7 multistages, 5-D
// block-private temporaries, a versioned field the wrapper allocates itself), compiled UNMODIFIED for
// sm_100a.  Fields are DENSE device buffers ni x nj x (nk+1), i fastest (the wrapper's own gz_0 gets that
// layout from the stand-in, and the kernels take one stride set for all ijk fields).
#include REF_TEXT

extern "C" int refcuda_update_dz_c(int ni, int nj, int nk, int halo, double dt, double* const* f9) {
  using namespace gridtools::dawn;
  domain dom(ni, nj, nk);
  dom.set_halos(halo, halo, halo, halo, 0, 0);
  dawn_generated::cuda::update_dz_c st(dom);
  st.set_dt(dt);
  storage_ijk_t s[9];
  for(int n = 0; n < 9; ++n)
    s[n] = storage_ijk_t(f9[n], ni, ni * nj);
  st.run(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8]);
  return (int)cudaDeviceSynchronize() + (int)cudaGetLastError();
}
