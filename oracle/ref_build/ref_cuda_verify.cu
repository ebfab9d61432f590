// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  The reference's own GPU field verifier
// (dawn/src/driver-includes/cuda_verify.hpp:84-196, 2 kernels + 8 thrust reductions), compiled UNMODIFIED
// for sm_100a from where it lies under /root/reference, behind a C entry point over caller-owned device
// buffers: the parity partner of the single-pass b200rt_verify_field.
#include <cstdio> // cuda_utils.hpp uses fprintf without including it
#include <driver-includes/cuda_verify.hpp>

#include <iostream>
#include <sstream>

extern "C" int refcuda_verify_field(int n, const double* dsl, const double* actual, double rel_tol, double abs_tol,
                                    double* out4, int* is_valid) {
  std::streambuf* keep = std::cout.rdbuf();
  std::ostringstream sink; // verify_field prints its metrics
  std::cout.rdbuf(sink.rdbuf());
  VerificationMetrics m = dawn::verify_field<double>(0, n, dsl, actual, "field", rel_tol, abs_tol, 0);
  std::cout.rdbuf(keep);
  out4[0] = m.maxRelErr, out4[1] = m.minRelErr, out4[2] = m.maxAbsErr, out4[3] = m.minAbsErr;
  *is_valid = m.isValid ? 1 : 0;
  return (int)cudaDeviceSynchronize() + (int)cudaGetLastError();
}
