// C++ driver in the shape of the reference's integration tests
// (gtclang/test/integration-test/CodeGen/hori_diff_type2_stencil_benchmark.cpp:40-73): domain with
// halos, storages, fillMath inputs, run the generated class, compare with the naive result.
// The generated b200 translation unit is #included exactly like the reference includes its generated
// files; the naive side is the oracle's C restatement (tests may link oracle/).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include GENERATED_FILE
#include <driver-includes/verify.hpp>

extern "C" void oracle_hori_diff_type2(int ni, int nj, int nk, int h, size_t sj, size_t sk, double* out,
                                       const double* in, const double* crlato, const double* crlatu,
                                       const double* hdmask, double* lap);

using namespace gridtools::dawn;

int main(int argc, char** argv) {
  const int I = argc > 1 ? std::atoi(argv[1]) : 12, J = argc > 2 ? std::atoi(argv[2]) : 12,
            K = argc > 3 ? std::atoi(argv[3]) : 10;
  domain dom(I + 2 * halo::value, J + 2 * halo::value, K);
  dom.set_halos(halo::value, halo::value, halo::value, halo::value, 0, 0);
  verifier verif(dom);

  meta_data_t meta_data(dom.isize(), dom.jsize(), dom.ksize() + 1);
  meta_data_j_t meta_data_j(1, dom.jsize(), 1);
  storage_t u(meta_data, "u"), hdmask(meta_data, "hdmask"), u_out(meta_data, "u_out"),
      u_out_naive(meta_data, "u_out_naive");
  storage_j_t crlato(meta_data_j, "crlato"), crlatu(meta_data_j, "crlatu");

  verif.fillMath(8.0, 2.0, 1.5, 1.5, 2.0, 4.0, u);
  verif.fillMath(5.0, 2.2, 1.7, 1.9, 2.0, 4.0, hdmask);
  verif.fillMath(6.5, 1.2, 1.7, 0.9, 2.1, 2.0, crlato);
  verif.fillMath(5.0, 2.2, 1.7, 0.9, 2.0, 1.0, crlatu);
  verif.fill(-1.0, u_out, u_out_naive);

  dawn_generated::b200::hori_diff_type2_stencil hd(dom);
  hd.run(u_out, u, crlato, crlatu, hdmask);
  hd.run(u_out, u, crlato, crlatu, hdmask); // a second run: timers accumulate, result unchanged

  // naive side on dense copies
  const int ni = dom.isize(), nj = dom.jsize(), nka = dom.ksize() + 1;
  std::vector<double> d_in((size_t)ni * nj * nka), d_mask(d_in.size()), d_out(d_in.size(), -1.0),
      d_lap(d_in.size(), 0.0), d_cro(nj), d_cru(nj);
  auto vu = gridtools::make_host_view<gridtools::access_mode::read_only>(u);
  auto vm = gridtools::make_host_view<gridtools::access_mode::read_only>(hdmask);
  auto vo = gridtools::make_host_view<gridtools::access_mode::read_only>(crlato);
  auto vc = gridtools::make_host_view<gridtools::access_mode::read_only>(crlatu);
  for(int k = 0; k < nka; ++k)
    for(int j = 0; j < nj; ++j)
      for(int i = 0; i < ni; ++i) {
        d_in[(size_t)i + (size_t)j * ni + (size_t)k * ni * nj] = vu(i, j, k);
        d_mask[(size_t)i + (size_t)j * ni + (size_t)k * ni * nj] = vm(i, j, k);
      }
  for(int j = 0; j < nj; ++j)
    d_cro[j] = vo(0, j, 0), d_cru[j] = vc(0, j, 0);
  oracle_hori_diff_type2(ni, nj, K, halo::value, ni, (size_t)ni * nj, d_out.data(), d_in.data(),
                         d_cro.data(), d_cru.data(), d_mask.data(), d_lap.data());
  auto vn = gridtools::make_host_view(u_out_naive);
  for(int k = 0; k < nka; ++k)
    for(int j = 0; j < nj; ++j)
      for(int i = 0; i < ni; ++i)
        vn(i, j, k) = d_out[(size_t)i + (size_t)j * ni + (size_t)k * ni * nj];

  bool ok = verif.verify(u_out, u_out_naive);
  // stricter than the reference's 1e-10: bit equality
  auto vg = gridtools::make_host_view<gridtools::access_mode::read_only>(u_out);
  long diff = 0;
  for(int k = 0; k < nka; ++k)
    for(int j = 0; j < nj; ++j)
      for(int i = 0; i < ni; ++i)
        diff += vg(i, j, k) != vn(i, j, k);
  // the same on a stream of the caller's (non-blocking: the legacy null stream does not wait for it):
  // host views and sync() must still see the finished result (the reference's run() synchronises
  // the device, CudaCodeGen.cpp:435-450; here the storage remembers the producing stream)
  void* user_stream = nullptr;
  if(b200rt_stream_create(&user_stream)) return 2;
  storage_t u_out2(meta_data, "u_out2");
  verif.fill(-1.0, u_out2);
  dawn_generated::b200::hori_diff_type2_stencil hd2(dom);
  hd2.set_stream(user_stream);
  hd2.run(u_out2, u, crlato, crlatu, hdmask);
  auto vg2 = gridtools::make_host_view<gridtools::access_mode::read_only>(u_out2);
  long diff2 = 0;
  for(int k = 0; k < nka; ++k)
    for(int j = 0; j < nj; ++j)
      for(int i = 0; i < ni; ++i)
        diff2 += vg2(i, j, k) != vn(i, j, k);
  std::printf("%s %s total_time=%.6f s, %ld differing values, %ld on a user stream\n", hd.get_name().c_str(),
              (ok && diff == 0 && diff2 == 0) ? "VERIFIED" : "FAILED", hd.get_total_time(), diff, diff2);
  return (ok && diff == 0 && diff2 == 0) ? 0 : 1;
}
