// End-to-end benchmark driver through the GENERATED C++ CLASS, written the way a Dawn user writes a driver
// (gtclang/test/integration-test/CodeGen/hori_diff_type2_stencil_benchmark.cpp:40-73): domain, storages,
// fillMath inputs, run().  Every step the host hands new inputs to the storages (write access to the host
// views marks them newer than the device copy), calls run() - which moves them to the device, launches and,
// as the reference's run() does with --run-with-sync, brings the result back - and reads the output on the
// host.  With set_host_pipeline(levels) run() streams the domain through the device in chunks of levels.
// Prints one JSON object; the result of the last step is compared with the oracle's C restatement (this
// binary is bench / test infrastructure: the only place generated code and oracle/ are linked together).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include GENERATED_FILE
#include <driver-includes/verify.hpp>

extern "C" void oracle_run_threaded(int kind, int threads, int ni, int nj, int nk, int h, size_t sj, size_t sk,
                                    double** fields, int nfields);

using namespace gridtools::dawn;

int main(int argc, char** argv) {
  const int I = argc > 1 ? std::atoi(argv[1]) : 1024, J = argc > 2 ? std::atoi(argv[2]) : 1024,
            K = argc > 3 ? std::atoi(argv[3]) : 80, steps = argc > 4 ? std::atoi(argv[4]) : 5,
            chunk = argc > 5 ? std::atoi(argv[5]) : 8, threads = argc > 6 ? std::atoi(argv[6]) : 8;
  domain dom(I + 2 * halo::value, J + 2 * halo::value, K);
  dom.set_halos(halo::value, halo::value, halo::value, halo::value, 0, 0);
  verifier verif(dom);
  meta_data_t meta_data(dom.isize(), dom.jsize(), dom.ksize() + 1);
  meta_data_j_t meta_data_j(1, dom.jsize(), 1);
  storage_t u(meta_data, "u"), hdmask(meta_data, "hdmask"), u_out(meta_data, "u_out");
  storage_j_t crlato(meta_data_j, "crlato"), crlatu(meta_data_j, "crlatu");
  verif.fillMath(8.0, 2.0, 1.5, 1.5, 2.0, 4.0, u);
  verif.fillMath(5.0, 2.2, 1.7, 1.9, 2.0, 4.0, hdmask);
  verif.fillMath(6.5, 1.2, 1.7, 0.9, 2.1, 2.0, crlato);
  verif.fillMath(5.0, 2.2, 1.7, 0.9, 2.0, 1.0, crlatu);
  verif.fill(-1.0, u_out);

  dawn_generated::b200::hori_diff_type2_stencil hd(dom);
  hd.set_host_pipeline(chunk);
  const int ni = dom.isize(), nj = dom.jsize(), nka = dom.ksize() + 1;
  auto touch_inputs = [&](int step) {
    // the host produces this step's inputs: one value per field changes (cheap), all four count as new
    auto vu = gridtools::make_host_view(u);
    auto vm = gridtools::make_host_view(hdmask);
    auto vo = gridtools::make_host_view(crlato);
    auto vc = gridtools::make_host_view(crlatu);
    vu(halo::value, halo::value, 0) += 1e-9 * step;
    vm(halo::value, halo::value, 0) += 0.0;
    vo(0, halo::value, 0) += 0.0;
    vc(0, halo::value, 0) += 0.0;
  };
  double sink = 0;
  touch_inputs(0);
  hd.run(u_out, u, crlato, crlatu, hdmask); // warm-up: allocates the device copies, uploads `out` once
  auto t0 = std::chrono::steady_clock::now();
  for(int s = 1; s <= steps; ++s) {
    touch_inputs(s);
    hd.run(u_out, u, crlato, crlatu, hdmask);
    auto vr = gridtools::make_host_view<gridtools::access_mode::read_only>(u_out);
    sink += vr(halo::value + 1, halo::value + 1, K - 1); // the result is on the host
  }
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps;

  // parity of the last step against the oracle on the same host data
  std::vector<double> d_in((size_t)ni * nj * nka), d_mask(d_in.size()), d_out(d_in.size(), -1.0), d_lap(d_in.size(), 0.0),
      d_cro(nj), d_cru(nj);
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
  double* fields[6] = {d_out.data(), d_in.data(), d_cro.data(), d_cru.data(), d_mask.data(), d_lap.data()};
  oracle_run_threaded(0, threads, ni, nj, K, halo::value, ni, (size_t)ni * nj, fields, 6);
  auto vg = gridtools::make_host_view<gridtools::access_mode::read_only>(u_out);
  long diff = 0;
  for(int k = 0; k < nka; ++k)
    for(int j = 0; j < nj; ++j)
      for(int i = 0; i < ni; ++i)
        diff += vg(i, j, k) != d_out[(size_t)i + (size_t)j * ni + (size_t)k * ni * nj];
  // the result as the INPUT of another run while it is newer on the device (non-pipelined run leaves it there):
  // the pipelined run must neither upload the stale host copy nor mark it coherent
  long diff2 = 0;
  {
    storage_t a(meta_data, "a"), b(meta_data, "b");
    verif.fill(-1.0, a, b);
    hd.set_host_pipeline(0);
    hd.run(a, u, crlato, crlatu, hdmask);        // a = f(u), whole storages; run() syncs a to the host
    {                                            // make the device copy of `a` the only good one
      auto ro = gridtools::make_host_view<gridtools::access_mode::read_only>(a);
      for(int k = 0; k < nka; ++k)
        for(int j = 0; j < nj; ++j)
          for(int i = 0; i < ni; ++i)
            ro(i, j, k) = 12345.0;               // host mirror scribbled over behind the storage's back
      a.mark_device_dirty(nullptr);
    }
    hd.set_host_pipeline(chunk);
    { auto vm2 = gridtools::make_host_view(hdmask); vm2(halo::value, halo::value, 0) += 0.0; } // hdmask newer on host
    hd.run(b, a, crlato, crlatu, hdmask);        // b = f(a) through the pipeline, a already on the device
    std::vector<double> d_a(d_out), d_b(d_in.size(), -1.0);
    std::fill(d_lap.begin(), d_lap.end(), 0.0);
    double* f2[6] = {d_b.data(), d_a.data(), d_cro.data(), d_cru.data(), d_mask.data(), d_lap.data()};
    oracle_run_threaded(0, threads, ni, nj, K, halo::value, ni, (size_t)ni * nj, f2, 6);
    auto vb = gridtools::make_host_view<gridtools::access_mode::read_only>(b);
    auto va = gridtools::make_host_view<gridtools::access_mode::read_only>(a);
    for(int k = 0; k < nka; ++k)
      for(int j = 0; j < nj; ++j)
        for(int i = 0; i < ni; ++i) {
          diff2 += vb(i, j, k) != d_b[(size_t)i + (size_t)j * ni + (size_t)k * ni * nj];
          diff2 += va(i, j, k) != d_a[(size_t)i + (size_t)j * ni + (size_t)k * ni * nj];
        }
  }
  diff += diff2;
  const double h2d = 2.0 * sizeof(double) * ((double)meta_data.padded_total_length()) * K / nka +
                     2.0 * sizeof(double) * meta_data_j.padded_total_length();
  const double d2h = sizeof(double) * ((double)meta_data.padded_total_length()) * K / nka;
  std::printf("{\"seconds_per_step\": %.6f, \"updates_per_s\": %.6e, \"steps\": %d, \"chunk_levels\": %d, "
              "\"h2d_bytes_per_step\": %.0f, \"d2h_bytes_per_step\": %.0f, \"differing_values\": %ld, "
              "\"launches\": %ld, \"sink\": %.3f}\n",
              sec, (double)I * J * K / sec, steps, chunk, h2d, d2h, diff, 0L, sink);
  return diff == 0 ? 0 : 1;
}
