// ensemble_bench.cpp -- BASELINE config 3 end to end: the reference's radiance model (test/simulate_radiances.cpp) for
// an ensemble of state vectors, recorded on all host cores through adept_cuda::record_ensemble (one adept::Stack per
// thread, plain Adept API) and differentiated by adept_cuda_jacobian_batch with the members sharded over the GPUs.
//
//   ensemble_bench <state.bin> <n_members> <n_devices> <out.bin> [n_threads=0] [reps=3] [chunks=4]
//
// <state.bin>: n_members x 26 doubles (surface temperature, 25 layer temperatures), written by the caller so that the
// CPU arm (the reference's own simulate_radiances() under OpenMP, oracle/ref_harness.cpp) sees the same ensemble.
// <out.bin>: 8 int64 {n_members, n_indep, n_dep, n_devices, threads, reps, rc, pinned} + 4 doubles {best record ms,
// best jacobian ms, best total ms, mean total ms} + the Jacobians [member][indep][dep] of the last repetition.
//
// The model function is the reference's own: simulate_radiances.cpp is #included (not copied) so that its file-static
// simulate_radiance_private() can be recorded member by member; the few lines around it restate what
// simulate_radiances() sets up before recording (test/simulate_radiances.cpp:73-128: oxygen density, emissivity).
#include <adept.h>
#include <simulate_radiances.cpp>   // from $(ADEPT_REF)/test: the model, unmodified

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "adept_cuda_ensemble.h"

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <state.bin> <n_members> <n_devices> <out.bin> [n_threads] [reps] [chunks]\n", argv[0]);
    return 2;
  }
  const int n = 25;
  const int64_t B = atoll(argv[2]);
  const int n_dev = atoi(argv[3]);
  const int threads = argc > 5 ? atoi(argv[5]) : 0;
  const int reps = argc > 6 ? atoi(argv[6]) : 3;
  const int chunks = argc > 7 ? atoi(argv[7]) : 4;
  std::vector<double> state(size_t(B) * (n + 1));
  {
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(state.data(), 8, state.size(), f) != state.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }
    fclose(f);
  }
  // what simulate_radiances() computes before it starts recording (test/simulate_radiances.cpp:73-81, :97-105, :113-115)
  static const adept::Real wavelength[2] = {0.006, 0.0061};
  static const adept::Real mass_abs_coefft[2] = {3.0e-5, 3.0e-3};
  static const adept::Real dz = 1000.0;
  std::vector<adept::Real> emissivity[2];
  for (int ichan = 0; ichan < 2; ichan++) {
    emissivity[ichan].resize(n);
    for (int i = 0; i < n; i++) {
      const adept::Real density_oxygen = 1.2 * 0.21 * (16.0 / 29.0) * exp(-(i * dz) / 8000.0);
      emissivity[ichan][i] = 1.0 - exp(-density_oxygen * mass_abs_coefft[ichan] * dz);
    }
  }
  std::vector<int> devs(n_dev);
  for (int k = 0; k < n_dev; ++k) devs[k] = k;
  adept_cuda_ctx* ctx = 0;
  int rc = adept_cuda_create(&ctx, devs.data(), n_dev);
  if (rc) { fprintf(stderr, "adept_cuda_create failed (%d)\n", rc); return 4; }
  adept_cuda_set_option(ctx, "batch_chunks", chunks);
  adept_cuda::EnsembleTape tape;
  double* jac = 0;
  double best_rec = 1e30, best_jac = 1e30, best_tot = 1e30, sum_tot = 0;
  int used_threads = threads > 0 ? threads : omp_get_max_threads();
  for (int rep = 0; rep < reps + 1 && !rc; ++rep) {   // the first repetition warms up (allocations, page-locking)
    const double t0 = now_ms();
    adept_cuda::record_ensemble(B, [&](int64_t b, adept::Stack& stack) {
      const double* x = &state[size_t(b) * (n + 1)];
      adept::aReal st = x[0];
      std::vector<adept::aReal> t(n);
      adept::aReal r[2];
      for (int i = 0; i < n; i++) t[i] = x[1 + i];
      stack.new_recording();
      for (int ichan = 0; ichan < 2; ichan++)
        r[ichan] = simulate_radiance_private(n, wavelength[ichan], &emissivity[ichan][0], st, &t[0]);
      stack.independent(st);
      stack.independent(&t[0], n);
      stack.dependent(r, 2);
    }, tape, threads);
    const double t1 = now_ms();
    if (!jac) {
      jac = static_cast<double*>(adept_cuda_host_alloc(size_t(B) * tape.indep.size() * tape.dep.size() * 8));
      if (!jac) { rc = 5; break; }
    }
    rc = adept_cuda::ensemble_jacobian(ctx, ADEPT_CUDA_REVERSE, tape, jac);   // 2 dependents < 26 independents: reverse
    const double t2 = now_ms();
    if (rc) { fprintf(stderr, "ensemble_jacobian: %s\n", adept_cuda_last_error(ctx)); break; }
    if (rep > 0) {
      best_rec = std::min(best_rec, t1 - t0);
      best_jac = std::min(best_jac, t2 - t1);
      best_tot = std::min(best_tot, t2 - t0);
      sum_tot += t2 - t0;
    }
  }
  FILE* f = fopen(argv[4], "wb");
  if (!f) return 6;
  const int64_t hdr[8] = {B, int64_t(tape.indep.size()), int64_t(tape.dep.size()), n_dev, used_threads, reps, rc,
                          tape.multipliers.pinned() ? 1 : 0};
  const double tm[4] = {best_rec, best_jac, best_tot, reps > 0 ? sum_tot / reps : 0.0};
  fwrite(hdr, 8, 8, f);
  fwrite(tm, 8, 4, f);
  if (!rc && jac) fwrite(jac, 8, size_t(B) * tape.indep.size() * tape.dep.size(), f);
  fclose(f);
  printf("ensemble_bench: %lld members on %d device(s), %d host threads: record %.2f ms, jacobian %.2f ms, total %.2f ms (best of %d), "
         "%.3e members/s, tape %lld statements / %lld operations, rc %d\n", (long long)B, n_dev, used_threads, best_rec, best_jac,
         best_tot, reps, B / (best_tot / 1e3), (long long)tape.n_statements(), (long long)tape.n_operations(), rc);
  if (jac) adept_cuda_host_free(jac);
  adept_cuda_destroy(ctx);
  return rc;
}
