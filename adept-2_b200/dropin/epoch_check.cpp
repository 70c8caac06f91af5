// epoch_check.cpp -- the drop-in's tape cache with adept_recording_epoch.patch applied to the reference headers
// (optimiser-style loops on ONE tape).  Plain Adept-2 API throughout.
//
//   epoch_check <n> <out.bin>
//
// 1. test/test_derivatives.cpp:25-29 pattern: one recording of y = f(x) (n inputs), then n forward() sweeps and n
//    reverse() sweeps on the same tape: with the patch the tape is uploaded and analysed ONCE (without it: 2n times).
// 2. the same program re-recorded at a different x -- a tape of IDENTICAL size at the same addresses -- must be
//    detected (the new multipliers must be used): the case a pointer/size guess gets wrong.
// 3. recording continues after a sweep (y2 = g(y)): the grown tape must be re-uploaded.
// Every sweep result is dumped; tests/test_gpu_dropin.py checks them against analytic derivatives and the upload counts.
#include <adept_arrays.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using adept::adouble;
using adept::Real;

extern "C" unsigned long long adept_cuda_dropin_uploads();

static adouble f(const adouble& x) { return sin(x) * x + exp(-x); }
static double dfdx(double x) { return cos(x) * x + sin(x) - exp(-x); }

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 16;
  const char* path = argc > 2 ? argv[2] : "epoch_dump.bin";
  FILE* out = fopen(path, "wb");
  if (!out) return 3;
  adept::Stack stack;
  std::vector<double> fwd(n), rev(n), fwd2(n), grown(n);
  std::vector<adouble> x(n), y(n);
  unsigned long long uploads[4];
  // 1. one tape, 2n sweeps
  for (int i = 0; i < n; ++i) x[i] = 0.3 + 0.1 * i;
  stack.new_recording();
  for (int i = 0; i < n; ++i) y[i] = f(x[i]);
  const unsigned long long u0 = adept_cuda_dropin_uploads();
  for (int i = 0; i < n; ++i) {
    stack.clear_gradients();
    x[i].set_gradient(1.0);
    stack.forward();
    y[i].get_gradient(fwd[i]);
  }
  for (int i = 0; i < n; ++i) {
    stack.clear_gradients();
    y[i].set_gradient(1.0);
    stack.reverse();
    x[i].get_gradient(rev[i]);
  }
  uploads[0] = adept_cuda_dropin_uploads() - u0;
  // 2. same program, same sizes, other values
  for (int i = 0; i < n; ++i) x[i] = 1.7 - 0.05 * i;
  stack.new_recording();
  for (int i = 0; i < n; ++i) y[i] = f(x[i]);
  const unsigned long long u1 = adept_cuda_dropin_uploads();
  for (int i = 0; i < n; ++i) {
    stack.clear_gradients();
    x[i].set_gradient(1.0);
    stack.forward();
    y[i].get_gradient(fwd2[i]);
  }
  uploads[1] = adept_cuda_dropin_uploads() - u1;
  // 3. the recording goes on after a sweep
  std::vector<adouble> z(n);
  for (int i = 0; i < n; ++i) z[i] = 3.0 * y[i] * y[i];
  const unsigned long long u2 = adept_cuda_dropin_uploads();
  for (int i = 0; i < n; ++i) {
    stack.clear_gradients();
    x[i].set_gradient(1.0);
    stack.forward();
    z[i].get_gradient(grown[i]);
  }
  uploads[2] = adept_cuda_dropin_uploads() - u2;
#ifdef ADEPT_RECORDING_EPOCH
  uploads[3] = 1;
#else
  uploads[3] = 0;
#endif
  fwrite(uploads, 8, 4, out);
  std::vector<double> want(4 * n);
  for (int i = 0; i < n; ++i) {
    const double xa = 0.3 + 0.1 * i, xb = 1.7 - 0.05 * i;
    want[i] = dfdx(xa);
    want[n + i] = dfdx(xa);
    want[2 * n + i] = dfdx(xb);
    const double yb = sin(xb) * xb + exp(-xb);
    want[3 * n + i] = 6.0 * yb * dfdx(xb);
  }
  fwrite(&fwd[0], 8, n, out); fwrite(&rev[0], 8, n, out); fwrite(&fwd2[0], 8, n, out); fwrite(&grown[0], 8, n, out);
  fwrite(&want[0], 8, 4 * n, out);
  fclose(out);
  printf("epoch_check: uploads %llu / %llu / %llu (patched headers: %llu)\n", uploads[0], uploads[1], uploads[2], uploads[3]);
  return 0;
}
