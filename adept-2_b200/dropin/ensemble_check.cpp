// ensemble_check.cpp -- exercises adept_cuda_ensemble.h with the plain Adept-2 API.
//
//   ensemble_check <n_members> <out.bin> [cpu|gpu] [n_threads]
//
// A small NONLINEAR model (so that the multipliers differ between members while the structure does not),
// nx inputs -> 3 outputs, is recorded for every member
//   (a) through adept_cuda::record_ensemble (all host cores, one Stack per thread), and
//   (b) serially, member after member on ONE Stack, copying each tape out with no helper code,
// and both are dumped so that the test can compare them.  The analytic Jacobian of every member is dumped too.
// In `gpu` mode the ensemble Jacobians computed by adept_cuda_jacobian_batch (reverse AND forward sweeps) are
// appended; `cpu` mode runs no kernel (the helper merely asks libadept_cuda for page-locked memory and falls back to
// the heap when there is no CUDA driver).
//
// The reference's own ensemble use case is test/simulate_radiances.cpp:52-146 (one call per state vector);
// the radiance model is linear in temperature, which is why this check uses a model of its own.
#include <adept.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "adept_cuda_ensemble.h"

using adept::adouble;
using adept::Real;

static const int NX = 8, NY = 3;

static void member_state(int64_t b, Real* x) {
  for (int i = 0; i < NX; ++i) x[i] = 0.3 + 0.05 * i + 0.001 * static_cast<Real>((b * 37 + i * 11) % 997);
}

// y_k = sum_i sin(w_ik x_i) * x_{(i+1) % NX}  +  exp(-x_k) * x_0 * x_0 ;   y_2 is then scaled in place (y *= c)
static Real weight(int i, int k) { return 0.5 + 0.25 * k + 0.1 * i; }

template <class T>
static void model(const T* x, T* y) {
  for (int k = 0; k < NY; ++k) {
    T acc = 0.0;
    for (int i = 0; i < NX; ++i) acc += sin(weight(i, k) * x[i]) * x[(i + 1) % NX];
    y[k] = acc + exp(-x[k]) * x[0] * x[0];
  }
  y[2] *= 1.5;
}

// dy_k/dx_j, column-major n_dep x n_indep: jac[k + j*NY]
static void analytic_jacobian(const Real* x, Real* jac) {
  for (int k = 0; k < NY; ++k) {
    const Real scale = (k == 2) ? 1.5 : 1.0;
    for (int j = 0; j < NX; ++j) {
      Real d = weight(j, k) * cos(weight(j, k) * x[j]) * x[(j + 1) % NX];   // through sin(w x_j)
      const int i = (j + NX - 1) % NX;                                      // x_j is the factor of term i = j-1
      d += sin(weight(i, k) * x[i]);
      if (j == k) d += -exp(-x[k]) * x[0] * x[0];
      if (j == 0) d += exp(-x[k]) * 2.0 * x[0];
      jac[k + j * NY] = scale * d;
    }
  }
}

static void record_member(int64_t b, adept::Stack& stack) {
  Real xv[NX];
  member_state(b, xv);
  adouble x[NX], y[NY];
  for (int i = 0; i < NX; ++i) x[i] = xv[i];
  stack.new_recording();
  model(x, y);
  stack.independent(x, NX);
  stack.dependent(y, NY);
}

// Reads the protected tape arrays of a Stack with no help from the ensemble header.
class PeekStack : public adept::Stack {
 public:
  void copy_out(std::vector<int32_t>& stmt, std::vector<int32_t>& idx, std::vector<double>& mult, int64_t& max_grad) const {
    stmt.resize(2 * n_statements_);
    for (int s = 0; s < static_cast<int>(n_statements_); ++s) {
      stmt[2 * s] = static_cast<int32_t>(statement_[s].index);
      stmt[2 * s + 1] = static_cast<int32_t>(statement_[s].end_plus_one);
    }
    idx.assign(index_, index_ + n_operations_);
    mult.assign(multiplier_, multiplier_ + n_operations_);
    max_grad = max_gradient_;
  }
};

template <class T>
static void put(FILE* f, const std::vector<T>& v) {
  if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
}
static void put(FILE* f, const adept_cuda::HostBuffer& v) {
  if (v.size()) fwrite(v.data(), sizeof(double), v.size(), f);
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <n_members> <out.bin> [cpu|gpu] [n_threads]\n", argv[0]);
    return 2;
  }
  const int64_t B = atoll(argv[1]);
  const bool gpu = argc > 3 && std::string(argv[3]) == "gpu";
  const int n_threads = argc > 4 ? atoi(argv[4]) : 0;

  // (a) the helper
  adept_cuda::EnsembleTape t;
  const double t_helper0 = omp_get_wtime();
  try {
    adept_cuda::record_ensemble(B, record_member, t, n_threads);
  } catch (const std::exception& e) {
    fprintf(stderr, "record_ensemble failed: %s\n", e.what());
    return 1;
  }
  const double helper_s = omp_get_wtime() - t_helper0;
  // a model with a data-dependent branch must be refused, not packed
  int64_t refused = 0;
  try {
    adept_cuda::EnsembleTape bad;
    adept_cuda::record_ensemble(4, [](int64_t b, adept::Stack& stack) {
      adouble x = 1.0 + b, y;
      stack.new_recording();
      y = (b == 2) ? adouble(x * x + x) : adouble(x * x);   // member 2 records one operation more
      stack.independent(x);
      stack.dependent(y);
    }, bad, 2);
  } catch (const adept_cuda::ensemble_structure_mismatch&) {
    refused = 1;
  }

  // (b) serial, one Stack, no helper
  std::vector<double> serial_mult(static_cast<size_t>(t.n_operations()) * B);   // [member][operation]
  std::vector<int32_t> stmt0, idx0;
  int64_t max_grad0 = 0, serial_same_structure = 1;
  const double t_serial0 = omp_get_wtime();
  {
    PeekStack s;
    for (int64_t b = 0; b < B; ++b) {
      record_member(b, s);
      std::vector<int32_t> st, ix;
      std::vector<double> mu;
      int64_t mg;
      s.copy_out(st, ix, mu, mg);
      if (b == 0) { stmt0 = st; idx0 = ix; max_grad0 = mg; }
      if (st != stmt0 || ix != idx0 || mg != max_grad0 || static_cast<int64_t>(mu.size()) != t.n_operations()) {
        serial_same_structure = 0;
        continue;
      }
      for (size_t o = 0; o < mu.size(); ++o) serial_mult[static_cast<size_t>(b) * mu.size() + o] = mu[o];
    }
  }
  const double serial_s = omp_get_wtime() - t_serial0;
  printf("members %lld  statements %lld  operations %lld  threads %d\n"
         "record_ensemble %.3f ms (%.3e members/s)   serial recording on one Stack %.3f ms (%.3e members/s)\n",
         static_cast<long long>(B), static_cast<long long>(t.n_statements() - 1), static_cast<long long>(t.n_operations()),
         n_threads > 0 ? n_threads : omp_get_max_threads(), 1e3 * helper_s, B / helper_s, 1e3 * serial_s, B / serial_s);
  std::vector<double> analytic(static_cast<size_t>(B) * NX * NY);
  for (int64_t b = 0; b < B; ++b) {
    Real xv[NX];
    member_state(b, xv);
    analytic_jacobian(xv, &analytic[static_cast<size_t>(b) * NX * NY]);
  }

  std::vector<double> jac_rev, jac_fwd;
  int64_t gpu_rc = -1;
  if (gpu) {
    adept_cuda_ctx* ctx = nullptr;
    int dev = 0;
    gpu_rc = adept_cuda_create(&ctx, &dev, 1);
    if (gpu_rc == 0) {
      jac_rev.assign(analytic.size(), 0.0);
      jac_fwd.assign(analytic.size(), 0.0);
      gpu_rc = adept_cuda::ensemble_jacobian(ctx, ADEPT_CUDA_REVERSE, t, jac_rev.data());
      if (gpu_rc == 0) gpu_rc = adept_cuda::ensemble_jacobian(ctx, ADEPT_CUDA_FORWARD, t, jac_fwd.data());
      if (gpu_rc) fprintf(stderr, "adept_cuda_jacobian_batch: %s\n", adept_cuda_last_error(ctx));
      adept_cuda_destroy(ctx);
    } else {
      fprintf(stderr, "adept_cuda_create failed (%lld)\n", static_cast<long long>(gpu_rc));
    }
    if (gpu_rc) { jac_rev.clear(); jac_fwd.clear(); }
  }

  FILE* f = fopen(argv[2], "wb");
  if (!f) return 3;
  const int64_t hdr[12] = {B, t.n_statements(), t.n_operations(), t.max_gradient, static_cast<int64_t>(t.indep.size()),
                           static_cast<int64_t>(t.dep.size()), refused, serial_same_structure, max_grad0,
                           static_cast<int64_t>(jac_rev.size()), gpu_rc, static_cast<int64_t>(stmt0.size() / 2)};
  fwrite(hdr, sizeof(int64_t), 12, f);
  put(f, t.statements); put(f, t.index); put(f, t.indep); put(f, t.dep); put(f, t.multipliers);
  put(f, stmt0); put(f, idx0); put(f, serial_mult); put(f, analytic); put(f, jac_rev); put(f, jac_fwd);
  fclose(f);
  return 0;
}
