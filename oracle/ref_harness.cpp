/* oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin C-ABI harness around the UNMODIFIED reference library (Adept 2.1.3),
 * compiled from the sources where they lie under $ADEPT_REF (/root/reference)
 * by oracle/Makefile into oracle/_ref/libadept_ref.so.  It is used
 *   - by tests/ to validate the C restatement (oracle/replay_oracle.c) and to
 *     produce golden vectors (oracle/make_golden.py),
 *   - by bench.py's cpu_baseline / --impl reference leg (the reference's own
 *     OpenMP Stack::jacobian timed on the host cores).
 * Nothing in the product path (adept-2_b200/, include/) may link or load it.
 *
 * What it exposes: "tape handles" that own one adept::Stack each.  A tape is
 * either RECORDED by real Adept expression templates (advection schemes of
 * benchmark/advection_schemes.h, test/algorithm.cpp, the radiance model of
 * test/simulate_radiances.cpp:22-49, the Rosenbrock cost of
 * test/test_minimizer.cpp:39-49,103-115) or INJECTED through the public
 * tape-writer API (Stack::register_gradients / new_recording / check_space /
 * push_rhs / push_lhs, include/adept/StackStorageOrig.h:46-121), and is then
 * replayed by the reference's own Stack::jacobian_forward/_reverse,
 * compute_adjoint, compute_tangent_linear.
 */
#include <adept_arrays.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using adept::adouble;
using adept::aReal;
using adept::Real;
using adept::Stack;
using adept::uIndex;

// Recorders compiled from the reference's own header with different NX
// (oracle/ref_advection.cpp, one object per NX).
extern "C" int ref_adv_record_nx100 (int scheme, int nt, double c, const double* q0, std::vector<uIndex>* indep, std::vector<uIndex>* dep);
extern "C" int ref_adv_record_nx128 (int scheme, int nt, double c, const double* q0, std::vector<uIndex>* indep, std::vector<uIndex>* dep);
extern "C" int ref_adv_record_nx4096(int scheme, int nt, double c, const double* q0, std::vector<uIndex>* indep, std::vector<uIndex>* dep);
// test/simulate_radiances.cpp, compiled unmodified into the harness (C++ linkage)
void simulate_radiances(int n, Real surface_temperature, const Real* temperature, Real radiance[2],
                        Real dradiance_dsurface_temperature[2], Real* dradiance_dtemperature);

// test/algorithm.cpp, compiled unmodified into the harness
adouble algorithm(const adouble x[2]);

namespace {

// Legal read access to the protected tape members: a pointer-to-member named
// through a derived class has the base class's type and may be applied to any
// Stack (include/adept/StackStorageOrig.h:155-163, include/adept/Stack.h:782-800).
struct Peek : public Stack {
  static const adept::internal::Statement* statements(const Stack& s) { return s.*(&Peek::statement_); }
  static const Real*   multipliers(const Stack& s) { return s.*(&Peek::multiplier_); }
  static const uIndex* indices(const Stack& s)     { return s.*(&Peek::index_); }
  static uIndex n_stmt(const Stack& s)  { return s.*(&Peek::n_statements_); }
  static uIndex n_ops(const Stack& s)   { return s.*(&Peek::n_operations_); }
  static uIndex max_grad(const Stack& s){ return s.*(&Peek::max_gradient_); }
  static const std::vector<uIndex>& indep(const Stack& s) { return s.*(&Peek::independent_index_); }
  static const std::vector<uIndex>& dep(const Stack& s)   { return s.*(&Peek::dependent_index_); }
};

// What Stack::independent/dependent need: anything with push_gradient_indices
// (include/adept/Stack.h:431-456, include/adept/Active.h:532-535).
struct SlotRef {
  uIndex slot;
  void push_gradient_indices(std::vector<uIndex>& v) const { v.push_back(slot); }
};

struct ActiveGuard {
  Stack* s; Stack* prev;
  explicit ActiveGuard(Stack* st) : s(st), prev(adept::active_stack()) {
    if (prev && prev != s) prev->deactivate();
    if (!s->is_active()) s->activate();
  }
  ~ActiveGuard() { s->deactivate(); if (prev && prev != s) prev->activate(); }
};

}  // namespace

struct ref_tape {
  Stack stack;
  uIndex injected_base;
  bool injected;
  std::string err;
  ref_tape() : stack(false), injected_base(0), injected(false) {}
};

extern "C" {

int ref_packet_size() { return ADEPT_REAL_PACKET_SIZE == 1 ? ADEPT_MULTIPASS_SIZE : ADEPT_REAL_PACKET_SIZE; }
int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
const char* ref_version() { return ADEPT_VERSION_STR; }

ref_tape* ref_tape_new() { try { return new ref_tape(); } catch (...) { return 0; } }
void ref_tape_free(ref_tape* t) { delete t; }
const char* ref_tape_error(ref_tape* t) { return t->err.c_str(); }

/* ---- inject a raw tape through the public writer API ------------------- */
int ref_tape_inject(ref_tape* t, const int32_t* stmt_pairs, int64_t n_stmt,
                    const double* mult, const int32_t* idx, int64_t n_ops,
                    int64_t n_slots,
                    const int32_t* indep, int64_t n_indep,
                    const int32_t* dep, int64_t n_dep) {
  try {
    ActiveGuard g(&t->stack);
    Stack& s = t->stack;
    if (!t->injected) {
      // slots [base, base+n_slots): base is 0 on a fresh Stack
      t->injected_base = s.register_gradients((uIndex)n_slots);
      t->injected = true;
    }
    if (t->injected_base != 0) { t->err = "inject needs a fresh tape handle"; return 2; }
    s.new_recording();   // pushes the sentinel {-1,0}; max_gradient_ = n_slots+1
    if (n_stmt < 1 || stmt_pairs[0] != -1 || stmt_pairs[1] != 0) {
      t->err = "statement[0] must be the {-1,0} sentinel"; return 3;
    }
    int64_t op = 0;
    for (int64_t ist = 1; ist < n_stmt; ++ist) {
      int64_t end = stmt_pairs[2*ist+1];
      s.check_space((uIndex)(end - op));
      for (; op < end; ++op) s.push_rhs(mult[op], (uIndex)idx[op]);
      s.push_lhs((uIndex)stmt_pairs[2*ist]);
    }
    if (op != n_ops) { t->err = "n_ops does not match last end_plus_one"; return 4; }
    for (int64_t i = 0; i < n_indep; ++i) { SlotRef r = {(uIndex)indep[i]}; s.independent(r); }
    for (int64_t i = 0; i < n_dep; ++i)   { SlotRef r = {(uIndex)dep[i]};   s.dependent(r); }
    return 0;
  } catch (const std::exception& e) { t->err = e.what(); return 1; }
}

/* change only the independent/dependent lists of an existing tape */
int ref_tape_set_io(ref_tape* t, const int32_t* indep, int64_t n_indep,
                    const int32_t* dep, int64_t n_dep) {
  Stack& s = t->stack;
  s.clear_independents(); s.clear_dependents();
  for (int64_t i = 0; i < n_indep; ++i) { SlotRef r = {(uIndex)indep[i]}; s.independent(r); }
  for (int64_t i = 0; i < n_dep; ++i)   { SlotRef r = {(uIndex)dep[i]};   s.dependent(r); }
  return 0;
}

/* ---- recorders: real Adept expression templates ------------------------- */

/* benchmark/advection_schemes.h:24-58 with the init of
   benchmark/autodiff_benchmark.cpp:294 (q0==NULL) ; scheme 0 = lax_wendroff, 1 = toon */
int ref_tape_record_advection(ref_tape* t, int scheme, int nx, int nt, double c, const double* q0) {
  try {
    ActiveGuard g(&t->stack);
    std::vector<uIndex> indep, dep;
    int rc;
    if      (nx == 100)  rc = ref_adv_record_nx100 (scheme, nt, c, q0, &indep, &dep);
    else if (nx == 128)  rc = ref_adv_record_nx128 (scheme, nt, c, q0, &indep, &dep);
    else if (nx == 4096) rc = ref_adv_record_nx4096(scheme, nt, c, q0, &indep, &dep);
    else { t->err = "advection recorder compiled for NX in {100,128,4096} only"; return 2; }
    if (rc) { t->err = "advection recorder failed"; return rc; }
    // the adoubles are gone now; re-declare the lists from the saved slots
    Stack& s = t->stack;
    s.clear_independents(); s.clear_dependents();
    for (size_t i = 0; i < indep.size(); ++i) { SlotRef r = {indep[i]}; s.independent(r); }
    for (size_t i = 0; i < dep.size(); ++i)   { SlotRef r = {dep[i]};   s.dependent(r); }
    return 0;
  } catch (const std::exception& e) { t->err = e.what(); return 1; }
}

/* test/algorithm.cpp:22-29 driven as in test/test_adept.cpp:147-195 */
int ref_tape_record_algorithm(ref_tape* t, double x0, double x1, double* y_out) {
  try {
    ActiveGuard g(&t->stack);
    Stack& s = t->stack;
    std::vector<uIndex> indep, dep;
    {
      adouble x[2] = {x0, x1};
      s.new_recording();
      adouble y = algorithm(x);
      if (y_out) *y_out = y.value();
      x[0].push_gradient_indices(indep); x[1].push_gradient_indices(indep);
      y.push_gradient_indices(dep);
    }
    for (size_t i = 0; i < indep.size(); ++i) { SlotRef r = {indep[i]}; s.independent(r); }
    for (size_t i = 0; i < dep.size(); ++i)   { SlotRef r = {dep[i]};   s.dependent(r); }
    return 0;
  } catch (const std::exception& e) { t->err = e.what(); return 1; }
}

/* The radiance model of test/simulate_radiances.cpp:22-49 recorded exactly as
   simulate_radiances() does (:84-128): independents st, t[0..n), dependents r[0..2). */
int ref_tape_record_radiances(ref_tape* t, int n, double surface_temperature,
                              const double* temperature, double* radiance_out) {
  try {
    ActiveGuard g(&t->stack);
    Stack& s = t->stack;
    static const Real wavelength[2] = {0.006, 0.0061};
    static const Real mass_abs_coefft[2] = {3.0e-5, 3.0e-3};
    static const Real dz = 1000.0;
    static const Real BOLTZMANN_CONSTANT = 1.380648813e-23;
    static const Real SPEED_OF_LIGHT = 299792458.0;
    std::vector<Real> density_oxygen(n), emissivity(n);
    std::vector<uIndex> indep, dep;
    {
      aReal st = surface_temperature;
      std::vector<aReal> tt(n);
      aReal r[2];
      for (int i = 0; i < n; i++) {
        Real altitude = i*dz;
        density_oxygen[i] = 1.2*0.21*(16.0/29.0)*exp(-altitude/8000.0);
        tt[i] = temperature[i];
      }
      s.new_recording();
      for (int ichan = 0; ichan < 2; ichan++) {
        for (int i = 0; i < n; i++)
          emissivity[i] = 1.0-exp(-density_oxygen[i]*mass_abs_coefft[ichan]*dz);
        aReal bt = st;
        for (int i = 0; i < n; i++)
          bt = bt*(1.0-emissivity[i]) + emissivity[i]*tt[i];
        Real wl = wavelength[ichan];
        r[ichan] = 2.0*SPEED_OF_LIGHT*BOLTZMANN_CONSTANT*bt/(wl*wl*wl*wl);
        if (radiance_out) radiance_out[ichan] = r[ichan].value();
      }
      st.push_gradient_indices(indep);
      for (int i = 0; i < n; i++) tt[i].push_gradient_indices(indep);
      r[0].push_gradient_indices(dep); r[1].push_gradient_indices(dep);
    }
    for (size_t i = 0; i < indep.size(); ++i) { SlotRef r = {indep[i]}; s.independent(r); }
    for (size_t i = 0; i < dep.size(); ++i)   { SlotRef r = {dep[i]};   s.dependent(r); }
    return 0;
  } catch (const std::exception& e) { t->err = e.what(); return 1; }
}

/* BASELINE config 3, CPU arm: the reference's OWN simulate_radiances() (test/simulate_radiances.cpp:52-146, compiled
   unmodified into this library) for every member of an ensemble, under OpenMP with one Stack per call -- the
   function makes its own adept::Stack, thread-local by the reference's design (include/adept/Stack.h:52-61).
   state: [member][n+1] = surface temperature, n layer temperatures.  jac_out: [member][(n+1)][2], i.e. per member the
   reference's default column-major 2 x (n+1) Jacobian (dependent index fastest). */
int ref_ensemble_radiances(int64_t n_members, int n, const double* state, double* jac_out, int n_threads,
                           double* seconds) {
  int failed = 0;
  const auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
  const int nt = n_threads > 0 ? n_threads : omp_get_num_procs();
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
  for (int64_t b = 0; b < n_members; ++b) {
    try {
      const double* x = state + b * (n + 1);
      Real rad[2], dst[2];
      std::vector<Real> dt(2 * n);
      simulate_radiances(n, x[0], x + 1, rad, dst, &dt[0]);
      double* j = jac_out + b * 2 * (n + 1);
      j[0] = dst[0]; j[1] = dst[1];
      for (int i = 0; i < n; ++i) { j[2 + 2 * i] = dt[2 * i]; j[3 + 2 * i] = dt[2 * i + 1]; }
    } catch (...) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      failed = 1;
    }
  }
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return failed;
}

/* RosenbrockN of test/test_minimizer.cpp:39-49 and :103-108: y = calc_y(x);
   cost = 0.5*sum(y*y).  independents = x, dependent = cost. */
int ref_tape_record_rosenbrock(ref_tape* t, int n, double x0, double* cost_out) {
  try {
    ActiveGuard g(&t->stack);
    Stack& s = t->stack;
    static const Real COST_SCALING = 1.0;
    std::vector<uIndex> indep, dep;
    {
      adept::Vector x(n); x = x0;
      adept::aVector xactive = x;
      s.new_recording();
      int nx = xactive.size();
      adept::aVector y((nx-1)*2);
      for (int ix = 0; ix < nx-1; ++ix) {
        y(ix*2)   = 10.0 * (xactive(ix+1)-xactive(ix)*xactive(ix));
        y(ix*2+1) = 1.0 - xactive(ix);
      }
      y *= sqrt(2.0 * COST_SCALING);
      aReal cost = 0.5*sum(y*y);
      if (cost_out) *cost_out = cost.value();
      xactive.push_gradient_indices(indep);
      cost.push_gradient_indices(dep);
    }
    for (size_t i = 0; i < indep.size(); ++i) { SlotRef r = {indep[i]}; s.independent(r); }
    for (size_t i = 0; i < dep.size(); ++i)   { SlotRef r = {dep[i]};   s.dependent(r); }
    return 0;
  } catch (const std::exception& e) { t->err = e.what(); return 1; }
}

/* ---- read the tape out -------------------------------------------------- */
void ref_tape_sizes(ref_tape* t, int64_t* n_stmt, int64_t* n_ops, int64_t* max_gradient,
                    int64_t* n_indep, int64_t* n_dep) {
  const Stack& s = t->stack;
  *n_stmt = Peek::n_stmt(s); *n_ops = Peek::n_ops(s); *max_gradient = Peek::max_grad(s);
  *n_indep = (int64_t)Peek::indep(s).size(); *n_dep = (int64_t)Peek::dep(s).size();
}

void ref_tape_copy(ref_tape* t, int32_t* stmt_pairs, double* mult, int32_t* idx,
                   int32_t* indep, int32_t* dep) {
  const Stack& s = t->stack;
  int64_t ns = Peek::n_stmt(s), no = Peek::n_ops(s);
  const adept::internal::Statement* st = Peek::statements(s);
  for (int64_t i = 0; i < ns; ++i) { stmt_pairs[2*i] = (int32_t)st[i].index; stmt_pairs[2*i+1] = (int32_t)st[i].end_plus_one; }
  std::memcpy(mult, Peek::multipliers(s), sizeof(double)*no);
  const uIndex* ix = Peek::indices(s);
  for (int64_t i = 0; i < no; ++i) idx[i] = (int32_t)ix[i];
  const std::vector<uIndex>& a = Peek::indep(s); const std::vector<uIndex>& b = Peek::dep(s);
  for (size_t i = 0; i < a.size(); ++i) indep[i] = (int32_t)a[i];
  for (size_t i = 0; i < b.size(); ++i) dep[i] = (int32_t)b[i];
}

/* ---- replay with the reference's own code ------------------------------- */
/* mode 0: jacobian_forward, 1: jacobian_reverse, 2: jacobian (auto).  n_threads
   follows Stack::set_max_jacobian_threads (adept/Stack.cpp:92-116): 1 = serial, 0 = all procs. */
int ref_tape_jacobian(ref_tape* t, int mode, double* out, int64_t dep_offset, int64_t indep_offset,
                      int n_threads, double* seconds) {
  try {
    Stack& s = t->stack;
    s.set_max_jacobian_threads(n_threads);
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    if (mode == 0)      s.jacobian_forward(out, (adept::Index)dep_offset, (adept::Index)indep_offset);
    else if (mode == 1) s.jacobian_reverse(out, (adept::Index)dep_offset, (adept::Index)indep_offset);
    else                s.jacobian(out, (adept::Index)dep_offset, (adept::Index)indep_offset);
    std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return 0;
  } catch (const adept::dependents_or_independents_not_identified& e) { t->err = e.what(); return 10; }
  catch (const std::exception& e) { t->err = e.what(); return 1; }
}

/* set_gradients / compute_* / get_gradients on the whole slot range [0,max_gradient).
   clear==1 first calls clear_gradients() so that the lazily zero-filled vector is
   re-initialised (include/adept/Stack.h:290-303, adept/Stack.cpp:515-531). */
int ref_tape_sweep(ref_tape* t, int adjoint, double* gradient_inout, int clear, double* seconds) {
  try {
    Stack& s = t->stack;
    uIndex g = Peek::max_grad(s);
    if (clear) s.clear_gradients();
    s.set_gradients((uIndex)0, g, gradient_inout);
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    if (adjoint) s.compute_adjoint(); else s.compute_tangent_linear();
    std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    s.get_gradients((uIndex)0, g, gradient_inout);
    return 0;
  } catch (const adept::gradients_not_initialized& e) { t->err = e.what(); return 11; }
  catch (const std::exception& e) { t->err = e.what(); return 1; }
}

/* compute_adjoint without any set_gradients: must raise gradients_not_initialized
   (adept/Stack.cpp:161-163) -> returns 11 */
int ref_tape_sweep_uninitialised(ref_tape* t, int adjoint) {
  try {
    Stack& s = t->stack;
    s.clear_gradients();
    if (adjoint) s.compute_adjoint(); else s.compute_tangent_linear();
    return 0;
  } catch (const adept::gradients_not_initialized& e) { t->err = e.what(); return 11; }
  catch (const std::exception& e) { t->err = e.what(); return 1; }
}

}  // extern "C"
