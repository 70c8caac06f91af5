// adept_cuda_ensemble.h -- host-side ensemble recording helper.
//
// The caller side of adept_cuda_jacobian_batch (include/adept_cuda.h): BASELINE config 3 is an ensemble of
// 10^5 state vectors pushed through ONE model (test/simulate_radiances.cpp:52-146 in the reference).  Every
// member records the same tape STRUCTURE (statements, operand indices, independent/dependent lists); only the
// multipliers differ.  At that size host recording, not replay, is the bottleneck, so the members are recorded
// on all host cores -- one adept::Stack per thread, the reference's own threading model (thread-local
// ADEPT_ACTIVE_STACK, include/adept/Stack.h:52-61; test/test_thread_safe.cpp:203-209) -- the structures are
// checked against member 0's, and the multipliers are packed [operation][member], the layout the batch entry
// point takes, straight from each thread's tape.
//
// Header only; compiled against the UNMODIFIED Adept-2 headers (plain Adept API inside `record`).
//
//   adept_cuda::EnsembleTape tape;
//   adept_cuda::record_ensemble(n_members, [&](int64_t b, adept::Stack& stack) {
//       adept::adouble x[N]; ... set values of member b ...
//       stack.new_recording();
//       ... the model ...
//       stack.independent(x, N); stack.dependent(y, M);
//   }, tape);
//   adept_cuda::ensemble_jacobian(ctx, ADEPT_CUDA_REVERSE, tape, jac /* [member][indep][dep], column-major per member */);
#ifndef ADEPT_CUDA_ENSEMBLE_H
#define ADEPT_CUDA_ENSEMBLE_H

#include <adept.h>
#include <omp.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/adept_cuda.h"

namespace adept_cuda {

class ensemble_structure_mismatch : public adept::exception {
 public:
  ensemble_structure_mismatch(const std::string& message = "ensemble members recorded different tape structures") {
    message_ = message;
  }
};

// A host array of doubles in PAGE-LOCKED memory (adept_cuda_host_alloc) when that is available, so that the engine's
// copy-in / sweep / copy-out pipeline really overlaps; plain heap memory otherwise.  Keeps its allocation when it
// is re-filled with the same or a smaller size (an ensemble is typically recorded again and again).
class HostBuffer {
 public:
  HostBuffer() : p_(0), n_(0), cap_(0), pinned_(false) {}
  ~HostBuffer() { release(); }
  HostBuffer(const HostBuffer&) = delete;
  HostBuffer& operator=(const HostBuffer&) = delete;
  void assign(size_t n, double value) {
    if (n > cap_) {
      release();
      p_ = static_cast<double*>(adept_cuda_host_alloc(n * sizeof(double)));
      pinned_ = p_ != 0;
      if (!p_) p_ = new double[n > 0 ? n : 1];
      cap_ = n;
    }
    n_ = n;
    for (size_t i = 0; i < n; ++i) p_[i] = value;
  }
  void resize_uninitialised(size_t n) {
    if (n > cap_) { n_ = 0; assign(n, 0.0); }
    n_ = n;
  }
  double* data() { return p_; }
  const double* data() const { return p_; }
  size_t size() const { return n_; }
  bool pinned() const { return pinned_; }
  double& operator[](size_t i) { return p_[i]; }
  const double& operator[](size_t i) const { return p_[i]; }

 private:
  void release() {
    if (p_) {
      if (pinned_) adept_cuda_host_free(p_);
      else delete[] p_;
    }
    p_ = 0;
    n_ = cap_ = 0;
  }
  double* p_;
  size_t n_, cap_;
  bool pinned_;
};

struct EnsembleTape {
  std::vector<int32_t> statements;   // {lhs slot, end_plus_one} pairs incl. the {-1,0} sentinel (Statement.h:24-30)
  std::vector<int32_t> index;        // operand slots
  std::vector<int32_t> indep, dep;   // slots of the independent / dependent variables
  int64_t max_gradient = 0;
  int64_t n_members = 0;
  HostBuffer multipliers;            // [operation][member], page-locked when possible
  int64_t n_statements() const { return static_cast<int64_t>(statements.size() / 2); }
  int64_t n_operations() const { return static_cast<int64_t>(index.size()); }
};

namespace detail {
// The tape arrays are protected members of adept::Stack's storage base (StackStorageOrig.h:155-163); a derived
// class may read them.  The object IS an adept::Stack: constructing it makes it the thread's active stack.
class RecordingStack : public adept::Stack {
 public:
  RecordingStack() : adept::Stack(true) {}
  int64_t n_stmt() const { return static_cast<int64_t>(n_statements_); }
  int64_t n_ops() const { return static_cast<int64_t>(n_operations_); }
  const adept::internal::Statement* stmts() const { return statement_; }
  const adept::Real* mults() const { return multiplier_; }
  const adept::uIndex* idxs() const { return index_; }
  int64_t max_grad() const { return static_cast<int64_t>(max_gradient_); }
  const std::vector<adept::uIndex>& indep() const { return independent_index_; }
  const std::vector<adept::uIndex>& dep() const { return dependent_index_; }
};
}  // namespace detail

// Records members [0, n_members) with `record(member, stack)` on n_threads host threads (0 = all cores) and fills
// `out`.  `record` must leave the member's tape on `stack` (new_recording ... independent/dependent) and must not
// keep active objects alive across calls.  Throws ensemble_structure_mismatch if a member's structure differs from
// member 0's (a data-dependent branch in the model); exceptions thrown by `record` are re-thrown on the caller.
template <class Record>
void record_ensemble(int64_t n_members, Record&& record, EnsembleTape& out, int n_threads = 0) {
  static_assert(sizeof(adept::Real) == sizeof(double), "libadept_cuda replays double-precision tapes");
  static_assert(sizeof(adept::internal::Statement) == 2 * sizeof(int32_t), "Statement is {int index, int end_plus_one}");
  out.statements.clear();
  out.index.clear();
  out.indep.clear();
  out.dep.clear();
  out.max_gradient = 0;
  out.multipliers.resize_uninitialised(0);   // keeps the (page-locked) allocation of an earlier recording
  out.n_members = n_members;
  if (n_members <= 0) return;
  // member 0 fixes the structure (recorded on the calling thread, with a stack of its own so that a stack the
  // caller may already have active is left alone: the reference allows one ACTIVE stack per thread, so the
  // caller's is deactivated for the duration and re-activated afterwards)
  adept::Stack* callers = adept::active_stack();
  if (callers) callers->deactivate();
  std::string failure;
  bool mismatch = false;
  try {
    {
      detail::RecordingStack s0;
      record(static_cast<int64_t>(0), static_cast<adept::Stack&>(s0));
      const int64_t S = s0.n_stmt(), O = s0.n_ops();
      out.statements.resize(static_cast<size_t>(2 * S));
      std::memcpy(out.statements.data(), s0.stmts(), static_cast<size_t>(S) * sizeof(adept::internal::Statement));
      out.index.resize(static_cast<size_t>(O));
      for (int64_t o = 0; o < O; ++o) out.index[static_cast<size_t>(o)] = static_cast<int32_t>(s0.idxs()[o]);
      out.indep.assign(s0.indep().begin(), s0.indep().end());
      out.dep.assign(s0.dep().begin(), s0.dep().end());
      out.max_gradient = s0.max_grad();
      out.multipliers.resize_uninitialised(static_cast<size_t>(O) * static_cast<size_t>(n_members));   // every element is written below
      for (int64_t o = 0; o < O; ++o) out.multipliers[static_cast<size_t>(o) * n_members] = s0.mults()[o];
    }
    const int64_t S = out.n_statements(), O = out.n_operations();
    const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
    // Members are recorded in blocks: a thread copies each member's multiplier array (contiguous) into a block
    // buffer [member in block][operation] and then transposes the block into out.multipliers[operation][member]
    // in runs of `bs` consecutive doubles -- writing one double per operation per member straight into the
    // member-minor matrix would touch O cache lines per member.
    int64_t bs = 64;
    while (bs > 8 && bs * O * 8 > (int64_t(2) << 20)) bs >>= 1;
    const int64_t n_blocks = (n_members + bs - 1) / bs;
#pragma omp parallel num_threads(nt)
    {
      // one Stack per thread, active on that thread for the lifetime of the region
      detail::RecordingStack s;
      std::vector<double> buf(static_cast<size_t>(bs) * static_cast<size_t>(O > 0 ? O : 1));
#pragma omp for schedule(static)
      for (int64_t k = 0; k < n_blocks; ++k) {
        bool stop;
#pragma omp atomic read
        stop = mismatch;
        if (stop) continue;
        const int64_t b_lo = k * bs > 1 ? k * bs : 1;   // member 0 is already in place
        const int64_t b_hi = (k + 1) * bs < n_members ? (k + 1) * bs : n_members;
        bool ok = true;
        for (int64_t b = b_lo; b < b_hi && ok; ++b) {
          try {
            record(b, static_cast<adept::Stack&>(s));
            bool same = s.n_stmt() == S && s.n_ops() == O && s.max_grad() == out.max_gradient &&
                        s.indep().size() == out.indep.size() && s.dep().size() == out.dep.size() &&
                        std::memcmp(s.stmts(), out.statements.data(), static_cast<size_t>(S) * sizeof(adept::internal::Statement)) == 0;
            for (int64_t o = 0; same && o < O; ++o) same = static_cast<int32_t>(s.idxs()[o]) == out.index[static_cast<size_t>(o)];
            for (size_t i = 0; same && i < out.indep.size(); ++i) same = static_cast<int32_t>(s.indep()[i]) == out.indep[i];
            for (size_t i = 0; same && i < out.dep.size(); ++i) same = static_cast<int32_t>(s.dep()[i]) == out.dep[i];
            if (!same) {
#pragma omp critical(adept_cuda_ensemble_failure)
              {
                if (failure.empty()) failure = "member " + std::to_string(b) + " recorded a tape structure different from member 0's";
                mismatch = true;
              }
              ok = false;
              break;
            }
            if (O > 0) std::memcpy(buf.data() + static_cast<size_t>(b - b_lo) * O, s.mults(), static_cast<size_t>(O) * sizeof(double));
          } catch (const std::exception& e) {
#pragma omp critical(adept_cuda_ensemble_failure)
            {
              if (failure.empty()) failure = std::string("member ") + std::to_string(b) + ": " + e.what();
              mismatch = true;
            }
            ok = false;
          }
        }
        if (!ok) continue;
        const int64_t nb = b_hi - b_lo;
        for (int64_t o = 0; o < O; ++o) {
          double* row = out.multipliers.data() + static_cast<size_t>(o) * n_members + b_lo;
          const double* src = buf.data() + o;
          for (int64_t j = 0; j < nb; ++j) row[j] = src[static_cast<size_t>(j) * O];
        }
      }
    }
  } catch (...) {
    if (callers) callers->activate();
    throw;
  }
  if (callers) callers->activate();
  if (mismatch) throw ensemble_structure_mismatch(failure);
}

// Every member's Jacobian on the GPU(s) of `ctx`: jac_out[b*n_dep*n_indep + i + j*n_dep] (the reference's default
// column-major layout per member).  Returns the status of adept_cuda_jacobian_batch.
inline int ensemble_jacobian(adept_cuda_ctx* ctx, int mode, const EnsembleTape& t, double* jac_out) {
  return adept_cuda_jacobian_batch(ctx, mode, t.statements.data(), t.n_statements(), t.index.data(), t.n_operations(),
                                   t.max_gradient, t.multipliers.data(), t.n_members, t.indep.data(),
                                   static_cast<int64_t>(t.indep.size()), t.dep.data(),
                                   static_cast<int64_t>(t.dep.size()), jac_out);
}

}  // namespace adept_cuda
#endif  // ADEPT_CUDA_ENSEMBLE_H
