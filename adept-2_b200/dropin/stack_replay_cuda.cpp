// stack_replay_cuda.cpp -- drop-in replacement for the REPLAY members of adept::Stack.
//
// Compiled against the UNMODIFIED Adept-2 headers and linked in place of the reference's
// adept/jacobian.cpp (whole file) and of two members of adept/Stack.cpp:
//
//   Stack::compute_adjoint()                 reference adept/Stack.cpp:139-164
//   Stack::compute_tangent_linear()          reference adept/Stack.cpp:170-190
//   Stack::jacobian_forward(Real*,Index,Index) const   reference adept/jacobian.cpp:204-315 (+ _openmp :127-194)
//   Stack::jacobian_reverse(Real*,Index,Index) const   reference adept/jacobian.cpp:462-647 (+ _openmp :331-452)
//   the six Array<2,Real,false> overloads    reference adept/jacobian.cpp:651-707
//
// Every signature is the one in include/adept/Stack.h; the bodies forward to the C-ABI of
// libadept_cuda (include/adept_cuda.h).  Recording, the gradient registry, set/get_gradients,
// print_* and everything inline in the headers are the reference's own code, untouched.
// Pre-checks throw the reference's own exception types at the same points; CUDA/NCCL failures
// surface as adept::cuda_replay_error (an adept::exception).  There is no CPU fallback.
//
// Context policy: one adept_cuda_ctx per HOST THREAD (the reference's threading model is one
// active Stack per thread, include/adept/Stack.h:52-61), created on first use on the devices
// listed in $ADEPT_CUDA_DEVICES (comma separated, default "0").
//
// When is the device copy of the tape still valid?  The UNMODIFIED reference keeps no recording
// counter, so with stock headers the tape is re-uploaded (and re-analysed) on every replay call;
// $ADEPT_CUDA_TAPE_CACHE=1 opts into reusing it while (Stack*, array pointers, n_statements_,
// n_operations_, max_gradient_) are unchanged -- a guess that fails when a program re-records a
// tape of identical size.  With adept_recording_epoch.patch applied to include/adept/Stack.h
// (4 lines: a per-thread-unique id taken in new_recording(); it defines ADEPT_RECORDING_EPOCH) the
// test is EXACT and always on: between two new_recording() calls the tape is append-only except for
// the index of its last statement (update_lhs, include/adept/Stack.h:168-176), so
// (epoch, n_statements_, n_operations_, last index, max_gradient_) identifies its contents.  Loops
// that sweep ONE tape many times (test/test_derivatives.cpp:25-29, an optimiser's line search) then
// pay one upload and one schedule build.
#include <adept_arrays.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/adept_cuda.h"

namespace adept {

class cuda_replay_error : public adept::exception {
 public:
  cuda_replay_error(const std::string& message = "CUDA tape replay failed") { message_ = message; }
};

namespace {

struct ThreadCtx {
  adept_cuda_ctx* ctx;
  // identity of the tape currently on the device
  const void* stack;
  const void* stmt;
  const void* mult;
  uIndex n_stmt, n_ops, max_gradient;
  uIndex rec_epoch;     // Stack::recording_epoch_ (patched headers only)
  long long last_index; // LHS of the last statement: update_lhs may change it in place
  bool valid;
  uint64_t epoch;
  uint64_t uploads;     // tape uploads so far (what the epoch patch saves; read by dropin_check)
  ThreadCtx() : ctx(0), stack(0), stmt(0), mult(0), n_stmt(0), n_ops(0), max_gradient(0), rec_epoch(0), last_index(0),
                valid(false), epoch(0), uploads(0) {}
  ~ThreadCtx() {
    if (ctx) adept_cuda_destroy(ctx);
  }
};

ThreadCtx& thread_ctx() {
  static thread_local ThreadCtx t;
  if (!t.ctx) {
    std::vector<int> devs;
    const char* env = std::getenv("ADEPT_CUDA_DEVICES");
    std::string s = env ? env : "0";
    size_t pos = 0;
    while (pos <= s.size()) {
      size_t comma = s.find(',', pos);
      if (comma == std::string::npos) comma = s.size();
      if (comma > pos) devs.push_back(std::atoi(s.substr(pos, comma - pos).c_str()));
      pos = comma + 1;
    }
    if (devs.empty()) devs.push_back(0);
    int rc = adept_cuda_create(&t.ctx, &devs[0], (int)devs.size());
    if (rc) {
      t.ctx = 0;
      throw cuda_replay_error("adept_cuda_create failed: no usable CUDA device (there is no CPU fallback)");
    }
    // P of the reference's "skip the statement when all P lanes are zero" rule (adept/jacobian.cpp:21, :392-407) as
    // THIS build of the headers defines it, so that results equal those of the libadept the program would otherwise
    // link even with non-finite multipliers; $ADEPT_CUDA_REF_BLOCK overrides (to match a libadept built differently)
    const int P = ADEPT_REAL_PACKET_SIZE == 1 ? ADEPT_MULTIPASS_SIZE : ADEPT_REAL_PACKET_SIZE;
    if (P >= 1 && P <= 32 && (P & (P - 1)) == 0) adept_cuda_set_option(t.ctx, "ref_block", P);
    const char* blk = std::getenv("ADEPT_CUDA_REF_BLOCK");
    if (blk) adept_cuda_set_option(t.ctx, "ref_block", std::atoi(blk));
    const char* sch = std::getenv("ADEPT_CUDA_SCHEDULE");
    if (sch) adept_cuda_set_option(t.ctx, "schedule", std::atoi(sch));
  }
  return t;
}

void check(ThreadCtx& t, int rc) {
  if (rc == ADEPT_CUDA_OK) return;
  const std::string msg = adept_cuda_last_error(t.ctx);
  switch (rc) {
    case ADEPT_CUDA_ERR_NOT_IDENTIFIED: throw dependents_or_independents_not_identified();
    case ADEPT_CUDA_ERR_GRAD_NOT_INIT: throw gradients_not_initialized();
    case ADEPT_CUDA_ERR_RANGE: throw gradient_out_of_range(msg);
    default: throw cuda_replay_error(msg);
  }
}

}  // namespace

// The device copy of this Stack's tape, uploading it when needed.  A protected-member reader in
// the shape of a Stack member: declared nowhere in the header, so it is a free function taking
// what the members below pass in.
static ThreadCtx& synced(const void* self, const internal::Statement* statement, const Real* multiplier,
                         const uIndex* index, uIndex n_statements, uIndex n_operations, uIndex max_gradient,
                         uIndex rec_epoch) {
  ThreadCtx& t = thread_ctx();
  const long long last_index = n_statements > 0 ? static_cast<long long>(statement[n_statements - 1].index) : 0;
#ifdef ADEPT_RECORDING_EPOCH
  // exact: see the header comment
  const bool same = t.valid && t.stack == self && t.rec_epoch == rec_epoch && t.n_stmt == n_statements &&
                    t.n_ops == n_operations && t.last_index == last_index && t.max_gradient == max_gradient;
#else
  static const bool cache = []() {
    const char* e = std::getenv("ADEPT_CUDA_TAPE_CACHE");
    return e && std::atoi(e) != 0;
  }();
  const bool same = cache && t.valid && t.stack == self && t.stmt == statement && t.mult == multiplier &&
                    t.n_stmt == n_statements && t.n_ops == n_operations && t.last_index == last_index &&
                    t.max_gradient == max_gradient;
#endif
  if (!same) {
    t.valid = false;
    ADEPT_STATIC_ASSERT(sizeof(internal::Statement) == 8 && sizeof(uIndex) == 4 && sizeof(Real) == 8,
                        LIBADEPT_CUDA_NEEDS_THE_DEFAULT_INT_INDEX_AND_DOUBLE_REAL);
    check(t, adept_cuda_upload_tape(t.ctx, statement, n_statements, multiplier, reinterpret_cast<const int32_t*>(index),
                                    n_operations, max_gradient, ++t.epoch));
    t.stack = self; t.stmt = statement; t.mult = multiplier;
    t.n_stmt = n_statements; t.n_ops = n_operations; t.max_gradient = max_gradient;
    t.rec_epoch = rec_epoch; t.last_index = last_index;
    t.valid = true;
    ++t.uploads;
  }
  return t;
}

#ifdef ADEPT_RECORDING_EPOCH
#define ADEPT_CUDA_REC_EPOCH recording_epoch_
#else
#define ADEPT_CUDA_REC_EPOCH 0
#endif

// Number of tape uploads this thread's context has made (diagnostic; dropin_check and epoch_check read it)
extern "C" unsigned long long adept_cuda_dropin_uploads() { return thread_ctx().uploads; }

// ---- single-direction sweeps ---------------------------------------------------------------
void Stack::compute_adjoint() {
  if (!gradients_are_initialized()) throw(gradients_not_initialized());   // adept/Stack.cpp:161-163
  ThreadCtx& t = synced(this, statement_, multiplier_, index_, n_statements_, n_operations_, max_gradient_, ADEPT_CUDA_REC_EPOCH);
  check(t, adept_cuda_adjoint(t.ctx, gradient_, 1, 0));                    // gradient_ holds the post-sweep state
}

void Stack::compute_tangent_linear() {
  if (!gradients_are_initialized()) throw(gradients_not_initialized());   // adept/Stack.cpp:187-189
  ThreadCtx& t = synced(this, statement_, multiplier_, index_, n_statements_, n_operations_, max_gradient_, ADEPT_CUDA_REC_EPOCH);
  check(t, adept_cuda_tangent_linear(t.ctx, gradient_, 1, 0));
}

// ---- Jacobians -----------------------------------------------------------------------------------
void Stack::jacobian_forward(Real* jacobian_out, Index dep_offset, Index indep_offset) const {
  if (independent_index_.empty() || dependent_index_.empty())             // adept/jacobian.cpp:208-210
    throw(dependents_or_independents_not_identified());
  ThreadCtx& t = synced(this, statement_, multiplier_, index_, n_statements_, n_operations_, max_gradient_, ADEPT_CUDA_REC_EPOCH);
  check(t, adept_cuda_jacobian(t.ctx, ADEPT_CUDA_FORWARD, reinterpret_cast<const int32_t*>(&independent_index_[0]),
                               n_independent(), reinterpret_cast<const int32_t*>(&dependent_index_[0]), n_dependent(),
                               jacobian_out, dep_offset, indep_offset));    // offsets <= 0 are defaulted by the engine
}

void Stack::jacobian_reverse(Real* jacobian_out, Index dep_offset, Index indep_offset) const {
  if (independent_index_.empty() || dependent_index_.empty())             // adept/jacobian.cpp:466-468
    throw(dependents_or_independents_not_identified());
  ThreadCtx& t = synced(this, statement_, multiplier_, index_, n_statements_, n_operations_, max_gradient_, ADEPT_CUDA_REC_EPOCH);
  check(t, adept_cuda_jacobian(t.ctx, ADEPT_CUDA_REVERSE, reinterpret_cast<const int32_t*>(&independent_index_[0]),
                               n_independent(), reinterpret_cast<const int32_t*>(&dependent_index_[0]), n_dependent(),
                               jacobian_out, dep_offset, indep_offset));
}

// The OpenMP variants are private helpers of the reference's jacobian.cpp; they stay declared in
// the header, so define them (the directions are partitioned over GPUs, not host threads).
void Stack::jacobian_forward_openmp(Real* jacobian_out, Index dep_offset, Index indep_offset) const {
  jacobian_forward(jacobian_out, dep_offset, indep_offset);
}
void Stack::jacobian_reverse_openmp(Real* jacobian_out, Index dep_offset, Index indep_offset) const {
  jacobian_reverse(jacobian_out, dep_offset, indep_offset);
}

// ---- Array<2> overloads (adept/jacobian.cpp:651-707): size check, then the raw-pointer calls with
//      the array's own strides ------------------------------------------------------------------
void Stack::jacobian(Array<2, Real, false> jac) const {
  if (jac.dimension(0) != n_dependent() || jac.dimension(1) != n_independent())
    throw size_mismatch("Jacobian matrix has wrong size");
  if (n_independent() <= n_dependent()) jacobian_forward(jac.data(), jac.offset(0), jac.offset(1));
  else jacobian_reverse(jac.data(), jac.offset(0), jac.offset(1));
}
void Stack::jacobian_forward(Array<2, Real, false> jac) const {
  if (jac.dimension(0) != n_dependent() || jac.dimension(1) != n_independent())
    throw size_mismatch("Jacobian matrix has wrong size");
  jacobian_forward(jac.data(), jac.offset(0), jac.offset(1));
}
void Stack::jacobian_reverse(Array<2, Real, false> jac) const {
  if (jac.dimension(0) != n_dependent() || jac.dimension(1) != n_independent())
    throw size_mismatch("Jacobian matrix has wrong size");
  jacobian_reverse(jac.data(), jac.offset(0), jac.offset(1));
}
Array<2, Real, false> Stack::jacobian() const {
  Array<2, Real, false> jac(n_dependent(), n_independent());
  if (n_independent() <= n_dependent()) jacobian_forward(jac.data(), jac.offset(0), jac.offset(1));
  else jacobian_reverse(jac.data(), jac.offset(0), jac.offset(1));
  return jac;
}
Array<2, Real, false> Stack::jacobian_forward() const {
  Array<2, Real, false> jac(n_dependent(), n_independent());
  jacobian_forward(jac.data(), jac.offset(0), jac.offset(1));
  return jac;
}
Array<2, Real, false> Stack::jacobian_reverse() const {
  Array<2, Real, false> jac(n_dependent(), n_independent());
  jacobian_reverse(jac.data(), jac.offset(0), jac.offset(1));
  return jac;
}

}  // namespace adept
