// adept_cuda_chain.h -- checkpoint-chained reverse sweeps on the GPU.
//
// The caller side of adept_cuda_chain_* (include/adept_cuda.h).  The reference differentiates a long simulation
// block by block, last block first (test/test_checkpoint.cpp:166-225):
//
//     stack.new_recording();  model(q_save[i], q);            // re-record block i
//     adept::set_gradients(q, NX, dJ_dq);                      // seeds = what block i+1 produced at its inputs
//     stack.reverse();                                         // Stack::compute_adjoint
//     adept::get_gradients(q_save[i], NX, dJ_dq);              // hand on
//
// With the plain drop-in every block pays a blocking tape upload and two max_gradient-long PCIe copies of
// gradient_.  ChainStack keeps the hand-off vector on the device and lets the GPU work on block i while the host
// records block i-1:
//
//     adept_cuda::ChainStack stack(ctx);                        // IS an adept::Stack (the thread's active stack)
//     stack.chain_begin(NX);
//     stack.new_recording();  model(q_save[n-1], q);  J = ...;
//     stack.chain_push_seeded(&J, &one, 1, q_save[n-1], NX);    // first block: J.set_gradient(1.0)
//     for (i = n-2; i >= 0; --i) {
//       stack.new_recording();  model(q_save[i], q);
//       stack.chain_push(q, q_save[i], NX);                     // seeds: the hand-off vector, on the device
//     }
//     stack.chain_end(dJ_dq);
//
// Header only; compiled against the UNMODIFIED Adept-2 headers.  Results are bit-identical to the loop above.
#ifndef ADEPT_CUDA_CHAIN_H
#define ADEPT_CUDA_CHAIN_H

#include <adept.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/adept_cuda.h"

namespace adept_cuda {

class chain_error : public adept::exception {
 public:
  chain_error(const std::string& message = "checkpoint chain failed") { message_ = message; }
};

// The tape arrays are protected members of adept::Stack's storage base (StackStorageOrig.h:155-163); a derived
// class may read them.
class ChainStack : public adept::Stack {
 public:
  explicit ChainStack(adept_cuda_ctx* ctx) : adept::Stack(true), ctx_(ctx), n_(0) {
    static_assert(sizeof(adept::Real) == sizeof(double), "libadept_cuda replays double-precision tapes");
    static_assert(sizeof(adept::internal::Statement) == 2 * sizeof(int32_t), "Statement is {int index, int end_plus_one}");
  }

  // h0: initial hand-off vector (NULL: zeros; irrelevant when the first block is pushed with explicit seeds)
  void chain_begin(int n_handoff, const adept::Real* h0 = 0) {
    n_ = n_handoff;
    check(adept_cuda_chain_begin(ctx_, n_handoff, h0));
  }

  // The block just recorded on this stack.  outputs[0..n): the active variables that receive the hand-off vector
  // (what adept::set_gradients(outputs, n, dJ) would seed); inputs[0..n): the active variables whose gradients are
  // handed on (what adept::get_gradients(inputs, n, dJ) would read).  Returns as soon as the tape is staged: the
  // next new_recording() may follow at once.
  template <class Active>
  void chain_push(const Active* outputs, const Active* inputs, int n) {
    slots(outputs, n, seed_);
    slots(inputs, n, out_);
    push(0);
  }
  // Same, seeded with explicit values (the first block: J.set_gradient(1.0))
  template <class Active>
  void chain_push_seeded(const Active* seed_vars, const adept::Real* seed_vals, int n_seed, const Active* inputs, int n) {
    slots(seed_vars, n_seed, seed_);
    slots(inputs, n, out_);
    push(seed_vals);
  }
  // Waits for the last block; dJ[0..n_handoff) = gradients at the last block's inputs
  void chain_end(adept::Real* dJ) { check(adept_cuda_chain_end(ctx_, dJ)); }

 private:
  template <class Active>
  static void slots(const Active* v, int n, std::vector<int32_t>& out) {
    out.resize(static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) out[static_cast<size_t>(i)] = static_cast<int32_t>(v[i].gradient_index());
  }
  void push(const adept::Real* seed_vals) {
    check(adept_cuda_chain_push(ctx_, statement_, static_cast<int64_t>(n_statements_), multiplier_,
                                reinterpret_cast<const int32_t*>(index_), static_cast<int64_t>(n_operations_),
                                static_cast<int64_t>(max_gradient_), seed_.empty() ? 0 : &seed_[0],
                                static_cast<int64_t>(seed_.size()), seed_vals, out_.empty() ? 0 : &out_[0],
                                static_cast<int64_t>(out_.size())));
  }
  void check(int rc) const {
    if (rc) throw chain_error(std::string("adept_cuda chain: ") + adept_cuda_last_error(ctx_));
  }
  adept_cuda_ctx* ctx_;
  int n_;
  std::vector<int32_t> seed_, out_;
};

}  // namespace adept_cuda
#endif  // ADEPT_CUDA_CHAIN_H
