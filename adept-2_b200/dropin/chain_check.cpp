// chain_check.cpp -- the reference's checkpointing example (test/test_checkpoint.cpp:136-225) three ways, with the
// plain Adept-2 API and the reference's own Toon scheme (benchmark/advection_schemes.h, included, NX = 100):
//
//   chain_check <nblocks> <nt> <out.bin> [gpu|record]
//
//   (a) `loop`  : exactly the reference's loop -- new_recording / set_gradients / reverse() / get_gradients per block
//                 -- on a plain adept::Stack linked against the drop-in (every reverse() is a GPU sweep with host
//                 gradient vectors);
//   (b) `chain` : the same blocks through adept_cuda::ChainStack (adept_cuda_chain.h): hand-off vector on the device,
//                 host recording of block i-1 overlapped with the GPU work on block i.
// Both dJ/dq_init vectors, the wall time of each variant and EVERY block's tape (so that the test can replay the
// chain through the oracle) are dumped to <out.bin>.  `record` makes no CUDA call: only the tapes are dumped.
#include <adept.h>
#include <advection_schemes.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "adept_cuda_chain.h"

using adept::adouble;
using adept::Real;

struct Dump {
  FILE* f;
  void i64(int64_t v) { fwrite(&v, 8, 1, f); }
  void raw(const void* p, size_t n) { if (n) fwrite(p, 1, n, f); }
};

// reads the protected tape arrays with no help from the chain header
class PeekStack : public adept::Stack {
 public:
  void dump_tape(Dump& d, const std::vector<int32_t>& seed, const std::vector<double>& seed_vals, const std::vector<int32_t>& out) const {
    d.i64(n_statements_); d.i64(n_operations_); d.i64(max_gradient_); d.i64(seed.size()); d.i64(seed_vals.size()); d.i64(out.size());
    d.raw(statement_, size_t(n_statements_) * 8);
    d.raw(multiplier_, size_t(n_operations_) * 8);
    d.raw(index_, size_t(n_operations_) * 4);
    d.raw(seed.data(), seed.size() * 4);
    d.raw(seed_vals.data(), seed_vals.size() * 8);
    d.raw(out.data(), out.size() * 4);
  }
};

template <class A>
static std::vector<int32_t> slots_of(const A* v, int n) {
  std::vector<int32_t> s(n);
  for (int i = 0; i < n; ++i) s[i] = int32_t(v[i].gradient_index());
  return s;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: chain_check <nblocks> <nt> <out.bin> [gpu|record]\n"); return 2; }
  const int nblocks = atoi(argv[1]), nt = atoi(argv[2]);
  const bool gpu = argc < 5 || strcmp(argv[4], "gpu") == 0;
  const double dt = 0.125, pi = 4.0 * atan(1.0);
  std::vector<double> q_init(NX);
  for (int i = 0; i < NX; i++) q_init[i] = (0.5 + 0.5 * sin((i * 2.0 * pi) / (NX - 1.5))) + 0.0001;   // test_checkpoint.cpp:62-64
  Dump d{fopen(argv[3], "wb")};
  if (!d.f) return 3;
  d.i64(nblocks); d.i64(nt); d.i64(NX); d.i64(gpu ? 1 : 0);
  std::vector<double> dJ_loop(NX, 0.0), dJ_chain(NX, 0.0);
  double ms_loop = 0, ms_chain = 0;
  const double one = 1.0;

  // ---- (a) the reference's loop; also dumps the tapes ----
  {
    PeekStack stack;
    std::vector<std::vector<adouble> > q_save(nblocks, std::vector<adouble>(NX));
    std::vector<adouble> q(NX);
    for (int i = 0; i < NX; i++) q_save[0][i] = q_init[i];
    for (int i = 0; i < nblocks - 1; i++) {
      stack.pause_recording();
      toon(nt, dt, &q_save[i][0], &q_save[i + 1][0]);
      stack.continue_recording();
    }
    const double t0 = now_ms();
    double rec_ms = 0;
    for (int i = nblocks - 1; i >= 0; i--) {
      const double tr = now_ms();
      stack.new_recording();
      toon(nt, dt, &q_save[i][0], &q[0]);
      adouble J = 0.0;
      std::vector<int32_t> seed;
      std::vector<double> seed_vals;
      if (i == nblocks - 1) {
        for (int k = 0; k < NX; k++) J += (q[k] - q_init[k]) * (q[k] - q_init[k]);
        seed = slots_of(&J, 1);
        seed_vals.assign(1, 1.0);
      } else {
        seed = slots_of(&q[0], NX);
      }
      rec_ms += now_ms() - tr;
      stack.dump_tape(d, seed, seed_vals, slots_of(&q_save[i][0], NX));
      if (gpu) {
        if (i == nblocks - 1) J.set_gradient(1.0);
        else adept::set_gradients(&q[0], NX, &dJ_loop[0]);
        stack.reverse();
        adept::get_gradients(&q_save[i][0], NX, &dJ_loop[0]);
      }
    }
    ms_loop = now_ms() - t0;
    d.i64(int64_t(rec_ms * 1000));
  }
  // ---- (b) the chain ----
  int chain_rc = 0;
  if (gpu) {
    adept_cuda_ctx* ctx = 0;
    const int dev = 0;
    chain_rc = adept_cuda_create(&ctx, &dev, 1);
    if (!chain_rc) {
      try {
        adept_cuda::ChainStack stack(ctx);
        std::vector<std::vector<adouble> > q_save(nblocks, std::vector<adouble>(NX));
        std::vector<adouble> q(NX);
        for (int i = 0; i < NX; i++) q_save[0][i] = q_init[i];
        for (int i = 0; i < nblocks - 1; i++) {
          stack.pause_recording();
          toon(nt, dt, &q_save[i][0], &q_save[i + 1][0]);
          stack.continue_recording();
        }
        const double t0 = now_ms();
        stack.chain_begin(NX);
        for (int i = nblocks - 1; i >= 0; i--) {
          stack.new_recording();
          toon(nt, dt, &q_save[i][0], &q[0]);
          if (i == nblocks - 1) {
            adouble J = 0.0;
            for (int k = 0; k < NX; k++) J += (q[k] - q_init[k]) * (q[k] - q_init[k]);
            stack.chain_push_seeded(&J, &one, 1, &q_save[i][0], NX);
          } else {
            stack.chain_push(&q[0], &q_save[i][0], NX);
          }
        }
        stack.chain_end(&dJ_chain[0]);
        ms_chain = now_ms() - t0;
      } catch (const std::exception& e) {
        fprintf(stderr, "chain_check: %s\n", e.what());
        chain_rc = 99;
      }
      adept_cuda_destroy(ctx);
    }
  }
  d.i64(chain_rc);
  d.i64(int64_t(ms_loop * 1000)); d.i64(int64_t(ms_chain * 1000));
  d.raw(dJ_loop.data(), NX * 8);
  d.raw(dJ_chain.data(), NX * 8);
  fclose(d.f);
  printf("chain_check: %d blocks x %d steps (NX=%d): loop %.2f ms, chain %.2f ms, rc %d\n", nblocks, nt, NX, ms_loop, ms_chain, chain_rc);
  return chain_rc;
}
