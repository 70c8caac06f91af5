// common.cuh -- shared types and sm_100a PTX helpers for libadept_cuda.
//
// Replay record: the device-side wire format of the tape.  The reference keeps three
// arrays (Statement{index,end_plus_one}, multiplier_, index_; include/adept/Statement.h:24-30,
// include/adept/StackStorageOrig.h:155-158).  On upload they are linearised into ONE stream
// of 16-byte records so that a tape segment is a single contiguous TMA bulk copy and the
// replay kernels never touch a statement table:
//
//   forward stream (tape order):   OP(m, slot)*  TAIL(lhs)      per statement
//   reverse stream (tape reversed): HEAD(lhs)  OP(m, slot)*     per statement, ops ascending
//
// A TAIL closes a statement in a forward sweep (store the accumulator, reset it to +0.0); a
// HEAD opens one in a reverse sweep (a = g[lhs]; g[lhs] = 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Device-side bounds checks of our own (this pool has no compute-sanitizer): `make CHECK=1` builds
// ../libadept_cuda_check.so with -DADEPT_BOUNDS_CHECK, every ADEPT_DCHECK becomes a device assert (a failed one traps
// the kernel and every later CUDA call fails loudly); $ADEPT_CUDA_LIB points the Python binding at that library, so
// the whole GPU test suite can be run against it.  In the product build the macro expands to nothing.
#ifdef ADEPT_BOUNDS_CHECK
#include <assert.h>
#define ADEPT_DCHECK(cond) assert(cond)
#else
#define ADEPT_DCHECK(cond) ((void)0)
#endif

namespace adept_b200 {

struct __align__(16) Rec {
  double  m;     // multiplier (OP); unused for HEAD/TAIL
  int32_t slot;  // gradient slot: operand (OP) or LHS (HEAD/TAIL)
  int32_t kind;  // REC_OP or REC_MARK
};
static_assert(sizeof(Rec) == 16, "Rec must be 16 bytes (TMA bulk-copy granule)");

enum : int32_t {
  REC_OP = 0,
  REC_MARK = 1,  // TAIL in a forward stream, HEAD in a reverse stream
  REC_ZMARK = 2  // fused reverse stream only: a HEAD whose sum starts from +0.0 (the target is an LHS of the step)
};

// ---------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS: UBLKCP) -- 1-D global -> shared.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  // make the initialised barrier visible to the async (TMA) proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream is
// still draining; pdl_trigger() lets OUR successor do the same, pdl_wait() blocks until the predecessor
// has completed and its writes are visible.  Everything before pdl_wait() must only touch data the
// predecessor does not write (kernel arguments, the immutable record streams).  Both are no-ops for a
// kernel launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Rounded, never-contracted fp64 multiply-add: parity with the reference built without FMA
// contraction ("multiply, round, add, round"; SURVEY.md section 7, oracle/replay_oracle.c).
__device__ __forceinline__ double mul_add_rn(double acc, double m, double g) {
  return __dadd_rn(acc, __dmul_rn(m, g));
}

}  // namespace adept_b200
