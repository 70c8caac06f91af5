// sweep1_kernels.cuh -- level-scheduled SINGLE-direction sweeps (Stack::compute_adjoint,
// adept/Stack.cpp:139-164; Stack::compute_tangent_linear, adept/Stack.cpp:170-190).
//
// With one direction there is nothing to vectorise across lanes, so parallelism comes from
// the statements of a level (device-side dependency analysis, tape_kernels.cuh): one thread
// replays one statement of the level-major forward stream (OP* TAIL).  No two statements of
// a level touch a common slot, so plain loads/stores suffice and every slot sees its
// accumulations in the reference's order.  Statements longer than `long_threshold` records
// (e.g. the 4(N-1)-operand `sum` of the Rosenbrock cost, include/adept/reduce.h:162-174) are
// skipped here and replayed by the cooperative "long" kernels below.
#pragma once
#include "common.cuh"

namespace adept_b200 {

// foff[j] .. foff[j+1]: records of scheduled statement j (foff[S] = R).
__global__ void sweep1_forward_level_kernel(const Rec* __restrict__ recs, const int64_t* __restrict__ foff,
                                            int64_t j0, int64_t j1, int64_t long_threshold, double* g) {
  const int64_t j = j0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= j1) return;
  const int64_t b = foff[j], e = foff[j + 1];
  if (e - b > long_threshold) return;
  double acc = 0.0;
  for (int64_t k = b; k < e - 1; ++k) {
    const Rec r = recs[k];
    acc = mul_add_rn(acc, r.m, g[r.slot]);
  }
  g[recs[e - 1].slot] = acc;
}

// Reverse step, kernel 1: a_s = g[lhs]; g[lhs] = 0 for the statements [j0, j0+n) of the step.
__global__ void sweep1_snapshot_kernel(const int32_t* __restrict__ sched_lhs, int64_t j0, int64_t n, double* g,
                                       double* __restrict__ abuf) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t lhs = sched_lhs[j0 + k];
  abuf[k] = g[lhs];
  g[lhs] = 0.0;
}
// Reverse step, kernel 2: one thread per target group [goff[gi], goff[gi+1]) of the target
// stream: HEAD(X) OP(m,k)*.  A zero a contributes nothing (adept/Stack.cpp:151).
__global__ void sweep1_target_kernel(const Rec* __restrict__ tgt, const int64_t* __restrict__ goff, int64_t g0,
                                     int64_t g1, double* g, const double* __restrict__ abuf) {
  const int64_t gi = g0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gi >= g1) return;
  const int64_t b = goff[gi], e = goff[gi + 1];
  const int32_t x = tgt[b].slot;
  double acc = g[x];
  for (int64_t k = b + 1; k < e; ++k) {
    const Rec r = tgt[k];
    const double a = abuf[r.slot];
    if (a != 0.0) acc = mul_add_rn(acc, r.m, a);
  }
  g[x] = acc;
}

// One long statement, forward, ONE CTA: the products m*g are formed in parallel (each is
// rounded on its own, exactly as the reference does without FMA), then one thread adds them
// in tape order -- the only order that reproduces the reference's bits.
constexpr int kLongTile = 4096;
// One CTA per long statement of the step: range[2*blockIdx.x] = {b, e} of its records.
__global__ void __launch_bounds__(512) sweep1_forward_long_kernel(const Rec* __restrict__ recs,
                                                                   const int64_t* __restrict__ range, double* g) {
  __shared__ double prod[kLongTile];
  __shared__ double acc_s;
  const int64_t b = range[2 * blockIdx.x], e = range[2 * blockIdx.x + 1];
  if (threadIdx.x == 0) acc_s = 0.0;
  __syncthreads();
  for (int64_t t0 = b; t0 < e - 1; t0 += kLongTile) {
    const int64_t n = (e - 1 - t0) < kLongTile ? (e - 1 - t0) : kLongTile;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      const Rec r = recs[t0 + k];
      prod[k] = __dmul_rn(r.m, g[r.slot]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double acc = acc_s;
      for (int64_t k = 0; k < n; ++k) acc = __dadd_rn(acc, prod[k]);
      acc_s = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) g[recs[e - 1].slot] = acc_s;
}

// list long statements of the schedule: out[3*k] = {j, first record, end record}, k = atomic++ (the count keeps
// running past max_out so that the host can size a second attempt)
__global__ void find_long_kernel(const int64_t* __restrict__ foff, int64_t S, int64_t long_threshold,
                                 int64_t* __restrict__ out, int* __restrict__ count, int max_out) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = t; j < S; j += stride) {
    const int64_t b = foff[j], e = foff[j + 1];
    if (e - b > long_threshold) {
      const int k = atomicAdd(count, 1);
      if (k < max_out) {
        out[3 * k] = j;
        out[3 * k + 1] = b;
        out[3 * k + 2] = e;
      }
    }
  }
}

}  // namespace adept_b200
