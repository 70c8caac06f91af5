// smallg_kernels.cuh -- Jacobian sweeps for tapes whose gradient list is small enough to live in
// SHARED MEMORY: max_gradient * Dc * 8 bytes (+ the reverse a-snapshot) <= ~200 KB for Dc = 32 or 64
// directions per CTA.  That is the regime of the reference's own benchmark and test programs
// (benchmark/autodiff_benchmark NX=100: 299 slots; test/test_thread_safe NX=128: 383 slots), where a
// step holds ~100 statements and the global-memory level kernels would spend their time in launches.
//
// One CTA owns Dc directions and walks ALL steps of the level schedule by itself: the barrier between
// steps is a __syncthreads, the gradient rows never leave the SM, and nothing but the record stream
// (L2-resident, read with warp-uniform loads) and the final Jacobian block touches global memory.
// A warp handles 32 directions of one statement (forward) or of one target slot (reverse) at a time;
// arithmetic and accumulation order are those of the reference (adept/jacobian.cpp:39-110, :382-431).
#pragma once
#include "common.cuh"

namespace adept_b200 {

struct SmallGArgs {
  // schedule (level-major forward stream / reverse target stream), all device pointers
  const Rec* fwd;               // OP* TAIL per scheduled statement
  const int64_t* foff;          // [S+1] record offset of every scheduled statement
  const Rec* tgt;               // HEAD(X) OP(m,k)* per target group
  const int64_t* goff;          // [n_groups+1]
  const int32_t* sched_lhs;     // [S]
  const int64_t* level_first;   // [n_levels+2]
  const int64_t* step_group;    // [n_levels+2]
  int32_t n_levels;
  // problem
  int64_t max_gradient;
  const int32_t* seed_slot;     // slot of direction d (global direction index)
  const int32_t* out_slot;      // slots to read out
  int64_t n_out;
  int64_t d_lo, d_hi;           // directions of this launch (global indices)
  int64_t d_base;               // direction index that maps to column 0 of `out`
  double* out;                  // out[i*stride_i + (d - d_base)*stride_d]
  int64_t stride_i, stride_d;
  int dc;                       // directions per CTA (multiple of 32)
  int max_width;                // widest step (rows of the reverse a-snapshot)
  int ref_block;
};

// shared memory: g[max_gradient][dc] then (reverse) abuf[max_width][dc]
template <bool REVERSE>
__global__ void __launch_bounds__(512) smallg_kernel(SmallGArgs A) {
  extern __shared__ __align__(16) double sm[];
  double* g = sm;
  double* abuf = sm + A.max_gradient * A.dc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int m = A.dc / 32;                 // 32-lane direction blocks per CTA
  const int blk = warp % m;                // my direction block
  const int wrow = warp / m, nrow = nwarps / m;   // my row among the warps sharing that block
  const int64_t dfirst = A.d_lo + static_cast<int64_t>(blockIdx.x) * A.dc;   // first direction of this CTA
  const int col = blk * 32 + lane;         // my column in the shared rows
  const int64_t d = dfirst + col;          // my global direction
  const bool ok = d < A.d_hi;

  for (int64_t i = threadIdx.x; i < A.max_gradient * A.dc; i += blockDim.x) g[i] = 0.0;
  __syncthreads();
  if (wrow == 0 && ok) g[static_cast<int64_t>(A.seed_slot[d]) * A.dc + col] = 1.0;
  __syncthreads();

  if (!REVERSE) {
    for (int32_t l = 1; l <= A.n_levels; ++l) {
      const int64_t j0 = A.level_first[l], j1 = A.level_first[l + 1];
      for (int64_t j = j0 + wrow; j < j1; j += nrow) {
        const int64_t b = A.foff[j], e = A.foff[j + 1];
        double acc = 0.0;
        for (int64_t k = b; k < e - 1; k += 8) {   // records are warp-uniform: eight loads in flight
          Rec rc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (k + u < e - 1) rc[u] = A.fwd[k + u];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (k + u < e - 1) acc = mul_add_rn(acc, rc[u].m, g[static_cast<int64_t>(rc[u].slot) * A.dc + col]);
        }
        g[static_cast<int64_t>(A.fwd[e - 1].slot) * A.dc + col] = acc;   // TAIL
      }
      __syncthreads();
    }
  } else {
    const uint32_t P = static_cast<uint32_t>(A.ref_block);
    const uint32_t grp_mask = (P >= 32 ? 0xffffffffu : ((1u << P) - 1u)) << ((lane / P) * P);
    for (int32_t l = A.n_levels; l >= 1; --l) {
      const int64_t j0 = A.level_first[l], n = A.level_first[l + 1] - j0;
      for (int64_t k = wrow; k < n; k += nrow) {                  // a_s = g[lhs]; g[lhs] = 0
        double* p = g + static_cast<int64_t>(A.sched_lhs[j0 + k]) * A.dc + col;
        abuf[k * A.dc + col] = *p;
        *p = 0.0;
      }
      __syncthreads();
      const int64_t g0 = A.step_group[l], g1 = A.step_group[l + 1];
      for (int64_t gi = g0 + wrow; gi < g1; gi += nrow) {         // gather per target, reference order
        const int64_t b = A.goff[gi], e = A.goff[gi + 1];
        double* p = g + static_cast<int64_t>(A.tgt[b].slot) * A.dc + col;
        double acc = *p;
        for (int64_t k = b + 1; k < e; k += 8) {
          Rec rc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (k + u < e) rc[u] = A.tgt[k + u];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (k + u < e) {
              const double a = abuf[static_cast<int64_t>(rc[u].slot) * A.dc + col];
              const uint32_t nz = __ballot_sync(0xffffffffu, ok && a != 0.0);
              if ((nz & grp_mask) != 0u) acc = mul_add_rn(acc, rc[u].m, a);
            }
        }
        *p = acc;
      }
      __syncthreads();
    }
  }
  // read out: out[i*stride_i + (d - d_base)*stride_d] = g[out_slot[i]][col]
  if (ok)
    for (int64_t i = wrow; i < A.n_out; i += nrow)
      A.out[i * A.stride_i + (d - A.d_base) * A.stride_d] = g[static_cast<int64_t>(A.out_slot[i]) * A.dc + col];
}

}  // namespace adept_b200
