// smallg_kernels.cuh -- Jacobian sweeps for tapes whose gradient list is small enough to live in
// SHARED MEMORY: max_gradient * Dc * 8 bytes (+ the reverse a-snapshot) <= ~150 KB for Dc = 32 or 64
// directions per CTA.  That is the regime of the reference's own benchmark and test programs
// (benchmark/autodiff_benchmark NX=100: 299 slots; test/test_thread_safe NX=128: 383 slots), where a
// step holds ~100 statements and the global-memory level kernels would spend their time in launches.
//
// One CTA owns Dc directions and walks ALL steps of the level schedule by itself: the barrier between
// steps is a __syncthreads, the gradient rows never leave the SM, and nothing but the record stream
// and the final Jacobian block touches global memory.  The records (and record offsets) of step l+1
// are prefetched into registers while step l is replayed out of shared memory, then parked in the
// other half of a double buffer; a step too large for the buffer is read straight from global memory.
// A warp handles 32 directions of one statement (forward) or of one target slot (reverse) at a time;
// arithmetic and accumulation order are those of the reference (adept/jacobian.cpp:39-110, :382-431).
#pragma once
#include "common.cuh"

namespace adept_b200 {

constexpr int kSgThreads = 512;
constexpr int kSgRecs = 1024;    // records of one step held in shared memory (2 per thread)
constexpr int kSgItems = 512;    // statements / target groups of one step (<= 2 offsets per thread)

struct SmallGArgs {
  // schedule (level-major forward stream / reverse target stream), all device pointers
  const Rec* fwd;               // OP* TAIL per scheduled statement
  const int64_t* foff;          // [S+1] record offset of every scheduled statement
  const Rec* tgt;               // HEAD(X) OP(m,k)* per target group
  const int64_t* goff;          // [n_groups+1]
  const int32_t* sched_lhs;     // [S]
  const int64_t* level_first;   // [n_levels+2]
  const int64_t* step_group;    // [n_levels+2]
  const longlong4* step_info;   // [n_levels+2] {first item, end item, first record, end record} of every step
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

struct SgStage {
  Rec rec[2][kSgRecs];
  int32_t off[2][kSgItems + 1];   // record offsets of the step's items, relative to the step's first record
  int32_t lhs[2][kSgItems];       // reverse: LHS slot of the step's statements
  int staged[2];                  // 1: the buffer holds the step
};

// dynamic shared memory: SgStage, then g[max_gradient][dc], then (reverse) abuf[max_width][dc]
template <bool REVERSE>
__global__ void __launch_bounds__(kSgThreads) smallg_kernel(SmallGArgs A) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  SgStage& S = *reinterpret_cast<SgStage*>(sm_raw);
  double* g = reinterpret_cast<double*>(sm_raw + ((sizeof(SgStage) + 15) / 16) * 16);
  double* abuf = g + A.max_gradient * A.dc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int m = A.dc / 32;                 // 32-lane direction blocks per CTA
  const int blk = warp % m;                // my direction block
  const int wrow = warp / m, nrow = nwarps / m;   // my row among the warps sharing that block
  const int64_t dfirst = A.d_lo + static_cast<int64_t>(blockIdx.x) * A.dc;   // first direction of this CTA
  const int col = blk * 32 + lane;         // my column in the shared rows
  const int64_t d = dfirst + col;          // my global direction
  const bool ok = d < A.d_hi;

  // item = statement (forward) or target group (reverse); its records are [ioff[i], ioff[i+1]) of `recs`
  const Rec* recs = REVERSE ? A.tgt : A.fwd;
  const int64_t* ioff = REVERSE ? A.goff : A.foff;

  for (int64_t i = threadIdx.x; i < A.max_gradient * A.dc; i += blockDim.x) g[i] = 0.0;
  if (threadIdx.x < 2) S.staged[threadIdx.x] = 0;
  __syncthreads();
  if (wrow == 0 && ok) g[static_cast<int64_t>(A.seed_slot[d]) * A.dc + col] = 1.0;

  // prefetch registers for the NEXT step
  Rec pre_r[kSgRecs / kSgThreads];
  int32_t pre_o[(kSgItems + kSgThreads) / kSgThreads];
  int32_t pre_l = 0;
  auto fetch = [&](const longlong4 si, int64_t j0, int64_t j1, bool& fits) {   // issue the loads of a step's records, offsets, LHS slots
    const int64_t i0 = si.x, i1 = si.y, r0 = si.z, r1 = si.w;
    fits = (i1 - i0) <= kSgItems && (r1 - r0) <= kSgRecs && (!REVERSE || (j1 - j0) <= kSgItems);
    if (!fits) return;
    if (REVERSE && j0 + threadIdx.x < j1) pre_l = A.sched_lhs[j0 + threadIdx.x];
#pragma unroll
    for (int u = 0; u < kSgRecs / kSgThreads; ++u) {
      const int64_t k = r0 + threadIdx.x + static_cast<int64_t>(u) * kSgThreads;
      if (k < r1) pre_r[u] = recs[k];
    }
#pragma unroll
    for (int u = 0; u < (kSgItems + kSgThreads) / kSgThreads; ++u) {
      const int64_t i = i0 + threadIdx.x + static_cast<int64_t>(u) * kSgThreads;
      if (i <= i1) pre_o[u] = static_cast<int32_t>(ioff[i] - r0);
    }
  };
  auto park = [&](const longlong4 si, int64_t j0, int64_t j1, int buf, bool fits) {   // registers -> shared buffer `buf`
    if (threadIdx.x == 0) S.staged[buf] = fits ? 1 : 0;
    if (!fits) return;
    if (REVERSE && j0 + threadIdx.x < j1) S.lhs[buf][threadIdx.x] = pre_l;
    const int64_t i0 = si.x, i1 = si.y;
    const int64_t nrec = si.w - si.z;
#pragma unroll
    for (int u = 0; u < kSgRecs / kSgThreads; ++u) {
      const int64_t k = threadIdx.x + static_cast<int64_t>(u) * kSgThreads;
      if (k < nrec) S.rec[buf][k] = pre_r[u];
    }
#pragma unroll
    for (int u = 0; u < (kSgItems + kSgThreads) / kSgThreads; ++u) {
      const int64_t i = threadIdx.x + static_cast<int64_t>(u) * kSgThreads;
      if (i <= i1 - i0) S.off[buf][i] = pre_o[u];
    }
  };

  const int32_t l_begin = REVERSE ? A.n_levels : 1, l_end = REVERSE ? 0 : A.n_levels + 1, l_step = REVERSE ? -1 : 1;
  int buf = 0;
  longlong4 si_cur = A.step_info[l_begin];
  longlong4 si_next = (l_begin + l_step != l_end) ? A.step_info[l_begin + l_step] : si_cur;
  // statement range [j0, j1) of the current / next step (reverse: the statements to snapshot)
  int64_t j0_cur = A.level_first[l_begin], j1_cur = A.level_first[l_begin + 1];
  int64_t j0_next = j0_cur, j1_next = j1_cur;
  if (l_begin + l_step != l_end) {
    j0_next = A.level_first[l_begin + l_step];
    j1_next = A.level_first[l_begin + l_step + 1];
  }
  {
    bool fits;
    fetch(si_cur, j0_cur, j1_cur, fits);
    park(si_cur, j0_cur, j1_cur, 0, fits);
  }
  __syncthreads();

  const uint32_t P = static_cast<uint32_t>(A.ref_block);
  const uint32_t grp_mask = (P >= 32 ? 0xffffffffu : ((1u << P) - 1u)) << ((lane / P) * P);

  for (int32_t l = l_begin; l != l_end; l += l_step) {
    const int32_t ln = l + l_step;
    bool next_fits = false;
    if (ln != l_end) fetch(si_next, j0_next, j1_next, next_fits);      // loads in flight while this step is replayed
    // the step after next: its table entries are independent loads, issued a whole step ahead
    const bool has2 = (ln != l_end && ln + l_step != l_end);
    const longlong4 si_next2 = has2 ? A.step_info[ln + l_step] : si_next;
    const int64_t j0_next2 = has2 ? A.level_first[ln + l_step] : j0_next;
    const int64_t j1_next2 = has2 ? A.level_first[ln + l_step + 1] : j1_next;

    const bool staged = S.staged[buf] != 0;
    const int64_t i0 = si_cur.x, i1 = si_cur.y;
    const int64_t rbase = staged ? 0 : si_cur.z;
    const Rec* R = staged ? &S.rec[buf][0] : recs + rbase;

    if (REVERSE) {
      const int64_t j0 = j0_cur, n = j1_cur - j0_cur;
      for (int64_t k = wrow; k < n; k += nrow) {                  // a_s = g[lhs]; g[lhs] = 0
        const int32_t lhs = staged ? S.lhs[buf][k] : A.sched_lhs[j0 + k];
        double* p = g + static_cast<int64_t>(lhs) * A.dc + col;
        abuf[k * A.dc + col] = *p;
        *p = 0.0;
      }
      __syncthreads();
    }
    for (int64_t i = i0 + wrow; i < i1; i += nrow) {
      const int64_t b = staged ? S.off[buf][i - i0] : (ioff[i] - rbase);
      const int64_t e = staged ? S.off[buf][i - i0 + 1] : (ioff[i + 1] - rbase);
      if (!REVERSE) {
        double acc = 0.0;
        for (int64_t k = b; k < e - 1; k += 4) {   // four records in flight: loads first, the ordered adds after
          Rec rc[4];
          double v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < e - 1) rc[u] = R[k + u];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < e - 1) v[u] = g[static_cast<int64_t>(rc[u].slot) * A.dc + col];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < e - 1) acc = mul_add_rn(acc, rc[u].m, v[u]);
        }
        g[static_cast<int64_t>(R[e - 1].slot) * A.dc + col] = acc;   // TAIL
      } else {
        double* p = g + static_cast<int64_t>(R[b].slot) * A.dc + col;  // HEAD(X)
        double acc = *p;
        for (int64_t k = b + 1; k < e; k += 4) {
          Rec rc[4];
          double v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < e) rc[u] = R[k + u];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < e) v[u] = abuf[static_cast<int64_t>(rc[u].slot) * A.dc + col];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < e) {
              const uint32_t nz = __ballot_sync(0xffffffffu, ok && v[u] != 0.0);
              if ((nz & grp_mask) != 0u) acc = mul_add_rn(acc, rc[u].m, v[u]);
            }
        }
        *p = acc;
      }
    }
    if (ln != l_end) park(si_next, j0_next, j1_next, buf ^ 1, next_fits);
    __syncthreads();
    buf ^= 1;
    si_cur = si_next;
    si_next = si_next2;
    j0_cur = j0_next;
    j1_cur = j1_next;
    j0_next = j0_next2;
    j1_next = j1_next2;
  }
  // read out: out[i*stride_i + (d - d_base)*stride_d] = g[out_slot[i]][col]
  if (ok)
    for (int64_t i = wrow; i < A.n_out; i += nrow)
      A.out[i * A.stride_i + (d - A.d_base) * A.stride_d] = g[static_cast<int64_t>(A.out_slot[i]) * A.dc + col];
}

}  // namespace adept_b200
