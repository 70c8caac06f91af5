// replay_kernels.cuh -- direction-vectorised tape replay for sm_100a.
//
// Replaces the hot loops of the reference:
//   forward : adept/jacobian.cpp:39-110   (jacobian_forward_kernel[_extra]) and adept/Stack.cpp:170-190
//   reverse : adept/jacobian.cpp:382-431  (sweep inside jacobian_reverse*)   and adept/Stack.cpp:139-164
//
// Layout: the gradient workspace is direction-major, g[slot*K + d] (the reference's
// g[index*MULTIPASS_SIZE + lane] with MULTIPASS_SIZE -> K).  One warp lane owns one direction
// d = 32*column_block + lane, so every gather/scatter of a slot row is one coalesced 256-byte
// access.  The tape arrives as a stream of 16-byte records (common.cuh); each CTA pulls its
// segment of the stream into shared memory with 1-D TMA bulk copies (cp.async.bulk +
// mbarrier transaction counts) through a ring of stages, and all warps of the CTA -- i.e. all
// the directions the CTA owns -- consume the same staged records.
//
// A CTA replays one "chunk" [chunk_begin[c], chunk_begin[c+1]) of the stream:
//   * tape-ordered schedule: ONE chunk = the whole tape; the sweep is strictly sequential.
//   * level schedule: the stream is permuted level-major by the device-side dependency
//     analysis (analysis.cuh); the chunks of one level are launched together and are mutually
//     independent (no two statements of a level touch a common slot).
//
// Each warp keeps U gradient loads in flight ("batch").  Hazards inside a batch (a record
// whose slot was written by an earlier record of the same batch) are resolved in registers by
// forwarding, so the result is exactly that of sequential execution; hazards between batches
// are ordered by the hardware (same thread, same address).  All arithmetic is
// __dmul_rn/__dadd_rn: never contracted, bit-identical to the reference built without FMA.
#pragma once
#include "common.cuh"

namespace adept_b200 {

constexpr int kBatch = 8;          // gradient loads in flight per warp
constexpr int kPieceRecs = 256;    // records per TMA stage (4 KB)
constexpr int kStages = 4;         // ring depth

struct ReplayArgs {
  const Rec* recs;             // record stream
  const int64_t* chunk_begin;  // chunk table (record offsets); chunk c = [chunk_begin[c], chunk_begin[c+1])
  int64_t first_chunk;         // chunk index of blockIdx.x == 0
  double* g;                   // workspace, [max_gradient][K]
  int64_t K;                   // row pitch in doubles (multiple of 32)
  int n_colblocks;             // 32-lane column blocks in this pass
  int64_t n_dirs;              // directions resident in this pass (lanes beyond are masked)
  int ref_block;               // P of the reference's skip test (reverse only)
};

struct StageRing {
  alignas(16) Rec buf[kStages][kPieceRecs];
  alignas(8) uint64_t full[kStages];
};

// Issue the TMA copy of piece `p` of the chunk into its ring slot.  One thread calls this.
__device__ __forceinline__ void issue_piece(StageRing& ring, const Rec* chunk, int64_t n_recs, int64_t p) {
  const int st = static_cast<int>(p % kStages);
  const int64_t off = p * kPieceRecs;
  const int64_t left = n_recs - off;
  const uint32_t cnt = static_cast<uint32_t>(left < kPieceRecs ? left : kPieceRecs);
  mbar_expect_tx(&ring.full[st], cnt * static_cast<uint32_t>(sizeof(Rec)));
  tma_bulk_g2s(&ring.buf[st][0], chunk + off, cnt * static_cast<uint32_t>(sizeof(Rec)), &ring.full[st]);
}

// ---------------------------------------------------------------------------------------
// Forward sweep.  CHECK=true adds the in-batch forwarding test (needed whenever records of
// a chunk may depend on each other: the tape-ordered schedule).  CHECK=false is for level
// chunks, where no OP reads a slot that any TAIL of the same level writes.
// ---------------------------------------------------------------------------------------
template <bool CHECK>
__global__ void __launch_bounds__(256) replay_forward_kernel(ReplayArgs A) {
  __shared__ StageRing ring;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int64_t c = A.first_chunk + blockIdx.x;
  const int64_t r0 = A.chunk_begin[c], r1 = A.chunk_begin[c + 1];
  const int64_t n_recs = r1 - r0;
  const Rec* chunk = A.recs + r0;
  const int64_t n_pieces = (n_recs + kPieceRecs - 1) / kPieceRecs;
  const int cb = blockIdx.y * nwarps + warp;
  const bool live = cb < A.n_colblocks;
  const bool lane_ok = static_cast<int64_t>(cb) * 32 + lane < A.n_dirs;
  double* gcol = A.g + static_cast<int64_t>(cb) * 32 + lane;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int64_t p = 0; p < n_pieces && p < kStages; ++p) issue_piece(ring, chunk, n_recs, p);

  double acc = 0.0;
  for (int64_t p = 0; p < n_pieces; ++p) {
    const int st = static_cast<int>(p % kStages);
    mbar_wait(&ring.full[st], static_cast<uint32_t>((p / kStages) & 1));
    const int64_t left = n_recs - p * kPieceRecs;
    const int n = static_cast<int>(left < kPieceRecs ? left : kPieceRecs);
    const Rec* r = &ring.buf[st][0];
    if (live) {
      for (int i = 0; i < n; i += kBatch) {
        Rec rc[kBatch];
        double v[kBatch];
        double stv[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (i + u < n) rc[u] = r[i + u];
          else { rc[u].m = 0.0; rc[u].slot = -1; rc[u].kind = REC_OP; }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          v[u] = 0.0;
          if (lane_ok && rc[u].kind == REC_OP && rc[u].slot >= 0) v[u] = gcol[static_cast<int64_t>(rc[u].slot) * A.K];
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          stv[u] = 0.0;
          if (rc[u].kind == REC_OP) {
            if (rc[u].slot >= 0) {
              double x = v[u];
              if (CHECK) {
#pragma unroll
                for (int w = 0; w < u; ++w)
                  if (rc[w].kind == REC_MARK && rc[w].slot == rc[u].slot) x = stv[w];
              }
              acc = mul_add_rn(acc, rc[u].m, x);
            }
          } else {  // TAIL: the statement is complete
            if (lane_ok) gcol[static_cast<int64_t>(rc[u].slot) * A.K] = acc;
            stv[u] = acc;
            acc = 0.0;
          }
        }
      }
    }
    __syncthreads();  // every warp is done with this stage
    if (threadIdx.x == 0 && p + kStages < n_pieces) issue_piece(ring, chunk, n_recs, p + kStages);
  }
}

// ---------------------------------------------------------------------------------------
// Reverse sweep.  HEAD: a = g[lhs]; g[lhs] = 0.  OP: g[slot] += m*a unless every lane of the
// reference's P-lane block holds a == 0 (adept/jacobian.cpp:392-407: the statement is skipped
// only when ALL lanes of the block are zero; otherwise all lanes are updated).  Forwarding is
// always on: a statement may name one slot several times, or its own LHS.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) replay_reverse_kernel(ReplayArgs A) {
  __shared__ StageRing ring;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int64_t c = A.first_chunk + blockIdx.x;
  const int64_t r0 = A.chunk_begin[c], r1 = A.chunk_begin[c + 1];
  const int64_t n_recs = r1 - r0;
  const Rec* chunk = A.recs + r0;
  const int64_t n_pieces = (n_recs + kPieceRecs - 1) / kPieceRecs;
  const int cb = blockIdx.y * nwarps + warp;
  const bool live = cb < A.n_colblocks;
  const bool lane_ok = static_cast<int64_t>(cb) * 32 + lane < A.n_dirs;
  double* gcol = A.g + static_cast<int64_t>(cb) * 32 + lane;
  // lanes of my reference block: P consecutive directions, P | 32
  const uint32_t P = static_cast<uint32_t>(A.ref_block);
  const uint32_t grp_mask = (P >= 32 ? 0xffffffffu : ((1u << P) - 1u)) << ((lane / P) * P);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int64_t p = 0; p < n_pieces && p < kStages; ++p) issue_piece(ring, chunk, n_recs, p);

  double a = 0.0;
  bool act = false;
  for (int64_t p = 0; p < n_pieces; ++p) {
    const int st = static_cast<int>(p % kStages);
    mbar_wait(&ring.full[st], static_cast<uint32_t>((p / kStages) & 1));
    const int64_t left = n_recs - p * kPieceRecs;
    const int n = static_cast<int>(left < kPieceRecs ? left : kPieceRecs);
    const Rec* r = &ring.buf[st][0];
    if (live) {
      for (int i = 0; i < n; i += kBatch) {
        Rec rc[kBatch];
        double v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (i + u < n) rc[u] = r[i + u];
          else { rc[u].m = 0.0; rc[u].slot = -1; rc[u].kind = REC_OP; }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          v[u] = 0.0;
          if (lane_ok && rc[u].slot >= 0) v[u] = gcol[static_cast<int64_t>(rc[u].slot) * A.K];
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (rc[u].slot < 0) continue;
          double x = v[u];
#pragma unroll
          for (int w = 0; w < u; ++w)
            if (rc[w].slot == rc[u].slot) x = v[w];  // v[w] holds what record w left in memory
          double* dst = gcol + static_cast<int64_t>(rc[u].slot) * A.K;
          if (rc[u].kind == REC_MARK) {  // HEAD
            a = x;
            const uint32_t nz = __ballot_sync(0xffffffffu, a != 0.0);
            act = (nz & grp_mask) != 0u;
            v[u] = 0.0;
            if (lane_ok) *dst = 0.0;
          } else {
            if (act) {
              x = mul_add_rn(x, rc[u].m, a);
              if (lane_ok) *dst = x;
            }
            v[u] = x;
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && p + kStages < n_pieces) issue_piece(ring, chunk, n_recs, p + kStages);
  }
}

// ---------------------------------------------------------------------------------------
// Level-scheduled reverse sweep, one step = two kernels (tape_kernels.cuh, "target stream"):
//   1. reverse_snapshot_kernel: for every statement of the step  a_s = g[lhs]; g[lhs] = 0,
//      with a_s parked in abuf[k][K] (k = rank of the statement inside the step).
//   2. reverse_target_kernel:   for every slot X named by the step's operations
//      g[X] = (..((g[X] + m1*a_s1) + m2*a_s2)..), contributions in the reference's order
//      (adept/jacobian.cpp:410-428: statements descending, operands ascending).
// A contribution is dropped when the contributing statement's a is zero in every lane of the
// reference's P-lane block (adept/jacobian.cpp:392-407), as in the scatter kernel above.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reverse_snapshot_kernel(const int32_t* __restrict__ sched_lhs, int64_t j0,
                                                               int64_t n_stmts, double* __restrict__ g,
                                                               double* __restrict__ abuf, int64_t K, int64_t n_dirs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
  const int64_t d = static_cast<int64_t>(blockIdx.y) * 32 + lane;
  if (k >= n_stmts || d >= n_dirs) return;
  double* p = g + static_cast<int64_t>(sched_lhs[j0 + k]) * K + d;
  abuf[k * K + d] = *p;
  *p = 0.0;
}

struct TargetArgs {
  const Rec* recs;             // target stream
  const int64_t* chunk_begin;  // its chunk table
  int64_t first_chunk;
  double* g;
  const double* abuf;
  int64_t K;
  int n_colblocks;
  int64_t n_dirs;
  int ref_block;
};

__global__ void __launch_bounds__(256) reverse_target_kernel(TargetArgs A) {
  __shared__ StageRing ring;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int64_t c = A.first_chunk + blockIdx.x;
  const int64_t r0 = A.chunk_begin[c], r1 = A.chunk_begin[c + 1];
  const int64_t n_recs = r1 - r0;
  const Rec* chunk = A.recs + r0;
  const int64_t n_pieces = (n_recs + kPieceRecs - 1) / kPieceRecs;
  const int cb = blockIdx.y * nwarps + warp;
  const bool live = cb < A.n_colblocks;
  const bool lane_ok = static_cast<int64_t>(cb) * 32 + lane < A.n_dirs;
  double* gcol = A.g + static_cast<int64_t>(cb) * 32 + lane;
  const double* acol = A.abuf + static_cast<int64_t>(cb) * 32 + lane;
  const uint32_t P = static_cast<uint32_t>(A.ref_block);
  const uint32_t grp_mask = (P >= 32 ? 0xffffffffu : ((1u << P) - 1u)) << ((lane / P) * P);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int64_t p = 0; p < n_pieces && p < kStages; ++p) issue_piece(ring, chunk, n_recs, p);

  double acc = 0.0;
  int32_t cur = -1;  // slot whose sum is in acc
  for (int64_t p = 0; p < n_pieces; ++p) {
    const int st = static_cast<int>(p % kStages);
    mbar_wait(&ring.full[st], static_cast<uint32_t>((p / kStages) & 1));
    const int64_t left = n_recs - p * kPieceRecs;
    const int n = static_cast<int>(left < kPieceRecs ? left : kPieceRecs);
    const Rec* r = &ring.buf[st][0];
    if (live) {
      for (int i = 0; i < n; i += kBatch) {
        Rec rc[kBatch];
        double v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (i + u < n) rc[u] = r[i + u];
          else { rc[u].m = 0.0; rc[u].slot = -1; rc[u].kind = REC_OP; }
        }
        // all loads of a step are independent of its stores: targets are distinct, abuf is read-only
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          v[u] = 0.0;
          if (lane_ok && rc[u].slot >= 0)
            v[u] = (rc[u].kind == REC_MARK) ? gcol[static_cast<int64_t>(rc[u].slot) * A.K]
                                            : acol[static_cast<int64_t>(rc[u].slot) * A.K];
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (rc[u].slot < 0) continue;
          if (rc[u].kind == REC_MARK) {  // next target: retire the previous one
            if (cur >= 0 && lane_ok) gcol[static_cast<int64_t>(cur) * A.K] = acc;
            cur = rc[u].slot;
            acc = v[u];
          } else {
            const double a = v[u];
            const uint32_t nz = __ballot_sync(0xffffffffu, a != 0.0);
            if ((nz & grp_mask) != 0u) acc = mul_add_rn(acc, rc[u].m, a);
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && p + kStages < n_pieces) issue_piece(ring, chunk, n_recs, p + kStages);
  }
  if (live && cur >= 0 && lane_ok) gcol[static_cast<int64_t>(cur) * A.K] = acc;
}

// ---------------------------------------------------------------------------------------
// Seeding and extraction (adept/jacobian.cpp:163-190 forward, :372-379 and :434-449 reverse).
// ---------------------------------------------------------------------------------------
// g[seed_slot[d0+j]*K + j] = 1.0 for the n directions of this pass.
__global__ void seed_kernel(double* g, int64_t K, const int32_t* seed_slot, int64_t d0, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) g[static_cast<int64_t>(seed_slot[d0 + j]) * K + j] = 1.0;
}

// out[i*stride_i + (d0+j)*stride_j] = g[out_slot[i]*K + j], i < n_out, j < n_dirs.
// 32x32 tiles through shared memory so that both the row reads and the writes coalesce
// whichever of stride_i / stride_j is the unit one.
__global__ void extract_kernel(const double* __restrict__ g, int64_t K, const int32_t* __restrict__ out_slot,
                               int64_t n_out, int n_dirs, int64_t d0, double* __restrict__ out,
                               int64_t stride_i, int64_t stride_j) {
  __shared__ double tile[32][33];
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int j0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r;
    const int j = j0 + threadIdx.x;
    if (i < n_out && j < n_dirs) tile[r][threadIdx.x] = g[static_cast<int64_t>(out_slot[i]) * K + j];
  }
  __syncthreads();
  if (stride_i < stride_j) {  // consecutive i are close in memory: lanes run over i
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int64_t i = i0 + threadIdx.x;
      const int j = j0 + r;
      if (i < n_out && j < n_dirs) out[i * stride_i + (d0 + j) * stride_j] = tile[threadIdx.x][r];
    }
  } else {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int64_t i = i0 + r;
      const int j = j0 + threadIdx.x;
      if (i < n_out && j < n_dirs) out[i * stride_i + (d0 + j) * stride_j] = tile[r][threadIdx.x];
    }
  }
}

}  // namespace adept_b200
