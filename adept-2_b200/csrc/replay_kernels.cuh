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
//     analysis (tape_kernels.cuh); the chunks of one step are launched together and are mutually
//     independent (RAW/WAR/WAW between slot versions are respected by the step numbers).
//     Level sweeps use forward_level_kernel / reverse_snapshot_kernel + reverse_target_kernel below
//     (two directions per lane, block sparsity); the kernels in this first part replay a stream
//     in tape order.
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
constexpr int kPieceRecs = 256;    // records per TMA stage (4 KB); level chunks (~128 records) fit one stage
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
  // ensemble mode (adept_cuda_jacobian_batch): lane L = member*dirs_padded + direction; the record's m field
  // holds the OPERATION INDEX and the multiplier is per member: bmult[op*n_members + member]
  const double* bmult;
  int64_t n_members;
  int dirs_padded;
};

struct StageRing {
  alignas(16) Rec buf[kStages][kPieceRecs];
  alignas(8) uint64_t full[kStages];
};

// Issue the TMA copy of piece `p` of the chunk into its ring slot.  One thread calls this.
// `slot_seq` is the position in the ring's issue sequence when a chunk is streamed more than once (default: p).
__device__ __forceinline__ void issue_piece(StageRing& ring, const Rec* chunk, int64_t n_recs, int64_t p, int64_t slot_seq = -1) {
  const int st = static_cast<int>((slot_seq < 0 ? p : slot_seq) % kStages);
  const int64_t off = p * kPieceRecs;
  const int64_t left = n_recs - off;
  const uint32_t cnt = static_cast<uint32_t>(left < kPieceRecs ? left : kPieceRecs);
  mbar_expect_tx(&ring.full[st], cnt * static_cast<uint32_t>(sizeof(Rec)));
  tma_bulk_g2s(&ring.buf[st][0], chunk + off, cnt * static_cast<uint32_t>(sizeof(Rec)), &ring.full[st]);
}

// ---------------------------------------------------------------------------------------
// Tape-ordered forward sweep.  CHECK=true adds the in-batch forwarding test (needed whenever
// records of a chunk may depend on each other, i.e. always in tape order).  BATCH=true is the
// ensemble mode of adept_cuda_jacobian_batch (per-member multipliers).
// ---------------------------------------------------------------------------------------
template <bool CHECK, bool BATCH>
__global__ void __launch_bounds__(256, 3) replay_forward_kernel(ReplayArgs A) {
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
  const double* mcol = nullptr;  // ensemble mode: this lane's member column of the multiplier matrix
  if (BATCH && lane_ok) mcol = A.bmult + (static_cast<int64_t>(cb) * 32 + lane) / A.dirs_padded;

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
              double m = rc[u].m;
              if (BATCH) m = lane_ok ? mcol[__double_as_longlong(rc[u].m) * A.n_members] : 0.0;
              acc = mul_add_rn(acc, m, x);
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
template <bool BATCH>
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
  const double* mcol = nullptr;
  if (BATCH && lane_ok) mcol = A.bmult + (static_cast<int64_t>(cb) * 32 + lane) / A.dirs_padded;
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
              double m = rc[u].m;
              if (BATCH) m = lane_ok ? mcol[__double_as_longlong(rc[u].m) * A.n_members] : 0.0;
              x = mul_add_rn(x, m, a);
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
// Structural sparsity.  A Jacobian row/column usually touches a small part of the tape (the
// reference exploits this in reverse mode by skipping statements whose adjoint block is zero,
// adept/jacobian.cpp:392-407).  Here the unit is (slot row, 32-lane column block):
//   rowact[slot*ncb + cb] == 0  =>  g[slot][32*cb .. 32*cb+31] is all +0.0
//   aflag[k*ncb + cb]     == 0  =>  the snapshot a_s of statement k is all zero in that block
// A zero block contributes exactly nothing in the reference either (the whole block is zero, so
// every P-lane sub-block is skipped), so eliding its loads, arithmetic and stores keeps all bits.
// V = doubles per lane (1 or 2).  With V = 2 a lane owns two adjacent directions and every row access is
// a 16-byte vector load/store (512 bytes per warp): twice the bytes in flight per warp for the same
// instruction count.  A "column block" is 32*V directions; rowact/aflag have one byte per block.
template <int V>
struct LaneVec;
template <>
struct LaneVec<1> {
  double x;
  __device__ __forceinline__ void load(const double* p) { x = *p; }
  __device__ __forceinline__ void store(double* p) const { *p = x; }
  __device__ __forceinline__ void zero() { x = 0.0; }
};
template <>
struct LaneVec<2> {
  double x, y;
  __device__ __forceinline__ void load(const double* p) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    x = t.x;
    y = t.y;
  }
  __device__ __forceinline__ void store(double* p) const { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
  __device__ __forceinline__ void zero() { x = 0.0; y = 0.0; }
};

template <>
struct LaneVec<4> {   // four adjacent directions per lane: 1024 bytes per warp and row (forward_balanced_kernel<4>)
  double x, y, z, w;
  __device__ __forceinline__ void load(const double* p) {
    const double2 a = reinterpret_cast<const double2*>(p)[0];
    const double2 b = reinterpret_cast<const double2*>(p)[1];
    x = a.x; y = a.y; z = b.x; w = b.y;
  }
  __device__ __forceinline__ void store(double* p) const {
    reinterpret_cast<double2*>(p)[0] = make_double2(x, y);
    reinterpret_cast<double2*>(p)[1] = make_double2(z, w);
  }
  __device__ __forceinline__ void zero() { x = 0.0; y = 0.0; z = 0.0; w = 0.0; }
};

// One thread looks at one (statement, column block) pair; the active pairs of the CTA are
// compacted in shared memory and moved one row segment per warp.
template <int V, int RB>
__global__ void __launch_bounds__(256) reverse_snapshot_kernel(const int32_t* __restrict__ sched_lhs, int64_t j0,
                                                               int64_t n_stmts, double* __restrict__ g,
                                                               double* __restrict__ abuf, int64_t K, int64_t n_dirs,
                                                               uint8_t* __restrict__ rowact,
                                                               uint8_t* __restrict__ aflag, int pitch, int ncb_live,
                                                               int cb_lo) {
  __shared__ int list[256];
  __shared__ int cnt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x;
  const int64_t total = n_stmts * ncb_live;
  pdl_wait();   // the previous step's target kernel wrote the rows and activity bytes read below
  {
    const int64_t t = t0 + threadIdx.x;
    if (t < total) {
      const int64_t k = t / ncb_live;
      const int cb = cb_lo + static_cast<int>(t - k * ncb_live);
      const int64_t lhs = sched_lhs[j0 + k];
      if (rowact[lhs * pitch + cb]) list[atomicAdd(&cnt, 1)] = threadIdx.x;
      else aflag[k * pitch + cb] = 0;  // the row is all zero: a_s = 0, nothing to move
    }
  }
  __syncthreads();
  const int n = cnt;
  const int nw = blockDim.x >> 5;
  // RB = row segments in flight per warp
  for (int q0 = warp * RB; q0 < n; q0 += nw * RB) {
    LaneVec<V> a[RB];
    double* p[RB];
    int64_t kk[RB], ll[RB];
    int cbs[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      p[u] = nullptr;
      if (q0 + u < n) {
        const int64_t t = t0 + list[q0 + u];
        kk[u] = t / ncb_live;
        cbs[u] = cb_lo + static_cast<int>(t - kk[u] * ncb_live);
        ll[u] = sched_lhs[j0 + kk[u]];
        p[u] = g + ll[u] * K + (static_cast<int64_t>(cbs[u]) * 32 + lane) * V;  // rows are padded to K: in bounds
        a[u].load(p[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      if (q0 + u >= n) break;
      const int64_t d = (static_cast<int64_t>(cbs[u]) * 32 + lane) * V;
      bool nzl = (a[u].x != 0.0) && (d < n_dirs);
      if (V == 2) nzl = nzl || ((reinterpret_cast<const double*>(&a[u])[V - 1] != 0.0) && (d + 1 < n_dirs));
      const bool any = __ballot_sync(0xffffffffu, nzl) != 0u;
      if (any) {
        a[u].store(abuf + kk[u] * K + d);
        LaneVec<V> z;
        z.zero();
        z.store(p[u]);
      }
      if (lane == 0) {
        aflag[kk[u] * pitch + cbs[u]] = any ? 1 : 0;
        rowact[ll[u] * pitch + cbs[u]] = 0;  // zeroed (or found zero)
      }
    }
  }
}

struct TargetArgs {
  const Rec* recs;             // target stream
  const int64_t* chunk_begin;  // its chunk table
  int64_t first_chunk;
  double* g;
  const double* abuf;
  uint8_t* rowact;
  const uint8_t* aflag;
  int64_t K;
  int n_colblocks;   // column blocks [cb_lo, n_colblocks) are handled by this launch
  int cb_lo;
  int flag_pitch;    // row pitch of rowact/aflag
  int64_t n_dirs;
  int ref_block;
  int cbs;           // reverse_fused_kernel: column blocks per CTA (<= kFusedCbs)
  int split;         // fused kernels: CTAs per (chunk, column group), each owning a range of the chunk's sub-tasks (<= 1: one)
  int sub;           // fused kernels: records per sub-task (<= 0: the whole piece is one)
  // reverse_fused_kernel: "tail" steps.  Steps of a chunk or two (boundary statements between two wide levels, scalar
  // chains) are not worth a launch of their own: the launch of the step before them carries them along, and the
  // CTA that finishes last (ticket counter *done) runs them one after the other.
  int n_tail;
  int64_t tail_c0[4], tail_c1[4];   // chunk range of every tail step, in sweep order
  unsigned int* done;
  unsigned long long* trace;   // profiling aid (built with -DADEPT_FUSED_TRACE only): 8 time stamps per launch
};
constexpr int kMaxTail = 4;
#ifdef ADEPT_FUSED_TRACE
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define FUSED_TRACE_MIN(k) do { if (A.trace && threadIdx.x == 0) atomicMin(&A.trace[k], gtimer()); } while (0)
#define FUSED_TRACE_MAX(k) do { if (A.trace && threadIdx.x == 0) atomicMax(&A.trace[k], gtimer()); } while (0)
#else
#define FUSED_TRACE_MIN(k) do { } while (0)
#define FUSED_TRACE_MAX(k) do { } while (0)
#endif

// The CTA first fetches, in one round trip and for all its warps, the activity byte of every
// record of the staged piece (pf); a warp whose column block has no active contribution in the
// piece skips it outright, so dead (target, column block) pairs cost no gradient traffic.
// (3 CTAs of 8 warps per SM: at 4 the register cap forces spills into the inner loop, measured 20% slower)
template <int V>
__global__ void __launch_bounds__(256, 3) reverse_target_kernel(TargetArgs A) {
  __shared__ StageRing ring;
  __shared__ uint8_t pf[kPieceRecs][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int64_t c = A.first_chunk + blockIdx.x;
  const int64_t r0 = A.chunk_begin[c], r1 = A.chunk_begin[c + 1];
  const int64_t n_recs = r1 - r0;
  const Rec* chunk = A.recs + r0;
  const int64_t n_pieces = (n_recs + kPieceRecs - 1) / kPieceRecs;
  const int cb0 = A.cb_lo + blockIdx.y * nwarps;
  const int cb = cb0 + warp;
  const int ncb = A.flag_pitch;
  const bool live = cb < A.n_colblocks;
  const int64_t d0 = (static_cast<int64_t>(cb) * 32 + lane) * V;  // first direction of this lane; rows are padded to K
  const bool ok0 = d0 < A.n_dirs, ok1 = d0 + 1 < A.n_dirs;
  double* gcol = A.g + d0;
  const double* acol = A.abuf + d0;
  uint8_t* ract = A.rowact + cb;
  // the reference's P-lane block of my direction(s): with V = 2 and P >= 2 a block is P/2 lanes
  const uint32_t P = static_cast<uint32_t>(A.ref_block);
  const uint32_t lanes_per_grp = (V == 2) ? (P >= 2 ? P / 2 : 1) : P;
  const uint32_t grp_mask =
      (lanes_per_grp >= 32 ? 0xffffffffu : ((1u << lanes_per_grp) - 1u)) << ((lane / lanes_per_grp) * lanes_per_grp);
  const bool per_component = (V == 2 && P == 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_trigger();
  if (threadIdx.x == 0)
    for (int64_t p = 0; p < n_pieces && p < kStages; ++p) issue_piece(ring, chunk, n_recs, p);
  pdl_wait();   // records (immutable) are already in flight; rows, snapshot and activity bytes come from the previous kernel

  LaneVec<V> acc;
  acc.zero();
  int32_t cur = -1;        // slot whose sum is (or will be) in acc
  bool acc_valid = false;  // acc holds g[cur] (+ contributions)
  bool dirty = false;      // acc differs from what memory holds
  for (int64_t p = 0; p < n_pieces; ++p) {
    const int st = static_cast<int>(p % kStages);
    mbar_wait(&ring.full[st], static_cast<uint32_t>((p / kStages) & 1));
    const int64_t left = n_recs - p * kPieceRecs;
    const int n = static_cast<int>(left < kPieceRecs ? left : kPieceRecs);
    const Rec* r = &ring.buf[st][0];
    // 1. activity bytes of the piece: thread t -> record t / nwarps, column block cb0 + t % nwarps
    for (int t = threadIdx.x; t < n * nwarps; t += blockDim.x) {
      const int i = t / nwarps, w = t - i * nwarps;
      const int32_t slot = r[i].slot;
      uint8_t f = 0;
      if (cb0 + w < A.n_colblocks && slot >= 0)
        f = (r[i].kind == REC_MARK) ? A.rowact[static_cast<int64_t>(slot) * ncb + cb0 + w]
                                    : A.aflag[static_cast<int64_t>(slot) * ncb + cb0 + w];
      pf[i][w] = f;
    }
    __syncthreads();
    if (live) {
      // does any contribution of this piece reach my column block?
      bool mine = false;
      for (int i = lane; i < n; i += 32) mine |= (r[i].kind == REC_OP && pf[i][warp] != 0);
      if (!__any_sync(0xffffffffu, mine)) {
        // nothing to add: retire the open target; remember the last HEAD for a group that goes on
        if (dirty) {
          acc.store(gcol + static_cast<int64_t>(cur) * A.K);
          if (lane == 0) ract[static_cast<int64_t>(cur) * ncb] = 1;
          dirty = false;
        }
        int last = -1;
        for (int i = lane; i < n; i += 32)
          if (r[i].kind == REC_MARK) last = i;
        last = __reduce_max_sync(0xffffffffu, last);
        if (last >= 0) {
          cur = r[last].slot;
          acc_valid = false;
        }
      } else {
        for (int i = 0; i < n; i += kBatch) {
          LaneVec<V> v[kBatch];
          // data of the active records only; all loads of a step are independent of its stores
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            v[u].zero();
            if (i + u < n && pf[i + u][warp]) {
              const int32_t slot = r[i + u].slot;
              v[u].load(((r[i + u].kind == REC_MARK) ? gcol : acol) + static_cast<int64_t>(slot) * A.K);
            }
          }
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            if (i + u >= n) break;
            const Rec rc = r[i + u];   // re-read from shared memory: keeps the batch's registers for data
            if (rc.kind == REC_MARK) {  // next target: retire the previous one
              if (dirty) {
                acc.store(gcol + static_cast<int64_t>(cur) * A.K);
                if (lane == 0) ract[static_cast<int64_t>(cur) * ncb] = 1;
              }
              cur = rc.slot;
              acc = v[u];  // +0.0 when the row is inactive
              acc_valid = true;
              dirty = false;
            } else if (pf[i + u][warp]) {
              if (!acc_valid) {  // the group began in a piece this warp skipped
                acc.zero();
                if (ract[static_cast<int64_t>(cur) * ncb]) acc.load(gcol + static_cast<int64_t>(cur) * A.K);
                acc_valid = true;
              }
              if (V == 1) {
                const uint32_t nz = __ballot_sync(0xffffffffu, ok0 && v[u].x != 0.0);
                if ((nz & grp_mask) != 0u) acc.x = mul_add_rn(acc.x, rc.m, v[u].x);
              } else {
                double* av = reinterpret_cast<double*>(&v[u]);
                double* cv = reinterpret_cast<double*>(&acc);
                const bool n0 = ok0 && av[0] != 0.0, n1 = ok1 && av[V - 1] != 0.0;
                bool act0, act1;
                if (per_component) {
                  act0 = n0;
                  act1 = n1;
                } else {
                  const uint32_t nz = __ballot_sync(0xffffffffu, n0 || n1);
                  act0 = act1 = (nz & grp_mask) != 0u;
                }
                if (act0) cv[0] = mul_add_rn(cv[0], rc.m, av[0]);
                if (act1) cv[V - 1] = mul_add_rn(cv[V - 1], rc.m, av[V - 1]);
              }
              dirty = true;
            }
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && p + kStages < n_pieces) issue_piece(ring, chunk, n_recs, p + kStages);
  }
  if (live && dirty) {
    acc.store(gcol + static_cast<int64_t>(cur) * A.K);
    if (lane == 0) ract[static_cast<int64_t>(cur) * ncb] = 1;
  }
}

// ---------------------------------------------------------------------------------------
// Level-scheduled reverse sweep, one step = ONE kernel (tape_kernels.cuh, "FUSED reverse stream").
// The workspace has 2*G rows (two per slot); records name physical rows.  Per target group:
//   MARK(row)   acc = g[row]   (the target's live row, updated in place)
//   ZMARK(row)  acc = +0.0     (the target is an LHS of this step: its NEW version goes to its other row)
//   OP(m, arow) acc += m * g[arow]   (arow = live row of the contributing statement's LHS, i.e. a_s)
// Within a step no row is both read by one group and written by another, so all loads of a step are
// independent of all its stores and one launch per step suffices.  rowact[row*ncb + cb] == 0 means the
// block is all +0.0 whatever the (possibly stale) memory holds; a ZMARK's bytes are cleared by the CTA
// that stages it, for all the column blocks the CTA owns, before any warp may set them again.
// Against snapshot + gather this saves a launch per step and, per statement and direction, the 8-byte
// read + 8-byte zero store of g[lhs] and the 8-byte write of the a_s copy.
// ---------------------------------------------------------------------------------------
//
// CTA shape.  A CTA owns one chunk of records and up to 32 column blocks (A.cbs).  Jacobians are
// structurally sparse (a Toon row touches about a quarter of the (statement, column block) pairs, and the
// reference skips the rest as well, adept/jacobian.cpp:392-407), so binding warp w to column block w would
// leave most warps idle while a few walk the dependent round trips.  Instead the CTA reads the activity
// bytes of all its (record, column block) pairs in one round trip, forms the mask of column blocks that
// receive any contribution from the chunk, and its warps take the ACTIVE column blocks in turn.
constexpr int kFusedCbs = 32;   // column blocks per CTA (<= 32: the activity mask is one word)

template <int V>
struct FusedState {
  LaneVec<V> acc;
  int32_t cur;     // row whose sum is (or will be) in acc
  bool cur_zero;   // the open group began with a ZMARK (its sum starts from +0.0)
  bool acc_valid;  // acc holds the start value of cur (+ contributions)
  bool dirty;      // acc differs from what memory holds
  __device__ __forceinline__ void reset() {
    acc.zero();
    cur = -1;
    cur_zero = false;
    acc_valid = false;
    dirty = false;
  }
};

struct FusedLane {
  double* gcol;     // &g[0][first direction of this lane]
  uint8_t* ract;    // &rowact[0][column block]
  int64_t K;
  int ncb;
  int lane;
  bool ok0, ok1;
  uint32_t grp_mask;
  bool per_component;
};

template <int V>
__device__ __forceinline__ void fused_retire(FusedState<V>& S, const FusedLane& L) {
  if (S.dirty) {
    if (!L.ok0) S.acc.zero();   // lanes beyond the last direction stay +0.0 in memory (fused_task relies on it)
    if (V == 2 && !L.ok1) reinterpret_cast<double*>(&S.acc)[V - 1] = 0.0;
    S.acc.store(L.gcol + static_cast<int64_t>(S.cur) * L.K);
    if (L.lane == 0) L.ract[static_cast<int64_t>(S.cur) * L.ncb] = 1;
    S.dirty = false;
  }
}

// records [0, n) of a staged piece for ONE column block; pfw[i] = activity byte of record i there
template <int V>
__device__ __forceinline__ void fused_piece(const Rec* __restrict__ r, int n, const uint8_t* __restrict__ pfw,
                                            FusedState<V>& S, const FusedLane& L) {
  for (int i = 0; i < n; i += kBatch) {
    LaneVec<V> v[kBatch];
    // rows of the active records only; all loads of a step are independent of its stores
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      v[u].zero();
      if (i + u < n && pfw[i + u]) v[u].load(L.gcol + static_cast<int64_t>(r[i + u].slot) * L.K);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (i + u >= n) break;
      const Rec rc = r[i + u];   // re-read from shared memory: keeps the batch's registers for data
      if (rc.kind != REC_OP) {   // next target: retire the previous one
        fused_retire<V>(S, L);
        S.cur = rc.slot;
        S.cur_zero = (rc.kind == REC_ZMARK);
        S.acc = v[u];  // +0.0 for a ZMARK and for an inactive live row
        S.acc_valid = true;
      } else if (pfw[i + u]) {
        if (!S.acc_valid) {  // the group began in a piece this warp skipped
          S.acc.zero();
          if (!S.cur_zero && L.ract[static_cast<int64_t>(S.cur) * L.ncb]) S.acc.load(L.gcol + static_cast<int64_t>(S.cur) * L.K);
          S.acc_valid = true;
        }
        if (V == 1) {
          const uint32_t nz = __ballot_sync(0xffffffffu, L.ok0 && v[u].x != 0.0);
          if ((nz & L.grp_mask) != 0u) S.acc.x = mul_add_rn(S.acc.x, rc.m, v[u].x);
        } else {
          double* av = reinterpret_cast<double*>(&v[u]);
          double* cv = reinterpret_cast<double*>(&S.acc);
          const bool n0 = L.ok0 && av[0] != 0.0, n1 = L.ok1 && av[V - 1] != 0.0;
          bool act0, act1;
          if (L.per_component) {
            act0 = n0;
            act1 = n1;
          } else {
            const uint32_t nz = __ballot_sync(0xffffffffu, n0 || n1);
            act0 = act1 = (nz & L.grp_mask) != 0u;
          }
          if (act0) cv[0] = mul_add_rn(cv[0], rc.m, av[0]);
          if (act1) cv[V - 1] = mul_add_rn(cv[V - 1], rc.m, av[V - 1]);
        }
        S.dirty = true;
      }
    }
  }
}

template <int V>
__device__ __forceinline__ FusedLane fused_lane(const TargetArgs& A, int cb, int lane) {
  FusedLane L;
  const int64_t d0 = (static_cast<int64_t>(cb) * 32 + lane) * V;  // first direction of this lane; rows are padded to K
  L.gcol = A.g + d0;
  L.ract = A.rowact + cb;
  L.K = A.K;
  L.ncb = A.flag_pitch;
  L.lane = lane;
  L.ok0 = d0 < A.n_dirs;
  L.ok1 = d0 + 1 < A.n_dirs;
  // the reference's P-lane block of my direction(s): with V = 2 and P >= 2 a block is P/2 lanes
  const uint32_t P = static_cast<uint32_t>(A.ref_block);
  const uint32_t lanes_per_grp = (V == 2) ? (P >= 2 ? P / 2 : 1) : P;
  L.grp_mask = (lanes_per_grp >= 32 ? 0xffffffffu : ((1u << lanes_per_grp) - 1u)) << ((lane / lanes_per_grp) * lanes_per_grp);
  L.per_component = (V == 2 && P == 1);
  return L;
}

// Single-piece fast path, one column block: the warp walks only the heads and the ACTIVE operations of the
// chunk, found by bit-scanning warp-uniform masks (32 records per mask word), eight row loads in flight.
// LP: how the reference's P-lane skip rule (adept/jacobian.cpp:392-407) maps to lanes holding V = 2 directions:
//   0  P = 1: per component;  1  P = 2: the lane's own pair;  2  P >= 4: P/2 adjacent lanes (warp ballot).
template <int V, int LP>
__device__ __forceinline__ void fused_task(const Rec* __restrict__ r, int i0, int i1, const uint8_t* __restrict__ pfw,
                                           const uint32_t* __restrict__ headm, uint4* __restrict__ wl,
                                           const FusedLane& L) {
  // records [i0, i1) of the staged piece: i0 is a head (or i0 == i1), i1 is a head or the end of the piece, so the
  // range holds whole target groups and ranges of one chunk can be walked by different warps (fused_subtask_range)
  // Per block of 32 records the lanes decode ONE record each (row byte offset = row * bytes per row, flags) and
  // park the heads and the active operations, compacted and in order, in the warp's list wl[]:
  //   x = low word of the byte offset | 1 (head) | 2 (row is active: load it), y = high word,
  //   (z, w) = multiplier bits (operation) or z = row (head)
  // so that the serial walk below costs one 16-byte shared-memory read and one 64-bit add per record.
  LaneVec<V> acc;
  acc.zero();
  char* const gbase = reinterpret_cast<char*>(L.gcol);
  const uint32_t Kb = static_cast<uint32_t>(L.K) * 8u;   // bytes per row: a multiple of 512, the low bits are free
  uint64_t cur_off = 0;
  uint32_t cur_row = 0u;
  bool dirty = false;
  const uint32_t lt_mask = (1u << L.lane) - 1u;
  const int nblk = (i1 + 31) >> 5;
  for (int blk = i0 >> 5; blk < nblk; ++blk) {
    const int i = (blk << 5) + L.lane;
    const bool valid = i >= i0 && i < i1;
    const bool a = valid && pfw[i] != 0;
    const uint32_t act = __ballot_sync(0xffffffffu, a);
    const uint32_t hm = headm[blk] & __ballot_sync(0xffffffffu, valid);
    const uint32_t work = act | hm;
    if (work == 0u) continue;
    if ((work >> L.lane) & 1u) {
      const Rec rc = r[i];
      const bool head = (hm >> L.lane) & 1u;
      const uint64_t off = static_cast<uint64_t>(static_cast<uint32_t>(rc.slot)) * Kb;
      uint4 e;
      e.x = static_cast<uint32_t>(off) | (head ? 1u : 0u) | (a ? 2u : 0u);
      e.y = static_cast<uint32_t>(off >> 32);
      const uint64_t mb = static_cast<uint64_t>(__double_as_longlong(rc.m));
      e.z = head ? static_cast<uint32_t>(rc.slot) : static_cast<uint32_t>(mb);
      e.w = static_cast<uint32_t>(mb >> 32);
      wl[__popc(work & lt_mask)] = e;
    }
    __syncwarp();
    const int cnt = __popc(work);
    for (int j = 0; j < cnt; j += kBatch) {
      LaneVec<V> v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        v[u].zero();
        if (j + u < cnt) {
          const uint2 e = *reinterpret_cast<const uint2*>(&wl[j + u]);
          if (e.x & 2u)
            v[u].load(reinterpret_cast<const double*>(gbase + ((static_cast<uint64_t>(e.y) << 32) | (e.x & ~15u))));
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (j + u >= cnt) break;
        const uint4 e = wl[j + u];
        if (e.x & 1u) {   // next target: retire the previous one
          if (dirty) {
            if (!L.ok0) acc.zero();   // lanes beyond the last direction stay +0.0 in memory (see below)
            if (V == 2 && !L.ok1) reinterpret_cast<double*>(&acc)[V - 1] = 0.0;
            acc.store(reinterpret_cast<double*>(gbase + cur_off));
            if (L.lane == 0) L.ract[static_cast<int64_t>(cur_row) * L.ncb] = 1;
            dirty = false;
          }
          cur_off = (static_cast<uint64_t>(e.y) << 32) | (e.x & ~15u);
          cur_row = e.z;
          acc = v[u];     // +0.0 for a ZMARK (never active) and for an inactive live row
        } else {          // an active contribution
          const double m = __longlong_as_double(static_cast<long long>((static_cast<uint64_t>(e.w) << 32) | e.z));
          double* av = reinterpret_cast<double*>(&v[u]);
          double* cv = reinterpret_cast<double*>(&acc);
          // padded lanes (direction >= n_dirs) always read +0.0: they are never seeded and are masked at every
          // store, so they cannot wake a reference block even when a non-finite multiplier makes them NaN
          const bool n0 = av[0] != 0.0, n1 = av[V - 1] != 0.0;
          bool act0, act1;
          if (LP == 0) {
            act0 = n0;
            act1 = n1;
          } else if (LP == 1) {
            act0 = act1 = n0 || n1;
          } else {
            const uint32_t nz = __ballot_sync(0xffffffffu, n0 || n1);
            act0 = act1 = (nz & L.grp_mask) != 0u;
          }
          if (act0) cv[0] = mul_add_rn(cv[0], m, av[0]);
          if (V == 2 && act1) cv[V - 1] = mul_add_rn(cv[V - 1], m, av[V - 1]);
          dirty = true;
        }
      }
    }
    __syncwarp();   // the list is rewritten for the next block
  }
  if (dirty) {
    if (!L.ok0) acc.zero();
    if (V == 2 && !L.ok1) reinterpret_cast<double*>(&acc)[V - 1] = 0.0;
    acc.store(reinterpret_cast<double*>(gbase + cur_off));
    if (L.lane == 0) L.ract[static_cast<int64_t>(cur_row) * L.ncb] = 1;
  }
}

// Sub-tasks of a staged piece: the piece is cut at the first head at or after every multiple of `sub` records.  A
// Jacobian's activity is banded (a chunk's targets either see all of a device's directions or none), so with few
// directions per device a step's work sits in a few CTAs whose warps each walk a whole chunk while most of the machine
// idles.  The host then (TargetArgs::split, ::sub) gives every chunk to `split` CTAs, each owning a range of the
// sub-tasks, and the warps of a CTA take (sub-task, active column block) pairs in turn: one or two batches of row
// loads per warp instead of eight.  With many directions the machine is full anyway and sub = kPieceRecs, split = 1
// keep the walks long (full batches).
__device__ __forceinline__ int fused_first_head(const uint32_t* __restrict__ headm, int pos, int n) {
  for (int blk = pos >> 5; (blk << 5) < n; ++blk) {
    uint32_t w = headm[blk];
    if (blk == (pos >> 5)) w &= 0xffffffffu << (pos & 31);
    if (w) {
      const int i = (blk << 5) + __ffs(w) - 1;
      return i < n ? i : n;
    }
  }
  return n;
}
// sub-task s of a piece of n records: [first head >= s*sub, first head >= (s+1)*sub)
__device__ __forceinline__ void fused_subtask_range(const uint32_t* __restrict__ headm, int n, int sub, int s, int& i0, int& i1) {
  i0 = (s == 0) ? 0 : fused_first_head(headm, s * sub, n);
  i1 = ((s + 1) * sub >= n) ? n : fused_first_head(headm, (s + 1) * sub, n);
}

constexpr int kPfPitch = kPieceRecs + 4;   // bytes per column block of the activity table (+4: conflict-free columns)

// The common case of both fused reverse kernels: a chunk that is ONE staged piece r[0, n).  The CTA owns the column
// blocks [cb0, cb0 + cbs) and part `part` of `nparts` of the piece's sub-tasks.  All threads of the CTA call this;
// between() runs after the activity bytes are in (a hook for the resident kernel's look-ahead).
template <int V, int LP, class Between>
__device__ __forceinline__ void fused_single_piece(const TargetArgs& A, const Rec* __restrict__ r, int n, int cb0, int cbs,
                                                   int part, int nparts, uint8_t* pf, uint32_t* headm, uint4 (*wlist)[32],
                                                   uint32_t* smask, Between between) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int ncb = A.flag_pitch;
  // head mask of every block of 32 records
  for (int i = threadIdx.x; i < ((n + 31) & ~31); i += blockDim.x) {
    const int32_t kind = i < n ? r[i].kind : REC_OP;
    const uint32_t h = __ballot_sync(0xffffffffu, kind != REC_OP);
    if (lane == 0) headm[i >> 5] = h;
  }
  const int sub = A.sub > 0 ? A.sub : kPieceRecs;
  const int nsub = (n + sub - 1) / sub;
  int s_lo = 0, s_hi = nsub, i0 = 0, i1 = n;
  if (nparts > 1) {   // my share of the sub-tasks and the records they cover
    __syncthreads();
    s_lo = part * nsub / nparts;
    s_hi = (part + 1) * nsub / nparts;
    i0 = (s_lo == 0) ? 0 : fused_first_head(headm, s_lo * sub, n);
    i1 = (s_hi * sub >= n) ? n : fused_first_head(headm, s_hi * sub, n);
    if (s_hi <= s_lo) i1 = i0;
  }
  ADEPT_DCHECK(0 <= i0 && i0 <= n && i1 <= n);
  // activity bytes: thread t -> record i0 + t / cbs, column block cb0 + t % cbs (consecutive bytes of a row)
  uint32_t seen = 0u;   // column blocks for which this thread saw an active contribution
  for (int t = threadIdx.x; t < (i1 - i0) * cbs; t += blockDim.x) {
    const int i = i0 + t / cbs, w = t % cbs;
    const int32_t row = r[i].slot;
    const int32_t kind = r[i].kind;
    uint8_t* b = A.rowact + static_cast<int64_t>(row) * ncb + cb0 + w;
    uint8_t f = 0;
    if (kind == REC_ZMARK) *b = 0;   // new version of an LHS slot: +0.0 until a contribution lands
    else f = *b;
    pf[w * kPfPitch + i] = f;
    if (kind == REC_OP && f) seen |= 1u << w;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) seen |= __shfl_xor_sync(0xffffffffu, seen, d);
  if (lane == 0 && seen) atomicOr(smask, seen);
  __syncthreads();
  between();
  const uint32_t mask = *smask;
  // tasks = (sub-task) x (active column block), handed to the warps in turn from the last warp down (warp 0 hosts the
  // thread that issues copies and looks ahead)
  const int nact = __popc(mask);
  for (int t = nwarps - 1 - warp; t < nact * (s_hi - s_lo); t += nwarps) {
    const int sb = s_lo + t / nact;
    const int w = __fns(mask, 0, t % nact + 1);
    int j0, j1;
    fused_subtask_range(headm, n, sub, sb, j0, j1);
    if (j0 >= j1) continue;
    const FusedLane L = fused_lane<V>(A, cb0 + w, lane);
    fused_task<V, LP>(r, j0, j1, &pf[w * kPfPitch], headm, &wlist[warp][0], L);
  }
}

template <int V, int LP>
__global__ void __launch_bounds__(256, 4) reverse_fused_kernel(TargetArgs A) {
  __shared__ StageRing ring;
  __shared__ __align__(4) uint8_t pf[kFusedCbs * kPfPitch];   // pf[w*kPfPitch + i]: record i, column block cb0 + w
  __shared__ uint32_t headm[kPieceRecs / 32];
  __shared__ uint4 wlist[8][32];   // per warp: the decoded heads and active operations of a block of 32 records
  __shared__ uint32_t smask;
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int ncb = A.flag_pitch;
  // work items of this CTA: its own (chunk, column group) and, if it turns out to be the last CTA of a launch
  // that carries tail steps, every (chunk, column group) of those steps, step after step
  const int split = A.split > 1 ? A.split : 1;   // CTAs per (chunk, column group): each owns a range of the chunk's sub-tasks
  int64_t c = A.first_chunk + blockIdx.x / split;
  int part = static_cast<int>(blockIdx.x % split), nparts = split;
  int cby = blockIdx.y;
  int tail = -1;   // -1: own item; >= 0: tail step being run
  FUSED_TRACE_MIN(0);
  for (;;) {
    if (threadIdx.x == 0) {
      if (tail >= 0)
        for (int s = 0; s < kStages; ++s) mbar_inval(&ring.full[s]);
      for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
      mbar_fence_init();
      smask = 0u;
    }
    __syncthreads();
    if (tail < 0) pdl_trigger();
    const int64_t r0 = A.chunk_begin[c], r1 = A.chunk_begin[c + 1];
    const int64_t n_recs = r1 - r0;
    const Rec* chunk = A.recs + r0;
    const int64_t n_pieces = (n_recs + kPieceRecs - 1) / kPieceRecs;
    const int cb0 = A.cb_lo + cby * A.cbs;
    const int cbs = min(A.cbs, A.n_colblocks - cb0);   // column blocks of this item
    if (n_pieces == 1) {
      // ---- the common case: the chunk is one staged piece; warps take the active column blocks in turn ----
      const int n = static_cast<int>(n_recs);
      if (threadIdx.x == 0) issue_piece(ring, chunk, n_recs, 0);
      if (tail < 0) pdl_wait();   // records (immutable) are already in flight; rows and activity bytes come from the previous step
      if (tail < 0) { FUSED_TRACE_MIN(1); FUSED_TRACE_MAX(2); }
      mbar_wait(&ring.full[0], 0u);
      const Rec* r = &ring.buf[0][0];
      fused_single_piece<V, LP>(A, r, n, cb0, cbs, part, nparts, pf, headm, wlist, &smask,
                                [&]() { if (tail < 0) FUSED_TRACE_MAX(3); });
    } else if (part != 0) {
      if (tail < 0) pdl_wait();   // a long chunk is not split: it belongs to the first of its CTAs (the others may still run tails)
    } else {
      // ---- a chunk longer than one piece (a target group of more than kPieceRecs contributions): the chunk is
      // streamed once per set of nwarps column blocks, warp w bound to one column block, state carried across pieces
      const int n_sets = (cbs + nwarps - 1) / nwarps;
      const int64_t total = n_pieces * n_sets;
      if (threadIdx.x == 0)
        for (int64_t it = 0; it < total && it < kStages; ++it) issue_piece(ring, chunk, n_recs, it % n_pieces, it);
      if (tail < 0) pdl_wait();
      FusedState<V> S;
      S.reset();
      for (int64_t it = 0; it < total; ++it) {
        const int64_t p = it % n_pieces;
        const int set = static_cast<int>(it / n_pieces);
        const int w = set * nwarps + warp;
        const bool live = w < cbs;
        const FusedLane L = fused_lane<V>(A, cb0 + (live ? w : 0), lane);
        if (p == 0) S.reset();
        const int st = static_cast<int>(it % kStages);
        mbar_wait(&ring.full[st], static_cast<uint32_t>((it / kStages) & 1));
        const int64_t left = n_recs - p * kPieceRecs;
        const int n = static_cast<int>(left < kPieceRecs ? left : kPieceRecs);
        const Rec* r = &ring.buf[st][0];
        for (int t = threadIdx.x; t < n * nwarps; t += blockDim.x) {
          const int i = t / nwarps, ww = t - i * nwarps;
          uint8_t f = 0;
          if (set * nwarps + ww < cbs) {
            uint8_t* b = A.rowact + static_cast<int64_t>(r[i].slot) * ncb + cb0 + set * nwarps + ww;
            if (r[i].kind == REC_ZMARK) *b = 0;
            else f = *b;
          }
          pf[ww * kPfPitch + i] = f;
        }
        __syncthreads();
        if (live) {
          bool mine = false;
          for (int i = lane; i < n; i += 32) mine |= (r[i].kind == REC_OP && pf[warp * kPfPitch + i] != 0);
          if (!__any_sync(0xffffffffu, mine)) {
            // nothing to add: retire the open target; remember the last head for a group that goes on
            fused_retire<V>(S, L);
            int last = -1;
            for (int i = lane; i < n; i += 32)
              if (r[i].kind != REC_OP) last = i;
            last = __reduce_max_sync(0xffffffffu, last);
            if (last >= 0) {
              S.cur = r[last].slot;
              S.cur_zero = (r[last].kind == REC_ZMARK);
              S.acc_valid = false;
            }
          } else {
            fused_piece<V>(r, n, &pf[warp * kPfPitch], S, L);
          }
          if (p == n_pieces - 1) fused_retire<V>(S, L);
        }
        __syncthreads();
        if (threadIdx.x == 0 && it + kStages < total) issue_piece(ring, chunk, n_recs, (it + kStages) % n_pieces, it + kStages);
      }
    }
    // ---- the next item, if any ----
    if (tail < 0) FUSED_TRACE_MAX(4);
    if (A.n_tail == 0) break;
    __syncthreads();   // every warp of this CTA has issued its stores
    if (tail < 0) FUSED_TRACE_MAX(5);
    if (tail < 0) {
      if (threadIdx.x == 0) {
        __threadfence();   // release: this CTA's rows and activity bytes
        s_last = (atomicAdd(A.done, 1u) == gridDim.x * gridDim.y - 1u) ? 1 : 0;
      }
      __syncthreads();
      if (!s_last) break;
      __threadfence();     // acquire: what every other CTA of the launch wrote (drops this SM's L1 lines as well)
      tail = 0;
      c = A.tail_c0[0];
      cby = 0;
      part = 0;     // a tail item is one chunk x one column group, whole
      nparts = 1;
    } else {
      if (++cby == static_cast<int>(gridDim.y)) {
        cby = 0;
        ++c;
      }
      if (c == A.tail_c1[tail]) {
        if (++tail == A.n_tail) break;
        __threadfence();   // the next tail step reads what this one wrote
        c = A.tail_c0[tail];
      }
    }
  }
  FUSED_TRACE_MAX(6);
}

// ---------------------------------------------------------------------------------------
// Level-scheduled forward sweep (one launch per step): statements of a step are mutually
// independent gathers, so loads are batched freely.  Same CTA shape, TMA staging, lane vectors
// and block-sparsity bytes as reverse_target_kernel: a statement whose operand blocks are all
// zero yields +0.0 -- exactly what the reference computes from finite multipliers -- so it is
// not evaluated (SPARSE is switched off by the host when the tape holds a non-finite multiplier,
// because 0*inf must then produce the reference's NaN).
// ---------------------------------------------------------------------------------------
struct ForwardArgs {
  const Rec* recs;             // level-major forward stream: OP* TAIL per statement
  const int64_t* chunk_begin;
  int64_t first_chunk;
  double* g;
  uint8_t* rowact;
  int64_t K;
  int n_colblocks;   // column blocks [cb_lo, n_colblocks) are handled by this launch
  int cb_lo;
  int flag_pitch;
  int cbs;           // forward_balanced_kernel: column blocks per CTA (<= kFwdCbs)
  int64_t n_rows;    // rows of the workspace (bounds checks of the -DADEPT_BOUNDS_CHECK build)
};

template <int V, bool SPARSE>
__global__ void __launch_bounds__(256, 3) forward_level_kernel(ForwardArgs A) {
  __shared__ StageRing ring;
  __shared__ uint8_t pf[kPieceRecs][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int64_t c = A.first_chunk + blockIdx.x;
  const int64_t r0 = A.chunk_begin[c], r1 = A.chunk_begin[c + 1];
  const int64_t n_recs = r1 - r0;
  const Rec* chunk = A.recs + r0;
  const int64_t n_pieces = (n_recs + kPieceRecs - 1) / kPieceRecs;
  const int cb0 = A.cb_lo + blockIdx.y * nwarps;
  const int cb = cb0 + warp;
  const int ncb = A.flag_pitch;
  const bool live = cb < A.n_colblocks;
  double* gcol = A.g + (static_cast<int64_t>(cb) * 32 + lane) * V;  // rows are padded to K: always in bounds
  uint8_t* ract = A.rowact + cb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_trigger();
  if (threadIdx.x == 0)
    for (int64_t p = 0; p < n_pieces && p < kStages; ++p) issue_piece(ring, chunk, n_recs, p);
  pdl_wait();   // records (immutable) are already in flight; gradient rows come from the previous step

  LaneVec<V> acc;
  acc.zero();
  bool open_active = false;  // the statement being summed has had an active operand
  for (int64_t p = 0; p < n_pieces; ++p) {
    const int st = static_cast<int>(p % kStages);
    mbar_wait(&ring.full[st], static_cast<uint32_t>((p / kStages) & 1));
    const int64_t left = n_recs - p * kPieceRecs;
    const int n = static_cast<int>(left < kPieceRecs ? left : kPieceRecs);
    const Rec* r = &ring.buf[st][0];
    if (SPARSE) {
      for (int t = threadIdx.x; t < n * nwarps; t += blockDim.x) {
        const int i = t / nwarps, w = t - i * nwarps;
        const int32_t slot = r[i].slot;
        uint8_t f = 0;
        if (cb0 + w < A.n_colblocks && slot >= 0) f = A.rowact[static_cast<int64_t>(slot) * ncb + cb0 + w];
        pf[i][w] = f;
      }
      __syncthreads();
    }
    if (live) {
      bool work = true;
      if (SPARSE && !open_active) {  // anything to add or to clear in my column block?
        bool mine = false;
        for (int i = lane; i < n; i += 32) mine |= (pf[i][warp] != 0);
        work = __any_sync(0xffffffffu, mine);
      }
      if (work) {
        for (int i = 0; i < n; i += kBatch) {
          LaneVec<V> v[kBatch];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            v[u].zero();
            if (i + u < n && r[i + u].kind == REC_OP && (!SPARSE || pf[i + u][warp])) {
              const int32_t slot = r[i + u].slot;
              if (slot >= 0) v[u].load(gcol + static_cast<int64_t>(slot) * A.K);
            }
          }
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            if (i + u >= n) break;
            const Rec rc = r[i + u];
            if (rc.kind == REC_OP) {
              if (rc.slot >= 0 && (!SPARSE || pf[i + u][warp])) {
                acc.x = mul_add_rn(acc.x, rc.m, v[u].x);
                if (V == 2) {
                  double* cv = reinterpret_cast<double*>(&acc);
                  cv[V - 1] = mul_add_rn(cv[V - 1], rc.m, reinterpret_cast<const double*>(&v[u])[V - 1]);
                }
                open_active = true;
              }
            } else {  // TAIL: the statement is complete
              if (!SPARSE || open_active) {
                acc.store(gcol + static_cast<int64_t>(rc.slot) * A.K);
                if (SPARSE && lane == 0) ract[static_cast<int64_t>(rc.slot) * ncb] = 1;
              } else if (pf[i + u][warp]) {  // result is +0.0 but the row still holds an older value
                acc.store(gcol + static_cast<int64_t>(rc.slot) * A.K);   // acc is +0.0 here
                if (lane == 0) ract[static_cast<int64_t>(rc.slot) * ncb] = 0;
              }
              acc.zero();
              open_active = false;
            }
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && p + kStages < n_pieces) issue_piece(ring, chunk, n_recs, p + kStages);
  }
}

// ---------------------------------------------------------------------------------------
// Level-scheduled forward sweep, balanced form (round 2).  Same stream, same arithmetic, same bits as
// forward_level_kernel above; different CTA shape.  ncu on the kernel above (BASELINE config 4, wide step): 27 warp
// instructions per (record, column block) -- every warp re-reads and re-decodes every record -- and issue slots 57 %
// busy on a kernel that should only wait for HBM.  Here, as in reverse_fused_kernel:
//   * a CTA owns one chunk (one staged piece of <= 256 records) and up to 32 column blocks (A.cbs); the activity
//     bytes of all its (record, column block) pairs are fetched in ONE round trip, and the warps take the column
//     blocks that have anything to do in turn;
//   * per block of 32 records the lanes decode ONE record each (row byte offset, flags, multiplier) and park the
//     statement ends and the ACTIVE operations, compacted and in order, in the warp's shared-memory list; the
//     serial walk then costs a 16-byte shared-memory read and a 64-bit add per record, eight 16-byte row loads in
//     flight per lane.
// Chunks longer than one piece (a statement of more than ~128 operands) are left to the kernel above: the host
// picks the kernel per step (Shard::level_long).
// ---------------------------------------------------------------------------------------
constexpr int kFwdCbs = 32;

template <int V, bool SPARSE>
__device__ __forceinline__ void forward_task(const Rec* __restrict__ r, int n, const uint8_t* __restrict__ pfw,
                                             const uint32_t* __restrict__ tailm, uint4* __restrict__ wl,
                                             double* __restrict__ gcol, uint8_t* __restrict__ ract, int64_t K, int ncb,
                                             int lane, int64_t n_rows) {
  // list entry: x = low word of the row's byte offset | 1 (statement end) | 2 (load the row) | 4 (statement end whose
  // row currently holds data), y = high word, (z, w) = multiplier bits (operation) or z = row (statement end)
  LaneVec<V> acc;
  acc.zero();
  bool open_active = false;   // the statement being summed has had an active operand
  char* const gbase = reinterpret_cast<char*>(gcol);
  const uint32_t Kb = static_cast<uint32_t>(K) * 8u;   // bytes per row: a multiple of 512, the low bits are free
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int nblk = (n + 31) >> 5;
  for (int blk = 0; blk < nblk; ++blk) {
    const int i = (blk << 5) + lane;
    const bool valid = i < n;
    const uint32_t tm = tailm[blk];
    const bool is_tail = (tm >> lane) & 1u;
    const uint8_t pfb = (SPARSE && valid) ? pfw[i] : 0;
    const bool a = valid && !is_tail && (!SPARSE || pfb != 0);
    const uint32_t act = __ballot_sync(0xffffffffu, a);
    const uint32_t work = act | tm;
    if (work == 0u) continue;
    if ((work >> lane) & 1u) {
      const Rec rc = r[i];
      const uint64_t off = static_cast<uint64_t>(static_cast<uint32_t>(rc.slot)) * Kb;
      uint4 e;
      e.x = static_cast<uint32_t>(off) | (is_tail ? 1u : 0u) | (a ? 2u : 0u) | ((is_tail && pfb) ? 4u : 0u);
      e.y = static_cast<uint32_t>(off >> 32);
      const uint64_t mb = static_cast<uint64_t>(__double_as_longlong(rc.m));
      e.z = is_tail ? static_cast<uint32_t>(rc.slot) : static_cast<uint32_t>(mb);
      e.w = static_cast<uint32_t>(mb >> 32);
      ADEPT_DCHECK(__popc(work & lt_mask) < 32 && rc.slot >= 0 && rc.slot < n_rows);
      wl[__popc(work & lt_mask)] = e;
    }
    __syncwarp();
    const int cnt = __popc(work);
    constexpr int NB = (V == 4) ? kBatch / 2 : kBatch;   // row loads in flight per lane: the same bytes for V = 2 and 4
    for (int j = 0; j < cnt; j += NB) {
      LaneVec<V> v[NB];
      // all loads of a step are independent of its stores (RAW/WAR between the statements of a step are excluded
      // by the schedule; a statement's own operands precede its end)
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        v[u].zero();
        if (j + u < cnt) {
          const uint2 e = *reinterpret_cast<const uint2*>(&wl[j + u]);
          if (e.x & 2u)
            v[u].load(reinterpret_cast<const double*>(gbase + ((static_cast<uint64_t>(e.y) << 32) | (e.x & ~15u))));
        }
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        if (j + u >= cnt) break;
        const uint4 e = wl[j + u];
        if (e.x & 1u) {   // the statement is complete
          double* dst = reinterpret_cast<double*>(gbase + ((static_cast<uint64_t>(e.y) << 32) | (e.x & ~15u)));
          if (!SPARSE || open_active) {
            acc.store(dst);
            if (SPARSE && lane == 0) ract[static_cast<int64_t>(e.z) * ncb] = 1;
          } else if (e.x & 4u) {   // the result is +0.0 but the row still holds an older value
            acc.store(dst);        // acc is +0.0 here
            if (lane == 0) ract[static_cast<int64_t>(e.z) * ncb] = 0;
          }
          acc.zero();
          open_active = false;
        } else {
          const double m = __longlong_as_double(static_cast<long long>((static_cast<uint64_t>(e.w) << 32) | e.z));
          double* cv = reinterpret_cast<double*>(&acc);
          const double* av = reinterpret_cast<const double*>(&v[u]);
#pragma unroll
          for (int k = 0; k < V; ++k) cv[k] = mul_add_rn(cv[k], m, av[k]);
          open_active = true;
        }
      }
    }
    __syncwarp();   // the list is rewritten for the next block
  }
}

// (V = 4 needs ~80 registers: three CTAs per SM instead of four, each with twice the bytes in flight per lane)
template <int V, bool SPARSE>
__global__ void __launch_bounds__(256, (V == 4 ? 3 : 4)) forward_balanced_kernel(ForwardArgs A) {
  __shared__ __align__(16) Rec recs[kPieceRecs];
  __shared__ __align__(8) uint64_t full;
  __shared__ __align__(4) uint8_t pf[kFwdCbs * kPfPitch];   // pf[w*kPfPitch + i]: record i, column block cb0 + w
  __shared__ uint32_t tailm[kPieceRecs / 32];
  __shared__ uint4 wlist[8][32];
  __shared__ uint32_t smask;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int ncb = A.flag_pitch;
  const int64_t c = A.first_chunk + blockIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(&full, 1);
    mbar_fence_init();
    smask = 0u;
  }
  __syncthreads();
  pdl_trigger();
  const int64_t r0 = A.chunk_begin[c], r1 = A.chunk_begin[c + 1];
  const int n = static_cast<int>(r1 - r0);   // <= kPieceRecs: the host sends longer chunks to forward_level_kernel
  const int cb0 = A.cb_lo + blockIdx.y * A.cbs;
  const int cbs = min(A.cbs, A.n_colblocks - cb0);
  ADEPT_DCHECK(n >= 1 && n <= kPieceRecs && cbs >= 1 && cbs <= kFwdCbs && (cb0 + cbs) * 32 * V <= A.K);
  if (threadIdx.x == 0) {
    const uint32_t bytes = static_cast<uint32_t>(n) * static_cast<uint32_t>(sizeof(Rec));
    mbar_expect_tx(&full, bytes);
    tma_bulk_g2s(&recs[0], A.recs + r0, bytes, &full);
  }
  pdl_wait();   // records (immutable) are already in flight; gradient rows and activity bytes come from the previous step
  mbar_wait(&full, 0u);
  const Rec* r = &recs[0];
  // statement-end mask of every block of 32 records
  for (int i = threadIdx.x; i < ((n + 31) & ~31); i += blockDim.x) {
    const int32_t kind = i < n ? r[i].kind : REC_OP;
    const uint32_t h = __ballot_sync(0xffffffffu, kind != REC_OP);
    if (lane == 0) tailm[i >> 5] = h;
  }
  uint32_t mask;
  if (SPARSE) {
    // activity bytes: thread t -> record t / cbs, column block cb0 + t % cbs (consecutive bytes of a row)
    uint32_t seen = 0u;   // column blocks in which this thread saw an active operand or a live result row
    for (int t = threadIdx.x; t < n * cbs; t += blockDim.x) {
      const int i = t / cbs, w = t - i * cbs;
      const uint8_t f = A.rowact[static_cast<int64_t>(r[i].slot) * ncb + cb0 + w];
      pf[w * kPfPitch + i] = f;
      if (f) seen |= 1u << w;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) seen |= __shfl_xor_sync(0xffffffffu, seen, d);
    if (lane == 0 && seen) atomicOr(&smask, seen);
    __syncthreads();
    mask = smask;
  } else {
    __syncthreads();
    mask = cbs >= 32 ? 0xffffffffu : ((1u << cbs) - 1u);
  }
  int k = 0;
  for (uint32_t mm = mask; mm; mm &= mm - 1u, ++k) {
    if (k % nwarps != warp) continue;
    const int w = __ffs(mm) - 1;
    const int cb = cb0 + w;
    forward_task<V, SPARSE>(r, n, &pf[w * kPfPitch], tailm, &wlist[warp][0],
                            A.g + (static_cast<int64_t>(cb) * 32 + lane) * V, A.rowact + cb, A.K, ncb, lane, A.n_rows);
  }
}

// ---------------------------------------------------------------------------------------
// Seeding and extraction (adept/jacobian.cpp:163-190 forward, :372-379 and :434-449 reverse).
// ---------------------------------------------------------------------------------------
// g[seed_slot[d0+j]*K + j] = 1.0 for the n directions of this pass.
__global__ void seed_kernel(double* g, int64_t K, const int32_t* seed_slot, int64_t d0, int n, uint8_t* rowact,
                            int ncb, int block_dirs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const int64_t slot = seed_slot[d0 + j];
    g[slot * K + j] = 1.0;
    if (rowact) rowact[slot * ncb + j / block_dirs] = 1;  // several lanes may store the same 1
  }
}

// out[i*stride_i + (d0+j)*stride_j] = g[out_slot[i]*K + j], i < n_out, j < n_dirs.
// 32x32 tiles through shared memory so that both the row reads and the writes coalesce
// whichever of stride_i / stride_j is the unit one.
// rowact != nullptr (fused reverse sweep): a block whose activity byte is 0 is all +0.0 WHATEVER memory
// holds -- a ZMARK only clears the byte of the row it opens, and when no active contribution lands on that
// row nothing is stored, so the memory still shows the version before last.  The sweep's own reads are all
// guarded by the byte; so is this one.
__global__ void extract_kernel(const double* __restrict__ g, int64_t K, const int32_t* __restrict__ out_slot,
                               int64_t n_out, int n_dirs, int64_t d0, double* __restrict__ out,
                               int64_t stride_i, int64_t stride_j, const uint8_t* __restrict__ rowact, int ncb,
                               int block_dirs) {
  __shared__ double tile[32][33];
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int j0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r;
    const int j = j0 + threadIdx.x;
    if (i < n_out && j < n_dirs) {
      const int64_t row = out_slot[i];
      double v = 0.0;
      if (!rowact || rowact[row * ncb + j / block_dirs]) v = g[row * K + j];
      tile[r][threadIdx.x] = v;
    }
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

// Ensemble seeding/extraction.  Lane L = b*Dp + d (b member, d direction, Dp = directions padded to the
// reference block).  Output per member is the reference's default column-major n_dep x n_indep layout.
__global__ void batch_seed_kernel(double* g, int64_t K, const int32_t* seed_slot, int n_dirs, int Dp, int64_t n_members) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_members * n_dirs) return;
  const int64_t b = t / n_dirs;
  const int d = static_cast<int>(t - b * n_dirs);
  g[static_cast<int64_t>(seed_slot[d]) * K + b * Dp + d] = 1.0;
}
// mode 0 (forward): directions are independents j, outputs are dependents i; mode 1: the other way round.
__global__ void batch_extract_kernel(const double* __restrict__ g, int64_t K, const int32_t* __restrict__ out_slot,
                                     int n_out, int n_dirs, int Dp, int64_t n_members, int mode, int64_t n_dep,
                                     int64_t n_indep, double* __restrict__ out) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per = static_cast<int64_t>(n_out) * n_dirs;
  if (t >= n_members * per) return;
  // consecutive threads: direction fastest, then member (coalesced reads of a gradient row), then out index
  const int64_t o = t / (n_members * n_dirs);
  const int64_t rem = t - o * n_members * n_dirs;
  const int64_t b = rem / n_dirs;
  const int d = static_cast<int>(rem - b * n_dirs);
  const double v = g[static_cast<int64_t>(out_slot[o]) * K + b * Dp + d];
  const int64_t i = mode == 0 ? o : d;  // dependent index
  const int64_t j = mode == 0 ? d : o;  // independent index
  out[b * n_dep * n_indep + i + j * n_dep] = v;
}

}  // namespace adept_b200
