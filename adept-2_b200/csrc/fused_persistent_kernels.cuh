// fused_persistent_kernels.cuh -- the fused level-scheduled REVERSE sweep as ONE resident kernel, for tapes whose
// steps are too narrow to be worth a launch each (BASELINE config 2, Toon nx=4096 nt=1000: 3126 dependent steps of
// ~225 chunks; one launch per step costs ~19 us of launch, drain and refill however few directions a GPU holds).
//
// What is replayed is exactly the fused stream of replay_kernels.cuh (MARK / ZMARK / OP over versioned rows,
// reverse_fused_kernel), item by item with the same device functions, so the arithmetic and every accumulation order are
// those of the per-step kernel -- and of the reference (adept/jacobian.cpp:382-431).  What changes is who waits for whom:
//
//   * An ITEM is (chunk, column group, part) of a step -- part: which of the chunk's T.split ranges of sub-tasks
//     (replay_kernels.cuh, fused_single_piece).  The G resident CTAs (cooperative launch: co-residency is guaranteed) take
//     the items of a step round robin, from a start that moves on after every step that fills the grid, so every CTA can
//     enumerate its own items of the whole sweep without any communication (PCursor).
//   * There is no grid-wide barrier.  Step l has parts(l) = min(G, items(l)) participating CTAs, which every CTA can
//     compute; a CTA counts itself in at cnt[l] after its last item of step l, and before its first item of step l it
//     waits until the step processed before l (the next higher non-empty one) has all its participants counted in.  By
//     induction that step's participants had waited for the one before, and so on: everything above l is complete.
//     CTAs without items in a step neither arrive nor wait, and a CTA that was the only participant of the previous
//     step goes on without touching global memory at all (the scalar chains between two wide steps).
//   * The records are immutable, so the TMA copy of a CTA's NEXT item is issued while the current item is processed
//     (once its activity bytes are in) -- also across a step boundary, i.e. before the wait.
//   CPU models: tests/test_persistent_cursor_model.py (items, counters, waits under random interleavings) and
//   tests/test_persistent_ring_model.py (ring slots and mbarrier phases).
//
// Memory model: a CTA's stores of a step are ordered before its arrival by bar.sync + fence.acq_rel.gpu (release); a
// waiter polls the counter with relaxed loads, then fence.acq_rel.gpu (acquire) + bar.sync order the loads of all its
// threads after what the counter stands for (the pattern of a cooperative-groups grid sync, per step and per
// participant instead of per grid).  A CTA that polls longer than spin_limit raises *abort and every CTA leaves at its
// next wait: the host reports an error instead of hanging the device.
//
// Measured (B200, BASELINE config 2; DESIGN.md section 5, "Narrow steps"): bit-identical, and SLOWER than one launch per
// step on four PDL-chained streams (512 directions: 48 vs 41 ms; 4096: 126-145 vs 83 ms) -- an item costs ~2.5 us even
// when its chunk lies outside the band of active targets.  Opt-in ("fused_persistent"); kept as the synchronisation
// skeleton of a step that first lists the active (sub-task, column block) pairs device-wide and then hands them out.
#pragma once
#include "replay_kernels.cuh"

namespace adept_b200 {

struct FusedPersistArgs {
  TargetArgs T;                 // stream, chunk table, workspace; cb_lo = 0, n_colblocks = all of the pass; cbs per item
  const int64_t* step_fchunk;   // [n_levels + 2] first fused chunk of every step (device copy)
  int32_t n_levels;
  int gy;                       // column groups of T.cbs column blocks
  unsigned long long* trace;    // profiling aid (-DADEPT_FUSED_TRACE): 4 clock stamps per item of CTA 0, first 8192 items
  unsigned int* cnt;            // [n_levels + 2] participants counted in per step (zeroed before the launch)
  unsigned int* abort;          // raised by a CTA that gave up waiting
  unsigned int spin_limit;
};

// A CTA's position in its own item sequence (thread 0 only, kept in shared memory)
struct PCursor {
  int64_t c0, n_items, it;   // step l: first chunk, items, my item index
  int32_t l;                 // step of the item; 0: no more items
  int32_t l_prev;            // the non-empty step the grid processes just before l (0: none)
  uint32_t parts_prev;       // its participants
  uint32_t rot;              // rotation of step l: CTA b starts at item (b + G - rot) % G
};

// What every thread must know when an item is finished
struct PItem {
  int64_t r0;        // the next item's records: [r0, r0 + n_recs) of the fused stream
  int64_t n_recs;
  int32_t cby;       // ... and column group
  int32_t part;      // ... and which of the chunk's T.split ranges of sub-tasks
  int32_t l;         // ... and step; 0: the CTA is done
  int32_t arrive;    // != 0: that was the CTA's last item of step `arrive`: count the CTA in
  int32_t wait_l;    // != 0: before the next item wait until cnt[wait_l] >= wait_n
  uint32_t wait_n;
};

// the item after the cursor's (or the first one, from the pseudo step n_levels + 1 with no items)
__device__ __forceinline__ void pcursor_advance(PCursor& cu, const FusedPersistArgs& P) {
  const uint32_t G = gridDim.x;
  cu.it += G;
  if (cu.it < cu.n_items) return;
  for (;;) {
    if (cu.n_items > 0) {   // leaving a non-empty step
      cu.l_prev = cu.l;
      cu.parts_prev = static_cast<uint32_t>(cu.n_items < static_cast<int64_t>(G) ? cu.n_items : G);
      // a step that fills the grid hands the start on to where it ended (balance); a narrower one does not, so that
      // a chain of narrow steps stays with the same CTAs (which then need not wait for anybody else)
      if (cu.n_items >= static_cast<int64_t>(G)) cu.rot = static_cast<uint32_t>((cu.rot + cu.n_items) % G);
    }
    int32_t l = cu.l - 1;
    int64_t c0 = 0, c1 = 0;
    while (l >= 1) {
      c0 = P.step_fchunk[l];
      c1 = P.step_fchunk[l + 1];
      if (c1 > c0) break;
      --l;
    }
    if (l < 1) {
      cu.l = 0;
      cu.n_items = 0;
      return;
    }
    cu.l = l;
    cu.c0 = c0;
    cu.n_items = (c1 - c0) * P.gy * (P.T.split > 1 ? P.T.split : 1);
    cu.it = (blockIdx.x + G - cu.rot) % G;
    if (cu.it < cu.n_items) return;
  }
}

__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_add_relaxed_gpu(unsigned int* p, unsigned int v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// the transition after an item of step l (l = 0: before the first item) to the cursor's item
__device__ __forceinline__ PItem pitem_from(const PCursor& cu, int32_t l, const FusedPersistArgs& P) {
  PItem t;
  t.l = cu.l;
  t.cby = 0;
  t.part = 0;
  t.r0 = 0;
  t.n_recs = 0;
  if (cu.l) {   // item index -> (chunk, column group, part), parts of a chunk adjacent
    const int split = P.T.split > 1 ? P.T.split : 1;
    const int64_t per_chunk = static_cast<int64_t>(P.gy) * split;
    const int64_t c = cu.c0 + cu.it / per_chunk;
    const int rem = static_cast<int>(cu.it % per_chunk);
    t.cby = rem / split;
    t.part = rem % split;
    t.r0 = P.T.chunk_begin[c];
    t.n_recs = P.T.chunk_begin[c + 1] - t.r0;
  }
  t.arrive = (cu.l != l) ? l : 0;
  const bool alone = (cu.l_prev == l && cu.parts_prev == 1u);   // I was the only participant of the step above
  t.wait_l = (cu.l != 0 && cu.l != l && cu.parts_prev != 0u && !alone) ? cu.l_prev : 0;
  t.wait_n = cu.parts_prev;
  return t;
}

#ifdef ADEPT_FUSED_TRACE
#define PERSIST_STAMP(q) do { if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && k < 8192u) P.trace[k * 4u + (q)] = gtimer(); } while (0)
#else
#define PERSIST_STAMP(q) do { } while (0)
#endif

template <int V, int LP>
__global__ void __launch_bounds__(256, 4) reverse_fused_persistent_kernel(FusedPersistArgs P) {
  __shared__ StageRing ring;
  __shared__ __align__(4) uint8_t pf[kFusedCbs * kPfPitch];
  __shared__ uint32_t headm[kPieceRecs / 32];
  __shared__ uint4 wlist[8][32];
  __shared__ uint32_t smask[2];
  __shared__ PCursor s_cu;
  __shared__ PItem s_tr[2];
  __shared__ int s_ok;
  __shared__ uint32_t s_pre;   // sequence number of the piece already in flight for the next item (thread 0)
  const TargetArgs& A = P.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int ncb = A.flag_pitch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
    mbar_fence_init();
    smask[0] = smask[1] = 0u;
    s_ok = 1;
    PCursor cu;
    cu.c0 = 0;
    cu.n_items = 0;
    cu.it = 0;
    cu.l = P.n_levels + 1;
    cu.l_prev = 0;
    cu.parts_prev = 0u;
    cu.rot = 0u;
    pcursor_advance(cu, P);
    s_cu = cu;
    const PItem t = pitem_from(cu, 0, P);
    s_tr[1] = t;
    s_pre = 0xffffffffu;
    // the first piece of the first item -- unless that item takes no piece at all (a long chunk belongs to the first of
    // its parts; the others pass it by: copying for them would leave a transaction nobody waits for in ring slot 0)
    if (cu.l && !(t.n_recs > kPieceRecs && t.part != 0)) {
      issue_piece(ring, A.recs + t.r0, t.n_recs, 0, 0);
      s_pre = 0u;
    }
  }
  __syncthreads();
  uint32_t pc = 0;   // pieces this CTA has consumed: sequence number of the next piece
  uint32_t k = 0;    // items this CTA has processed
  PItem tr = s_tr[1];
  bool ok = true;
  for (;;) {
    // ---- between two items: count the CTA in, wait for the step above the next item ----
    if (tr.arrive && threadIdx.x == 0) {
      fence_acq_rel_gpu();   // release: this CTA's rows and activity bytes (ordered before by the bar.sync that ended the item)
      red_add_relaxed_gpu(&P.cnt[tr.arrive], 1u);
    }
    if (tr.l == 0) break;
    PERSIST_STAMP(0);
    if (tr.wait_l) {
      if (threadIdx.x == 0) {
        const unsigned int* ctr = &P.cnt[tr.wait_l];
        unsigned int spins = 0;
        int good = 1;
        while (ld_relaxed_gpu(ctr) < tr.wait_n) {
          if ((++spins & 63u) == 0u && (spins > P.spin_limit || ld_relaxed_gpu(P.abort) != 0u)) {
            atomicExch(P.abort, 1u);
            good = 0;
            break;
          }
        }
        fence_acq_rel_gpu();   // acquire: what the participants of the steps above wrote
        s_ok = good;
      }
      __syncthreads();
      if (!s_ok) {
        ok = false;
        break;
      }
    }
    // ---- the item ----
    PERSIST_STAMP(1);
    const int cby = tr.cby;
    const int32_t l = tr.l;
    const int64_t n_recs = tr.n_recs;
    const Rec* chunk = A.recs + tr.r0;
    const int64_t n_pieces = (n_recs + kPieceRecs - 1) / kPieceRecs;
    const int cb0 = cby * A.cbs;
    const int cbs = min(A.cbs, A.n_colblocks - cb0);
    const int n_sets = (cbs + nwarps - 1) / nwarps;
    // pieces this item takes through the ring (a long chunk is not split: it belongs to the first of its items)
    const int64_t total = (n_pieces == 1) ? 1 : (tr.part == 0 ? n_pieces * n_sets : 0);
    ADEPT_DCHECK(cbs >= 1 && n_recs >= 1);
    // thread 0: my own pieces (the first kStages of them, less the one that is already in flight) ...
    if (threadIdx.x == 0)
      for (int64_t it = 0; it < total && it < kStages; ++it)
        if (!(it == 0 && s_pre == pc)) issue_piece(ring, chunk, n_recs, it % n_pieces, pc + it);
    // ... and, off the critical path (single piece: after the activity bytes are in), the item after this one: what
    // happens between the two, and its first piece -- the records are immutable and its ring slot was last used
    // kStages pieces ago
    auto look_ahead = [&]() {
      PCursor cu = s_cu;
      pcursor_advance(cu, P);
      s_cu = cu;
      const PItem t = pitem_from(cu, l, P);
      s_tr[k & 1u] = t;
      if (t.l != 0 && total < kStages && !(t.n_recs > kPieceRecs && t.part != 0)) {
        issue_piece(ring, A.recs + t.r0, t.n_recs, 0, pc + total);
        s_pre = pc + static_cast<uint32_t>(total);
      }
      smask[(k + 1u) & 1u] = 0u;   // the next item's mask (its last reader finished an item ago)
    };
    if (n_pieces == 1) {
      // ---- the common case: one staged piece; warps take the ACTIVE column blocks in turn (reverse_fused_kernel) ----
      const int n = static_cast<int>(n_recs);
      const int st = static_cast<int>(pc % kStages);
      mbar_wait(&ring.full[st], (pc / kStages) & 1u);
      const Rec* r = &ring.buf[st][0];
      fused_single_piece<V, LP>(A, r, n, cb0, cbs, tr.part, A.split > 1 ? A.split : 1, pf, headm, wlist, &smask[k & 1u], [&]() {
        PERSIST_STAMP(2);
        if (threadIdx.x == 0) look_ahead();
      });
    } else if (tr.part != 0) {
      if (threadIdx.x == 0) look_ahead();   // a long chunk is not split: it belongs to the first of its items
    } else {
      // ---- a target group longer than one piece: streamed once per set of nwarps column blocks (reverse_fused_kernel) ----
      if (threadIdx.x == 0) look_ahead();
      FusedState<V> S;
      S.reset();
      for (int64_t it = 0; it < total; ++it) {
        const int64_t p = it % n_pieces;
        const int set = static_cast<int>(it / n_pieces);
        const int w = set * nwarps + warp;
        const bool live = w < cbs;
        const FusedLane L = fused_lane<V>(A, cb0 + (live ? w : 0), lane);
        if (p == 0) S.reset();
        const uint32_t seq = pc + static_cast<uint32_t>(it);
        const int st = static_cast<int>(seq % kStages);
        mbar_wait(&ring.full[st], (seq / kStages) & 1u);
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
        if (threadIdx.x == 0 && it + kStages < total)
          issue_piece(ring, chunk, n_recs, (it + kStages) % n_pieces, pc + it + kStages);
      }
    }
    __syncthreads();   // the item is finished: stores issued, shared tables free, s_tr[k & 1] written
    PERSIST_STAMP(3);
    pc += static_cast<uint32_t>(total);
    tr = s_tr[k & 1u];
    ++k;
  }
  if (!ok && threadIdx.x == 0 && s_pre == pc) mbar_wait(&ring.full[pc % kStages], (pc / kStages) & 1u);   // no copy may outlive the CTA
}

}  // namespace adept_b200
