// tape_kernels.cuh -- device-side tape preparation: validation, linearisation into record
// streams, dependency analysis (levels) and level-major re-emission.
//
// Input is the reference's tape exactly as recorded (include/adept/Statement.h:24-30,
// include/adept/StackStorageOrig.h:155-163): stmt[s] = {lhs slot, end_plus_one} with the
// {-1,0} sentinel at s = 0; statement s >= 1 owns operations [stmt[s-1].end, stmt[s].end).
//
// Notation: S = n_stmt-1 real statements, O = n_ops, R = O + S records.
//   forward tape-order position of op o of statement s : o + (s-1)
//   forward tape-order position of the TAIL of s       : end[s] + (s-1)
//   reverse tape-order base of s (its HEAD)            : (S-s) + (O-end[s]); op o follows at
//                                                        base + 1 + (o - end[s-1])
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace adept_b200 {
namespace cg = cooperative_groups;

// statement (>= 1) that owns operation o: smallest s with end[s] > o
__device__ __forceinline__ int64_t stmt_of_op(const int2* __restrict__ stmt, int64_t n_stmt, int64_t o) {
  int64_t lo = 1, hi = n_stmt - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (static_cast<int64_t>(stmt[mid].y) > o) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

// flags[0] |= 1 on any malformed entry (slot out of range, ends not monotone), |= 2 on a non-finite multiplier.
// counts[0..2] += number of multipliers equal to 0, +1, -1 (Stack::fraction_multipliers_equal_to,
// include/adept/Stack.h:707-715); counts may be null.
__global__ void validate_kernel(const int2* __restrict__ stmt, int64_t n_stmt, const int32_t* __restrict__ idx,
                                const double* __restrict__ mult, int64_t n_ops, int64_t max_gradient, int* flags,
                                unsigned long long* counts = nullptr) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  bool bad = false, nonfinite = false;
  unsigned int n0 = 0, n1 = 0, nm1 = 0;
  for (int64_t o = t; o < n_ops; o += stride) {
    const int32_t x = idx[o];
    if (x < 0 || x >= max_gradient) bad = true;
    if (mult) {
      const double m = mult[o];
      if (!isfinite(m)) nonfinite = true;
      n0 += (m == 0.0);
      n1 += (m == 1.0);
      nm1 += (m == -1.0);
    }
  }
  if (counts) {
    n0 = __reduce_add_sync(0xffffffffu, n0);
    n1 = __reduce_add_sync(0xffffffffu, n1);
    nm1 = __reduce_add_sync(0xffffffffu, nm1);
    if ((threadIdx.x & 31) == 0) {
      if (n0) atomicAdd(&counts[0], static_cast<unsigned long long>(n0));
      if (n1) atomicAdd(&counts[1], static_cast<unsigned long long>(n1));
      if (nm1) atomicAdd(&counts[2], static_cast<unsigned long long>(nm1));
    }
  }
  for (int64_t s = t; s < n_stmt; s += stride) {
    const int2 st = stmt[s];
    if (s == 0) {
      if (st.y != 0) bad = true;
    } else {
      if (st.x < 0 || st.x >= max_gradient) bad = true;
      if (st.y < stmt[s - 1].y || st.y > n_ops) bad = true;
      if (s == n_stmt - 1 && st.y != n_ops) bad = true;
    }
  }
  if (bad) atomicOr(flags, 1);
  if (nonfinite) atomicOr(flags, 2);  // 0*inf must stay NaN: forward sweeps then evaluate every statement
}

// ---------------------------------------------------------------------------------------
// Tape compaction on upload (option "remove_null"): operations whose multiplier is zero are dropped, exactly what
// the reference does AT RECORDING TIME when built with -DADEPT_REMOVE_NULL_STATEMENTS
// (include/adept/StackStorageOrig.h:46-65: push_rhs stores the pair only `if (multiplier != 0.0)`).  Statements keep
// their place (a statement left without operations still overwrites its LHS with zero).  keep[o] = 1 for the
// operations that stay; pos = exclusive scan of keep.
// ---------------------------------------------------------------------------------------
__global__ void compact_keep_kernel(const double* __restrict__ mult, int64_t n_ops, int32_t* __restrict__ keep) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t o = t; o < n_ops; o += stride) keep[o] = (mult[o] != 0.0) ? 1 : 0;   // NaN != 0: kept, like the reference
}
__global__ void compact_ops_kernel(const double* __restrict__ mult, const int32_t* __restrict__ idx, int64_t n_ops,
                                   const int32_t* __restrict__ keep, const int32_t* __restrict__ pos,
                                   double* __restrict__ mult_out, int32_t* __restrict__ idx_out) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t o = t; o < n_ops; o += stride)
    if (keep[o]) {
      mult_out[pos[o]] = mult[o];
      idx_out[pos[o]] = idx[o];
    }
}
// new end of every statement = number of kept operations before its old end (pos has n_ops + 1 entries)
__global__ void compact_stmt_kernel(int2* __restrict__ stmt, int64_t n_stmt, const int32_t* __restrict__ pos) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t s = t; s < n_stmt; s += stride) stmt[s].y = pos[stmt[s].y];
}

// Tape-order streams.  One thread per record of the forward stream.
__global__ void build_streams_kernel(const int2* __restrict__ stmt, int64_t n_stmt,
                                     const double* __restrict__ mult, const int32_t* __restrict__ idx,
                                     int64_t n_ops, Rec* __restrict__ fwd, Rec* __restrict__ rev,
                                     int32_t* __restrict__ rec_stmt) {
  const int64_t S = n_stmt - 1;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t o = t; o < n_ops; o += stride) {
    const int64_t s = stmt_of_op(stmt, n_stmt, o);
    Rec r;
    r.m = mult ? mult[o] : __longlong_as_double(o);  // ensemble mode: the record carries the operation index
    r.slot = idx[o];
    r.kind = REC_OP;
    const int64_t fpos = o + (s - 1);
    fwd[fpos] = r;
    if (rec_stmt) rec_stmt[fpos] = static_cast<int32_t>(s);
    const int64_t base = (S - s) + (n_ops - stmt[s].y);
    rev[base + 1 + (o - stmt[s - 1].y)] = r;
  }
  for (int64_t s = 1 + t; s < n_stmt; s += stride) {
    Rec r;
    r.m = 0.0;
    r.slot = stmt[s].x;
    r.kind = REC_MARK;
    const int64_t fpos = static_cast<int64_t>(stmt[s].y) + (s - 1);
    fwd[fpos] = r;
    if (rec_stmt) rec_stmt[fpos] = static_cast<int32_t>(s);
    rev[(S - s) + (n_ops - stmt[s].y)] = r;
  }
}

// ---------------------------------------------------------------------------------------
// Dependency analysis (no counterpart in the reference, which replays strictly in tape order).
//
// A statement may leave its tape position as long as no bit of any sweep changes.  That holds
// under these constraints on the step number of each statement (oracle/replay_oracle.c,
// oracle_levels mode 2, is the sequential statement of the same rule):
//   RAW  a reader of a slot version runs after its writer;
//   WAR  the next writer of the slot runs after every reader of the old version;
//   WAW  writers of a slot keep their order;
//   MONO readers of one slot version have non-decreasing steps in tape order.
// Forward sweeps are gathers (each statement adds its own operands in tape order), so RAW/WAR/
// WAW suffice.  Reverse sweeps walk the steps downwards and, inside a step, add the
// contributions to each slot in (statement descending, operand ascending) order (the "target
// stream" below); MONO makes that the reference's accumulation order across steps too.
//
// The tape is cut into windows of `window` statements; ONE WARP analyses one window
// sequentially (lanes share the operands of a statement), keeping per-slot (writer step,
// max reader step) in an open-addressing hash table in global memory.  Steps are local to the
// window; the global step is (window, local step) in lexicographic order.
// ---------------------------------------------------------------------------------------
struct __align__(16) SlotEntry {
  int32_t key;  // slot, -1 = empty   (tables are memset to 0xFF: key = wl = rl = -1)
  int32_t pad;
  int32_t wl;   // step of the last writer   (<= 0: none in this window)          } 8-byte aligned pair,
  int32_t rl;   // max step of the readers of the current version (<= 0: none)    } written together
};

// Window boundaries.  A window starts every `window` statements, and a statement with more than
// `huge` operands (e.g. the 4(N-1)-operand `sum` of a cost function, include/adept/reduce.h:162-174) is a
// window of its own: windows are totally ordered, so such a statement needs no table look-ups at all
// (its local step is 1) and never stalls the warp that walks its neighbours.
__global__ void window_flags_kernel(const int2* __restrict__ stmt, int64_t n_stmt, int64_t window, int64_t huge,
                                    int32_t* __restrict__ flag) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t s = t; s < n_stmt; s += stride) {
    int f = 0;
    if (s >= 1) {
      const bool big = static_cast<int64_t>(stmt[s].y) - stmt[s - 1].y > huge;
      const bool prev_big = s >= 2 && static_cast<int64_t>(stmt[s - 1].y) - stmt[s - 2].y > huge;
      f = ((s - 1) % window == 0) || big || prev_big;
    }
    flag[s] = f;
  }
}
// win_lo[w] = first statement of window w (wid = inclusive scan of flag, minus 1); win_lo[n_windows] = n_stmt
__global__ void window_starts_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ incl, int64_t n_stmt,
                                     int64_t n_windows, int64_t* __restrict__ win_lo) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t s = 1 + t; s < n_stmt; s += stride)
    if (flag[s]) win_lo[incl[s] - 1] = s;
  if (t == 0) win_lo[n_windows] = n_stmt;
}

// capacity (power of two >= 2 * touches) of every window's table; a one-statement window needs none
__global__ void window_caps_kernel(const int2* __restrict__ stmt, const int64_t* __restrict__ win_lo, int64_t n_windows,
                                   int64_t* __restrict__ cap) {
  const int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w > n_windows) return;
  if (w == n_windows) { cap[w] = 0; return; }
  const int64_t lo = win_lo[w], hi = win_lo[w + 1] - 1;
  if (lo == hi) { cap[w] = 0; return; }
  const int64_t touches = (static_cast<int64_t>(stmt[hi].y) - stmt[lo - 1].y) + (hi - lo + 1);
  int64_t c = 64;
  while (c < 2 * touches) c <<= 1;
  cap[w] = c;
}

__device__ __forceinline__ uint32_t slot_hash(int32_t x) {
  uint32_t h = static_cast<uint32_t>(x) * 0x9E3779B1u;
  return h ^ (h >> 15);
}
// Find (or insert) the entry of slot x and return its (writer step, reader step).  One 16-byte
// L2 read in the common case; the CAS is only needed for a slot's first touch in the window.
__device__ __forceinline__ SlotEntry* table_entry(SlotEntry* T, uint32_t mask, int32_t x, int32_t& wl, int32_t& rl) {
  uint32_t h = slot_hash(x) & mask;
  while (true) {
    const int4 e = __ldcg(reinterpret_cast<const int4*>(&T[h]));
    if (e.x == x) {
      wl = e.z;
      rl = e.w;
      return &T[h];
    }
    if (e.x == -1) {
      const int32_t k = atomicCAS(&T[h].key, -1, x);
      if (k == -1 || k == x) {  // fresh (possibly inserted by a sibling lane just now): no writer, no reader
        wl = -1;
        rl = -1;
        return &T[h];
      }
    }
    h = (h + 1) & mask;
  }
}

// One statement, the whole warp: lanes share its operands (any number of them) and its LHS.  Returns its local step.
__device__ __forceinline__ int32_t window_walk_one(SlotEntry* T, uint32_t mask, const int32_t* __restrict__ idx, int64_t o0,
                                                   int64_t o1, int32_t lhs, int lane, int32_t* __restrict__ oplev) {
  const int64_t nops = o1 - o0;
  int32_t lv = 1;
  if (nops < 32) {
    // one pass: lane < nops handles an operand, lane == nops the LHS
    const int32_t x = (lane < nops) ? idx[o0 + lane] : (lane == nops ? lhs : -1);
    SlotEntry* e = nullptr;
    if (x >= 0) {
      int32_t wl, rl;
      e = table_entry(T, mask, x, wl, rl);
      lv = (lane < nops) ? max(wl + 1, rl)          // RAW (strict), MONO (non-strict); empty = -1
                         : max(wl, rl) + 1;         // WAW, WAR (strict)
      lv = max(lv, 1);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) lv = max(lv, __shfl_xor_sync(0xffffffffu, lv, d));
    if (lane < nops) __stcg(&e->rl, lv);            // lv >= the old rl by MONO
    __syncwarp();
    if (lane == nops) __stcg(reinterpret_cast<int2*>(&e->wl), make_int2(lv, 0));   // the write retires the old version and its readers
  } else {
    for (int64_t o = o0 + lane; o < o1; o += 32) {
      int32_t wl, rl;
      table_entry(T, mask, idx[o], wl, rl);
      lv = max(lv, max(wl + 1, rl));
    }
    int32_t wl, rl;
    SlotEntry* eL = table_entry(T, mask, lhs, wl, rl);
    lv = max(lv, max(wl, rl) + 1);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) lv = max(lv, __shfl_xor_sync(0xffffffffu, lv, d));
    __syncwarp();
    for (int64_t o = o0 + lane; o < o1; o += 32) {
      int32_t a, b;
      SlotEntry* e = table_entry(T, mask, idx[o], a, b);
      __stcg(&e->rl, lv);
    }
    __syncwarp();
    if (lane == 0) __stcg(reinterpret_cast<int2*>(&eL->wl), make_int2(lv, 0));
  }
  __syncwarp();
  if (oplev)
    for (int64_t o = o0 + lane; o < o1; o += 32) oplev[o] = lv;
  return lv;
}

// The walk of a window is a chain of dependent table look-ups, one L2/DRAM round trip per statement (the tables of
// all windows together are far larger than L2).  Two things shorten the chain:
//   * Consecutive statements rarely touch a common slot, so the warp looks up FOUR statements at a time -- eight lanes
//     each: up to seven operands and the LHS -- and commits the longest prefix of them that is conflict-free: a
//     statement that shares ANY slot with an earlier statement of the round (found with one match.any over the 32
//     slot ids) must see that statement's updates, so it and its successors are looked up again in the next round.
//   * The statements and operand slots of the next two rounds are already in registers (a sliding window of twelve
//     statements in three tiers of four, refilled by loads that are issued BEHIND the round's table probe), so the
//     only memory latency on the critical path of a round is the probe itself.
// The result is exactly that of the one-at-a-time walk (oracle_levels mode 2).
struct WalkTier {
  int32_t lhs, beg, end, x;   // the statement of my group; x = the slot my lane looks up (-1: none)
  bool ok;                    // the statement exists and has at most seven operands
  bool exists;
};
// statements base+g (g = 0..3) -> a tier; prev_end = end of statement base-1 (ignored when base-1 does not exist)
__device__ __forceinline__ WalkTier walk_load_tier(const int2* __restrict__ stmt, const int32_t* __restrict__ idx, int64_t base,
                                                   int64_t hi, int32_t prev_end, int g, int q) {
  WalkTier t;
  const int64_t sg = base + g;
  int2 st = make_int2(0, prev_end);
  if (q == 0 && sg <= hi) st = stmt[sg];
  const int32_t e0 = __shfl_sync(0xffffffffu, st.y, 0), e1 = __shfl_sync(0xffffffffu, st.y, 8),
                e2 = __shfl_sync(0xffffffffu, st.y, 16);
  t.exists = sg <= hi;
  t.lhs = __shfl_sync(0xffffffffu, st.x, g * 8);
  t.end = __shfl_sync(0xffffffffu, st.y, g * 8);
  t.beg = g == 0 ? prev_end : (g == 1 ? e0 : (g == 2 ? e1 : e2));
  const int32_t nops = t.end - t.beg;
  t.ok = t.exists && nops <= 7;
  t.x = -1;
  if (t.ok) t.x = q < nops ? idx[t.beg + q] : (q == nops ? t.lhs : -1);
  return t;
}
__device__ __forceinline__ WalkTier walk_pick(const WalkTier& a, const WalkTier& b, int j, int q) {
  // tier entry number j of the concatenation (a, b): groups 0..3 from a, 4..7 from b
  const int src = ((j & 3) << 3) + q;
  WalkTier r;
  const bool hi4 = j >= 4;
  const int32_t la = __shfl_sync(0xffffffffu, a.lhs, src), lb = __shfl_sync(0xffffffffu, b.lhs, src);
  const int32_t ba = __shfl_sync(0xffffffffu, a.beg, src), bb = __shfl_sync(0xffffffffu, b.beg, src);
  const int32_t ea = __shfl_sync(0xffffffffu, a.end, src), eb = __shfl_sync(0xffffffffu, b.end, src);
  const int32_t xa = __shfl_sync(0xffffffffu, a.x, src), xb = __shfl_sync(0xffffffffu, b.x, src);
  const int fa = __shfl_sync(0xffffffffu, (a.ok ? 1 : 0) | (a.exists ? 2 : 0), src);
  const int fb = __shfl_sync(0xffffffffu, (b.ok ? 1 : 0) | (b.exists ? 2 : 0), src);
  r.lhs = hi4 ? lb : la;
  r.beg = hi4 ? bb : ba;
  r.end = hi4 ? eb : ea;
  r.x = hi4 ? xb : xa;
  const int f = hi4 ? fb : fa;
  r.ok = f & 1;
  r.exists = f & 2;
  return r;
}

__global__ void __launch_bounds__(128) window_levels_kernel(const int2* __restrict__ stmt, int64_t n_stmt,
                                                            const int32_t* __restrict__ idx,
                                                            const int64_t* __restrict__ win_lo,
                                                            int64_t n_windows, SlotEntry* __restrict__ tables,
                                                            const int64_t* __restrict__ tab_off,
                                                            int32_t* __restrict__ level, int32_t* __restrict__ win_depth,
                                                            int32_t* __restrict__ oplev) {
  // oplev[o] = local step of the statement that owns operation o: the packing pass reads it instead of searching
  // the statement table for every operation (left unwritten for one-statement windows: their step is 1)
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_windows) return;
  const int64_t lo = win_lo[w];
  const int64_t hi = win_lo[w + 1] - 1;
  if (lo == hi) {  // a window of one statement: step 1, nothing to look up
    if (lane == 0) {
      level[lo] = 1;
      win_depth[w] = 1;
    }
    return;
  }
  SlotEntry* T = tables + tab_off[w];
  const uint32_t mask = static_cast<uint32_t>(tab_off[w + 1] - tab_off[w]) - 1u;
  const int g = lane >> 3, q = lane & 7;
  int32_t depth = 0;
  int64_t s = lo;
  bool warm = false;
  WalkTier c, n;   // statements s..s+3 and s+4..s+7
  int32_t o0 = stmt[lo - 1].y;   // end of statement s-1 (operation offsets fit 32 bits: n_ops + n_statements < 2^31)
  while (s <= hi) {
    if (!warm) {
      c = walk_load_tier(stmt, idx, s, hi, o0, g, q);
      n = walk_load_tier(stmt, idx, s + 4, hi, __shfl_sync(0xffffffffu, c.end, 24), g, q);
      warm = true;
    }
    const uint32_t okm = __ballot_sync(0xffffffffu, q == 0 && c.ok);
    const int ngrp = (okm & 1u) ? ((okm >> 8 & 1u) ? ((okm >> 16 & 1u) ? ((okm >> 24 & 1u) ? 4 : 3) : 2) : 1) : 0;
    if (ngrp == 0) {   // statement s has eight operands or more: the whole warp takes it
      const int32_t b0 = __shfl_sync(0xffffffffu, c.beg, 0), e0 = __shfl_sync(0xffffffffu, c.end, 0);
      const int32_t lhs0 = __shfl_sync(0xffffffffu, c.lhs, 0);
      const int32_t lv = window_walk_one(T, mask, idx, b0, e0, lhs0, lane, oplev);
      if (lane == 0) level[s] = lv;
      depth = max(depth, lv);
      o0 = e0;
      ++s;
      warm = false;
      continue;
    }
    // 1. the round's table probe goes out first ...
    const int32_t x = (g < ngrp) ? c.x : -1;
    uint32_t h = 0;
    int4 probe = make_int4(-1, 0, -1, -1);
    if (x >= 0) {
      h = slot_hash(x) & mask;
      probe = __ldcg(reinterpret_cast<const int4*>(&T[h]));
    }
    // 2. ... then the loads that refill the window two rounds ahead (statements s+8..s+11 and their operand slots)
    const WalkTier f = walk_load_tier(stmt, idx, s + 8, hi, __shfl_sync(0xffffffffu, n.end, 24), g, q);
    // 3. the probe's answer (open addressing: walk on from h when the first entry belongs to another slot).
    //    (A variant that enters new slots with plain stores and sorts out clashing entries with two more match.any
    //    was measured 2x SLOWER than this CAS path -- 119 vs 58 ms on the 1e7-statement tape: match.any over 32
    //    distinct values costs about as much as the DRAM round trip it was meant to save.)
    SlotEntry* e = nullptr;
    int32_t v = 1;
    if (x >= 0) {
      int32_t wl, rl;
      if (probe.x == x) {
        e = &T[h];
        wl = probe.z;
        rl = probe.w;
      } else {
        e = table_entry(T, mask, x, wl, rl);
      }
      const int32_t nops = c.end - c.beg;
      v = (q < nops) ? max(wl + 1, rl) : max(wl, rl) + 1;   // RAW (strict), MONO (non-strict) | WAW, WAR (strict)
      v = max(v, 1);
    }
    v = max(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = max(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = max(v, __shfl_xor_sync(0xffffffffu, v, 4));   // the statement's step, in all eight lanes of its group
    // a slot shared with a statement of an EARLIER group of the round?  One match.any over the 32 slot ids (measured
    // against 24 shuffles + compares: 58 vs 70 ms for the walk of the 1e7-statement tape)
    const uint32_t lower = (g == 0) ? 0u : ((1u << (g * 8)) - 1u);   // the lanes of the groups before mine
    const uint32_t same = __match_any_sync(0xffffffffu, x);
    const uint32_t cm = __ballot_sync(0xffffffffu, x >= 0 && (same & lower) != 0u);
    int ncommit = ngrp;
    if (cm) ncommit = min(ncommit, (__ffs(cm) - 1) >> 3);   // the first group that shares a slot with an earlier one (>= 1)
    const bool commit = g < ncommit;
    const int32_t my_nops = c.end - c.beg;
    if (commit && x >= 0 && q < my_nops) {
      ADEPT_DCHECK(c.beg + q < c.end && s + g <= hi);
      __stcg(&e->rl, v);                                                     // readers of the live version
      oplev[c.beg + q] = v;
    }
    __syncwarp();
    if (commit && q == my_nops) {                                            // the write retires the old version and its readers
      __stcg(reinterpret_cast<int2*>(&e->wl), make_int2(v, 0));
      level[s + g] = v;
    }
    __syncwarp();
    int32_t vm = commit ? v : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) vm = max(vm, __shfl_xor_sync(0xffffffffu, vm, d));
    depth = max(depth, vm);
    // 4. slide the window by the statements that were committed
    const WalkTier c2 = walk_pick(c, n, ncommit + g, q);
    const WalkTier n2 = walk_pick(n, f, ncommit + g, q);      // entry (ncommit + 4 + g) of (c, n, f) = entry (ncommit + g) of (n, f)
    o0 = __shfl_sync(0xffffffffu, c.end, (ncommit - 1) * 8);
    c = c2;
    n = n2;
    s += ncommit;
  }
  if (lane == 0) win_depth[w] = depth;
}

// Same analysis for tapes with a SMALL gradient list (max_gradient * 8 bytes fits shared memory, e.g.
// the 12 287 slots of Toon nx=4096): the per-slot (writer step, reader step) table is a direct-indexed
// array in shared memory instead of a hash table in L2, which takes the L2 round trip out of every
// statement.  One warp (= one CTA) per window; the window's final table is written to wstate[w][slot]
// for the packing pass.
//
// The walk is a chain of dependent shared-memory round trips (operand -> table entry -> max -> table
// update), ~150 cycles per statement; a global load inside that chain would cost several times as much.
// So the window's statements and operands are streamed into shared memory ahead of the walk with
// cp.async (no registers held): tiles of kWinTile statements, statement tiles two ahead (their ends size
// the operand copy), operand tiles one ahead.  A tile with more than kWinIdxCap operands (a long
// statement) reads its operands from global memory instead.
constexpr int kWinTile = 128;
constexpr int kWinIdxCap = 4096;
constexpr size_t kWinStageBytes = size_t(3) * (kWinTile + 1) * sizeof(int2) + size_t(2) * kWinIdxCap * sizeof(int32_t);

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(32) window_levels_smem_kernel(const int2* __restrict__ stmt, const int32_t* __restrict__ idx,
                                                                const int64_t* __restrict__ win_lo, int64_t n_windows,
                                                                int64_t max_gradient, int2* __restrict__ wstate,
                                                                int32_t* __restrict__ level,
                                                                int32_t* __restrict__ win_depth) {
  extern __shared__ int2 tab[];   // tab[slot] = {writer step, max reader step}; 0 = none; then the staging buffers
  int2* const sbuf = tab + max_gradient;                                        // [3][kWinTile + 1]
  int32_t* const ibuf = reinterpret_cast<int32_t*>(sbuf + 3 * (kWinTile + 1));  // [2][kWinIdxCap]
  const int64_t w = blockIdx.x;
  const int lane = threadIdx.x;
  if (w >= n_windows) return;
  const int64_t lo = win_lo[w];
  const int64_t hi = win_lo[w + 1] - 1;
  int2* out = wstate + w * max_gradient;
  if (lo == hi) {  // a window of one statement: step 1; its table is never read (single-statement update path)
    if (lane == 0) {
      level[lo] = 1;
      win_depth[w] = 1;
    }
    return;
  }
  for (int64_t i = lane; i < max_gradient; i += 32) tab[i] = make_int2(0, 0);
  const int64_t n_tiles = (hi - lo + 1 + kWinTile - 1) / kWinTile;
  // statement tile t -> sbuf[t % 3][0 .. cnt]: entry 0 is the statement BEFORE the tile (its end = first operand)
  auto issue_stmts = [&](int64_t t) {
    if (t < n_tiles) {
      const int64_t s0 = lo + t * kWinTile;
      const int cnt = static_cast<int>(min(static_cast<int64_t>(kWinTile), hi + 1 - s0));
      int2* dst = sbuf + (t % 3) * (kWinTile + 1);
      for (int k = lane; k <= cnt; k += 32) cp_async8(dst + k, stmt + s0 - 1 + k);
    }
    cp_async_commit();
  };
  // operand tile t -> ibuf[t % 2] (statement tile t must have arrived)
  auto issue_ops = [&](int64_t t) {
    if (t < n_tiles) {
      const int64_t s0 = lo + t * kWinTile;
      const int cnt = static_cast<int>(min(static_cast<int64_t>(kWinTile), hi + 1 - s0));
      const int2* sb = sbuf + (t % 3) * (kWinTile + 1);
      const int64_t o0 = sb[0].y, n = static_cast<int64_t>(sb[cnt].y) - o0;
      if (n <= kWinIdxCap) {
        int32_t* dst = ibuf + (t % 2) * kWinIdxCap;
        for (int64_t k = lane; k < n; k += 32) cp_async4(dst + k, idx + o0 + k);
      }
    }
    cp_async_commit();
  };
  issue_stmts(0);
  issue_stmts(1);
  cp_async_wait<1>();
  __syncwarp();
  issue_ops(0);
  int32_t depth = 0;
  for (int64_t t = 0; t < n_tiles; ++t) {
    issue_stmts(t + 2);   // pending, oldest first: S(t+1), I(t), S(t+2)
    cp_async_wait<2>();   // S(t+1) has arrived
    __syncwarp();
    issue_ops(t + 1);     // pending: I(t), S(t+2), I(t+1)
    cp_async_wait<2>();   // I(t) has arrived
    __syncwarp();
    const int64_t s0 = lo + t * kWinTile;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kWinTile), hi + 1 - s0));
    const int2* sb = sbuf + (t % 3) * (kWinTile + 1);
    const int64_t tile_o0 = sb[0].y;
    const bool staged = static_cast<int64_t>(sb[cnt].y) - tile_o0 <= kWinIdxCap;
    const int32_t* ib = ibuf + (t % 2) * kWinIdxCap;
    int64_t o0 = tile_o0;
    const int tile_lo32 = static_cast<int>(tile_o0);   // ends of a staged tile are compared in 32 bits (wrap-safe)
    int o0r = 0;          // o0 - tile_o0
    int32_t mylv = 0;     // lane (k-1) & 31 keeps the step of statement k: one coalesced store per 32 statements
    for (int k = 1; k <= cnt; ++k) {
      const int2 st = sb[k];
      const int o1r = st.y - tile_lo32;
      const int nops = o1r - o0r;
      int32_t lv = 1;
      if (staged && nops < 32) {
        // the common case, kept to ~30 instructions: a single warp walks the window, so every instruction of this
        // dependent chain is paid in full
        const int32_t lhs = st.x;
        const int32_t x = (lane < nops) ? ib[o0r + lane] : lhs;
        const int2 e = tab[x];
        int32_t v = (lane < nops) ? max(e.x + 1, e.y)    // RAW (strict), MONO (non-strict)
                                  : max(e.x, e.y) + 1;   // WAW, WAR (strict)
        v = (lane <= nops) ? max(v, 1) : 1;
        lv = __reduce_max_sync(0xffffffffu, v);
        if (lane < nops && x != lhs) tab[x].y = lv;      // readers of the live version (a self reference is retired below)
        if (lane == nops) tab[lhs] = make_int2(lv, 0);   // the write retires the old version and its readers
      } else {
        const int32_t lhs = st.x;
        const int64_t o1 = o0 + nops;
        for (int64_t o = o0 + lane; o < o1; o += 32) {
          const int2 e = tab[staged ? ib[o - tile_o0] : idx[o]];
          lv = max(lv, max(e.x + 1, e.y));
        }
        const int2 eL = tab[lhs];
        lv = max(lv, max(eL.x, eL.y) + 1);
        lv = __reduce_max_sync(0xffffffffu, lv);
        __syncwarp();
        for (int64_t o = o0 + lane; o < o1; o += 32) tab[staged ? ib[o - tile_o0] : idx[o]].y = lv;
        __syncwarp();
        if (lane == 0) tab[lhs] = make_int2(lv, 0);
      }
      __syncwarp();
      if (lane == ((k - 1) & 31)) mylv = lv;
      if ((k & 31) == 0 || k == cnt) {
        const int kb = (k - 1) & ~31;   // statements kb+1 .. k of the tile
        if (kb + lane < k) level[s0 + kb + lane] = mylv;
      }
      depth = max(depth, lv);
      o0 += nops;
      o0r = o1r;
    }
    __syncwarp();   // the buffers of tile t are rewritten two and three tiles from now
  }
  cp_async_wait<0>();
  for (int64_t i = lane; i < max_gradient; i += 32) out[i] = tab[i];
  if (lane == 0) win_depth[w] = depth;
}
// packing update from a dense window state (see window_update_kernel)
__global__ void window_update_dense_kernel(const int2* __restrict__ state, int64_t max_gradient,
                                           const int32_t* __restrict__ gmap, int32_t* __restrict__ gw,
                                           int32_t* __restrict__ gr) {
  for (int64_t x = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; x < max_gradient;
       x += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int2 e = state[x];
    if (e.x > 0) {
      gw[x] = gmap[e.x - 1];
      gr[x] = e.y > 0 ? gmap[e.y - 1] : 0;
    } else if (e.y > 0) {
      gr[x] = max(gr[x], gmap[e.y - 1]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Window packing.  Local steps are turned into global ones by mapping every local step l of a window
// to a global step g[l], strictly increasing in l and at least the largest lower bound that earlier
// windows impose on a statement of step l (RAW/WAR/WAW/MONO), summarised per slot in gw (global step
// of the last writer) and gr (max global step of the readers of the live version).  Windows whose statements do not depend on their predecessors
// share steps (Rosenbrock: 613 windows -> a handful of steps); dependent ones stack up as before.
// Windows are visited in order (two small launches each); the work inside a window is parallel.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t stmt_of_op_range(const int2* __restrict__ stmt, int64_t lo, int64_t hi, int64_t o) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (static_cast<int64_t>(stmt[mid].y) > o) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}
// need[l-1] = the largest lower bound that earlier windows impose on a statement of local step l.
// Consecutive statements of a tape have unrelated local steps, so the elements of a warp rarely agree on l: the
// maxima are first formed per CTA in shared memory (a window has a few dozen local steps; kNeedSmem of them fit)
// and only one global atomic per (CTA, local step) is issued -- a window's 3e5 touches would otherwise queue up
// behind a few dozen addresses in L2 (measured: 131 us per window of 65 536 statements, 85 % of the packing pass).
constexpr int kNeedSmem = 4096;
__global__ void __launch_bounds__(256) window_need_kernel(const int2* __restrict__ stmt, const int32_t* __restrict__ idx,
                                                          int64_t lo, int64_t hi, int64_t o0, int64_t o1,
                                                          const int32_t* __restrict__ level, const int32_t* __restrict__ gw,
                                                          const int32_t* __restrict__ gr, int32_t* need, int32_t depth) {
  __shared__ int32_t sneed[kNeedSmem];
  const int ns = depth < kNeedSmem ? depth : kNeedSmem;
  for (int k = threadIdx.x; k < ns; k += blockDim.x) sneed[k] = 0;
  __syncthreads();
  const int64_t n_ops = o1 - o0, total = n_ops + (hi - lo + 1);
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int32_t bound, l;
    if (t < n_ops) {
      const int64_t o = o0 + t;
      const int64_t s = stmt_of_op_range(stmt, lo, hi, o);
      const int32_t x = idx[o];
      bound = max(gw[x] + 1, gr[x]);
      l = level[s];
    } else {
      const int64_t s = lo + (t - n_ops);
      const int32_t x = stmt[s].x;
      bound = max(gw[x], gr[x]) + 1;
      l = level[s];
    }
    if (bound > 1) {
      if (l <= ns) atomicMax(&sneed[l - 1], bound);
      else atomicMax(&need[l - 1], bound);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ns; k += blockDim.x)
    if (sneed[k] > 1) atomicMax(&need[k], sneed[k]);
}
// gmap[l-1] = global step of local step l: strictly increasing, at least need[l-1].  With u[l] = gmap[l] - l the
// recurrence gmap[l] = max(need[l], gmap[l-1] + 1) is a running maximum, u[l] = max(need[l] - l, u[l-1]), u[-1] = 1:
// one warp scans it 32 local steps at a time.
__global__ void window_map_kernel(const int32_t* __restrict__ need, int32_t depth, int32_t* __restrict__ gmap) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int32_t carry = 1;
  for (int32_t l0 = 0; l0 < depth; l0 += 32) {
    const int32_t l = l0 + lane;
    int32_t u = l < depth ? need[l] - l : INT32_MIN;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, u, d);
      if (lane >= d) u = max(u, o);
    }
    u = max(u, carry);
    if (l < depth) gmap[l] = u + l;
    carry = __shfl_sync(0xffffffffu, u, 31);
  }
}
// fold the window's final per-slot state (its hash table, local steps) into the global summary
__global__ void window_update_kernel(const SlotEntry* __restrict__ T, int64_t cap, const int32_t* __restrict__ gmap,
                                     int32_t* __restrict__ gw, int32_t* __restrict__ gr) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < cap;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const SlotEntry e = T[i];
    if (e.key < 0) continue;
    if (e.wl > 0) {  // written in this window: its last writer and the readers after it are what is live now
      gw[e.key] = gmap[e.wl - 1];
      gr[e.key] = e.rl > 0 ? gmap[e.rl - 1] : 0;
    } else if (e.rl > 0) {  // only read: more readers of the version that was already live
      gr[e.key] = max(gr[e.key], gmap[e.rl - 1]);
    }
  }
}
// The whole packing pass as ONE cooperative kernel: windows in order, two grid-wide barriers per window
// (need | barrier | map (every CTA, redundantly, for itself) + update | barrier) instead of three dependent launches
// from the host.  Same arithmetic as window_need_kernel / window_map_kernel / window_update*_kernel above.
struct PackArgs {
  const int2* stmt;
  const int32_t* idx;
  const int64_t* win_lo;    // [n_windows+1] first statement of each window
  const int64_t* win_op;    // [n_windows+1] first operation of each window
  const int32_t* win_base;  // [n_windows+1] first entry of each window in need/gmap
  const int64_t* tab_off;   // [n_windows+1] hash-table offsets (hash variant)
  const SlotEntry* tables;  // hash variant
  const int2* wstate;       // dense variant: [n_windows][max_gradient]
  int64_t max_gradient;
  int64_t n_windows;
  const int32_t* level;     // local steps
  int32_t* gw;
  int32_t* gr;
  int32_t* need;
  int32_t* gmap;
  const int32_t* oplev;     // local step of every operation's statement (hash variant), or null: search the statement table
  int dense;
};

__global__ void __launch_bounds__(256) window_pack_kernel(PackArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ int32_t sneed[kNeedSmem];   // phase A: the CTA's maxima; phase B/C: this window's local -> global step map
  const int64_t gtid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t gsize = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t w = 0; w < A.n_windows; ++w) {
    const int64_t lo = A.win_lo[w], hi = A.win_lo[w + 1] - 1, o0 = A.win_op[w], o1 = A.win_op[w + 1];
    const int32_t base = A.win_base[w], depth = A.win_base[w + 1] - base;
    int32_t* need = A.need + base;
    int32_t* gmap = A.gmap + base;
    const int ns = depth < kNeedSmem ? depth : kNeedSmem;
    // ---- A. need[l-1] = the largest lower bound earlier windows impose on a statement of local step l ----
    for (int k = threadIdx.x; k < ns; k += blockDim.x) sneed[k] = 0;
    __syncthreads();
    const int64_t n_ops = o1 - o0, total = n_ops + (hi - lo + 1);
    for (int64_t t = gtid; t < total; t += gsize) {
      int32_t bound, l;
      if (t < n_ops) {
        const int64_t o = o0 + t;
        const int32_t x = A.idx[o];
        bound = max(__ldcg(&A.gw[x]) + 1, __ldcg(&A.gr[x]));   // written by other SMs in the previous window's phase C
        l = (lo == hi) ? 1 : (A.oplev ? A.oplev[o] : A.level[stmt_of_op_range(A.stmt, lo, hi, o)]);
      } else {
        const int64_t s = lo + (t - n_ops);
        const int32_t x = A.stmt[s].x;
        bound = max(__ldcg(&A.gw[x]), __ldcg(&A.gr[x])) + 1;
        l = A.level[s];
      }
      ADEPT_DCHECK(l >= 1 && l <= depth);
      if (bound > 1) {
        if (l <= ns) atomicMax(&sneed[l - 1], bound);
        else atomicMax(&need[l - 1], bound);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < ns; k += blockDim.x)
      if (sneed[k] > 1) atomicMax(&need[k], sneed[k]);
    grid.sync();
    // ---- B. gmap[l-1] = max(need[l-1], gmap[l-2] + 1): a running maximum of need[l] - l (see window_map_kernel);
    //         every CTA scans it for itself into shared memory, CTA 0 also publishes it ----
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      int32_t carry = 1;
      for (int32_t l0 = 0; l0 < depth; l0 += 32) {
        const int32_t l = l0 + lane;
        int32_t u = l < depth ? __ldcg(&need[l]) - l : INT32_MIN;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int32_t o = __shfl_up_sync(0xffffffffu, u, d);
          if (lane >= d) u = max(u, o);
        }
        u = max(u, carry);
        if (l < depth) {
          if (l < kNeedSmem) sneed[l] = u + l;
          if (blockIdx.x == 0 || l >= kNeedSmem) gmap[l] = u + l;   // steps beyond the shared table are read back from global
        }
        carry = __shfl_sync(0xffffffffu, u, 31);
      }
    }
    __syncthreads();
    if (depth > kNeedSmem) grid.sync();   // (never with the default window length) readers of gmap[l >= kNeedSmem] wait for every CTA
    auto G = [&](int32_t l1) -> int32_t {
      ADEPT_DCHECK(l1 >= 1 && l1 <= depth);
      return l1 <= kNeedSmem ? sneed[l1 - 1] : __ldcg(&gmap[l1 - 1]);
    };
    // ---- C. fold the window's final per-slot state into the global summary ----
    if (lo == hi) {
      // a window of one (huge) statement has no table: its operands are readers at its step, then its LHS is written
      const int32_t step = G(1);
      for (int64_t o = o0 + gtid; o < o1; o += gsize) atomicMax(&A.gr[A.idx[o]], step);
      grid.sync();
      if (gtid == 0) {
        const int32_t lhs = A.stmt[lo].x;
        A.gw[lhs] = step;
        A.gr[lhs] = 0;
      }
    } else if (A.dense) {
      const int2* state = A.wstate + w * A.max_gradient;
      for (int64_t x = gtid; x < A.max_gradient; x += gsize) {
        const int2 e = state[x];
        if (e.x > 0) {
          A.gw[x] = G(e.x);
          A.gr[x] = e.y > 0 ? G(e.y) : 0;
        } else if (e.y > 0) {
          A.gr[x] = max(__ldcg(&A.gr[x]), G(e.y));
        }
      }
    } else {
      const SlotEntry* T = A.tables + A.tab_off[w];
      const int64_t cap = A.tab_off[w + 1] - A.tab_off[w];
      for (int64_t i = gtid; i < cap; i += gsize) {
        const SlotEntry e = T[i];
        if (e.key < 0) continue;
        if (e.wl > 0) {   // written in this window: its last writer and the readers after it are what is live now
          A.gw[e.key] = G(e.wl);
          A.gr[e.key] = e.rl > 0 ? G(e.rl) : 0;
        } else if (e.rl > 0) {   // only read: more readers of the version that was already live
          A.gr[e.key] = max(__ldcg(&A.gr[e.key]), G(e.rl));
        }
      }
    }
    grid.sync();
  }
}

// a window of one (huge) statement has no table: its operands are readers at its step ...
__global__ void window_update_single_ops_kernel(const int32_t* __restrict__ idx, int64_t o0, int64_t o1,
                                                const int32_t* __restrict__ gmap, int32_t* __restrict__ gr) {
  const int32_t step = gmap[0];
  for (int64_t o = o0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < o1;
       o += static_cast<int64_t>(gridDim.x) * blockDim.x)
    atomicMax(&gr[idx[o]], step);
}
// ... and its LHS is written at that step (after the reads: a self reference reads the old version)
__global__ void window_update_single_lhs_kernel(const int2* __restrict__ stmt, int64_t s, const int32_t* __restrict__ gmap,
                                                int32_t* __restrict__ gw, int32_t* __restrict__ gr) {
  const int32_t lhs = stmt[s].x;
  gw[lhs] = gmap[0];
  gr[lhs] = 0;
}
// out[w] = first operation of window w (= end of the statement before its first one)
__global__ void gather_ends_kernel(const int2* __restrict__ stmt, const int64_t* __restrict__ win_lo, int64_t n,
                                   int64_t* __restrict__ out) {
  const int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w < n) out[w] = stmt[win_lo[w] - 1].y;
}
// wide steps are cut into sub-steps of at most `cap` statements (always legal: a step's statements
// are independent, and the reverse target groups are formed per sub-step in schedule order)
__global__ void split_steps_kernel(uint32_t* __restrict__ sched_step, const uint32_t* __restrict__ sched_stmt, int64_t S,
                                   const int64_t* __restrict__ level_first, const int32_t* __restrict__ new_base,
                                   int64_t cap, int32_t* __restrict__ level) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = t; j < S; j += stride) {
    const uint32_t l = sched_step[j];
    const uint32_t nl = static_cast<uint32_t>(new_base[l] + (j - level_first[l]) / cap);
    sched_step[j] = nl;
    level[sched_stmt[j]] = static_cast<int32_t>(nl);
  }
}

// global step of every statement: gmap[step_base[window] + local step - 1]
__global__ void global_steps_kernel(int32_t* __restrict__ level, int64_t n_stmt, const int32_t* __restrict__ wincl,
                                    const int32_t* __restrict__ step_base, const int32_t* __restrict__ gmap) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t s = 1 + t; s < n_stmt; s += stride) level[s] = gmap[step_base[wincl[s] - 1] + level[s] - 1];
  if (t == 0) level[0] = 0;
}

// ---------------------------------------------------------------------------------------
// Level-major emission.
// ---------------------------------------------------------------------------------------
// keys for the stable sort of statements by level: key[j] = level[j+1], val[j] = j+1
__global__ void level_keys_kernel(const int32_t* __restrict__ level, int64_t S, uint32_t* __restrict__ key,
                                  uint32_t* __restrict__ val) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = t; j < S; j += stride) {
    key[j] = static_cast<uint32_t>(level[j + 1]);
    val[j] = static_cast<uint32_t>(j + 1);
  }
}

// nrec[j] for the scheduled statement order; level_first[l] = first j of level l (1-based), and
// level_first[n_levels+1] = S.
__global__ void sched_sizes_kernel(const int2* __restrict__ stmt, const uint32_t* __restrict__ sched_stmt,
                                   const uint32_t* __restrict__ sched_level, int64_t S, int32_t n_levels,
                                   int64_t* __restrict__ nrec, int64_t* __restrict__ level_first) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = t; j < S; j += stride) {
    const uint32_t s = sched_stmt[j];
    nrec[j] = static_cast<int64_t>(stmt[s].y) - stmt[s - 1].y + 1;
    if (j == 0 || sched_level[j] != sched_level[j - 1]) level_first[sched_level[j]] = j;
  }
  if (t == 0) {
    level_first[0] = 0;
    level_first[n_levels + 1] = S;
  }
}

// chunk flags: statement j opens a chunk if it is the first of its level or crosses a
// chunk_recs boundary of the stream.
__global__ void chunk_flags_kernel(const int64_t* __restrict__ foff, const uint32_t* __restrict__ sched_level,
                                   int64_t S, int64_t chunk_recs, int32_t* __restrict__ flag) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = t; j < S; j += stride) {
    int f = 0;
    if (j == 0 || sched_level[j] != sched_level[j - 1]) f = 1;
    else if (foff[j] / chunk_recs != foff[j - 1] / chunk_recs) f = 1;
    flag[j] = f;
  }
}

// steps that hold a statement of more than `thresh` records (its chunk may exceed one staged piece): flag[step] = 1
__global__ void mark_long_levels_kernel(const int64_t* __restrict__ foff, const uint32_t* __restrict__ sched_level, int64_t S,
                                        int64_t thresh, uint8_t* __restrict__ flag) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = t; j < S; j += stride)
    if (foff[j + 1] - foff[j] > thresh) flag[sched_level[j]] = 1;
}

// Guard of the balanced forward kernel's one-staged-piece assumption: out[0] = the longest chunk (in records) of any
// step that is NOT flagged long.  The host refuses the schedule when that exceeds the staging buffer -- a wrong flag
// must fail loudly at upload, not write past shared memory in a sweep.
__global__ void max_short_chunk_kernel(const int64_t* __restrict__ level_chunk, const int64_t* __restrict__ fwd_chunk,
                                       const uint8_t* __restrict__ level_long, int32_t n_levels,
                                       unsigned long long* __restrict__ out) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  unsigned long long m = 0;
  for (int64_t l = 1 + t; l <= n_levels; l += stride) {
    if (level_long[l]) continue;
    for (int64_t c = level_chunk[l]; c < level_chunk[l + 1]; ++c) {
      const unsigned long long len = static_cast<unsigned long long>(fwd_chunk[c + 1] - fwd_chunk[c]);
      m = len > m ? len : m;
    }
  }
  if (m) atomicMax(out, m);
}

// chunk tables: fwd_chunk[c] = foff[j] for the c-th flagged j; fwd_chunk[n_chunks] = R;
// rev_chunk[c'] = R - fwd_chunk[n_chunks - c']; level_chunk[l] = chunk id of level_first[l].
__global__ void chunk_tables_kernel(const int64_t* __restrict__ foff, const int32_t* __restrict__ flag,
                                    const int32_t* __restrict__ chunk_id /* exclusive scan of flag */, int64_t S,
                                    int64_t R, int64_t n_chunks, const int64_t* __restrict__ level_first,
                                    int32_t n_levels, int64_t* __restrict__ fwd_chunk, int64_t* __restrict__ rev_chunk,
                                    int64_t* __restrict__ level_chunk) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = t; j < S; j += stride) {
    if (flag[j]) {
      const int64_t c = chunk_id[j];
      fwd_chunk[c] = foff[j];
      if (rev_chunk) rev_chunk[n_chunks - c] = R - foff[j];
    }
  }
  for (int64_t l = 1 + t; l <= n_levels; l += stride) level_chunk[l] = chunk_id[level_first[l]];
  if (t == 0) {
    fwd_chunk[n_chunks] = R;
    if (rev_chunk) rev_chunk[0] = 0;
    level_chunk[0] = 0;
    level_chunk[n_levels + 1] = n_chunks;
  }
}

// Level-major record streams.  One thread per destination record of the forward stream; the
// mirrored record of the reverse stream is written by the same thread.
//   forward : statements in schedule order j = 0..S-1, each  OP* TAIL
//   reverse : statements in order j = S-1..0,          each  HEAD OP*  (ops ascending)
__global__ void emit_level_streams_kernel(const int2* __restrict__ stmt, const double* __restrict__ mult,
                                          const int32_t* __restrict__ idx, const uint32_t* __restrict__ sched_stmt,
                                          const int64_t* __restrict__ foff, int64_t S, int64_t R,
                                          Rec* __restrict__ fwd, Rec* __restrict__ rev) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t q = t; q < R; q += stride) {
    // largest j with foff[j] <= q
    int64_t lo = 0, hi = S - 1;
    while (lo < hi) {
      const int64_t mid = (lo + hi + 1) >> 1;
      if (foff[mid] <= q) lo = mid;
      else hi = mid - 1;
    }
    const int64_t j = lo;
    const uint32_t s = sched_stmt[j];
    const int64_t o0 = stmt[s - 1].y, o1 = stmt[s].y;
    const int64_t nrec = o1 - o0 + 1;
    const int64_t k = q - foff[j];
    const int64_t rbase = R - foff[j] - nrec;  // HEAD position in the reverse stream
    Rec r;
    if (k == nrec - 1) {
      r.m = 0.0;
      r.slot = stmt[s].x;
      r.kind = REC_MARK;
      fwd[q] = r;
      if (rev) rev[rbase] = r;
    } else {
      r.m = mult[o0 + k];
      r.slot = idx[o0 + k];
      r.kind = REC_OP;
      fwd[q] = r;
      if (rev) rev[rbase + 1 + k] = r;
    }
  }
}

// schedule look-ups: sched_lhs[j] = LHS slot of the j-th scheduled statement;
// pos_in_sched[s] = j.
__global__ void sched_lookup_kernel(const int2* __restrict__ stmt, const uint32_t* __restrict__ sched_stmt, int64_t S,
                                    int32_t* __restrict__ sched_lhs, int32_t* __restrict__ pos_in_sched) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = t; j < S; j += stride) {
    const uint32_t s = sched_stmt[j];
    sched_lhs[j] = stmt[s].x;
    pos_in_sched[s] = static_cast<int32_t>(j);
  }
}

// ---------------------------------------------------------------------------------------
// Reverse "target stream".  In a reverse sweep every operation (m, slot X) of statement s
// contributes m*a_s to g[X].  Inside one step several statements may name the same X, so the
// step is replayed as a GATHER per target slot: contributions to X are listed in the
// reference's order (statement descending, operand ascending) behind a HEAD(X) record:
//     HEAD(X)  OP(m, k)*      k = rank of the contributing statement inside its step
// a_s is read from the snapshot buffer abuf[k] that the step's first kernel fills.
// Built with two stable radix sorts of the operations: by slot, then by step.
// ---------------------------------------------------------------------------------------
// operations in reference reverse order: rank(o) = O - end[s] + (o - end[s-1])
__global__ void target_keys_kernel(const int2* __restrict__ stmt, int64_t n_stmt, const int32_t* __restrict__ idx,
                                   int64_t n_ops, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t o = t; o < n_ops; o += stride) {
    const int64_t s = stmt_of_op(stmt, n_stmt, o);
    const int64_t rank = n_ops - stmt[s].y + (o - stmt[s - 1].y);
    key[rank] = static_cast<uint32_t>(idx[o]);
    val[rank] = static_cast<uint32_t>(o);
  }
}
// second sort key: the step of the operation's statement
__global__ void target_step_keys_kernel(const int2* __restrict__ stmt, int64_t n_stmt, const uint32_t* __restrict__ val,
                                        int64_t n_ops, const int32_t* __restrict__ level, uint32_t* __restrict__ key) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n_ops; i += stride)
    key[i] = static_cast<uint32_t>(level[stmt_of_op(stmt, n_stmt, val[i])]);
}
// group boundaries: a new (step, slot) pair
__global__ void target_flags_kernel(const uint32_t* __restrict__ step_key, const uint32_t* __restrict__ val,
                                    const int32_t* __restrict__ idx, int64_t n_ops, int32_t* __restrict__ flag) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n_ops; i += stride) {
    int f = 0;
    if (i == 0 || step_key[i] != step_key[i - 1] || idx[val[i]] != idx[val[i - 1]]) f = 1;
    flag[i] = f;
  }
}
// emit the stream; incl[i] = inclusive scan of flag.  goff[g] = record offset of group g's HEAD,
// gstep[g] = its step.
__global__ void target_emit_kernel(const int2* __restrict__ stmt, int64_t n_stmt, const double* __restrict__ mult,
                                   const int32_t* __restrict__ idx, const uint32_t* __restrict__ step_key,
                                   const uint32_t* __restrict__ val, const int32_t* __restrict__ flag,
                                   const int32_t* __restrict__ incl, int64_t n_ops,
                                   const int32_t* __restrict__ pos_in_sched, const int64_t* __restrict__ level_first,
                                   Rec* __restrict__ tgt, int64_t* __restrict__ goff, int32_t* __restrict__ gstep) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n_ops; i += stride) {
    const int64_t o = val[i];
    const int64_t s = stmt_of_op(stmt, n_stmt, o);
    const int64_t pos = i + incl[i];
    Rec r;
    r.m = mult[o];
    r.slot = static_cast<int32_t>(pos_in_sched[s] - level_first[step_key[i]]);
    r.kind = REC_OP;
    tgt[pos] = r;
    if (flag[i]) {
      Rec h;
      h.m = 0.0;
      h.slot = idx[o];
      h.kind = REC_MARK;
      tgt[pos - 1] = h;
      goff[incl[i] - 1] = pos - 1;
      gstep[incl[i] - 1] = static_cast<int32_t>(step_key[i]);
    }
  }
}
// chunking of the target stream: group g opens a chunk if it is the first of its step or its
// HEAD crosses a chunk_recs boundary.  step_group[step] = first group of the step (steps
// without operations keep -1 and are patched on the host).
__global__ void target_chunk_flags_kernel(const int64_t* __restrict__ goff, const int32_t* __restrict__ gstep,
                                          int64_t n_groups, int64_t chunk_recs, int32_t* __restrict__ cflag,
                                          int64_t* __restrict__ step_group) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g = t; g < n_groups; g += stride) {
    int f = 0;
    if (g == 0 || gstep[g] != gstep[g - 1]) {
      f = 1;
      step_group[gstep[g]] = g;
    } else if (goff[g] / chunk_recs != goff[g - 1] / chunk_recs) {
      f = 1;
    }
    cflag[g] = f;
  }
}
// tchunk[c] = goff[g] for the c-th flagged group; group_chunk[g] = chunk id of group g
__global__ void target_chunk_table_kernel(const int64_t* __restrict__ goff, const int32_t* __restrict__ cflag,
                                          const int32_t* __restrict__ cincl, int64_t n_groups, int64_t n_tchunks,
                                          int64_t R_t, int64_t* __restrict__ tchunk) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g = t; g < n_groups; g += stride)
    if (cflag[g]) tchunk[cincl[g] - 1] = goff[g];
  if (t == 0) tchunk[n_tchunks] = R_t;
}

// step tables of the target stream, on the device: patch steps without operations (they own
// no group) to the next step's first group, then translate to chunk ids.  One thread: n_steps
// is small.
__global__ void target_step_tables_kernel(int64_t* __restrict__ step_group, int32_t n_steps, int64_t n_groups,
                                          const int32_t* __restrict__ cincl, int64_t n_tchunks,
                                          int64_t* __restrict__ step_tchunk) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  step_group[n_steps + 1] = n_groups;
  for (int32_t l = n_steps; l >= 1; --l)
    if (step_group[l] < 0) step_group[l] = step_group[l + 1];
  step_group[0] = 0;
  for (int32_t l = 1; l <= n_steps + 1; ++l) {
    const int64_t g = step_group[l];
    step_tchunk[l] = (g < n_groups) ? (cincl[g] - 1) : n_tchunks;
  }
  step_tchunk[0] = 0;
}

// ---------------------------------------------------------------------------------------
// FUSED reverse stream ("versioned rows"): one kernel per step instead of snapshot + gather.
//
// Every slot X owns two physical rows, row(X, p) = X + p*G.  The live value of X sits in row(X, p)
// with p = (number of statements LATER in the tape whose LHS is X) mod 2: the live row flips each time
// the reverse sweep passes a writer of X.  A step then only READS live rows and WRITES either a
// target's own live row in place (X is not an LHS of the step, so no other target of the step reads
// it) or the OTHER row of an LHS slot, so that a_s = g[row(lhs_s, p)] stays intact for every target of
// the step that still needs it:
//     ZMARK(row(lhs_s, p^1))  OP(m, row(lhs_s, p))*      target X = lhs_s: starts from +0.0; only
//                                                         self references of s can follow
//     MARK (row(X, p))        OP(m, row(lhs_s', p'))*    any other target: starts from its live value
// Per slot the contributions keep the reference's order (statement descending, operand ascending;
// adept/jacobian.cpp:410-428) exactly as in the two-kernel form.  tests/test_fused_stream_model.py is
// the sequential statement of this construction, checked against the oracle.
//
// Elements are the R = O + S references of the tape in the reference's reverse sweep order; element
// id: operation o -> o, LHS of statement s -> O + (s-1).
// ---------------------------------------------------------------------------------------
__global__ void fused_keys_kernel(const int2* __restrict__ stmt, int64_t n_stmt, const int32_t* __restrict__ idx,
                                  int64_t n_ops, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  const int64_t S = n_stmt - 1;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t o = t; o < n_ops; o += stride) {
    const int64_t s = stmt_of_op(stmt, n_stmt, o);
    const int64_t pos = (S - s) + (n_ops - stmt[s].y) + 1 + (o - stmt[s - 1].y);
    key[pos] = static_cast<uint32_t>(idx[o]);
    val[pos] = static_cast<uint32_t>(o);
  }
  for (int64_t s = 1 + t; s < n_stmt; s += stride) {
    const int64_t pos = (S - s) + (n_ops - stmt[s].y);
    key[pos] = static_cast<uint32_t>(stmt[s].x);
    val[pos] = static_cast<uint32_t>(n_ops + (s - 1));
  }
}
// isl[i] = 1 when element i is an LHS reference
__global__ void fused_islhs_kernel(const uint32_t* __restrict__ val, int64_t n, int64_t n_ops, int32_t* __restrict__ isl) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n; i += stride) isl[i] = (static_cast<int64_t>(val[i]) >= n_ops) ? 1 : 0;
}
// elements sorted by (slot, sweep order); E = exclusive scan of isl.  e0[slot] = E at the slot's first element.
__global__ void fused_segstart_kernel(const uint32_t* __restrict__ key, const int32_t* __restrict__ E, int64_t n,
                                      int32_t* __restrict__ e0) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n; i += stride)
    if (i == 0 || key[i] != key[i - 1]) e0[key[i]] = E[i];
}
// par[id] = (writers of the element's slot already passed by the sweep) & 1; fpar[slot] = parity after the sweep
__global__ void fused_parity_kernel(const uint32_t* __restrict__ key, const uint32_t* __restrict__ val,
                                    const int32_t* __restrict__ isl, const int32_t* __restrict__ E,
                                    const int32_t* __restrict__ e0, int64_t n, uint8_t* __restrict__ par,
                                    uint8_t* __restrict__ fpar) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n; i += stride) {
    const uint32_t x = key[i];
    const int32_t cnt = E[i] - e0[x];
    par[val[i]] = static_cast<uint8_t>(cnt & 1);
    if (i == n - 1 || key[i + 1] != x) fpar[x] = static_cast<uint8_t>((cnt + isl[i]) & 1);
  }
}
__device__ __forceinline__ int64_t fused_stmt_of(const int2* __restrict__ stmt, int64_t n_stmt, int64_t n_ops, int64_t id) {
  return id < n_ops ? stmt_of_op(stmt, n_stmt, id) : id - n_ops + 1;
}
// second sort key: the step of the element's statement
__global__ void fused_step_keys_kernel(const int2* __restrict__ stmt, int64_t n_stmt, int64_t n_ops,
                                       const uint32_t* __restrict__ val, int64_t n, const int32_t* __restrict__ level,
                                       uint32_t* __restrict__ key) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n; i += stride) key[i] = static_cast<uint32_t>(level[fused_stmt_of(stmt, n_stmt, n_ops, val[i])]);
}
__device__ __forceinline__ int32_t fused_slot_of(const int2* __restrict__ stmt, const int32_t* __restrict__ idx,
                                                 int64_t n_ops, int64_t id) {
  return id < n_ops ? idx[id] : stmt[id - n_ops + 1].x;
}
// elements sorted by (step, slot, sweep order): flag = a new (step, slot) group; hflag = the group needs an
// inserted MARK head (it opens with an operation; an LHS reference is its own head)
__global__ void fused_flags_kernel(const int2* __restrict__ stmt, const int32_t* __restrict__ idx, int64_t n_ops,
                                   const uint32_t* __restrict__ step_key, const uint32_t* __restrict__ val, int64_t n,
                                   int32_t* __restrict__ flag, int32_t* __restrict__ hflag, int* __restrict__ bad) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n; i += stride) {
    int f = 0;
    if (i == 0 || step_key[i] != step_key[i - 1] ||
        fused_slot_of(stmt, idx, n_ops, val[i]) != fused_slot_of(stmt, idx, n_ops, val[i - 1]))
      f = 1;
    const bool is_op = static_cast<int64_t>(val[i]) < n_ops;
    flag[i] = f;
    hflag[i] = (f && is_op) ? 1 : 0;
    if (!is_op && !f) atomicOr(bad, 1);   // cannot happen with a schedule that obeys RAW/WAR/WAW
  }
}
// emit the stream; gincl/hincl = inclusive scans of flag/hflag.  goff[g] = record offset of group g's head,
// gstep[g] = its step.
__global__ void fused_emit_kernel(const int2* __restrict__ stmt, int64_t n_stmt, const double* __restrict__ mult,
                                  const int32_t* __restrict__ idx, int64_t n_ops, const uint32_t* __restrict__ step_key,
                                  const uint32_t* __restrict__ val, const int32_t* __restrict__ flag,
                                  const int32_t* __restrict__ hflag, const int32_t* __restrict__ gincl,
                                  const int32_t* __restrict__ hincl, int64_t n, const uint8_t* __restrict__ par,
                                  int64_t G, Rec* __restrict__ out, int64_t* __restrict__ goff,
                                  int32_t* __restrict__ gstep) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = t; i < n; i += stride) {
    const int64_t id = val[i];
    const int64_t pos = i + hincl[i];
    const int64_t s = fused_stmt_of(stmt, n_stmt, n_ops, id);
    const int64_t lhs = stmt[s].x;
    const int64_t lp = par[n_ops + s - 1];     // parity of the LHS's live row when the sweep reaches s
    Rec r;
    if (id < n_ops) {
      r.m = mult[id];
      r.slot = static_cast<int32_t>(lhs + lp * G);
      r.kind = REC_OP;
      out[pos] = r;
      if (hflag[i]) {
        Rec h;
        h.m = 0.0;
        h.slot = static_cast<int32_t>(static_cast<int64_t>(idx[id]) + static_cast<int64_t>(par[id]) * G);
        h.kind = REC_MARK;
        out[pos - 1] = h;
        goff[gincl[i] - 1] = pos - 1;
        gstep[gincl[i] - 1] = static_cast<int32_t>(step_key[i]);
      }
    } else {
      r.m = 0.0;
      r.slot = static_cast<int32_t>(lhs + (lp ^ 1) * G);
      r.kind = REC_ZMARK;
      out[pos] = r;
      if (flag[i]) {
        goff[gincl[i] - 1] = pos;
        gstep[gincl[i] - 1] = static_cast<int32_t>(step_key[i]);
      }
    }
  }
}
// rows[i] = slot[i] + fpar[slot[i]]*G: where the swept value of each output slot ends up
__global__ void fused_out_rows_kernel(const int32_t* __restrict__ slot, int64_t n, const uint8_t* __restrict__ fpar,
                                      int64_t G, int32_t* __restrict__ rows) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) rows[i] = static_cast<int32_t>(static_cast<int64_t>(slot[i]) + static_cast<int64_t>(fpar[slot[i]]) * G);
}

// per-step table for the shared-memory-resident kernel: {first item, end item, first record, end record}
__global__ void step_info_kernel(const int64_t* __restrict__ first, const int64_t* __restrict__ off, int32_t n_levels,
                                 longlong4* __restrict__ info) {
  const int64_t l = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (l > n_levels + 1) return;
  if (l == 0 || l == n_levels + 1) { info[l] = make_longlong4(0, 0, 0, 0); return; }
  const int64_t i0 = first[l], i1 = first[l + 1];
  info[l] = make_longlong4(i0, i1, off[i0], off[i1]);
}

}  // namespace adept_b200
