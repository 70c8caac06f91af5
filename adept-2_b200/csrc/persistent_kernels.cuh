// persistent_kernels.cuh -- the level-scheduled REVERSE sweep as ONE cooperative kernel.
//
// The per-step path (replay_kernels.cuh) launches two kernels per step; a tape like Toon
// nx=4096 nt=1000 has ~3100 steps of ~4000 statements, so launch latency and the drain/fill of
// every small grid are a large part of each step, and they do not shrink when the directions are
// split over more GPUs.  Here a grid that fills the machine once walks ALL steps:
//
//     for step = n_steps .. 1:   snapshot phase  ->  grid.sync()  ->  target phase  ->  grid.sync()
//
// Each phase is the body of the corresponding per-step kernel, executed over "virtual blocks"
// (CTA b takes virtual blocks b, b + gridDim.x, ...).  Data that other CTAs write between phases
// (gradient rows, the a-snapshot, the activity bytes) is read with ld.global.cg (L2), never
// through L1.  The TMA ring and its mbarriers live for the whole kernel; stage and phase parity
// follow a running piece counter.  Arithmetic and accumulation order are those of the per-step
// kernels, i.e. the reference's (adept/jacobian.cpp:382-431).
#pragma once
#include <cooperative_groups.h>

#include "replay_kernels.cuh"

namespace adept_b200 {

struct PersistArgs {
  TargetArgs T;
  const int32_t* sched_lhs;     // LHS slot of every scheduled statement
  const int64_t* level_first;   // [n_levels+2] first schedule position of each step (device copy)
  const int64_t* step_tchunk;   // [n_levels+2] first target chunk of each step (device copy)
  int32_t n_levels;
  int gy;                       // column groups = ceil(n_colblocks / warps per CTA)
};

template <int V>
__device__ __forceinline__ void lv_load_cg(LaneVec<V>& v, const double* p) {
  if (V == 1) {
    v.x = __ldcg(p);
  } else {
    const double2 t = __ldcg(reinterpret_cast<const double2*>(p));
    v.x = t.x;
    reinterpret_cast<double*>(&v)[V - 1] = t.y;
  }
}
template <int V>
__device__ __forceinline__ void lv_store_cg(const LaneVec<V>& v, double* p) {
  if (V == 1) __stcg(p, v.x);
  else __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, reinterpret_cast<const double*>(&v)[V - 1]));
}

template <int V>
__global__ void __launch_bounds__(256, 3) reverse_persistent_kernel(PersistArgs P) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ StageRing ring;
  __shared__ uint8_t pf[kPieceRecs][8];
  __shared__ int list[256];
  __shared__ int cnt;
  const TargetArgs& A = P.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int ncb = A.flag_pitch;
  const uint32_t Pb = static_cast<uint32_t>(A.ref_block);
  const uint32_t lanes_per_grp = (V == 2) ? (Pb >= 2 ? Pb / 2 : 1) : Pb;
  const uint32_t grp_mask =
      (lanes_per_grp >= 32 ? 0xffffffffu : ((1u << lanes_per_grp) - 1u)) << ((lane / lanes_per_grp) * lanes_per_grp);
  const bool per_component = (V == 2 && Pb == 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t pc = 0;  // pieces this CTA has consumed so far: stage = pc % kStages, parity = (pc / kStages) & 1

  for (int32_t l = P.n_levels; l >= 1; --l) {
    // ---------------- phase A: a_s = g[lhs]; g[lhs] = 0 for the statements of the step ----------------
    const int64_t j0 = P.level_first[l];
    const int64_t n_stmts = P.level_first[l + 1] - j0;
    const int64_t total = n_stmts * A.n_colblocks;
    const int64_t nvb_a = (total + blockDim.x - 1) / blockDim.x;
    for (int64_t vb = blockIdx.x; vb < nvb_a; vb += gridDim.x) {
      if (threadIdx.x == 0) cnt = 0;
      __syncthreads();
      const int64_t t0 = vb * blockDim.x;
      {
        const int64_t t = t0 + threadIdx.x;
        if (t < total) {
          const int64_t k = t / A.n_colblocks;
          const int cb = static_cast<int>(t - k * A.n_colblocks);
          const int64_t lhs = P.sched_lhs[j0 + k];
          if (__ldcg(&A.rowact[lhs * ncb + cb])) list[atomicAdd(&cnt, 1)] = threadIdx.x;
          else const_cast<uint8_t*>(A.aflag)[k * ncb + cb] = 0;
        }
      }
      __syncthreads();
      const int n = cnt;
      for (int q = warp; q < n; q += nwarps) {
        const int64_t t = t0 + list[q];
        const int64_t k = t / A.n_colblocks;
        const int cb = static_cast<int>(t - k * A.n_colblocks);
        const int64_t lhs = P.sched_lhs[j0 + k];
        const int64_t d = (static_cast<int64_t>(cb) * 32 + lane) * V;
        double* p = A.g + lhs * A.K + d;
        LaneVec<V> a;
        lv_load_cg<V>(a, p);
        bool nzl = (a.x != 0.0) && (d < A.n_dirs);
        if (V == 2) nzl = nzl || ((reinterpret_cast<const double*>(&a)[V - 1] != 0.0) && (d + 1 < A.n_dirs));
        const bool any = __ballot_sync(0xffffffffu, nzl) != 0u;
        if (any) {
          lv_store_cg<V>(a, const_cast<double*>(A.abuf) + k * A.K + d);
          LaneVec<V> z;
          z.zero();
          lv_store_cg<V>(z, p);
        }
        if (lane == 0) {
          const_cast<uint8_t*>(A.aflag)[k * ncb + cb] = any ? 1 : 0;
          A.rowact[lhs * ncb + cb] = 0;
        }
      }
      __syncthreads();
    }
    grid.sync();

    // ---------------- phase B: per-target gather in the reference's accumulation order ----------------
    const int64_t c0 = P.step_tchunk[l], c1 = P.step_tchunk[l + 1];
    const int64_t nvb_b = (c1 - c0) * P.gy;
    for (int64_t vb = blockIdx.x; vb < nvb_b; vb += gridDim.x) {
      const int64_t c = c0 + vb / P.gy;
      const int cb0 = static_cast<int>(vb % P.gy) * nwarps;
      const int cb = cb0 + warp;
      const int64_t r0 = A.chunk_begin[c], r1 = A.chunk_begin[c + 1];
      const int64_t n_recs = r1 - r0;
      const Rec* chunk = A.recs + r0;
      const uint32_t n_pieces = static_cast<uint32_t>((n_recs + kPieceRecs - 1) / kPieceRecs);
      const bool live = cb < A.n_colblocks;
      const int64_t d0 = (static_cast<int64_t>(cb) * 32 + lane) * V;
      const bool ok0 = d0 < A.n_dirs, ok1 = d0 + 1 < A.n_dirs;
      double* gcol = A.g + d0;
      const double* acol = A.abuf + d0;
      uint8_t* ract = A.rowact + cb;

      if (threadIdx.x == 0) {
        for (uint32_t p = 0; p < n_pieces && p < kStages; ++p) {
          const int st = static_cast<int>((pc + p) % kStages);
          const int64_t off = static_cast<int64_t>(p) * kPieceRecs;
          const int64_t left = n_recs - off;
          const uint32_t cntb = static_cast<uint32_t>(left < kPieceRecs ? left : kPieceRecs) * static_cast<uint32_t>(sizeof(Rec));
          mbar_expect_tx(&ring.full[st], cntb);
          tma_bulk_g2s(&ring.buf[st][0], chunk + off, cntb, &ring.full[st]);
        }
      }
      LaneVec<V> acc;
      acc.zero();
      int32_t cur = -1;
      bool acc_valid = false, dirty = false;
      for (uint32_t p = 0; p < n_pieces; ++p) {
        const uint32_t q = pc + p;
        const int st = static_cast<int>(q % kStages);
        mbar_wait(&ring.full[st], (q / kStages) & 1u);
        const int64_t left = n_recs - static_cast<int64_t>(p) * kPieceRecs;
        const int n = static_cast<int>(left < kPieceRecs ? left : kPieceRecs);
        const Rec* r = &ring.buf[st][0];
        for (int t = threadIdx.x; t < n * nwarps; t += blockDim.x) {
          const int i = t / nwarps, w = t - i * nwarps;
          const int32_t slot = r[i].slot;
          uint8_t f = 0;
          if (cb0 + w < A.n_colblocks && slot >= 0)
            f = (r[i].kind == REC_MARK) ? __ldcg(&A.rowact[static_cast<int64_t>(slot) * ncb + cb0 + w])
                                        : __ldcg(&A.aflag[static_cast<int64_t>(slot) * ncb + cb0 + w]);
          pf[i][w] = f;
        }
        __syncthreads();
        if (live) {
          bool mine = false;
          for (int i = lane; i < n; i += 32) mine |= (r[i].kind == REC_OP && pf[i][warp] != 0);
          if (!__any_sync(0xffffffffu, mine)) {
            if (dirty) {
              lv_store_cg<V>(acc, gcol + static_cast<int64_t>(cur) * A.K);
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
#pragma unroll
              for (int u = 0; u < kBatch; ++u) {
                v[u].zero();
                if (i + u < n && pf[i + u][warp]) {
                  const int32_t slot = r[i + u].slot;
                  lv_load_cg<V>(v[u], ((r[i + u].kind == REC_MARK) ? gcol : acol) + static_cast<int64_t>(slot) * A.K);
                }
              }
#pragma unroll
              for (int u = 0; u < kBatch; ++u) {
                if (i + u >= n) break;
                const Rec rc = r[i + u];
                if (rc.kind == REC_MARK) {
                  if (dirty) {
                    lv_store_cg<V>(acc, gcol + static_cast<int64_t>(cur) * A.K);
                    if (lane == 0) ract[static_cast<int64_t>(cur) * ncb] = 1;
                  }
                  cur = rc.slot;
                  acc = v[u];
                  acc_valid = true;
                  dirty = false;
                } else if (pf[i + u][warp]) {
                  if (!acc_valid) {
                    acc.zero();
                    if (__ldcg(&ract[static_cast<int64_t>(cur) * ncb])) lv_load_cg<V>(acc, gcol + static_cast<int64_t>(cur) * A.K);
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
        if (threadIdx.x == 0 && p + kStages < n_pieces) {
          const uint32_t pn = p + kStages;
          const int stn = static_cast<int>((pc + pn) % kStages);
          const int64_t off = static_cast<int64_t>(pn) * kPieceRecs;
          const int64_t lf = n_recs - off;
          const uint32_t cntb = static_cast<uint32_t>(lf < kPieceRecs ? lf : kPieceRecs) * static_cast<uint32_t>(sizeof(Rec));
          mbar_expect_tx(&ring.full[stn], cntb);
          tma_bulk_g2s(&ring.buf[stn][0], chunk + off, cntb, &ring.full[stn]);
        }
      }
      if (live && dirty) {
        lv_store_cg<V>(acc, gcol + static_cast<int64_t>(cur) * A.K);
        if (lane == 0) ract[static_cast<int64_t>(cur) * ncb] = 1;
      }
      pc += n_pieces;
    }
    grid.sync();
  }
}

}  // namespace adept_b200
