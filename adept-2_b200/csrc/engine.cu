// engine.cu -- libadept_cuda: context, tape upload, schedule construction, replay drivers
// and the C-ABI of include/adept_cuda.h.  Device code lives in *_kernels.cuh.
//
// Drop-in boundary: the entry points below stand in for the bodies of
//   Stack::jacobian_forward / jacobian_reverse   (reference adept/jacobian.cpp:127-647)
//   Stack::compute_tangent_linear / compute_adjoint (reference adept/Stack.cpp:139-190)
// Host code here is plumbing (memory, streams, NCCL); there is no CPU replay path.
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/adept_cuda.h"
#include "common.cuh"
#include "replay_kernels.cuh"
#include "persistent_kernels.cuh"
#include "fused_persistent_kernels.cuh"
#include "smallg_kernels.cuh"
#include "sweep1_kernels.cuh"
#include "tape_kernels.cuh"

using namespace adept_b200;

namespace {

// Tuning constants (measured on B200)
constexpr int64_t kChunkRecs = 64;                    // records per level chunk = one CTA's share of a step (best of 16..256 on Toon nx=4096)
constexpr int64_t kLongThreshold = 1024;              // records; longer statements get cooperative kernels in single-direction sweeps
constexpr int64_t kMinLevelStatements = 4096;         // shorter tapes are replayed in tape order (no schedule is built)
constexpr int64_t kDefaultWindow = 65536;             // statements per dependency-analysis window
constexpr int64_t kHugeStatement = 65536;             // operands; a longer statement is an analysis window of its own
constexpr int64_t kMaxStepWidth = int64_t(1) << 20;   // statements per step (wider steps are split; bounds the reverse a-snapshot)
constexpr double kWideStepCtas = 4096.0;              // average CTAs per step above which one CUDA stream suffices
constexpr size_t kStageBytes = size_t(32) << 20;      // pinned staging buffer for pageable uploads (two of them)
constexpr size_t kPoolKeepBytes = size_t(8) << 30;    // analysis temporaries kept between uploads up to this size
constexpr int64_t kTailItems = 4;                     // (chunks x column groups) up to which a step rides along as a tail
constexpr int kFwdCbsDefault = 16;                    // forward_balanced_kernel: column blocks per CTA (kFwdCbs = 32 is the limit)
constexpr size_t kPoolSlots = 48;                     // analysis temporaries (Shard::pool)

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct LongStmt {
  int64_t j;      // position in the schedule
  int32_t level;
  int64_t b, e;   // record range in the level-major forward stream
};

struct Shard {
  int dev = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  ncclComm_t comm = nullptr;
  int sm_count = 148;
  // pinned staging ring for pageable uploads
  void* stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  // the tape as recorded
  Buf stmt, mult, idx;
  int64_t n_stmt = 0, n_ops = 0, max_gradient = 0;
  int64_t n_ops_recorded = 0;   // before "remove_null" compaction
  // the multiplier array may still be on its way (own copy stream) while the level analysis, which reads only
  // statement_ and index_, already runs: mult_pending until the shard's stream has waited for mult_ev
  cudaStream_t mult_stream = nullptr;
  cudaEvent_t mult_ev = nullptr;
  bool mult_pending = false;
  bool have_tape = false;       // the recorded arrays are on this device
  bool have_schedule = false;   // build_schedule has run for them (validation, level analysis, forward stream)
  bool have_tape_streams = false;
  bool has_nonfinite = false;  // some multiplier is inf/NaN: forward block-sparsity is then off
  // tape-order streams
  Buf fwd_t, rev_t, one_chunk;
  int64_t R = 0;
  // level schedule
  bool have_levels = false;
  Buf level, fwd_l, fwd_chunk, foff, sched_lhs;   // level-major forward stream and its tables
  Buf tgt, tchunk, goff;                           // reverse target stream, its chunks and groups (built on first need)
  Buf pos_in_sched;                                // schedule position of every statement (input of the target stream)
  bool have_target = false;
  // fused reverse stream (versioned rows, one kernel per step; built on first need)
  bool have_fused = false;
  Buf ftgt, fchunk, fpar, out_rows, done, trace;
  Buf d_step_fchunk, pcnt, pabort;    // resident fused kernel: device copy of step_fchunk, per-step counters, abort word
  bool persist_checked = false;       // a resident fused kernel ran since the last shard_finish: look at pabort
  int64_t max_step_fchunks = 0;       // widest step of the fused stream, in chunks
  int64_t max_fchunk_recs = 0;        // longest chunk of the fused stream, in records (> kPieceRecs: streamed in pieces)
  int64_t n_fchunks = 0, R_f = 0, chunk_recs = 0;
  std::vector<int64_t> step_fchunk;   // [n_levels+2] first fused chunk of each step
  Buf d_level_first, d_step_tchunk, d_step_group;  // device copies of the step tables (persistent / small-G kernels)
  Buf sg_fwd_info, sg_rev_info;                    // per-step {items, records} ranges for the small-G kernel
  int32_t n_levels = 0;                            // number of steps
  int64_t n_chunks = 0, n_tchunks = 0, n_groups = 0, R_t = 0, max_step_width = 0;
  std::vector<int64_t> level_chunk;   // [n_levels+2] first forward chunk of each step
  std::vector<int64_t> level_first;   // [n_levels+2] first schedule position of each step
  std::vector<int64_t> step_group;    // [n_levels+2] first target group of each step
  std::vector<int64_t> step_tchunk;   // [n_levels+2] first target chunk of each step
  std::vector<int64_t> step_recs;       // [n_levels+2] forward records (operations + statements) of each step
  std::vector<uint8_t> level_long;      // [n_levels+2] the step holds a chunk that may exceed one staged piece
  std::vector<std::string> prof_log;    // option "profile_every": one line per launch handed to the profiler
  std::vector<LongStmt> long_stmts;     // sorted by schedule position (hence by step)
  Buf long_tab;                         // device copy: {b, e} record range of every long statement, same order
  // workspace
  Buf g, abuf, rowact, aflag, seed_slots, out_slots, out, g1, a1, scratch, agree_buf;
  // ensemble pipeline (adept_cuda_jacobian_batch): two buffer sets, copy-in and copy-out streams
  Buf bg[2], bmult[2], bout[2];
  cudaStream_t bcopy = nullptr, bback = nullptr;
  cudaEvent_t b_in[2] = {nullptr, nullptr}, b_done[2] = {nullptr, nullptr}, b_out[2] = {nullptr, nullptr};
  // optional per-launch timing of the dominant replay kernel (option "time_kernels")
  std::vector<cudaEvent_t> kev;
  size_t kev_used = 0;
  // the replay kernels of EVERY pass of a Jacobian are bracketed by an event pair on the shard's stream: their sum
  // is "sweep_ms", measured inside whatever region the caller times (PDL and streams as in production)
  std::vector<cudaEvent_t> pev;
  size_t pev_used = 0;
  double sweep_launches = 0;   // launches of the replay kernel(s) between those events, last Jacobian
  std::vector<Buf> pool;  // analysis temporaries (build_schedule)
  std::vector<cudaStream_t> aux;      // extra streams: independent direction ranges of one sweep
  std::vector<cudaEvent_t> aux_ev;
  cudaEvent_t fork_ev = nullptr;
  // per-shard bookkeeping (shards may be driven by worker threads)
  std::string err;
  double launches = 0;
  std::map<std::string, double> stat;
};

}  // namespace

namespace {
// Checkpoint-chained replay (reference test/test_checkpoint.cpp:166-225): a sequence of tapes, each
// swept ONCE in reverse, the gradients at block k's inputs becoming the seeds at block k+1's outputs.  The caller
// re-records into the same host arrays right after a push, so a push copies the block into one of two pinned host
// slots and returns; a worker thread moves it to one of two DEVICE tape slots (copy stream), builds the schedule,
// seeds from the device-resident hand-off vector, sweeps and gathers the next hand-off -- while the caller records
// the next block.  The upload of block k+1 is issued before the worker waits for the sweep of block k.
struct ChainBlock {
  int slot = 0;
  int64_t n_stmt = 0, n_ops = 0, max_gradient = 0, n_seed = 0, n_out = 0;
  bool seed_from_handoff = false;
  bool uploaded = false;   // its H2D has already been issued (pipelined behind the previous block's sweep)
};
struct ChainHostSlot {
  void* p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // stmt, mult, idx, seed slots, seed values, out slots
  size_t cap[6] = {0, 0, 0, 0, 0, 0};
  bool busy = false;
};
struct Chain {
  bool active = false;
  int64_t n_handoff = 0;
  Buf handoff, owner;
  Buf dtape[2][3], dseed_slots[2], dseed_vals[2], dout_slots[2];
  ChainHostSlot hs[2];
  cudaStream_t copy = nullptr;
  cudaEvent_t h2d_done[2] = {nullptr, nullptr};
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<ChainBlock> q;
  bool stop = false;
  int rc = 0;
  std::string err;
  int64_t blocks = 0;
  double busy_ms = 0, sweep_ms = 0, analysis_ms = 0;
};
}  // namespace

struct adept_cuda_ctx {
  std::vector<Shard> sh;
  Chain chain;
  int world = 1;       // total ranks (devices over all processes)
  int rank0 = 0;       // global rank of sh[0]
  bool multi_process = false;
  uint64_t epoch = 0;
  // options
  int opt_schedule = 0, opt_ref_block = 2, opt_warps = 8;
  int64_t opt_pass_dirs = 0, opt_workspace_bytes = 0, opt_window = 0;
  int opt_time_kernels = 0;
  int opt_persistent = 0;  // 1: reverse level sweep as one cooperative kernel (measured slower than two launches per step)
  int64_t opt_chunk_recs = 0;  // records per level chunk (0 = kChunkRecs)
  int opt_pdl = 1;      // programmatic dependent launch between the per-step kernels of a level sweep
  int opt_streams = 4;  // CUDA streams over which the direction ranges of a level-scheduled reverse sweep are spread
  int opt_smallg = 1;   // use the shared-memory-resident kernel when the gradient list fits
  int opt_fused = 1;    // level-scheduled reverse sweep as ONE kernel per step over versioned rows (0: snapshot + gather)
  int opt_cbs = 0;      // fused kernel: column blocks per CTA (0 = automatic)
  int opt_tail = 1;     // fused kernel: steps of a chunk or two ride along with the previous step's launch
  int opt_fused_persistent = 0;    // fused reverse sweep as ONE resident kernel with per-step participant counters
                                   // (fused_persistent_kernels.cuh): 1 always, 0 never, -1 when the steps are narrow
  int opt_auto_split = 1;          // 1: few directions on the device -> the fused reverse kernels share a chunk between two CTAs
  int opt_split = 0;               // fused reverse kernels: CTAs per (chunk, column group) (0 = automatic, 1..16)
  int opt_sub = 0;                 // ... and records per sub-task (0 = automatic, 4..256)
  int64_t opt_spin_limit = int64_t(1) << 24;   // polls after which a waiting CTA of that kernel gives up (error, not a hang)
  int opt_fwd_v4 = 0;         // balanced forward kernel with FOUR doubles per lane (1024-byte row segments per warp)
  int opt_overlap_upload = 1; // multipliers in page-locked memory are uploaded on their own stream while the level analysis runs
  int opt_coop_pack = 1;      // window packing as one cooperative kernel (0: three launches per window)
  int opt_batch_chunks = 4;   // ensemble entry point: chunks per device through the copy-in / sweep / copy-out pipeline
  int opt_wide_streams = 0;   // > 0: this many streams for the forward level sweep even when its steps are wide (experiments)
  int opt_fwd_balanced = 1;   // forward level sweep: balanced CTA shape (forward_balanced_kernel); 0: forward_level_kernel only
  int opt_remove_null = 0;    // 1: drop zero-multiplier operations on upload (the reference's ADEPT_REMOVE_NULL_STATEMENTS build)
  int opt_profile_every = 0;  // > 0: cudaProfilerStart/Stop around every Nth launch of a level sweep (ncu --profile-from-start off)
  // stats
  std::map<std::string, double> stat;
  std::string err;
  double total_launches() const {
    double n = 0;
    for (const Shard& s : sh) n += s.launches;
    return n;
  }
};

namespace {

// Errors are written to whatever std::string* named `errp` is in scope: the context's in the
// C-ABI functions, the shard's own inside per-shard work (which may run on worker threads).
int fail(std::string* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) *e = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(errp, ADEPT_CUDA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                             \
  } while (0)
#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess)                                                                         \
      return fail(errp, ADEPT_CUDA_ERR_NCCL, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(r_), \
                  __FILE__, __LINE__);                                                             \
  } while (0)
// inside ncclGroupStart/ncclGroupEnd: remember the first failure, keep going, so that the group is always closed
#define NKG(var, call)                                   \
  do {                                                   \
    ncclResult_t r_ = (call);                            \
    if (r_ != ncclSuccess && (var) == ncclSuccess) (var) = r_; \
  } while (0)
#define RET(call)                \
  do {                           \
    int rc_ = (call);            \
    if (rc_) return rc_;         \
  } while (0)

int ensure(std::string* errp, Buf& b, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (b.bytes >= bytes) return 0;
  if (b.p) { CK(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return fail(errp, ADEPT_CUDA_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  }
  b.bytes = bytes;
  return 0;
}
void release(Buf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}

inline int grid_for(int64_t n, int block, int sm_count) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = int64_t(sm_count) * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return int(g);
}

int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
  int b = 1;
  while ((int64_t(1) << b) < n && b < 32) ++b;
  return b;
}

// H2D of one array: pinned sources are DMA'd directly; pageable ones go through a two-slot
// pinned ring so the host memcpy of piece k+1 overlaps the DMA of piece k.
int upload(Shard& s, void* dst, const void* src, size_t bytes) {
  std::string* errp = &s.err;
  if (bytes == 0) return 0;
  cudaPointerAttributes at;
  bool pinned = false;
  if (cudaPointerGetAttributes(&at, src) == cudaSuccess) pinned = (at.type == cudaMemoryTypeHost);
  else cudaGetLastError();
  if (pinned) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s.stream));
    return 0;
  }
  for (int k = 0; k < 2; ++k)
    if (!s.stage[k]) {
      CK(cudaMallocHost(&s.stage[k], kStageBytes));
      CK(cudaEventCreateWithFlags(&s.stage_ev[k], cudaEventDisableTiming));
    }
  size_t off = 0;
  int k = 0;
  while (off < bytes) {
    const size_t n = std::min(kStageBytes, bytes - off);
    CK(cudaEventSynchronize(s.stage_ev[k]));
    {  // host copy into the pinned slot, split over a few threads (one core moves ~10 GB/s, PCIe 5 takes 50+)
      char* dstp = static_cast<char*>(s.stage[k]);
      const char* srcp = static_cast<const char*>(src) + off;
      const int nt = n >= (size_t(8) << 20) ? 4 : 1;
      const size_t part = (n + nt - 1) / nt;
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) {
        const size_t o = size_t(t) * part;
        if (o < n) th.emplace_back([dstp, srcp, o, part, n]() { std::memcpy(dstp + o, srcp + o, std::min(part, n - o)); });
      }
      std::memcpy(dstp, srcp, std::min(part, n));
      for (std::thread& t : th) t.join();
    }
    CK(cudaMemcpyAsync(static_cast<char*>(dst) + off, s.stage[k], n, cudaMemcpyHostToDevice, s.stream));
    CK(cudaEventRecord(s.stage_ev[k], s.stream));
    off += n;
    k ^= 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// Schedule construction on one shard (tape already on the device).
// ---------------------------------------------------------------------------------------
int sort_pairs(std::string* errp, Shard& s, Buf& k0, Buf& k1, Buf& v0, Buf& v1, Buf& tmp, int64_t n, int nbits) {
  cub::DoubleBuffer<uint32_t> dk(k0.as<uint32_t>(), k1.as<uint32_t>());
  cub::DoubleBuffer<uint32_t> dv(v0.as<uint32_t>(), v1.as<uint32_t>());
  size_t tb = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, n, 0, nbits, s.stream));
  RET(ensure(errp, tmp, tb));
  CK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, n, 0, nbits, s.stream));
  if (dk.Current() != k0.as<uint32_t>()) std::swap(k0, k1);
  if (dv.Current() != v0.as<uint32_t>()) std::swap(v0, v1);
  return 0;  // sorted keys in k0, values in v0
}

template <class T>
int exclusive_sum(std::string* errp, Shard& s, Buf& tmp, const T* in, T* out, int64_t n) {
  size_t tb = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, s.stream));
  RET(ensure(errp, tmp, tb));
  CK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, n, s.stream));
  return 0;
}
template <class T>
int inclusive_sum(std::string* errp, Shard& s, Buf& tmp, const T* in, T* out, int64_t n) {
  size_t tb = 0;
  CK(cub::DeviceScan::InclusiveSum(nullptr, tb, in, out, n, s.stream));
  RET(ensure(errp, tmp, tb));
  CK(cub::DeviceScan::InclusiveSum(tmp.p, tb, in, out, n, s.stream));
  return 0;
}

// Tape-order record streams (one chunk = the whole tape), built on first need: the tape-ordered kernels use
// them; a tape that is replayed through its level schedule never pays for them (16 bytes x 2 per record).
int ensure_tape_streams(Shard& s) {
  std::string* errp = &s.err;
  if (s.have_tape_streams) return 0;
  const int64_t S = s.n_stmt - 1, O = s.n_ops, R = O + S;
  if (R <= 0) { s.have_tape_streams = true; return 0; }
  CK(cudaSetDevice(s.dev));
  RET(ensure(errp, s.fwd_t, size_t(R) * sizeof(Rec)));
  RET(ensure(errp, s.rev_t, size_t(R) * sizeof(Rec)));
  build_streams_kernel<<<grid_for(std::max(O, S), 256, s.sm_count), 256, 0, s.stream>>>(
      s.stmt.as<int2>(), s.n_stmt, s.mult.as<double>(), s.idx.as<int32_t>(), O, s.fwd_t.as<Rec>(),
      s.rev_t.as<Rec>(), nullptr);
  s.launches += 1;
  CK(cudaGetLastError());
  s.have_tape_streams = true;
  return 0;
}

// The analysis temporaries (tens of GB for a 1e8-statement tape: the per-window hash tables, sort buffers) survive the
// call so that the next upload of a similar tape pays no cudaMalloc/cudaFree (measured: freeing and re-allocating them
// costs more than the analysis itself).  They are given back when the device runs short of memory -- here, when less
// than a quarter of it is left, and by the replay drivers when a Jacobian workspace would otherwise need more passes
// (release_pool).
size_t pool_bytes(const Shard& s) {
  size_t total = 0;
  for (const Buf& b : s.pool) total += b.bytes;
  return total;
}
void release_pool(Shard& s) {
  for (Buf& b : s.pool) release(b);
}
struct PoolGuard {
  Shard& s;
  ~PoolGuard() {
    if (pool_bytes(s) <= kPoolKeepBytes) return;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < total_b / 4) release_pool(s);
  }
};

// The shard's stream waits for an overlapped multiplier upload (adept_cuda_upload_tape), then the multipliers are
// screened: non-finite values (forward block sparsity is then off) and the fractions equal to 0, +1, -1.
int settle_multipliers(const adept_cuda_ctx* c, Shard& s) {
  std::string* errp = &s.err;
  if (s.mult_pending) {
    CK(cudaStreamWaitEvent(s.stream, s.mult_ev, 0));
    s.mult_pending = false;
  }
  const int64_t O = s.n_ops;
  RET(ensure(errp, s.scratch, 64));
  CK(cudaMemsetAsync(s.scratch.p, 0, 64, s.stream));
  if (O > 0) {
    validate_kernel<<<grid_for(O, 256, s.sm_count), 256, 0, s.stream>>>(
        s.stmt.as<int2>(), 0, s.idx.as<int32_t>(), s.mult.as<double>(), O, s.max_gradient, s.scratch.as<int>(),
        reinterpret_cast<unsigned long long*>(s.scratch.as<char>() + 8));
    s.launches += 1;
  }
  struct { int flags; int pad; unsigned long long n0, n1, nm1; } vh;
  CK(cudaMemcpyAsync(&vh, s.scratch.p, sizeof vh, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  s.has_nonfinite = (vh.flags & 2) != 0;
  // Stack::fraction_multipliers_equal_to (include/adept/Stack.h:707-715) for 0, +1, -1, of the tape as recorded
  s.stat["frac_mult_zero"] = O > 0 ? double(vh.n0) / double(O) : 0.0;
  s.stat["frac_mult_one"] = O > 0 ? double(vh.n1) / double(O) : 0.0;
  s.stat["frac_mult_minus_one"] = O > 0 ? double(vh.nm1) / double(O) : 0.0;
  s.stat["n_zero_multipliers"] = double(vh.n0);
  return 0;
}

int build_schedule(const adept_cuda_ctx* c, Shard& s) {
  std::string* errp = &s.err;
  CK(cudaSetDevice(s.dev));
  const int64_t S = s.n_stmt - 1, O = s.n_ops, R = O + S;
  s.R = R;
  s.have_schedule = false;
  s.have_levels = false;
  s.have_tape_streams = false;
  s.have_target = false;
  s.have_fused = false;
  s.n_levels = 0;
  s.n_chunks = 0;
  s.n_tchunks = 0;
  s.n_fchunks = 0;
  s.n_groups = 0;
  s.max_step_width = 0;
  s.long_stmts.clear();
  if (s.n_stmt < 1) return fail(errp, ADEPT_CUDA_ERR_ARG, "tape has no sentinel statement");
  if (R >= (int64_t(1) << 31) - 1)
    return fail(errp, ADEPT_CUDA_ERR_LIMIT, "tape too long: n_ops + n_statements = %lld >= 2^31", (long long)R);
  const int T = 256;
  auto tp = std::chrono::steady_clock::now();
  // 1. validate the STRUCTURE (a bad slot would otherwise be a wild device access).  The multipliers may still be on
  //    their way (overlapped upload): they are screened later, just before the first kernel that reads them -- unless
  //    compaction needs them now.
  RET(ensure(errp, s.scratch, 64));
  CK(cudaMemsetAsync(s.scratch.p, 0, 64, s.stream));
  validate_kernel<<<grid_for(std::max(O, s.n_stmt), T, s.sm_count), T, 0, s.stream>>>(
      s.stmt.as<int2>(), s.n_stmt, s.idx.as<int32_t>(), nullptr, O, s.max_gradient, s.scratch.as<int>(), nullptr);
  s.launches += 1;
  int flags = 0;
  CK(cudaMemcpyAsync(&flags, s.scratch.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  if (flags & 1)
    return fail(errp, ADEPT_CUDA_ERR_RANGE,
                "malformed tape: a slot index is outside [0,max_gradient) or statement ends are not monotone");
  s.stat["ops_removed"] = 0.0;
  s.stat["n_ops_recorded"] = double(O);
  bool mult_settled = false;
  if (c->opt_remove_null) {
    RET(settle_multipliers(c, s));
    mult_settled = true;
  }
  struct { unsigned long long n0; } vh = {mult_settled ? (unsigned long long)s.stat["n_zero_multipliers"] : 0ull};
  // 1b. tape compaction (option "remove_null"): drop the zero-multiplier operations, as a reference built with
  //     -DADEPT_REMOVE_NULL_STATEMENTS would have done while recording; everything below sees the compacted tape
  if (c->opt_remove_null && vh.n0 > 0) {
    s.pool.resize(kPoolSlots);
    Buf &keep = s.pool[41], &pos = s.pool[42], &tmpc = s.pool[9], &m2 = s.pool[43], &i2 = s.pool[44];
    const int64_t O2 = O - int64_t(vh.n0);
    RET(ensure(errp, keep, size_t(O + 1) * 4));
    RET(ensure(errp, pos, size_t(O + 1) * 4));
    RET(ensure(errp, m2, size_t(std::max<int64_t>(O2, 1)) * 8));
    RET(ensure(errp, i2, size_t(std::max<int64_t>(O2, 1)) * 4));
    CK(cudaMemsetAsync(keep.as<int32_t>() + O, 0, 4, s.stream));
    compact_keep_kernel<<<grid_for(O, T, s.sm_count), T, 0, s.stream>>>(s.mult.as<double>(), O, keep.as<int32_t>());
    RET(exclusive_sum(errp, s, tmpc, keep.as<int32_t>(), pos.as<int32_t>(), O + 1));
    compact_ops_kernel<<<grid_for(O, T, s.sm_count), T, 0, s.stream>>>(s.mult.as<double>(), s.idx.as<int32_t>(), O, keep.as<int32_t>(),
                                                                      pos.as<int32_t>(), m2.as<double>(), i2.as<int32_t>());
    compact_stmt_kernel<<<grid_for(s.n_stmt, T, s.sm_count), T, 0, s.stream>>>(s.stmt.as<int2>(), s.n_stmt, pos.as<int32_t>());
    s.launches += 3;
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaGetLastError());
    std::swap(s.mult, m2);     // the recorded arrays go back to the pool
    std::swap(s.idx, i2);
    s.n_ops = O2;
    const double f0 = s.stat["frac_mult_zero"], f1 = s.stat["frac_mult_one"], fm1 = s.stat["frac_mult_minus_one"];
    const int rc = build_schedule(c, s);   // once more on the compacted tape (no zero multipliers left: one level deep)
    s.stat["frac_mult_zero"] = f0;         // the statistics describe the tape as recorded
    s.stat["frac_mult_one"] = f1;
    s.stat["frac_mult_minus_one"] = fm1;
    s.stat["ops_removed"] = double(vh.n0);
    s.stat["n_ops_recorded"] = double(O);
    return rc;
  }
  s.stat["an_validate_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp).count();
  tp = std::chrono::steady_clock::now();
  // 2. tape-order streams
  const int64_t one_chunk_host[2] = {0, R};
  RET(ensure(errp, s.one_chunk, sizeof one_chunk_host));
  CK(cudaMemcpyAsync(s.one_chunk.p, one_chunk_host, sizeof one_chunk_host, cudaMemcpyHostToDevice, s.stream));
  CK(cudaStreamSynchronize(s.stream));  // one_chunk_host is a stack variable
  if (S == 0) {
    if (!mult_settled) RET(settle_multipliers(c, s));
    s.have_schedule = true;
    return 0;
  }
  s.have_tape_streams = false;
  const bool want_levels = (c->opt_schedule != 1) && (S >= kMinLevelStatements || c->opt_schedule == 2);
  if (!want_levels) {
    if (!mult_settled) RET(settle_multipliers(c, s));
    RET(ensure_tape_streams(s));
    s.have_schedule = true;
    return 0;
  }

  // Temporaries live in a per-shard pool that survives the call (re-uploads of similar tapes then cost
  // no cudaMalloc/cudaFree, which dominated and jittered the analysis time); a pool above
  // kPoolKeepBytes is released on exit.
  PoolGuard pool_guard{s};
  s.pool.resize(kPoolSlots);
  Buf &cap = s.pool[0], &tab_off = s.pool[1], &tables = s.pool[2], &wdepth = s.pool[3], &wbase = s.pool[4],
      &k0 = s.pool[5], &k1 = s.pool[6], &v0 = s.pool[7], &v1 = s.pool[8], &tmp = s.pool[9], &nrec = s.pool[10],
      &lfirst = s.pool[11], &flag = s.pool[12], &cid = s.pool[13], &lchunk = s.pool[14], &lo = s.pool[22], &lc = s.pool[23];
  Buf& pos_in_sched = s.pos_in_sched;

  auto lap = [&](const char* key) {
    cudaStreamSynchronize(s.stream);
    auto now = std::chrono::steady_clock::now();
    s.stat[key] = std::chrono::duration<double, std::milli>(now - tp).count();
    tp = now;
  };
  // 3. dependency analysis: one warp per window of statements (tape_kernels.cuh)
  const int64_t W = c->opt_window > 0 ? c->opt_window : kDefaultWindow;
  Buf &wflag = s.pool[24], &wincl = s.pool[25], &win_lo = s.pool[26];
  RET(ensure(errp, wflag, size_t(s.n_stmt) * 4));
  RET(ensure(errp, wincl, size_t(s.n_stmt) * 4));
  window_flags_kernel<<<grid_for(s.n_stmt, T, s.sm_count), T, 0, s.stream>>>(s.stmt.as<int2>(), s.n_stmt, W,
                                                                            kHugeStatement, wflag.as<int32_t>());
  s.launches += 1;
  RET(inclusive_sum(errp, s, tmp, wflag.as<int32_t>(), wincl.as<int32_t>(), s.n_stmt));
  int32_t n_windows32 = 0;
  CK(cudaMemcpyAsync(&n_windows32, wincl.as<int32_t>() + S, 4, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  const int64_t n_windows = n_windows32;
  RET(ensure(errp, win_lo, size_t(n_windows + 1) * 8));
  window_starts_kernel<<<grid_for(s.n_stmt, T, s.sm_count), T, 0, s.stream>>>(wflag.as<int32_t>(), wincl.as<int32_t>(),
                                                                             s.n_stmt, n_windows, win_lo.as<int64_t>());
  s.launches += 1;
  RET(ensure(errp, cap, size_t(n_windows + 1) * 8));
  RET(ensure(errp, tab_off, size_t(n_windows + 1) * 8));
  // small gradient lists: direct-indexed table in shared memory, dense per-window state for the packing pass
  const size_t dense_tab = size_t(s.max_gradient) * 8;
  const size_t dense_smem = dense_tab + kWinStageBytes;   // table + cp.async staging of statements and operands
  const bool dense = dense_tab <= (size_t(160) << 10) && size_t(n_windows) * dense_tab <= (size_t(2) << 30);
  Buf& wstate = s.pool[33];
  Buf& oplev = s.pool[46];   // hash variant: local step of every operation's statement, for the packing pass
  RET(ensure(errp, s.level, size_t(s.n_stmt) * 4));
  RET(ensure(errp, wdepth, size_t(n_windows + 1) * 4));
  RET(ensure(errp, wbase, size_t(n_windows + 1) * 4));
  CK(cudaMemsetAsync(wdepth.p, 0, size_t(n_windows + 1) * 4, s.stream));
  RET(ensure(errp, tab_off, size_t(n_windows + 1) * 8));
  if (dense) {
    RET(ensure(errp, wstate, size_t(n_windows) * dense_tab));
    CK(cudaFuncSetAttribute(window_levels_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dense_smem)));
    window_levels_smem_kernel<<<unsigned(n_windows), 32, dense_smem, s.stream>>>(
        s.stmt.as<int2>(), s.idx.as<int32_t>(), win_lo.as<int64_t>(), n_windows, s.max_gradient, wstate.as<int2>(),
        s.level.as<int32_t>(), wdepth.as<int32_t>());
    s.launches += 1;
    CK(cudaMemsetAsync(tab_off.p, 0, size_t(n_windows + 1) * 8, s.stream));
  } else {
    RET(ensure(errp, cap, size_t(n_windows + 1) * 8));
    window_caps_kernel<<<unsigned((n_windows + 1 + T - 1) / T), T, 0, s.stream>>>(s.stmt.as<int2>(), win_lo.as<int64_t>(),
                                                                                  n_windows, cap.as<int64_t>());
    s.launches += 1;
    RET(exclusive_sum(errp, s, tmp, cap.as<int64_t>(), tab_off.as<int64_t>(), n_windows + 1));
    int64_t total_entries = 0;
    CK(cudaMemcpyAsync(&total_entries, tab_off.as<int64_t>() + n_windows, 8, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    RET(ensure(errp, tables, size_t(std::max<int64_t>(total_entries, 1)) * sizeof(SlotEntry)));
    CK(cudaMemsetAsync(tables.p, 0xFF, size_t(total_entries) * sizeof(SlotEntry), s.stream));
    RET(ensure(errp, oplev, size_t(std::max<int64_t>(O, 1)) * 4));
    window_levels_kernel<<<unsigned((n_windows * 32 + 127) / 128), 128, 0, s.stream>>>(
        s.stmt.as<int2>(), s.n_stmt, s.idx.as<int32_t>(), win_lo.as<int64_t>(), n_windows, tables.as<SlotEntry>(),
        tab_off.as<int64_t>(), s.level.as<int32_t>(), wdepth.as<int32_t>(), oplev.as<int32_t>());
    s.launches += 1;
  }
  lap("an_window_ms");
  // 3b. window packing: local steps -> global steps against a per-slot summary of all earlier windows
  int32_t n_levels = 0;
  Buf &gw = s.pool[27], &gr = s.pool[28], &wopl = s.pool[29], &need = s.pool[31], &gmap = s.pool[32];
  {
    RET(exclusive_sum(errp, s, tmp, wdepth.as<int32_t>(), wbase.as<int32_t>(), n_windows + 1));
    std::vector<int64_t> h_lo(n_windows + 1), h_tab(n_windows + 1), h_op(n_windows + 1);
    std::vector<int32_t> h_base(n_windows + 1);
    RET(ensure(errp, wopl, size_t(n_windows + 1) * 8));
    gather_ends_kernel<<<unsigned((n_windows + 1 + T - 1) / T), T, 0, s.stream>>>(s.stmt.as<int2>(), win_lo.as<int64_t>(),
                                                                                 n_windows + 1, wopl.as<int64_t>());
    s.launches += 1;
    CK(cudaMemcpyAsync(h_lo.data(), win_lo.p, size_t(n_windows + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(h_tab.data(), tab_off.p, size_t(n_windows + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(h_op.data(), wopl.p, size_t(n_windows + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(h_base.data(), wbase.p, size_t(n_windows + 1) * 4, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    const int64_t total_depth = h_base[n_windows];
    RET(ensure(errp, gw, size_t(s.max_gradient) * 4));
    RET(ensure(errp, gr, size_t(s.max_gradient) * 4));
    RET(ensure(errp, need, size_t(total_depth + 1) * 4));
    RET(ensure(errp, gmap, size_t(total_depth + 1) * 4));
    CK(cudaMemsetAsync(gw.p, 0, size_t(s.max_gradient) * 4, s.stream));
    CK(cudaMemsetAsync(gr.p, 0, size_t(s.max_gradient) * 4, s.stream));
    CK(cudaMemsetAsync(need.p, 0, size_t(total_depth + 1) * 4, s.stream));
    // the whole pass as one cooperative kernel (tape_kernels.cuh, window_pack_kernel); the per-window launches below
    // remain as the fallback when a cooperative launch is not possible (option "coop_pack" = 0 forces them: tests)
    bool packed = false;
    if (c->opt_coop_pack) {
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_pack_kernel, 256, 0) == cudaSuccess && per_sm >= 1) {
        PackArgs PA;
        PA.stmt = s.stmt.as<int2>();
        PA.idx = s.idx.as<int32_t>();
        PA.win_lo = win_lo.as<int64_t>();
        PA.win_op = wopl.as<int64_t>();
        PA.win_base = wbase.as<int32_t>();
        PA.tab_off = tab_off.as<int64_t>();
        PA.tables = tables.as<SlotEntry>();
        PA.wstate = wstate.as<int2>();
        PA.max_gradient = s.max_gradient;
        PA.n_windows = n_windows;
        PA.level = s.level.as<int32_t>();
        PA.gw = gw.as<int32_t>();
        PA.gr = gr.as<int32_t>();
        PA.need = need.as<int32_t>();
        PA.gmap = gmap.as<int32_t>();
        PA.oplev = dense ? nullptr : oplev.as<int32_t>();
        PA.dense = dense ? 1 : 0;
        void* args[] = {&PA};
        const unsigned grid = unsigned(s.sm_count * std::min(per_sm, 4));
        if (cudaLaunchCooperativeKernel((void*)window_pack_kernel, dim3(grid), dim3(256), args, 0, s.stream) == cudaSuccess) {
          packed = true;
          s.launches += 1;
        } else {
          cudaGetLastError();
        }
      } else {
        cudaGetLastError();
      }
    }
    for (int64_t w = 0; w < n_windows && !packed; ++w) {
      const int64_t lo = h_lo[w], hi = h_lo[w + 1] - 1, o0 = h_op[w], o1 = h_op[w + 1];
      const int32_t depth = h_base[w + 1] - h_base[w];
      int32_t* need_w = need.as<int32_t>() + h_base[w];
      int32_t* gmap_w = gmap.as<int32_t>() + h_base[w];
      const int64_t total = (o1 - o0) + (hi - lo + 1);
      // ~1200 elements per CTA: few enough CTAs for few global atomics, enough of them to fill the machine
      const int need_grid = int(std::max<int64_t>(1, std::min<int64_t>(int64_t(s.sm_count) * 4, (total + 1023) / 1024)));
      window_need_kernel<<<need_grid, T, 0, s.stream>>>(
          s.stmt.as<int2>(), s.idx.as<int32_t>(), lo, hi, o0, o1, s.level.as<int32_t>(), gw.as<int32_t>(), gr.as<int32_t>(),
          need_w, depth);
      window_map_kernel<<<1, 32, 0, s.stream>>>(need_w, depth, gmap_w);
      if (lo == hi) {
        window_update_single_ops_kernel<<<grid_for(std::max<int64_t>(o1 - o0, 1), T, s.sm_count), T, 0, s.stream>>>(
            s.idx.as<int32_t>(), o0, o1, gmap_w, gr.as<int32_t>());
        window_update_single_lhs_kernel<<<1, 1, 0, s.stream>>>(s.stmt.as<int2>(), lo, gmap_w, gw.as<int32_t>(), gr.as<int32_t>());
        s.launches += 4;
      } else {
        if (dense) {
          window_update_dense_kernel<<<grid_for(s.max_gradient, T, s.sm_count), T, 0, s.stream>>>(
              wstate.as<int2>() + w * s.max_gradient, s.max_gradient, gmap_w, gw.as<int32_t>(), gr.as<int32_t>());
        } else {
          const int64_t cap_w = h_tab[w + 1] - h_tab[w];
          window_update_kernel<<<grid_for(cap_w, T, s.sm_count), T, 0, s.stream>>>(tables.as<SlotEntry>() + h_tab[w], cap_w,
                                                                                   gmap_w, gw.as<int32_t>(), gr.as<int32_t>());
        }
        s.launches += 3;
      }
    }
    std::vector<int32_t> h_map(total_depth + 1);
    CK(cudaMemcpyAsync(h_map.data(), gmap.p, size_t(total_depth) * 4, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (int64_t w = 0; w < n_windows; ++w)
      if (h_base[w + 1] > h_base[w]) n_levels = std::max(n_levels, h_map[h_base[w + 1] - 1]);
  }
  global_steps_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(s.level.as<int32_t>(), s.n_stmt, wincl.as<int32_t>(),
                                                                      wbase.as<int32_t>(), gmap.as<int32_t>());
  s.launches += 1;
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaGetLastError());
  s.n_levels = n_levels;
  if (n_levels < 1) return fail(errp, ADEPT_CUDA_ERR_CUDA, "dependency analysis produced no steps");

  lap("an_pack_ms");
  // one CTA's share of a step: 64 records when steps are narrow (shorter critical path per launch), 128 when they
  // are wide (measured: Toon nx=4096 171 vs 196 ms; synthetic forward 537 vs 462 ms)
  const int64_t chunk_recs = c->opt_chunk_recs > 0 ? c->opt_chunk_recs : (R / std::max(1, n_levels) >= 32768 ? 2 * kChunkRecs : kChunkRecs);
  s.chunk_recs = chunk_recs;
  // 4. statements sorted by step (stable), record offsets, chunk table, level-major forward stream
  const int64_t nmax = std::max(S, O);
  RET(ensure(errp, k0, size_t(nmax) * 4));
  RET(ensure(errp, k1, size_t(nmax) * 4));
  RET(ensure(errp, v0, size_t(nmax) * 4));
  RET(ensure(errp, v1, size_t(nmax) * 4));
  level_keys_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(s.level.as<int32_t>(), S, k0.as<uint32_t>(),
                                                                    v0.as<uint32_t>());
  s.launches += 1;
  RET(sort_pairs(errp, s, k0, k1, v0, v1, tmp, S, bits_for(int64_t(n_levels) + 1)));
  // k0 = step of scheduled statement j, v0 = its statement id
  RET(ensure(errp, nrec, size_t(S + 1) * 8));
  RET(ensure(errp, s.foff, size_t(S + 1) * 8));
  RET(ensure(errp, lfirst, size_t(n_levels + 2) * 8));
  CK(cudaMemsetAsync(nrec.p, 0, size_t(S + 1) * 8, s.stream));
  sched_sizes_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(
      s.stmt.as<int2>(), v0.as<uint32_t>(), k0.as<uint32_t>(), S, n_levels, nrec.as<int64_t>(),
      lfirst.as<int64_t>());
  s.launches += 1;
  {  // steps wider than kMaxStepWidth are cut into sub-steps (bounds the reverse snapshot buffer)
    std::vector<int64_t> lf(size_t(n_levels) + 2);
    CK(cudaMemcpyAsync(lf.data(), lfirst.p, size_t(n_levels + 2) * 8, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    int64_t widest = 0;
    for (int32_t l = 1; l <= n_levels; ++l) widest = std::max(widest, lf[l + 1] - lf[l]);
    if (widest > kMaxStepWidth) {
      std::vector<int32_t> nb(size_t(n_levels) + 2, 0);
      int32_t next = 1;
      for (int32_t l = 1; l <= n_levels; ++l) {
        nb[l] = next;
        next += int32_t((lf[l + 1] - lf[l] + kMaxStepWidth - 1) / kMaxStepWidth);
      }
      Buf& nbd = s.pool[30];
      RET(ensure(errp, nbd, size_t(n_levels + 2) * 4));
      CK(cudaMemcpyAsync(nbd.p, nb.data(), size_t(n_levels + 2) * 4, cudaMemcpyHostToDevice, s.stream));
      split_steps_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(k0.as<uint32_t>(), v0.as<uint32_t>(), S,
                                                                         lfirst.as<int64_t>(), nbd.as<int32_t>(),
                                                                         kMaxStepWidth, s.level.as<int32_t>());
      s.launches += 1;
      CK(cudaStreamSynchronize(s.stream));  // nb is a host vector
      n_levels = next - 1;
      s.n_levels = n_levels;
      RET(ensure(errp, lfirst, size_t(n_levels + 2) * 8));
      sched_sizes_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(
          s.stmt.as<int2>(), v0.as<uint32_t>(), k0.as<uint32_t>(), S, n_levels, nrec.as<int64_t>(),
          lfirst.as<int64_t>());
      s.launches += 1;
    }
  }
  RET(exclusive_sum(errp, s, tmp, nrec.as<int64_t>(), s.foff.as<int64_t>(), S + 1));
  RET(ensure(errp, flag, size_t(nmax + 1) * 4));
  RET(ensure(errp, cid, size_t(S + 1) * 4));
  CK(cudaMemsetAsync(flag.p, 0, size_t(S + 1) * 4, s.stream));
  chunk_flags_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(s.foff.as<int64_t>(), k0.as<uint32_t>(), S,
                                                                     chunk_recs, flag.as<int32_t>());
  s.launches += 1;
  {  // steps whose chunks may be longer than one staged piece go to the streaming forward kernel
    Buf& ll = s.pool[45];
    RET(ensure(errp, ll, size_t(n_levels) + 2));
    CK(cudaMemsetAsync(ll.p, 0, size_t(n_levels) + 2, s.stream));
    mark_long_levels_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(s.foff.as<int64_t>(), k0.as<uint32_t>(), S,
                                                                            int64_t(kPieceRecs) - chunk_recs, ll.as<uint8_t>());
    s.launches += 1;
    s.level_long.assign(size_t(n_levels) + 2, 0);
    CK(cudaMemcpyAsync(s.level_long.data(), ll.p, size_t(n_levels) + 2, cudaMemcpyDeviceToHost, s.stream));
  }
  RET(exclusive_sum(errp, s, tmp, flag.as<int32_t>(), cid.as<int32_t>(), S + 1));
  int32_t n_chunks32 = 0;
  CK(cudaMemcpyAsync(&n_chunks32, cid.as<int32_t>() + S, 4, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  const int64_t n_chunks = n_chunks32;
  s.n_chunks = n_chunks;
  RET(ensure(errp, s.fwd_chunk, size_t(n_chunks + 1) * 8));
  RET(ensure(errp, lchunk, size_t(n_levels + 2) * 8));
  chunk_tables_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(
      s.foff.as<int64_t>(), flag.as<int32_t>(), cid.as<int32_t>(), S, R, n_chunks, lfirst.as<int64_t>(), n_levels,
      s.fwd_chunk.as<int64_t>(), nullptr, lchunk.as<int64_t>());
  s.launches += 1;
  {  // the balanced forward kernel stages a chunk in ONE piece: verify that every chunk of a step not flagged long fits
    RET(ensure(errp, s.scratch, 64));
    CK(cudaMemsetAsync(s.scratch.p, 0, 64, s.stream));
    max_short_chunk_kernel<<<grid_for(n_levels, T, s.sm_count), T, 0, s.stream>>>(
        lchunk.as<int64_t>(), s.fwd_chunk.as<int64_t>(), s.pool[45].as<uint8_t>(), n_levels,
        reinterpret_cast<unsigned long long*>(s.scratch.p));
    s.launches += 1;
    unsigned long long longest = 0;
    CK(cudaMemcpyAsync(&longest, s.scratch.p, 8, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    s.stat["max_short_chunk_recs"] = double(longest);
    if (longest > (unsigned long long)kPieceRecs)
      return fail(errp, ADEPT_CUDA_ERR_CUDA, "internal: a %llu-record chunk in a step not flagged long (staging holds %d)", longest,
                  kPieceRecs);
  }
  RET(ensure(errp, s.fwd_l, size_t(R) * sizeof(Rec)));
  if (!mult_settled) {   // the first reader of the multipliers: wait for an overlapped upload, screen them
    RET(settle_multipliers(c, s));
    mult_settled = true;
  }
  emit_level_streams_kernel<<<grid_for(R, T, s.sm_count), T, 0, s.stream>>>(
      s.stmt.as<int2>(), s.mult.as<double>(), s.idx.as<int32_t>(), v0.as<uint32_t>(), s.foff.as<int64_t>(), S, R,
      s.fwd_l.as<Rec>(), nullptr);
  s.launches += 1;
  RET(ensure(errp, s.sched_lhs, size_t(S) * 4));
  RET(ensure(errp, pos_in_sched, size_t(s.n_stmt) * 4));
  sched_lookup_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(s.stmt.as<int2>(), v0.as<uint32_t>(), S,
                                                                      s.sched_lhs.as<int32_t>(),
                                                                      pos_in_sched.as<int32_t>());
  s.launches += 1;
  s.level_chunk.assign(size_t(n_levels) + 2, 0);
  s.level_first.assign(size_t(n_levels) + 2, 0);
  CK(cudaMemcpyAsync(s.level_chunk.data(), lchunk.p, size_t(n_levels + 2) * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaMemcpyAsync(s.level_first.data(), lfirst.p, size_t(n_levels + 2) * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  for (int32_t l = 1; l <= n_levels; ++l)
    s.max_step_width = std::max(s.max_step_width, s.level_first[l + 1] - s.level_first[l]);

  lap("an_forward_ms");
  // 5. the reverse streams (two-kernel target stream, fused stream) are built on first need:
  //    ensure_target_stream / ensure_fused_stream
  RET(ensure(errp, s.d_level_first, size_t(n_levels + 2) * 8));
  CK(cudaMemcpyAsync(s.d_level_first.p, s.level_first.data(), size_t(n_levels + 2) * 8, cudaMemcpyHostToDevice, s.stream));
  RET(ensure(errp, s.sg_fwd_info, size_t(n_levels + 2) * 32));
  step_info_kernel<<<unsigned((n_levels + 2 + T - 1) / T), T, 0, s.stream>>>(s.d_level_first.as<int64_t>(), s.foff.as<int64_t>(),
                                                                           n_levels, s.sg_fwd_info.as<longlong4>());
  s.launches += 1;
  {
    std::vector<longlong4> info(size_t(n_levels) + 2);
    CK(cudaMemcpyAsync(info.data(), s.sg_fwd_info.p, info.size() * 32, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    s.step_recs.assign(size_t(n_levels) + 2, 0);
    for (int32_t l = 1; l <= n_levels; ++l) s.step_recs[l] = info[l].w - info[l].z;
  }
  CK(cudaStreamSynchronize(s.stream));
  // 6. long statements (rare): listed for the single-direction forward sweep.  The list grows to whatever the tape
  //    needs (a recorded dense mat-vec has one long statement per row); {j, b, e} triples come back in one copy.
  {
    int64_t cap_long = 4096;
    int nl = 0;
    std::vector<int64_t> trip;
    for (int attempt = 0; attempt < 2; ++attempt) {
      RET(ensure(errp, lo, size_t(cap_long) * 24));
      RET(ensure(errp, lc, 16));
      CK(cudaMemsetAsync(lc.p, 0, 16, s.stream));
      find_long_kernel<<<grid_for(S, T, s.sm_count), T, 0, s.stream>>>(s.foff.as<int64_t>(), S, kLongThreshold,
                                                                       lo.as<int64_t>(), lc.as<int>(), int(cap_long));
      s.launches += 1;
      CK(cudaMemcpyAsync(&nl, lc.p, 4, cudaMemcpyDeviceToHost, s.stream));
      CK(cudaStreamSynchronize(s.stream));
      if (nl <= cap_long) break;
      cap_long = nl;
    }
    trip.resize(size_t(nl) * 3);
    if (nl > 0) {
      CK(cudaMemcpyAsync(trip.data(), lo.p, size_t(nl) * 24, cudaMemcpyDeviceToHost, s.stream));
      CK(cudaStreamSynchronize(s.stream));
    }
    std::vector<int> order(nl);
    for (int k = 0; k < nl; ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return trip[size_t(a) * 3] < trip[size_t(b) * 3]; });
    std::vector<int64_t> tab(size_t(std::max(nl, 1)) * 2);
    for (int k = 0; k < nl; ++k) {
      LongStmt L;
      L.j = trip[size_t(order[k]) * 3];
      // step of schedule position j: last l with level_first[l] <= j
      L.level = int32_t(std::upper_bound(s.level_first.begin() + 1, s.level_first.begin() + n_levels + 1, L.j) -
                        s.level_first.begin()) - 1;
      L.b = trip[size_t(order[k]) * 3 + 1];
      L.e = trip[size_t(order[k]) * 3 + 2];
      tab[size_t(k) * 2] = L.b;
      tab[size_t(k) * 2 + 1] = L.e;
      s.long_stmts.push_back(L);
    }
    RET(ensure(errp, s.long_tab, tab.size() * 8));
    CK(cudaMemcpyAsync(s.long_tab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, s.stream));
    CK(cudaStreamSynchronize(s.stream));   // tab is a host vector
  }
  CK(cudaGetLastError());
  s.have_levels = true;
  s.have_schedule = true;
  return 0;
}

// The schedule of the tape on this shard, built on first need.  adept_cuda_upload_tape builds it eagerly in a
// one-process context; in a one-process-per-GPU job (adept_cuda_join) it is built by adept_cuda_share_tape on ALL
// ranks at the same time, after the broadcast -- not on the root first and on its peers afterwards.
int ensure_schedule(adept_cuda_ctx* c, Shard& s) {
  std::string* errp = &s.err;
  if (s.have_schedule) return 0;
  if (!s.have_tape) return fail(errp, ADEPT_CUDA_ERR_NO_TAPE, "no tape uploaded");
  auto t0 = std::chrono::steady_clock::now();
  RET(build_schedule(c, s));
  CK(cudaStreamSynchronize(s.stream));
  s.stat["analysis_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

void publish_schedule_stats(adept_cuda_ctx* c) {
  const Shard& s = c->sh[0];
  c->stat["n_levels"] = s.n_levels;
  c->stat["n_chunks"] = double(s.n_chunks);
  c->stat["n_records"] = double(s.R);
  c->stat["have_levels"] = s.have_levels ? 1.0 : 0.0;
  c->stat["kernel_launches"] = c->total_launches();
  for (const char* k : {"analysis_ms", "an_validate_ms", "an_window_ms", "an_pack_ms", "an_forward_ms", "an_target_ms", "an_fused_ms",
                        "frac_mult_zero", "frac_mult_one", "frac_mult_minus_one", "ops_removed", "n_ops_recorded",
                        "max_short_chunk_recs"}) {
    auto it = s.stat.find(k);
    if (it != s.stat.end()) c->stat[k] = it->second;
  }
}

// One-process-per-GPU jobs: every rank contributes its local status; all ranks get the worst one, so that a
// failure on one rank (a bad argument, an allocation that did not fit) makes EVERY rank leave before the next
// collective instead of leaving its peers waiting in it.
int agree(adept_cuda_ctx* c, int local_rc) {
  if (!c->multi_process || c->world <= 1) return local_rc;
  Shard& s = c->sh[0];
  std::string* errp = &c->err;
  CK(cudaSetDevice(s.dev));
  RET(ensure(errp, s.agree_buf, 16));
  int v = local_rc;
  CK(cudaMemcpyAsync(s.agree_buf.p, &v, sizeof v, cudaMemcpyHostToDevice, s.stream));
  NK(ncclAllReduce(s.agree_buf.p, s.agree_buf.p, 1, ncclInt, ncclMax, s.comm, s.stream));
  CK(cudaMemcpyAsync(&v, s.agree_buf.p, sizeof v, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  if (v && !local_rc) return fail(errp, v, "another rank failed this collective call (status %d)", v);
  return local_rc;
}



// chunk and step tables of a reverse stream made of target groups (shared by the two-kernel target stream and
// the fused stream): goff[g] = record offset of group g (goff[n_groups] = n_recs), gstep[g] = its step.
int build_group_tables(Shard& s, int64_t n_groups, int64_t n_recs, const int64_t* goff, const int32_t* gstep, Buf& chunk_tab,
                       int64_t* n_chunks_out, std::vector<int64_t>& step_group_h, std::vector<int64_t>& step_chunk_h,
                       Buf* d_step_group) {
  std::string* errp = &s.err;
  const int T = 256;
  const int32_t n_levels = s.n_levels;
  Buf &tmp = s.pool[9], &cflag = s.pool[18], &cincl = s.pool[19], &sgroup = s.pool[20], &stchunk = s.pool[21];
  RET(ensure(errp, cflag, size_t(n_groups + 1) * 4));
  RET(ensure(errp, cincl, size_t(n_groups + 1) * 4));
  RET(ensure(errp, sgroup, size_t(n_levels + 2) * 8));
  RET(ensure(errp, stchunk, size_t(n_levels + 2) * 8));
  CK(cudaMemsetAsync(sgroup.p, 0xFF, size_t(n_levels + 2) * 8, s.stream));
  target_chunk_flags_kernel<<<grid_for(n_groups, T, s.sm_count), T, 0, s.stream>>>(goff, gstep, n_groups, s.chunk_recs,
                                                                                  cflag.as<int32_t>(), sgroup.as<int64_t>());
  s.launches += 1;
  RET(inclusive_sum(errp, s, tmp, cflag.as<int32_t>(), cincl.as<int32_t>(), n_groups));
  int32_t n_chunks32 = 0;
  CK(cudaMemcpyAsync(&n_chunks32, cincl.as<int32_t>() + (n_groups - 1), 4, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  const int64_t n_chunks = n_chunks32;
  *n_chunks_out = n_chunks;
  RET(ensure(errp, chunk_tab, size_t(n_chunks + 1) * 8));
  target_chunk_table_kernel<<<grid_for(n_groups, T, s.sm_count), T, 0, s.stream>>>(goff, cflag.as<int32_t>(), cincl.as<int32_t>(),
                                                                                  n_groups, n_chunks, n_recs,
                                                                                  chunk_tab.as<int64_t>());
  target_step_tables_kernel<<<1, 1, 0, s.stream>>>(sgroup.as<int64_t>(), n_levels, n_groups, cincl.as<int32_t>(), n_chunks,
                                                   stchunk.as<int64_t>());
  s.launches += 2;
  step_group_h.assign(size_t(n_levels) + 2, 0);
  step_chunk_h.assign(size_t(n_levels) + 2, 0);
  CK(cudaMemcpyAsync(step_group_h.data(), sgroup.p, size_t(n_levels + 2) * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaMemcpyAsync(step_chunk_h.data(), stchunk.p, size_t(n_levels + 2) * 8, cudaMemcpyDeviceToHost, s.stream));
  if (d_step_group) {
    RET(ensure(errp, *d_step_group, size_t(n_levels + 2) * 8));
    CK(cudaMemcpyAsync(d_step_group->p, sgroup.p, size_t(n_levels + 2) * 8, cudaMemcpyDeviceToDevice, s.stream));
  }
  CK(cudaStreamSynchronize(s.stream));
  return 0;
}

// Reverse target stream of the two-kernel level sweep (snapshot + gather), the shared-memory kernel, the persistent
// kernel and the single-direction adjoint: operations sorted by (step, slot, reference reverse order).  Built on
// first need: a tape that is only swept forward, or backward through the fused stream, never pays for it.
int ensure_target_stream(const adept_cuda_ctx* c, Shard& s) {
  std::string* errp = &s.err;
  if (s.have_target) return 0;
  if (!s.have_levels) return fail(errp, ADEPT_CUDA_ERR_ARG, "no level schedule: no target stream");
  CK(cudaSetDevice(s.dev));
  const int T = 256;
  const int64_t O = s.n_ops;
  const int32_t n_levels = s.n_levels;
  auto t0 = std::chrono::steady_clock::now();
  PoolGuard pool_guard{s};
  s.pool.resize(kPoolSlots);
  Buf &k0 = s.pool[5], &k1 = s.pool[6], &v0 = s.pool[7], &v1 = s.pool[8], &tmp = s.pool[9], &flag = s.pool[12],
      &incl = s.pool[16], &gstep = s.pool[17];
  s.step_group.assign(size_t(n_levels) + 2, 0);
  s.step_tchunk.assign(size_t(n_levels) + 2, 0);
  s.n_groups = 0;
  s.n_tchunks = 0;
  s.R_t = 0;
  RET(ensure(errp, s.d_step_tchunk, size_t(n_levels + 2) * 8));
  RET(ensure(errp, s.d_step_group, size_t(n_levels + 2) * 8));
  RET(ensure(errp, s.sg_rev_info, size_t(n_levels + 2) * 32));
  if (O > 0) {
    for (Buf* b : {&k0, &k1, &v0, &v1}) RET(ensure(errp, *b, size_t(O) * 4));
    RET(ensure(errp, flag, size_t(O + 1) * 4));
    target_keys_kernel<<<grid_for(O, T, s.sm_count), T, 0, s.stream>>>(s.stmt.as<int2>(), s.n_stmt,
                                                                       s.idx.as<int32_t>(), O, k0.as<uint32_t>(),
                                                                       v0.as<uint32_t>());
    s.launches += 1;
    RET(sort_pairs(errp, s, k0, k1, v0, v1, tmp, O, bits_for(s.max_gradient)));
    target_step_keys_kernel<<<grid_for(O, T, s.sm_count), T, 0, s.stream>>>(
        s.stmt.as<int2>(), s.n_stmt, v0.as<uint32_t>(), O, s.level.as<int32_t>(), k0.as<uint32_t>());
    s.launches += 1;
    RET(sort_pairs(errp, s, k0, k1, v0, v1, tmp, O, bits_for(int64_t(n_levels) + 1)));
    target_flags_kernel<<<grid_for(O, T, s.sm_count), T, 0, s.stream>>>(k0.as<uint32_t>(), v0.as<uint32_t>(),
                                                                        s.idx.as<int32_t>(), O, flag.as<int32_t>());
    s.launches += 1;
    RET(ensure(errp, incl, size_t(O) * 4));
    RET(inclusive_sum(errp, s, tmp, flag.as<int32_t>(), incl.as<int32_t>(), O));
    int32_t n_groups32 = 0;
    CK(cudaMemcpyAsync(&n_groups32, incl.as<int32_t>() + (O - 1), 4, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    const int64_t n_groups = n_groups32;
    const int64_t R_t = O + n_groups;
    s.n_groups = n_groups;
    s.R_t = R_t;
    RET(ensure(errp, s.tgt, size_t(R_t) * sizeof(Rec)));
    RET(ensure(errp, s.goff, size_t(n_groups + 1) * 8));
    RET(ensure(errp, gstep, size_t(n_groups + 1) * 4));
    target_emit_kernel<<<grid_for(O, T, s.sm_count), T, 0, s.stream>>>(
        s.stmt.as<int2>(), s.n_stmt, s.mult.as<double>(), s.idx.as<int32_t>(), k0.as<uint32_t>(), v0.as<uint32_t>(),
        flag.as<int32_t>(), incl.as<int32_t>(), O, s.pos_in_sched.as<int32_t>(), s.d_level_first.as<int64_t>(),
        s.tgt.as<Rec>(), s.goff.as<int64_t>(), gstep.as<int32_t>());
    s.launches += 1;
    CK(cudaMemcpyAsync(s.goff.as<int64_t>() + n_groups, &s.R_t, 8, cudaMemcpyHostToDevice, s.stream));
    RET(build_group_tables(s, n_groups, R_t, s.goff.as<int64_t>(), gstep.as<int32_t>(), s.tchunk, &s.n_tchunks, s.step_group,
                           s.step_tchunk, nullptr));
  }
  CK(cudaMemcpyAsync(s.d_step_group.p, s.step_group.data(), size_t(n_levels + 2) * 8, cudaMemcpyHostToDevice, s.stream));
  CK(cudaMemcpyAsync(s.d_step_tchunk.p, s.step_tchunk.data(), size_t(n_levels + 2) * 8, cudaMemcpyHostToDevice, s.stream));
  if (O > 0)
    step_info_kernel<<<unsigned((n_levels + 2 + T - 1) / T), T, 0, s.stream>>>(s.d_step_group.as<int64_t>(), s.goff.as<int64_t>(),
                                                                             n_levels, s.sg_rev_info.as<longlong4>());
  else
    CK(cudaMemsetAsync(s.sg_rev_info.p, 0, size_t(n_levels + 2) * 32, s.stream));
  s.launches += 1;
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaGetLastError());
  s.stat["an_target_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  s.have_target = true;
  return 0;
}

__global__ void max_chunk_len_kernel(const int64_t* __restrict__ chunk_begin, int64_t n_chunks, unsigned long long* out) {
  unsigned long long m = 0;
  for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < n_chunks; c += int64_t(gridDim.x) * blockDim.x)
    m = max(m, static_cast<unsigned long long>(chunk_begin[c + 1] - chunk_begin[c]));
  m = __reduce_max_sync(0xffffffffu, static_cast<unsigned int>(m));   // chunks are shorter than 2^31 records (R_f < 2^31)
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// Fused reverse stream over versioned rows (tape_kernels.cuh, "FUSED reverse stream"; the sequential statement of
// the construction is tests/test_fused_stream_model.py).  Built on first need.
int ensure_fused_stream(const adept_cuda_ctx* c, Shard& s) {
  std::string* errp = &s.err;
  if (s.have_fused) return 0;
  if (!s.have_levels) return fail(errp, ADEPT_CUDA_ERR_ARG, "no level schedule: no fused stream");
  CK(cudaSetDevice(s.dev));
  const int T = 256;
  const int64_t S = s.n_stmt - 1, O = s.n_ops, n = O + S, G = s.max_gradient;
  const int32_t n_levels = s.n_levels;
  auto t0 = std::chrono::steady_clock::now();
  PoolGuard pool_guard{s};
  s.pool.resize(kPoolSlots);
  Buf &k0 = s.pool[5], &k1 = s.pool[6], &v0 = s.pool[7], &v1 = s.pool[8], &tmp = s.pool[9], &flag = s.pool[12],
      &gincl = s.pool[16], &gstep = s.pool[17], &isl = s.pool[34], &E = s.pool[35], &e0 = s.pool[36], &par = s.pool[37],
      &hflag = s.pool[38], &hincl = s.pool[39], &goff = s.pool[40];
  for (Buf* b : {&k0, &k1, &v0, &v1, &isl, &E, &hflag, &hincl, &gincl}) RET(ensure(errp, *b, size_t(n) * 4));
  RET(ensure(errp, flag, size_t(n + 1) * 4));
  RET(ensure(errp, e0, size_t(G) * 4));
  RET(ensure(errp, par, size_t(n)));
  RET(ensure(errp, s.fpar, size_t(G)));
  // references in sweep order, sorted by slot: writers already passed -> row parities
  fused_keys_kernel<<<grid_for(std::max(O, S), T, s.sm_count), T, 0, s.stream>>>(s.stmt.as<int2>(), s.n_stmt, s.idx.as<int32_t>(),
                                                                                 O, k0.as<uint32_t>(), v0.as<uint32_t>());
  s.launches += 1;
  RET(sort_pairs(errp, s, k0, k1, v0, v1, tmp, n, bits_for(G)));
  fused_islhs_kernel<<<grid_for(n, T, s.sm_count), T, 0, s.stream>>>(v0.as<uint32_t>(), n, O, isl.as<int32_t>());
  s.launches += 1;
  RET(exclusive_sum(errp, s, tmp, isl.as<int32_t>(), E.as<int32_t>(), n));
  CK(cudaMemsetAsync(s.fpar.p, 0, size_t(G), s.stream));
  fused_segstart_kernel<<<grid_for(n, T, s.sm_count), T, 0, s.stream>>>(k0.as<uint32_t>(), E.as<int32_t>(), n, e0.as<int32_t>());
  fused_parity_kernel<<<grid_for(n, T, s.sm_count), T, 0, s.stream>>>(k0.as<uint32_t>(), v0.as<uint32_t>(), isl.as<int32_t>(),
                                                                      E.as<int32_t>(), e0.as<int32_t>(), n, par.as<uint8_t>(),
                                                                      s.fpar.as<uint8_t>());
  s.launches += 2;
  // sorted by (step, slot, sweep order): groups, inserted heads, emission
  fused_step_keys_kernel<<<grid_for(n, T, s.sm_count), T, 0, s.stream>>>(s.stmt.as<int2>(), s.n_stmt, O, v0.as<uint32_t>(), n,
                                                                         s.level.as<int32_t>(), k0.as<uint32_t>());
  s.launches += 1;
  RET(sort_pairs(errp, s, k0, k1, v0, v1, tmp, n, bits_for(int64_t(n_levels) + 1)));
  RET(ensure(errp, s.scratch, 64));
  CK(cudaMemsetAsync(s.scratch.p, 0, 64, s.stream));
  fused_flags_kernel<<<grid_for(n, T, s.sm_count), T, 0, s.stream>>>(s.stmt.as<int2>(), s.idx.as<int32_t>(), O, k0.as<uint32_t>(),
                                                                     v0.as<uint32_t>(), n, flag.as<int32_t>(),
                                                                     hflag.as<int32_t>(), s.scratch.as<int>());
  s.launches += 1;
  RET(inclusive_sum(errp, s, tmp, flag.as<int32_t>(), gincl.as<int32_t>(), n));
  RET(inclusive_sum(errp, s, tmp, hflag.as<int32_t>(), hincl.as<int32_t>(), n));
  int32_t n_groups32 = 0, n_heads32 = 0, bad = 0;
  CK(cudaMemcpyAsync(&n_groups32, gincl.as<int32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaMemcpyAsync(&n_heads32, hincl.as<int32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaMemcpyAsync(&bad, s.scratch.p, 4, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  if (bad) return fail(errp, ADEPT_CUDA_ERR_CUDA, "fused stream: an LHS reference does not open its group (schedule violates WAR/RAW)");
  const int64_t n_groups = n_groups32;
  s.R_f = n + int64_t(n_heads32);
  if (s.R_f >= (int64_t(1) << 31) - 1) return fail(errp, ADEPT_CUDA_ERR_LIMIT, "fused stream too long");
  RET(ensure(errp, s.ftgt, size_t(s.R_f) * sizeof(Rec)));
  RET(ensure(errp, goff, size_t(n_groups + 1) * 8));
  RET(ensure(errp, gstep, size_t(n_groups + 1) * 4));
  fused_emit_kernel<<<grid_for(n, T, s.sm_count), T, 0, s.stream>>>(
      s.stmt.as<int2>(), s.n_stmt, s.mult.as<double>(), s.idx.as<int32_t>(), O, k0.as<uint32_t>(), v0.as<uint32_t>(),
      flag.as<int32_t>(), hflag.as<int32_t>(), gincl.as<int32_t>(), hincl.as<int32_t>(), n, par.as<uint8_t>(), G,
      s.ftgt.as<Rec>(), goff.as<int64_t>(), gstep.as<int32_t>());
  s.launches += 1;
  CK(cudaMemcpyAsync(goff.as<int64_t>() + n_groups, &s.R_f, 8, cudaMemcpyHostToDevice, s.stream));
  std::vector<int64_t> step_group_unused;
  RET(build_group_tables(s, n_groups, s.R_f, goff.as<int64_t>(), gstep.as<int32_t>(), s.fchunk, &s.n_fchunks, step_group_unused,
                         s.step_fchunk, nullptr));
  RET(ensure(errp, s.scratch, 64));
  CK(cudaMemsetAsync(s.scratch.p, 0, 8, s.stream));
  max_chunk_len_kernel<<<grid_for(s.n_fchunks, T, s.sm_count), T, 0, s.stream>>>(s.fchunk.as<int64_t>(), s.n_fchunks,
                                                                                 s.scratch.as<unsigned long long>());
  unsigned long long longest_chunk = 0;
  CK(cudaMemcpyAsync(&longest_chunk, s.scratch.p, 8, cudaMemcpyDeviceToHost, s.stream));
  RET(ensure(errp, s.d_step_fchunk, size_t(n_levels + 2) * 8));
  CK(cudaMemcpyAsync(s.d_step_fchunk.p, s.step_fchunk.data(), size_t(n_levels + 2) * 8, cudaMemcpyHostToDevice, s.stream));
  s.max_step_fchunks = 0;
  for (int32_t l = 1; l <= n_levels; ++l) s.max_step_fchunks = std::max(s.max_step_fchunks, s.step_fchunk[l + 1] - s.step_fchunk[l]);
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaGetLastError());
  s.max_fchunk_recs = int64_t(longest_chunk);
  s.stat["max_fused_chunk_recs"] = double(longest_chunk);
  s.stat["an_fused_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  s.stat["fused_records"] = double(s.R_f);
  s.have_fused = true;
  return 0;
}

// ---------------------------------------------------------------------------------------
// Replay drivers
// ---------------------------------------------------------------------------------------
struct SweepPlan {
  bool use_levels;
  int warps;
};

SweepPlan plan_sweep(const adept_cuda_ctx* c, const Shard& s, int64_t n_dirs) {
  SweepPlan p;
  p.use_levels = false;
  if (c->opt_schedule == 1 || !s.have_levels) p.use_levels = false;
  else if (c->opt_schedule == 2) p.use_levels = true;
  else {
    // the level schedule pays when a level offers more parallel work than one launch costs
    const double stmts_per_level = double(s.n_stmt - 1) / std::max(1, s.n_levels);
    p.use_levels = stmts_per_level >= 16.0;
  }
  p.warps = c->opt_warps;
  if (p.warps < 1) p.warps = 1;
  if (p.warps > 8) p.warps = 8;
  const int64_t ncb = (n_dirs + 31) / 32;
  if (!p.use_levels) {
    // tape-ordered: one chunk, so spread the column blocks over as many CTAs as possible
    const int64_t per = std::max<int64_t>(1, ncb / (2 * s.sm_count));
    p.warps = int(std::min<int64_t>(p.warps, per));
  } else {
    if (ncb < p.warps) p.warps = int(ncb);
  }
  if (p.warps < 1) p.warps = 1;
  return p;
}

// Launch with programmatic stream serialization (PDL, common.cuh): the kernel may begin while its predecessor
// in the stream drains; it calls pdl_wait() before touching anything the predecessor writes.
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// next event of the per-launch timing pool, recorded on the shard's stream
int mark(Shard& s) {
  std::string* errp = &s.err;
  if (s.kev_used == s.kev.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    s.kev.push_back(e);
  }
  CK(cudaEventRecord(s.kev[s.kev_used++], s.stream));
  return 0;
}

// next event of the per-pass pool, recorded on the shard's stream
int mark_pass(Shard& s) {
  std::string* errp = &s.err;
  if (s.pev_used == s.pev.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    s.pev.push_back(e);
  }
  CK(cudaEventRecord(s.pev[s.pev_used++], s.stream));
  return 0;
}

// One sweep over the resident pass (g already zeroed and seeded).  mode: 0 forward, 1 reverse.
int launch_sweep(const adept_cuda_ctx* c, Shard& s, int mode, double* g, int64_t K, int64_t n_dirs,
                 const SweepPlan& plan, bool fused, int V_fwd = 2) {
  std::string* errp = &s.err;
  if (s.R == 0) return 0;
  ReplayArgs A{};
  A.g = g;
  A.K = K;
  A.n_colblocks = int((n_dirs + 31) / 32);
  A.n_dirs = n_dirs;
  A.ref_block = c->opt_ref_block;
  const int threads = plan.warps * 32;
  const unsigned gy = unsigned((A.n_colblocks + plan.warps - 1) / plan.warps);
  s.stat["sweep_streams"] = 1;
  if (!plan.use_levels) {
    RET(ensure_tape_streams(s));
    A.recs = (mode == 0 ? s.fwd_t : s.rev_t).as<Rec>();
    A.chunk_begin = s.one_chunk.as<int64_t>();
    A.first_chunk = 0;
    if (c->opt_time_kernels) RET(mark(s));
    if (mode == 0) replay_forward_kernel<true, false><<<dim3(1, gy), threads, 0, s.stream>>>(A);
    else replay_reverse_kernel<false><<<dim3(1, gy), threads, 0, s.stream>>>(A);
    if (c->opt_time_kernels) RET(mark(s));
    s.launches += 1;
  } else if (mode == 0) {
    const int V = V_fwd;   // doubles per lane: 2, or 4 (forward_balanced_kernel<4>: 1024-byte row segments per warp)
    ForwardArgs F;
    F.recs = s.fwd_l.as<Rec>();
    F.chunk_begin = s.fwd_chunk.as<int64_t>();
    F.g = g;
    F.rowact = s.rowact.as<uint8_t>();
    F.K = K;
    F.n_colblocks = int((n_dirs + 32 * V - 1) / (32 * V));
    F.flag_pitch = int(K / (32 * V));
    F.cb_lo = 0;
    F.n_rows = s.max_gradient;
    int fwarps = std::max(1, std::min(plan.warps, F.n_colblocks));
    if (F.n_colblocks >= 16 && F.n_colblocks / fwarps < c->opt_streams)
      fwarps = std::max(1, std::min(fwarps, F.n_colblocks / std::max(1, c->opt_streams)));
    const bool sparse = !s.has_nonfinite;
    const bool balanced = c->opt_fwd_balanced != 0;
    s.stat["used_fwd_balanced"] = balanced ? 1.0 : 0.0;
    s.stat["fwd_lane_doubles"] = double(V);
    // balanced kernel: a CTA covers `cbs` column blocks and its warps take the active ones in turn; enough column
    // groups are kept for the streams below
    // 16 column blocks per CTA (two per warp) measured best on BASELINE config 4's shape: 8 and 16 within 1 %, 32 about 4 % slower
    // ... and when a device holds few directions (one rank of eight: 16 column blocks), smaller CTAs on more streams:
    // 1024 directions of config 4: 16 per CTA 949 ms, 8 on two streams 771 ms, 4 on four streams 701 ms
    const int cbs_default = V == 4 ? kFwdCbsDefault / 2 : kFwdCbsDefault;   // the same directions per CTA
    int cbs = cbs_default;
    if (c->opt_cbs > 0) cbs = int(std::min<int64_t>(c->opt_cbs, kFwdCbs));
    else if (F.n_colblocks < cbs_default * c->opt_streams)
      cbs = std::max(std::min(2, F.n_colblocks), (F.n_colblocks + c->opt_streams - 1) / std::max(1, c->opt_streams));
    cbs = std::max(1, std::min(cbs, kFwdCbs));
    F.cbs = cbs;
    const int bwarps = std::max(1, std::min(plan.warps, cbs));
    // independent direction ranges on separate streams, one issuing host thread each (see the reverse sweep)
    const int all_cb = F.n_colblocks;
    const int gwidth = balanced ? cbs : fwarps;   // column blocks per stream-partition unit
    const int groups = (all_cb + gwidth - 1) / gwidth;
    int nstr = std::max(1, std::min<int>(c->opt_streams, groups));
    // only narrow steps leave the machine idle between launches; wide ones would just re-read the tape per stream
    // wide steps: with the balanced kernel independent direction ranges on separate streams still pay (the tail of one
    // range's step is filled by another's next step; config 4's shape, 4096 directions: 1 stream 2.79 s, 2: 2.69 s,
    // 4: 2.63 s; the record stream is re-read once per range, < 1 % of the traffic); the old kernel keeps one stream
    if (double(s.n_chunks) / std::max(1, s.n_levels) * groups >= kWideStepCtas && !balanced) nstr = 1;
    if (c->opt_wide_streams > 0) nstr = std::max(1, std::min<int>(c->opt_wide_streams, groups));
    if (c->opt_time_kernels || c->opt_profile_every) nstr = 1;
    s.stat["sweep_streams"] = nstr;
    while (int(s.aux.size()) < nstr - 1) {
      cudaStream_t st;
      cudaEvent_t ev;
      CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      s.aux.push_back(st);
      s.aux_ev.push_back(ev);
    }
    if (nstr > 1) {
      CK(cudaEventRecord(s.fork_ev, s.stream));
      for (int i = 1; i < nstr; ++i) CK(cudaStreamWaitEvent(s.aux[i - 1], s.fork_ev, 0));
    }
    std::vector<int> rcs(nstr, 0);
    std::vector<double> nl(nstr, 0.0);
    const bool pdl = c->opt_pdl && !c->opt_time_kernels;
    auto run_stream = [&](int i) {
      if (cudaSetDevice(s.dev) != cudaSuccess) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
      const int g_lo = int(int64_t(groups) * i / nstr), g_hi = int(int64_t(groups) * (i + 1) / nstr);
      if (g_hi <= g_lo) return;
      cudaStream_t st = (i == 0) ? s.stream : s.aux[i - 1];
      ForwardArgs Fi = F;
      Fi.cb_lo = g_lo * gwidth;
      Fi.n_colblocks = std::min(all_cb, g_hi * gwidth);
      const int ncols = Fi.n_colblocks - Fi.cb_lo;
      for (int32_t l = 1; l <= s.n_levels; ++l) {
        const int64_t c0 = s.level_chunk[l], c1 = s.level_chunk[l + 1];
        if (c1 <= c0) continue;
        Fi.first_chunk = c0;
        const bool use_bal = balanced && !s.level_long[l];
        const dim3 grid(unsigned(c1 - c0), unsigned(use_bal ? (ncols + cbs - 1) / cbs : (ncols + fwarps - 1) / fwarps));
        const dim3 block(unsigned((use_bal ? bwarps : fwarps) * 32));
        if (c->opt_time_kernels && mark(s)) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
        const bool prof = c->opt_profile_every > 0 && (int64_t(nl[i]) % c->opt_profile_every) == 0;
        if (prof) {
            cudaStreamSynchronize(st);
            cudaProfilerStart();
        }
        cudaError_t le;
        if (use_bal && V == 4) {
          if (sparse) le = launch_pdl(forward_balanced_kernel<4, true>, grid, block, st, pdl && !prof, Fi);
          else le = launch_pdl(forward_balanced_kernel<4, false>, grid, block, st, pdl && !prof, Fi);
        } else if (use_bal) {
          if (sparse) le = launch_pdl(forward_balanced_kernel<2, true>, grid, block, st, pdl && !prof, Fi);
          else le = launch_pdl(forward_balanced_kernel<2, false>, grid, block, st, pdl && !prof, Fi);
        } else {   // (V == 2 here: shard_enqueue keeps four doubles per lane for tapes without long steps)
          if (sparse) le = launch_pdl(forward_level_kernel<2, true>, grid, block, st, pdl && !prof, Fi);
          else le = launch_pdl(forward_level_kernel<2, false>, grid, block, st, pdl && !prof, Fi);
        }
        if (le != cudaSuccess) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
        if (prof) {
          cudaStreamSynchronize(st);
          cudaProfilerStop();
          char line[160];
          snprintf(line, sizeof line, "fwd step %d stmts %lld recs %lld dirs %lld grid %u %u", int(l),
                   (long long)(s.level_first[l + 1] - s.level_first[l]), (long long)s.step_recs[l],
                   (long long)n_dirs, grid.x, grid.y);
          s.prof_log.push_back(line);
        }
        if (c->opt_time_kernels && mark(s)) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
        nl[i] += 1;
      }
      if (cudaGetLastError() != cudaSuccess) rcs[i] = ADEPT_CUDA_ERR_CUDA;
    };
    if (nstr == 1) {
      run_stream(0);
    } else {
      std::vector<std::thread> th;
      for (int i = 1; i < nstr; ++i) th.emplace_back(run_stream, i);
      run_stream(0);
      for (std::thread& t : th) t.join();
    }
    for (int i = 0; i < nstr; ++i) {
      s.launches += nl[i];
      if (rcs[i]) return fail(errp, rcs[i], "launch failed in the forward level sweep");
    }
    if (nstr > 1) {
      for (int i = 1; i < nstr; ++i) {
        CK(cudaEventRecord(s.aux_ev[i - 1], s.aux[i - 1]));
        CK(cudaStreamWaitEvent(s.stream, s.aux_ev[i - 1], 0));
      }
    }
  } else {
    // reverse: steps downwards.  Fused: one kernel per step over versioned rows (reverse_fused_kernel).  Otherwise:
    // per step snapshot a_s and zero the LHS slots, then gather per target.
    constexpr int V = 2;   // a lane owns two adjacent directions: 16-byte row accesses
    if (!fused) RET(ensure_target_stream(c, s));
    TargetArgs B{};
    B.recs = (fused ? s.ftgt : s.tgt).as<Rec>();
    B.chunk_begin = (fused ? s.fchunk : s.tchunk).as<int64_t>();
    B.g = g;
    B.abuf = fused ? nullptr : s.abuf.as<double>();
    B.rowact = s.rowact.as<uint8_t>();
    B.aflag = fused ? nullptr : s.aflag.as<uint8_t>();
    B.K = K;
    B.n_colblocks = int((n_dirs + 32 * V - 1) / (32 * V));
    B.flag_pitch = int(K / (32 * V));
    B.n_dirs = n_dirs;
    B.ref_block = c->opt_ref_block;
    B.cb_lo = 0;
    // warps per CTA: with few column blocks (few directions on this GPU) use narrower CTAs so that there are still
    // several independent direction ranges to spread over the streams
    int twarps = std::max(1, std::min(plan.warps, B.n_colblocks));
    if (!fused && !c->opt_persistent && B.n_colblocks >= 16 && B.n_colblocks / twarps < c->opt_streams)
      twarps = std::max(1, std::min(twarps, B.n_colblocks / std::max(1, c->opt_streams)));
    const unsigned tgy = unsigned((B.n_colblocks + twarps - 1) / twarps);
    // fused kernel: a CTA covers `cbs` column blocks and its warps take the active ones in turn; enough column
    // groups are kept for the streams below
    int cbs = kFusedCbs;
    if (c->opt_cbs > 0) cbs = int(std::min<int64_t>(c->opt_cbs, kFusedCbs));
    else if (B.n_colblocks < kFusedCbs * c->opt_streams)
      cbs = std::max(std::min(8, B.n_colblocks), (B.n_colblocks + c->opt_streams - 1) / std::max(1, c->opt_streams));
    cbs = std::max(1, std::min(cbs, kFusedCbs));
    B.cbs = cbs;
    // Work distribution of the fused kernels.  A Jacobian's activity is banded: of a step's chunks only those whose
    // targets lie within reach of this device's directions have anything to do, and there all column blocks at once.
    // When the device holds few directions (one rank of several) a step is a few busy CTAs, each warp walking a whole
    // chunk (BASELINE config 2 at 512 directions: 18 us of dependent row loads per step on a mostly idle machine).
    // Then every chunk goes to `split` CTAs and is cut into sub-tasks of `sub` records (replay_kernels.cuh).
    {
      int split = 1, sub = 0;
      const double ctas = double(s.n_fchunks) / std::max(1, s.n_levels) * ((B.n_colblocks + cbs - 1) / cbs);
      const double resident = 4.0 * s.sm_count;
      // measured on config 2 (sweep of one Jacobian, ms; profiles/r02s_*): 512 directions 41.3 / 32.7 / 37.0 / 51.4 for
      // split 1 / 2 / 4 / 8, 1024 directions 41.2 / 37.1 / 48.9 / 80.6, 2048 directions 48.5 / 60.6 / 92.8: every extra
      // CTA costs its launch, staging and activity round trip, so two per chunk, and only with few directions
      if (c->opt_auto_split && fused && B.n_colblocks <= 16 && ctas > 0 && ctas < resident && s.max_fchunk_recs <= kPieceRecs) split = 2;
      if (c->opt_split > 0) split = c->opt_split;
      if (split > 1) sub = int(std::max<int64_t>(8, (s.chunk_recs > 0 ? s.chunk_recs : kChunkRecs) / split));
      if (c->opt_sub > 0) sub = c->opt_sub;
      B.split = fused ? split : 1;
      B.sub = fused ? sub : 0;
      s.stat["fused_split"] = double(B.split);
      s.stat["fused_sub"] = double(B.sub);
    }
    if (fused) twarps = std::max(1, std::min(plan.warps, cbs));
    const int gwidth = fused ? cbs : twarps;   // column blocks per CTA row
    if (fused && c->opt_fused_persistent != 0 && !c->opt_time_kernels && !c->opt_profile_every && s.n_fchunks > 0) {
      // ONE resident kernel for the whole pass (fused_persistent_kernels.cuh) when the steps are narrow: items of a
      // step = chunks x column groups; a step that does not fill the resident grid twice is all launch, drain and refill
      int pcbs = c->opt_cbs > 0 ? int(std::min<int64_t>(c->opt_cbs, kFusedCbs)) : std::min(kFusedCbs, B.n_colblocks);
      pcbs = std::max(1, pcbs);
      const int pgy = (B.n_colblocks + pcbs - 1) / pcbs;
      const int pwarps = std::max(1, std::min(plan.warps, pcbs));
      int per_sm = 0;
      cudaError_t oe;
      if (B.ref_block == 1) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reverse_fused_persistent_kernel<V, 0>, pwarps * 32, 0);
      else if (B.ref_block == 2) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reverse_fused_persistent_kernel<V, 1>, pwarps * 32, 0);
      else oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reverse_fused_persistent_kernel<V, 2>, pwarps * 32, 0);
      CK(oe);
      const int64_t resident = int64_t(per_sm) * s.sm_count;
      const double items_avg = double(s.n_fchunks) / std::max(1, s.n_levels) * pgy;
      const bool narrow = items_avg < 2.0 * double(resident);
      if (resident >= 1 && (c->opt_fused_persistent > 0 || narrow)) {
        FusedPersistArgs PA{};
        PA.T = B;
        PA.T.cbs = pcbs;
        PA.T.cb_lo = 0;
        PA.T.n_tail = 0;
        PA.T.done = nullptr;
        PA.step_fchunk = s.d_step_fchunk.as<int64_t>();
        PA.n_levels = s.n_levels;
        PA.gy = pgy;
        RET(ensure(errp, s.pcnt, size_t(s.n_levels + 2) * sizeof(unsigned int)));
        if (!s.pabort.p) {
          RET(ensure(errp, s.pabort, 128));
          CK(cudaMemsetAsync(s.pabort.p, 0, 128, s.stream));
        }
        CK(cudaMemsetAsync(s.pcnt.p, 0, size_t(s.n_levels + 2) * sizeof(unsigned int), s.stream));
        PA.cnt = s.pcnt.as<unsigned int>();
        PA.abort = s.pabort.as<unsigned int>();
        PA.spin_limit = static_cast<unsigned int>(c->opt_spin_limit);
        const int64_t G = std::max<int64_t>(1, std::min(resident, s.max_step_fchunks * pgy * PA.T.split));
#ifdef ADEPT_FUSED_TRACE
        const char* ptrace_file = getenv("ADEPT_CUDA_TRACE_FILE");
        if (ptrace_file && *ptrace_file) {
          RET(ensure(errp, s.trace, size_t(8192) * 4 * 8));
          CK(cudaMemsetAsync(s.trace.p, 0, size_t(8192) * 4 * 8, s.stream));
          PA.trace = s.trace.as<unsigned long long>();
        }
#endif
        void* args[] = {&PA};
        const void* fn = B.ref_block == 1   ? (const void*)reverse_fused_persistent_kernel<V, 0>
                         : B.ref_block == 2 ? (const void*)reverse_fused_persistent_kernel<V, 1>
                                            : (const void*)reverse_fused_persistent_kernel<V, 2>;
        if (cudaLaunchCooperativeKernel(fn, dim3(unsigned(G)), dim3(unsigned(pwarps * 32)), args, 0, s.stream) != cudaSuccess) {
          cudaGetLastError();   // no cooperative launch on this device / in this context: one launch per step, below
          goto per_step;
        }
        s.launches += 1;
        s.persist_checked = true;
        s.stat["sweep_streams"] = 1;
        s.stat["used_fused_persistent"] = 1.0;
        s.stat["persistent_grid"] = double(G);
        s.stat["persistent_cbs"] = double(pcbs);
        CK(cudaGetLastError());
#ifdef ADEPT_FUSED_TRACE
        if (PA.trace) {   // profiling aid: the clock stamps of CTA 0's items (wait begins, wait ends, activity bytes in, item done)
          CK(cudaStreamSynchronize(s.stream));
          std::vector<unsigned long long> h(size_t(8192) * 4);
          CK(cudaMemcpy(h.data(), s.trace.p, h.size() * 8, cudaMemcpyDeviceToHost));
          if (FILE* f = fopen(ptrace_file, "w")) {
            for (size_t q = 0; q < 8192 && h[q * 4 + 3]; ++q) fprintf(f, "%llu %llu %llu %llu\n", h[q * 4], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]);
            fclose(f);
          }
        }
#endif
        return 0;
      }
    }
  per_step:
    if (!fused && c->opt_persistent && s.n_tchunks > 0) {
      // one cooperative kernel walks all steps (persistent_kernels.cuh)
      PersistArgs PA;
      PA.T = B;
      PA.sched_lhs = s.sched_lhs.as<int32_t>();
      PA.level_first = s.d_level_first.as<int64_t>();
      PA.step_tchunk = s.d_step_tchunk.as<int64_t>();
      PA.n_levels = s.n_levels;
      PA.gy = int(tgy);
      int per_sm = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reverse_persistent_kernel<V>, twarps * 32, 0));
      if (per_sm >= 1) {
        void* args[] = {&PA};
        if (c->opt_time_kernels) RET(mark(s));
        CK(cudaLaunchCooperativeKernel((void*)reverse_persistent_kernel<V>, dim3(unsigned(s.sm_count * per_sm)),
                                       dim3(unsigned(twarps * 32)), args, 0, s.stream));
        if (c->opt_time_kernels) RET(mark(s));
        s.launches += 1;
        CK(cudaGetLastError());
        return 0;
      }
    }
    // The steps of one direction range are strictly ordered, but direction ranges are independent: split the
    // column blocks over a few CUDA streams so that the launch gap and the tail of one range's step are filled
    // by another range's kernels.
    const int all_cb = B.n_colblocks;
    int nstr = std::max(1, std::min<int>(c->opt_streams, (all_cb + gwidth - 1) / gwidth));
    // only narrow steps leave the machine idle between launches; wide ones would just re-read the tape per stream
    if (double(fused ? s.n_fchunks : s.n_tchunks) / std::max(1, s.n_levels) * ((all_cb + gwidth - 1) / gwidth) >= kWideStepCtas ||
        c->opt_time_kernels || c->opt_profile_every)
      nstr = 1;
    s.stat["sweep_streams"] = nstr;
    while (int(s.aux.size()) < nstr - 1) {
      cudaStream_t st;
      cudaEvent_t ev;
      CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      s.aux.push_back(st);
      s.aux_ev.push_back(ev);
    }
#ifdef ADEPT_FUSED_TRACE
    constexpr size_t kTraceLaunches = 4096;
    const char* trace_file = getenv("ADEPT_CUDA_TRACE_FILE");
    const bool trace_on = fused && trace_file && *trace_file;
    if (trace_on) {
      std::vector<unsigned long long> init(kTraceLaunches * 8, 0ull);
      for (size_t k = 0; k < kTraceLaunches; ++k) init[k * 8 + 0] = init[k * 8 + 1] = ~0ull;
      RET(ensure(errp, s.trace, init.size() * 8));
      CK(cudaMemcpy(s.trace.p, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
    }
#endif
    if (fused && c->opt_tail) {  // ticket counters of the launches that carry tail steps
      const size_t nb = size_t(nstr) * size_t(s.n_levels + 1) * sizeof(unsigned int);
      RET(ensure(errp, s.done, nb));
      CK(cudaMemsetAsync(s.done.p, 0, nb, s.stream));
    }
    if (nstr > 1) {
      CK(cudaEventRecord(s.fork_ev, s.stream));
      for (int i = 1; i < nstr; ++i) CK(cudaStreamWaitEvent(s.aux[i - 1], s.fork_ev, 0));
    }
    // column groups (of twarps blocks) per stream; one host thread per stream issues that stream's launches
    // (thousands of small launches per sweep: a single thread would be the bottleneck)
    const int groups = (all_cb + gwidth - 1) / gwidth;
    std::vector<int> rcs(nstr, 0);
    std::vector<double> nl(nstr, 0.0);
    const bool pdl = c->opt_pdl && !c->opt_time_kernels;
    auto run_stream = [&](int i) {
      if (cudaSetDevice(s.dev) != cudaSuccess) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
      const int g_lo = int(int64_t(groups) * i / nstr), g_hi = int(int64_t(groups) * (i + 1) / nstr);
      if (g_hi <= g_lo) return;
      const int cb_lo = g_lo * gwidth, cb_hi = std::min(all_cb, g_hi * gwidth);
      cudaStream_t st = (i == 0) ? s.stream : s.aux[i - 1];
      const int ncols = cb_hi - cb_lo;
      TargetArgs Bi = B;
      Bi.cb_lo = cb_lo;
      Bi.n_colblocks = cb_hi;
      for (int32_t l = s.n_levels; l >= 1; --l) {
        const int64_t j0 = s.level_first[l], n = s.level_first[l + 1] - j0;
        if (n <= 0) continue;
        if (fused) {
          const int64_t c0 = s.step_fchunk[l], c1 = s.step_fchunk[l + 1];
          if (c1 <= c0) continue;
          Bi.first_chunk = c0;
          const int32_t l_top = l;
          // steps of a chunk or two that follow ride along with this launch (run by its last CTA)
          Bi.n_tail = 0;
          while (c->opt_tail && Bi.n_tail < kMaxTail && l - 1 >= 1) {
            const int64_t t0 = s.step_fchunk[l - 1], t1 = s.step_fchunk[l];
            if ((t1 - t0) * int64_t(g_hi - g_lo) > kTailItems) break;
            if (t1 > t0) {
              Bi.tail_c0[Bi.n_tail] = t0;
              Bi.tail_c1[Bi.n_tail] = t1;
              ++Bi.n_tail;
            }
            --l;
          }
          Bi.done = Bi.n_tail ? s.done.as<unsigned int>() + size_t(i) * size_t(s.n_levels + 1) + size_t(l) : nullptr;
#ifdef ADEPT_FUSED_TRACE
          Bi.trace = (i == 0 && trace_on && nl[0] < kTraceLaunches) ? s.trace.as<unsigned long long>() + size_t(nl[0]) * 8 : nullptr;
#endif
          if (c->opt_time_kernels && mark(s)) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
          const dim3 fgrid(unsigned((c1 - c0) * Bi.split), unsigned(g_hi - g_lo)), fblock(twarps * 32);
          const bool prof = c->opt_profile_every > 0 && (int64_t(nl[i]) % c->opt_profile_every) == 0;
          if (prof) {
            cudaStreamSynchronize(st);
            cudaProfilerStart();
          }
          cudaError_t le;
          if (B.ref_block == 1) le = launch_pdl(reverse_fused_kernel<V, 0>, fgrid, fblock, st, pdl && !prof, Bi);
          else if (B.ref_block == 2) le = launch_pdl(reverse_fused_kernel<V, 1>, fgrid, fblock, st, pdl && !prof, Bi);
          else le = launch_pdl(reverse_fused_kernel<V, 2>, fgrid, fblock, st, pdl && !prof, Bi);
          if (le != cudaSuccess) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
          if (prof) {
            cudaStreamSynchronize(st);
            cudaProfilerStop();
            // the launch covers step l_top .. l (tails ride along): algorithmic work = their forward records
            long long recs = 0, stmts = 0;
            for (int32_t q = l; q <= l_top; ++q) { recs += s.step_recs[q]; stmts += s.level_first[q + 1] - s.level_first[q]; }
            char line[160];
            snprintf(line, sizeof line, "rev step %d stmts %lld recs %lld dirs %lld grid %u %u tails %d", int(l_top), stmts, recs,
                     (long long)n_dirs, fgrid.x, fgrid.y, Bi.n_tail);
            s.prof_log.push_back(line);
          }
          if (c->opt_time_kernels && mark(s)) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
          nl[i] += 1;
          continue;
        }
        const int64_t c0 = s.step_tchunk[l], c1 = s.step_tchunk[l + 1];
        if (launch_pdl(reverse_snapshot_kernel<V, 4>, dim3(unsigned((n * ncols + 255) / 256)), dim3(256), st, pdl,
                       s.sched_lhs.as<int32_t>(), j0, n, g, s.abuf.as<double>(), K, n_dirs, s.rowact.as<uint8_t>(),
                       s.aflag.as<uint8_t>(), B.flag_pitch, ncols, cb_lo) != cudaSuccess) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
        nl[i] += 1;
        if (c1 > c0) {
          Bi.first_chunk = c0;
          if (c->opt_time_kernels && mark(s)) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
          if (launch_pdl(reverse_target_kernel<V>, dim3(unsigned(c1 - c0), unsigned(g_hi - g_lo)), dim3(twarps * 32), st, pdl, Bi) !=
              cudaSuccess) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
          if (c->opt_time_kernels && mark(s)) { rcs[i] = ADEPT_CUDA_ERR_CUDA; return; }
          nl[i] += 1;
        }
      }
      if (cudaGetLastError() != cudaSuccess) rcs[i] = ADEPT_CUDA_ERR_CUDA;
    };
    const auto t_issue = std::chrono::steady_clock::now();
    if (nstr == 1) {
      run_stream(0);
    } else {
      std::vector<std::thread> th;
      for (int i = 1; i < nstr; ++i) th.emplace_back(run_stream, i);
      run_stream(0);
      for (std::thread& t : th) t.join();
    }
    s.stat["host_issue_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_issue).count();
    for (int i = 0; i < nstr; ++i) {
      s.launches += nl[i];
      if (rcs[i]) return fail(errp, rcs[i], "launch failed in the reverse level sweep: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (nstr > 1) {
      for (int i = 1; i < nstr; ++i) {
        CK(cudaEventRecord(s.aux_ev[i - 1], s.aux[i - 1]));
        CK(cudaStreamWaitEvent(s.stream, s.aux_ev[i - 1], 0));
      }
    }
#ifdef ADEPT_FUSED_TRACE
    if (trace_on) {   // profiling aid: dump the time stamps of stream 0's launches
      CK(cudaStreamSynchronize(s.stream));
      std::vector<unsigned long long> h(kTraceLaunches * 8);
      CK(cudaMemcpy(h.data(), s.trace.p, h.size() * 8, cudaMemcpyDeviceToHost));
      if (FILE* f = fopen(trace_file, "w")) {
        for (size_t k = 0; k < kTraceLaunches && k < size_t(nl[0]); ++k) {
          for (int q = 0; q < 8; ++q) fprintf(f, "%llu ", h[k * 8 + q]);
          fprintf(f, "\n");
        }
        fclose(f);
      }
    }
#endif
  }
  CK(cudaGetLastError());
  return 0;
}

// Enqueue (no host synchronisation) directions [d_lo, d_hi) of the Jacobian on one shard;
// results go to `out` (device memory of this shard) as out[i*stride_i + (d - d_base)*stride_d],
// i over the non-swept variable list.  shard_finish() waits and collects the timings.
int shard_enqueue(const adept_cuda_ctx* c, Shard& s, int mode, const int32_t* seed_slots_dev,
                  const int32_t* out_slots_dev, int64_t n_out, int64_t d_lo, int64_t d_hi, int64_t d_base,
                  double* out, int64_t stride_i, int64_t stride_d) {
  std::string* errp = &s.err;
  CK(cudaSetDevice(s.dev));
  const int64_t D = d_hi - d_lo;
  s.kev_used = 0;
  s.pev_used = 0;
  s.sweep_launches = 0;
  CK(cudaEventRecord(s.ev[0], s.stream));
  s.stat["n_passes"] = 0;
  s.stat["used_fused_persistent"] = 0.0;
  // tapes whose gradient list fits shared memory: one CTA per 32/64 directions walks all steps by itself
  if (D > 0 && c->opt_smallg && s.have_levels && c->opt_schedule != 1 && s.max_step_width < (int64_t(1) << 20)) {
    const size_t kSmemCap = size_t(220) << 10;
    const size_t stage_bytes = ((sizeof(SgStage) + 15) / 16) * 16;
    auto need = [&](int dc) {
      return stage_bytes + (size_t(s.max_gradient) + (mode == 1 ? size_t(s.max_step_width) : 0)) * size_t(dc) * 8;
    };
    int dc = 0;
    if (need(32) <= kSmemCap) dc = 32;
    if (need(64) <= kSmemCap && (D + 31) / 32 > 2 * s.sm_count) dc = 64;
    if (dc) {
      if (mode == 1) RET(ensure_target_stream(c, s));
      SmallGArgs G;
      G.fwd = s.fwd_l.as<Rec>();
      G.foff = s.foff.as<int64_t>();
      G.tgt = s.tgt.as<Rec>();
      G.goff = s.goff.as<int64_t>();
      G.sched_lhs = s.sched_lhs.as<int32_t>();
      G.level_first = s.d_level_first.as<int64_t>();
      G.step_group = s.d_step_group.as<int64_t>();
      G.step_info = (mode == 0 ? s.sg_fwd_info : s.sg_rev_info).as<longlong4>();
      G.n_levels = s.n_levels;
      G.max_gradient = s.max_gradient;
      G.seed_slot = seed_slots_dev;
      G.out_slot = out_slots_dev;
      G.n_out = n_out;
      G.d_lo = d_lo;
      G.d_hi = d_hi;
      G.d_base = d_base;
      G.out = out;
      G.stride_i = stride_i;
      G.stride_d = stride_d;
      G.dc = dc;
      G.max_width = int(s.max_step_width);
      G.ref_block = c->opt_ref_block;
      const size_t smem = need(dc);
      const unsigned grid = unsigned((D + dc - 1) / dc);
      RET(mark_pass(s));
      if (c->opt_time_kernels) RET(mark(s));
      if (mode == 0) {
        CK(cudaFuncSetAttribute(smallg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        smallg_kernel<false><<<grid, 512, smem, s.stream>>>(G);
      } else {
        CK(cudaFuncSetAttribute(smallg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        smallg_kernel<true><<<grid, 512, smem, s.stream>>>(G);
      }
      if (c->opt_time_kernels) RET(mark(s));
      RET(mark_pass(s));
      s.launches += 1;
      s.sweep_launches += 1;
      s.stat["pass_dirs"] = double(D);
      s.stat["n_passes"] = 1;
      s.stat["used_schedule"] = 3.0;   // level schedule, shared-memory-resident gradients
      s.stat["warps_per_cta"] = 16;
      CK(cudaEventRecord(s.ev[1], s.stream));
      CK(cudaGetLastError());
      return 0;
    }
  }
  if (D > 0) {
    // fused reverse sweep: two rows per slot and no snapshot buffer
    const SweepPlan plan0 = plan_sweep(c, s, D);
    const bool fused = plan0.use_levels && mode == 1 && c->opt_fused && !c->opt_persistent && s.n_ops > 0 &&
                       s.max_gradient < (int64_t(1) << 30);
    const int64_t n_rows = fused ? 2 * s.max_gradient : s.max_gradient;
    // the reverse stream this sweep needs (built on first use; before the workspace is sized, since it takes memory)
    if (fused) RET(ensure_fused_stream(c, s));
    else if (plan0.use_levels && mode == 1) RET(ensure_target_stream(c, s));
    // pass size from the workspace budget
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget0 =
        c->opt_workspace_bytes > 0 ? size_t(c->opt_workspace_bytes) : size_t(double(free_b + s.g.bytes + s.abuf.bytes) * 0.6);
    // per direction: a gradient column plus (reverse, two-kernel level sweep) a column of the a-snapshot buffer
    const size_t per_dir = size_t(n_rows + (fused ? 0 : s.max_step_width)) * 8;
    size_t budget = budget0;
    if (c->opt_workspace_bytes == 0 && pool_bytes(s) > 0) {
      // would this Jacobian need fewer passes without the kept analysis temporaries?  Then give them back.
      auto passes = [&](size_t b) {
        const int64_t md = std::max<int64_t>(32, (int64_t(b / per_dir) / 32) * 32);
        return (D + md - 1) / md;
      };
      if (passes(budget + size_t(double(pool_bytes(s)) * 0.6)) < passes(budget)) {
        release_pool(s);
        CK(cudaMemGetInfo(&free_b, &total_b));
        budget = size_t(double(free_b + s.g.bytes + s.abuf.bytes) * 0.6);
      }
    }
    int64_t max_dirs = int64_t(budget / per_dir);
    max_dirs = (max_dirs / 32) * 32;
    if (max_dirs < 32) max_dirs = 32;
    int64_t pass = std::min<int64_t>(((D + 31) / 32) * 32, max_dirs);
    if (pass < D) {  // several passes: make them equal
      const int64_t np = (D + pass - 1) / pass;
      pass = std::min(pass, (((D + np - 1) / np + 63) / 64) * 64);
    }
    if (c->opt_pass_dirs > 0) pass = std::min<int64_t>(pass, ((c->opt_pass_dirs + 31) / 32) * 32);
    // doubles per lane in the level kernels: 2; 4 for the balanced forward kernel when asked for ("fwd_v4") and the tape
    // has no step that must go to forward_level_kernel (the two kernels of one sweep must agree on the block width)
    bool any_long = false;
    for (uint8_t f : s.level_long) any_long |= (f != 0);
    const int V = (mode == 0 && plan0.use_levels && c->opt_fwd_v4 && c->opt_fwd_balanced && !any_long) ? 4 : 2;
    const int64_t K = ((pass + 32 * V - 1) / (32 * V)) * (32 * V);   // rows padded so that lane vectors never leave a row
    RET(ensure(errp, s.g, size_t(n_rows) * size_t(K) * 8));
    double* g = s.g.as<double>();
    int n_passes = 0;
    const SweepPlan plan = plan_sweep(c, s, std::min(D, pass));
    const int ncb_pass = int(K / (32 * V));
    const bool sparse = plan.use_levels;
    const int32_t* out_rows_dev = out_slots_dev;
    if (sparse) {
      if (mode == 1 && !fused) RET(ensure(errp, s.abuf, size_t(s.max_step_width) * size_t(K) * 8));
      RET(ensure(errp, s.rowact, size_t(n_rows) * size_t(ncb_pass)));
      if (!fused) RET(ensure(errp, s.aflag, size_t(std::max<int64_t>(s.max_step_width, 1)) * size_t(ncb_pass)));
    }
    if (fused) {  // the swept value of slot X ends up in row X + fpar[X]*G
      RET(ensure(errp, s.out_rows, size_t(n_out) * 4));
      fused_out_rows_kernel<<<unsigned((n_out + 255) / 256), 256, 0, s.stream>>>(out_slots_dev, n_out, s.fpar.as<uint8_t>(),
                                                                                 s.max_gradient, s.out_rows.as<int32_t>());
      s.launches += 1;
      out_rows_dev = s.out_rows.as<int32_t>();
    }
    for (int64_t d0 = d_lo; d0 < d_hi; d0 += pass) {
      const int64_t n = std::min(pass, d_hi - d0);
      CK(cudaMemsetAsync(g, 0, size_t(n_rows) * size_t(K) * 8, s.stream));
      if (sparse) CK(cudaMemsetAsync(s.rowact.p, 0, size_t(n_rows) * size_t(ncb_pass), s.stream));
      seed_kernel<<<unsigned((n + 255) / 256), 256, 0, s.stream>>>(g, K, seed_slots_dev, d0, int(n),
                                                                    sparse ? s.rowact.as<uint8_t>() : nullptr, ncb_pass, 32 * V);
      s.launches += 1;
      RET(mark_pass(s));
      {
        const double l0 = s.launches;
        RET(launch_sweep(c, s, mode, g, K, n, plan, fused, V));
        s.sweep_launches += s.launches - l0;
      }
      RET(mark_pass(s));
      const dim3 eg(unsigned((n + 31) / 32), unsigned((n_out + 31) / 32));
      extract_kernel<<<eg, dim3(32, 8), 0, s.stream>>>(g, K, out_rows_dev, n_out, int(n), d0 - d_base, out,
                                                        stride_i, stride_d, fused ? s.rowact.as<uint8_t>() : nullptr,
                                                        ncb_pass, 32 * V);
      s.launches += 1;
      ++n_passes;
    }
    s.stat["pass_dirs"] = double(pass);
    s.stat["n_passes"] = double(n_passes);
    s.stat["used_schedule"] = plan.use_levels ? 2.0 : 1.0;
    s.stat["used_fused"] = fused ? 1.0 : 0.0;
    s.stat["warps_per_cta"] = double(plan.warps);
  }
  CK(cudaEventRecord(s.ev[1], s.stream));
  CK(cudaGetLastError());
  return 0;
}

int shard_finish(Shard& s) {
  std::string* errp = &s.err;
  CK(cudaSetDevice(s.dev));
  CK(cudaStreamSynchronize(s.stream));
  if (s.persist_checked) {   // the resident fused kernel: did a CTA give up waiting?
    s.persist_checked = false;
    unsigned int aborted = 0;
    CK(cudaMemcpy(&aborted, s.pabort.p, sizeof aborted, cudaMemcpyDeviceToHost));
    if (aborted) {
      CK(cudaMemset(s.pabort.p, 0, 128));
      return fail(errp, ADEPT_CUDA_ERR_CUDA, "resident fused reverse kernel: a CTA waited longer than spin_limit polls for the step above");
    }
  }
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]));
  s.stat["replay_ms"] = ms;                       // zero-fill + seed + sweep + extraction, all passes
  double sweep = 0;
  for (size_t k = 0; k + 1 < s.pev_used; k += 2) {
    CK(cudaEventElapsedTime(&ms, s.pev[k], s.pev[k + 1]));
    sweep += ms;
  }
  s.stat["sweep_ms"] = sweep;                     // the replay kernels alone, summed over all passes
  s.stat["sweep_launches"] = s.sweep_launches;    // their launch count
  // dominant kernel (forward: replay_forward_kernel, reverse: reverse_target_kernel / replay_reverse_kernel),
  // every launch timed on its own when "time_kernels" is on
  double dom = 0, dq[4] = {0, 0, 0, 0};
  const size_t n_dom = s.kev_used / 2;
  for (size_t k = 0; k + 1 < s.kev_used; k += 2) {
    CK(cudaEventElapsedTime(&ms, s.kev[k], s.kev[k + 1]));
    dom += ms;
    dq[std::min<size_t>(3, (k / 2) * 4 / std::max<size_t>(n_dom, 1))] += ms;   // by quarter of the sweep, in issue order
  }
  if (!s.prof_log.empty()) {   // option "profile_every": which launches the profiler saw, in order
    if (const char* f = getenv("ADEPT_CUDA_PROFILE_LOG"))
      if (FILE* fp = fopen(f, "a")) {
        for (const std::string& l : s.prof_log) fprintf(fp, "%s\n", l.c_str());
        fclose(fp);
      }
    s.prof_log.clear();
  }
  s.stat["dominant_ms"] = dom;
  for (int q = 0; q < 4; ++q) s.stat[std::string("dominant_q") + char('1' + q) + "_ms"] = dq[q];
  s.stat["dominant_launches"] = double(s.kev_used / 2);
  return 0;
}

// Run fn(shard index) for every local shard: inline for one shard, one host thread per shard
// otherwise (thousands of per-level launches per GPU must not be serialised behind each other).
template <class F>
int over_shards(adept_cuda_ctx* c, F fn) {
  const size_t n = c->sh.size();
  std::vector<int> rc(n, 0);
  if (n == 1) {
    rc[0] = fn(0);
  } else {
    std::vector<std::thread> th;
    for (size_t k = 0; k < n; ++k) th.emplace_back([&, k]() { rc[k] = fn(k); });
    for (std::thread& t : th) t.join();
  }
  for (size_t k = 0; k < n; ++k)
    if (rc[k]) {
      c->err = c->sh[k].err;
      return rc[k];
    }
  return 0;
}

// strided re-layout on the root: dst[i*si + d*sd] = src[d*n_out + i]
__global__ void relayout_kernel(const double* __restrict__ src, int64_t n_out, int64_t n_dirs, double* __restrict__ dst,
                                int64_t si, int64_t sd) {
  __shared__ double tile[32][33];
  const int64_t i0 = int64_t(blockIdx.x) * 32, d0 = int64_t(blockIdx.y) * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t d = d0 + r, i = i0 + threadIdx.x;
    if (d < n_dirs && i < n_out) tile[r][threadIdx.x] = src[d * n_out + i];
  }
  __syncthreads();
  if (si <= sd) {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int64_t d = d0 + r, i = i0 + threadIdx.x;
      if (d < n_dirs && i < n_out) dst[i * si + d * sd] = tile[r][threadIdx.x];
    }
  } else {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int64_t d = d0 + threadIdx.x, i = i0 + r;
      if (d < n_dirs && i < n_out) dst[i * si + d * sd] = tile[threadIdx.x][r];
    }
  }
}

// host-side scatter of a direction-major block in the reference's write order (direction
// blocks of P, then out index, then lane): only matters when user strides make elements overlap
void scatter_host(const std::vector<double>& dm, int64_t D, int64_t n_out, double* out, int64_t stride_i,
                  int64_t stride_d, int64_t P) {
  for (int64_t d0 = 0; d0 < D; d0 += P)
    for (int64_t i = 0; i < n_out; ++i)
      for (int64_t d = d0; d < std::min(D, d0 + P); ++d) out[i * stride_i + d * stride_d] = dm[size_t(d) * n_out + i];
}

int do_jacobian(adept_cuda_ctx* c, int mode, const int32_t* indep, int64_t n_indep, const int32_t* dep,
                int64_t n_dep, double* jac_out, bool out_on_device, int64_t dep_offset, int64_t indep_offset) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  if (mode != ADEPT_CUDA_FORWARD && mode != ADEPT_CUDA_REVERSE) return fail(errp, ADEPT_CUDA_ERR_ARG, "mode must be 0 or 1");
  Shard& root = c->sh[0];
  if (c->chain.active) return fail(errp, ADEPT_CUDA_ERR_ARG, "a checkpoint chain is open on this context (adept_cuda_chain_end first)");
  const bool is_root_proc = (c->rank0 == 0);
  // pre-checks first, agreed between the ranks of a one-process-per-GPU job BEFORE anybody enters a collective
  const int pre = [&]() -> int {
  if (!root.have_tape) return fail(errp, ADEPT_CUDA_ERR_NO_TAPE, "no tape uploaded");
  // reference pre-check (adept/jacobian.cpp:208-210, :466-468)
  if (n_indep <= 0 || n_dep <= 0)
    return fail(errp, ADEPT_CUDA_ERR_NOT_IDENTIFIED, "dependent or independent variables not identified");
  if (!indep || !dep) return fail(errp, ADEPT_CUDA_ERR_ARG, "null index list");
  if (is_root_proc && !jac_out) return fail(errp, ADEPT_CUDA_ERR_ARG, "null output");
  for (int64_t i = 0; i < n_indep; ++i)
    if (indep[i] < 0 || indep[i] >= root.max_gradient) return fail(errp, ADEPT_CUDA_ERR_RANGE, "independent slot out of range");
  for (int64_t i = 0; i < n_dep; ++i)
    if (dep[i] < 0 || dep[i] >= root.max_gradient) return fail(errp, ADEPT_CUDA_ERR_RANGE, "dependent slot out of range");
  return over_shards(c, [&](size_t k) -> int { return ensure_schedule(c, c->sh[k]); });   // built on first need
  }();
  RET(agree(c, pre));
  // offset defaulting exactly as adept/jacobian.cpp:215-220
  if (dep_offset <= 0) dep_offset = n_indep;
  if (indep_offset <= 0) indep_offset = n_dep;

  const int64_t D = (mode == 0) ? n_indep : n_dep;          // directions
  const int64_t n_out = (mode == 0) ? n_dep : n_indep;      // entries per direction
  const int32_t* seeds = (mode == 0) ? indep : dep;
  const int32_t* outs = (mode == 0) ? dep : indep;
  const int64_t stride_i = (mode == 0) ? dep_offset : indep_offset;  // user stride of the out-list index
  const int64_t stride_d = (mode == 0) ? indep_offset : dep_offset;  // user stride of the direction index

  // direction partition over all ranks, in multiples of 32 (keeps the reference's P-blocks whole)
  const int W = c->world;
  int64_t blk = (D + W - 1) / W;
  blk = ((blk + 31) / 32) * 32;
  auto lo_of = [&](int r) { return std::min<int64_t>(D, int64_t(r) * blk); };

  auto t_begin = std::chrono::steady_clock::now();
  // index lists to every local shard
  for (Shard& s : c->sh) {
    CK(cudaSetDevice(s.dev));
    RET(ensure(errp, s.seed_slots, size_t(D) * 4));
    RET(ensure(errp, s.out_slots, size_t(n_out) * 4));
    CK(cudaMemcpyAsync(s.seed_slots.p, seeds, size_t(D) * 4, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.out_slots.p, outs, size_t(n_out) * 4, cudaMemcpyHostToDevice, s.stream));
  }

  const bool compact_user = (stride_i == 1 && stride_d == n_out) || (stride_d == 1 && stride_i == D);
  const bool dm_is_final = (stride_i == 1 && stride_d == n_out);
  if (W == 1) {
    // single GPU: extract straight into the final layout
    Shard& s = root;
    double* dev_out = nullptr;
    int64_t si = stride_i, sd = stride_d;
    if (out_on_device) {
      dev_out = jac_out;
    } else {
      RET(ensure(errp, s.out, size_t(D) * size_t(n_out) * 8));
      dev_out = s.out.as<double>();
      if (!compact_user) { si = 1; sd = n_out; }
    }
    int rc = shard_enqueue(c, s, mode, s.seed_slots.as<int32_t>(), s.out_slots.as<int32_t>(), n_out, 0, D, 0, dev_out, si, sd);
    if (!rc) rc = shard_finish(s);
    if (rc) { c->err = s.err; return rc; }
    if (!out_on_device) {
      auto t0 = std::chrono::steady_clock::now();
      if (compact_user) {
        CK(cudaMemcpyAsync(jac_out, dev_out, size_t(D) * size_t(n_out) * 8, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
      } else {
        std::vector<double> tmp(size_t(D) * size_t(n_out));
        CK(cudaMemcpyAsync(tmp.data(), dev_out, tmp.size() * 8, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        scatter_host(tmp, D, n_out, jac_out, stride_i, stride_d, c->opt_ref_block);
      }
      c->stat["d2h_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
  } else {
    if (!c->multi_process && !out_on_device && compact_user) {
      // one process, host output in one of the two compact layouts: every device extracts its block of directions
      // in the FINAL layout and copies it straight into the caller's buffer over its own PCIe link -- no gather
      // through device 0, no re-layout pass
      RET(over_shards(c, [&](size_t k) -> int {
        Shard& s = c->sh[k];
        std::string* errp = &s.err;
        const int64_t lo = lo_of(int(k)), hi = lo_of(int(k) + 1), nd = hi - lo;
        CK(cudaSetDevice(s.dev));
        if (nd <= 0) { s.stat["replay_ms"] = 0; s.stat["sweep_ms"] = 0; return 0; }
        RET(ensure(errp, s.out, size_t(nd) * size_t(n_out) * 8));
        const int64_t si = dm_is_final ? 1 : nd, sd = dm_is_final ? n_out : 1;
        RET(shard_enqueue(c, s, mode, s.seed_slots.as<int32_t>(), s.out_slots.as<int32_t>(), n_out, lo, hi, lo,
                          s.out.as<double>(), si, sd));
        if (dm_is_final)
          CK(cudaMemcpyAsync(jac_out + lo * n_out, s.out.p, size_t(nd) * n_out * 8, cudaMemcpyDeviceToHost, s.stream));
        else
          CK(cudaMemcpy2DAsync(jac_out + lo, size_t(D) * 8, s.out.p, size_t(nd) * 8, size_t(nd) * 8, size_t(n_out),
                               cudaMemcpyDeviceToHost, s.stream));
        return shard_finish(s);
      }));
      c->stat["d2h_ms"] = 0.0;   // overlapped with the other devices' sweeps
      double replay_ms = 0, sweep_ms = 0;
      for (Shard& s : c->sh) {
        replay_ms = std::max(replay_ms, s.stat["replay_ms"]);
        sweep_ms = std::max(sweep_ms, s.stat["sweep_ms"]);
      }
      for (const auto& kv : root.stat) c->stat[kv.first] = kv.second;
      c->stat["replay_ms"] = replay_ms;
      c->stat["sweep_ms"] = sweep_ms;
      c->stat["jacobian_wall_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
      c->stat["kernel_launches"] = c->total_launches();
      return 0;
    }
    // multi GPU, general case: every rank fills a direction-major block dm[(d-lo)*n_out + i]; the blocks are
    // gathered on global rank 0 (ncclSend/ncclRecv over NVLink) and re-laid out there.
    const int enq = over_shards(c, [&](size_t k) -> int {
      Shard& s = c->sh[k];
      std::string* errp = &s.err;
      const int r = c->rank0 + int(k);
      const int64_t lo = lo_of(r), hi = lo_of(r + 1);
      CK(cudaSetDevice(s.dev));
      const bool is_root = (r == 0);
      RET(ensure(errp, s.out, size_t(is_root ? D : std::max<int64_t>(hi - lo, 1)) * size_t(n_out) * 8));
      return shard_enqueue(c, s, mode, s.seed_slots.as<int32_t>(), s.out_slots.as<int32_t>(), n_out, lo, hi,
                           is_root ? 0 : lo, s.out.as<double>(), 1, n_out);
    });
    RET(agree(c, enq));   // nobody enters the gather unless everybody can
    {
      ncclResult_t nr = ncclSuccess;
      NKG(nr, ncclGroupStart());
      for (size_t k = 0; k < c->sh.size(); ++k) {
        Shard& s = c->sh[k];
        const int r = c->rank0 + int(k);
        if (r == 0) {
          for (int q = 1; q < W; ++q) {
            const int64_t lo = lo_of(q), hi = lo_of(q + 1);
            if (hi > lo)
              NKG(nr, ncclRecv(s.out.as<double>() + lo * n_out, size_t(hi - lo) * n_out, ncclDouble, q, s.comm, s.stream));
          }
        } else {
          const int64_t lo = lo_of(r), hi = lo_of(r + 1);
          if (hi > lo) NKG(nr, ncclSend(s.out.as<double>(), size_t(hi - lo) * n_out, ncclDouble, 0, s.comm, s.stream));
        }
      }
      NKG(nr, ncclGroupEnd());
      if (nr != ncclSuccess) return fail(errp, ADEPT_CUDA_ERR_NCCL, "gather of the Jacobian blocks failed: %s", ncclGetErrorString(nr));
    }
    RET(over_shards(c, [&](size_t k) -> int { return shard_finish(c->sh[k]); }));
    if (is_root_proc) {
      Shard& s = root;
      CK(cudaSetDevice(s.dev));
      auto t0 = std::chrono::steady_clock::now();
      double* final_dev = s.out.as<double>();
      if (out_on_device) {
        if (dm_is_final) {
          CK(cudaMemcpyAsync(jac_out, s.out.p, size_t(D) * n_out * 8, cudaMemcpyDeviceToDevice, s.stream));
        } else {
          const dim3 g(unsigned((n_out + 31) / 32), unsigned((D + 31) / 32));
          relayout_kernel<<<g, dim3(32, 8), 0, s.stream>>>(s.out.as<double>(), n_out, D, jac_out, stride_i, stride_d);
          s.launches += 1;
        }
        CK(cudaStreamSynchronize(s.stream));
      } else {
        if (!dm_is_final && compact_user) {
          RET(ensure(errp, s.scratch, size_t(D) * n_out * 8));
          const dim3 g(unsigned((n_out + 31) / 32), unsigned((D + 31) / 32));
          relayout_kernel<<<g, dim3(32, 8), 0, s.stream>>>(s.out.as<double>(), n_out, D, s.scratch.as<double>(),
                                                           stride_i, stride_d);
          s.launches += 1;
          final_dev = s.scratch.as<double>();
        }
        if (dm_is_final || compact_user) {
          CK(cudaMemcpyAsync(jac_out, final_dev, size_t(D) * n_out * 8, cudaMemcpyDeviceToHost, s.stream));
          CK(cudaStreamSynchronize(s.stream));
        } else {
          std::vector<double> tmp(size_t(D) * size_t(n_out));
          CK(cudaMemcpyAsync(tmp.data(), s.out.p, tmp.size() * 8, cudaMemcpyDeviceToHost, s.stream));
          CK(cudaStreamSynchronize(s.stream));
          scatter_host(tmp, D, n_out, jac_out, stride_i, stride_d, c->opt_ref_block);
        }
      }
      c->stat["d2h_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
  }
  // this process's view: the slowest local shard
  double replay_ms = 0, sweep_ms = 0;
  for (Shard& s : c->sh) {
    replay_ms = std::max(replay_ms, s.stat["replay_ms"]);
    sweep_ms = std::max(sweep_ms, s.stat["sweep_ms"]);
  }
  for (const auto& kv : root.stat) c->stat[kv.first] = kv.second;
  c->stat["replay_ms"] = replay_ms;
  c->stat["sweep_ms"] = sweep_ms;
  c->stat["jacobian_wall_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  c->stat["kernel_launches"] = c->total_launches();
  return 0;
}


// Enqueue (no host synchronisation) one single-direction sweep of the device vector g[max_gradient] on the
// shard's stream.  adjoint: Stack::compute_adjoint (adept/Stack.cpp:139-164), else compute_tangent_linear (:170-190).
int enqueue_sweep1(adept_cuda_ctx* c, Shard& s, int adjoint, double* g, int flags, bool* ordered_out) {
  std::string* errp = &s.err;
  const bool ordered = (flags & ADEPT_CUDA_SWEEP_ORDERED) || !s.have_levels || c->opt_schedule == 1;
  if (ordered_out) *ordered_out = ordered;
  if (s.R <= 0) return 0;
  if (ordered) {
    RET(ensure_tape_streams(s));
    ReplayArgs A{};
    A.g = g;
    A.K = 1;
    A.n_colblocks = 1;
    A.n_dirs = 1;  // K = 1: lanes 1..31 are masked off
    A.ref_block = 1;
    A.recs = (adjoint ? s.rev_t : s.fwd_t).as<Rec>();
    A.chunk_begin = s.one_chunk.as<int64_t>();
    A.first_chunk = 0;
    if (adjoint) replay_reverse_kernel<false><<<dim3(1, 1), 32, 0, s.stream>>>(A);
    else replay_forward_kernel<true, false><<<dim3(1, 1), 32, 0, s.stream>>>(A);
    s.launches += 1;
  } else {
    const Rec* recs = s.fwd_l.as<Rec>();
    const int64_t* foff = s.foff.as<int64_t>();
    if (!adjoint) {
      size_t lf = 0;  // cursor into long_stmts (sorted by schedule position, hence by step)
      for (int32_t l = 1; l <= s.n_levels; ++l) {
        const int64_t j0 = s.level_first[l], j1 = s.level_first[l + 1];
        if (j1 > j0) {
          sweep1_forward_level_kernel<<<unsigned((j1 - j0 + 255) / 256), 256, 0, s.stream>>>(recs, foff, j0, j1,
                                                                                             kLongThreshold, g);
          s.launches += 1;
        }
        size_t le = lf;
        while (le < s.long_stmts.size() && s.long_stmts[le].level == l) ++le;
        if (le > lf) {   // one CTA per long statement of this step
          sweep1_forward_long_kernel<<<unsigned(le - lf), 512, 0, s.stream>>>(recs, s.long_tab.as<int64_t>() + 2 * lf, g);
          s.launches += 1;
          lf = le;
        }
      }
    } else {
      RET(ensure_target_stream(c, s));
      RET(ensure(errp, s.a1, size_t(std::max<int64_t>(s.max_step_width, 1)) * 8));
      for (int32_t l = s.n_levels; l >= 1; --l) {
        const int64_t j0 = s.level_first[l], n = s.level_first[l + 1] - j0;
        if (n <= 0) continue;
        sweep1_snapshot_kernel<<<unsigned((n + 255) / 256), 256, 0, s.stream>>>(s.sched_lhs.as<int32_t>(), j0, n, g,
                                                                                s.a1.as<double>());
        s.launches += 1;
        const int64_t g0 = s.step_group[l], g1 = s.step_group[l + 1];
        if (g1 > g0) {
          sweep1_target_kernel<<<unsigned((g1 - g0 + 255) / 256), 256, 0, s.stream>>>(
              s.tgt.as<Rec>(), s.goff.as<int64_t>(), g0, g1, g, s.a1.as<double>());
          s.launches += 1;
        }
      }
    }
  }
  CK(cudaGetLastError());
  return 0;
}

// single-direction sweep on shard 0: gradient vector in/out on the host (or in place on the device)
int do_sweep1(adept_cuda_ctx* c, int adjoint, double* gradient_host, int initialized, int flags, bool on_device = false) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  Shard& s = c->sh[0];
  if (c->chain.active) return fail(errp, ADEPT_CUDA_ERR_ARG, "a checkpoint chain is open on this context (adept_cuda_chain_end first)");
  if (!initialized) return fail(errp, ADEPT_CUDA_ERR_GRAD_NOT_INIT, "gradients not initialized");
  if (!s.have_tape) return fail(errp, ADEPT_CUDA_ERR_NO_TAPE, "no tape uploaded");
  if (!gradient_host) return fail(errp, ADEPT_CUDA_ERR_ARG, "null gradient vector");
  CK(cudaSetDevice(s.dev));
  if (int rc = ensure_schedule(c, s)) { c->err = s.err; return rc; }
  const int64_t G = s.max_gradient;
  double* g = gradient_host;   // device-resident gradient vector: swept in place
  if (!on_device) {
    RET(ensure(errp, s.g1, size_t(G) * 8));
    g = s.g1.as<double>();
    if (int rc = upload(s, g, gradient_host, size_t(G) * 8)) { c->err = s.err; return rc; }
  }
  CK(cudaEventRecord(s.ev[0], s.stream));
  bool ordered = false;
  if (int rc = enqueue_sweep1(c, s, adjoint, g, flags, &ordered)) { c->err = s.err; return rc; }
  CK(cudaEventRecord(s.ev[1], s.stream));
  CK(cudaGetLastError());
  if (!on_device) CK(cudaMemcpyAsync(gradient_host, g, size_t(G) * 8, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]));
  c->stat["replay_ms"] = ms;
  c->stat["sweep_ms"] = ms;
  c->stat["used_schedule"] = ordered ? 1.0 : 2.0;
  c->stat["kernel_launches"] = c->total_launches();
  return 0;
}

int init_shard(Shard& s, int dev) {
  std::string* errp = &s.err;
  s.dev = dev;
  CK(cudaSetDevice(dev));
  CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) CK(cudaEventCreate(&s.ev[k]));
  CK(cudaEventCreateWithFlags(&s.fork_ev, cudaEventDisableTiming));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  s.sm_count = prop.multiProcessorCount;
  if (prop.major < 10)
    return fail(errp, ADEPT_CUDA_ERR_CUDA, "device %d is sm_%d%d; libadept_cuda is built for sm_100a only", dev, prop.major,
                prop.minor);
  return 0;
}

void free_shard(Shard& s) {
  cudaSetDevice(s.dev);
  for (Buf* b : {&s.stmt, &s.mult, &s.idx, &s.fwd_t, &s.rev_t, &s.one_chunk, &s.level, &s.fwd_l, &s.fwd_chunk,
                 &s.foff, &s.long_tab, &s.sched_lhs, &s.tgt, &s.tchunk, &s.goff, &s.pos_in_sched, &s.ftgt, &s.fchunk, &s.fpar, &s.out_rows, &s.done, &s.d_level_first, &s.d_step_tchunk, &s.d_step_group, &s.sg_fwd_info, &s.sg_rev_info, &s.abuf, &s.rowact, &s.aflag, &s.a1, &s.g, &s.seed_slots, &s.out_slots, &s.out, &s.g1, &s.scratch, &s.agree_buf})
    release(*b);
  for (int k = 0; k < 2; ++k) {
    if (s.stage[k]) cudaFreeHost(s.stage[k]);
    if (s.stage_ev[k]) cudaEventDestroy(s.stage_ev[k]);
  }
  for (int k = 0; k < 2; ++k)
    if (s.ev[k]) cudaEventDestroy(s.ev[k]);
  for (int q = 0; q < 2; ++q) {
    release(s.bg[q]); release(s.bmult[q]); release(s.bout[q]);
    if (s.b_in[q]) cudaEventDestroy(s.b_in[q]);
    if (s.b_done[q]) cudaEventDestroy(s.b_done[q]);
    if (s.b_out[q]) cudaEventDestroy(s.b_out[q]);
  }
  if (s.mult_stream) cudaStreamDestroy(s.mult_stream);
  if (s.mult_ev) cudaEventDestroy(s.mult_ev);
  if (s.bcopy) cudaStreamDestroy(s.bcopy);
  if (s.bback) cudaStreamDestroy(s.bback);
  for (cudaEvent_t e : s.kev) cudaEventDestroy(e);
  for (cudaEvent_t e : s.pev) cudaEventDestroy(e);
  for (cudaEvent_t e : s.aux_ev) cudaEventDestroy(e);
  for (cudaStream_t st : s.aux) cudaStreamDestroy(st);
  if (s.fork_ev) cudaEventDestroy(s.fork_ev);
  for (Buf& b : s.pool) release(b);
  if (s.comm) ncclCommDestroy(s.comm);
  if (s.stream) cudaStreamDestroy(s.stream);
}

// ---------------------------------------------------------------------------------------
// Checkpoint-chained replay
// ---------------------------------------------------------------------------------------
// set_gradients semantics (include/adept/Stack.h:290-306): values are stored one after the other, so when a slot is
// named twice the LAST value wins.  owner[slot] = largest i naming it (atomicMax), then thread i stores iff it owns.
__global__ void chain_owner_kernel(const int32_t* __restrict__ slots, int64_t n, int32_t* __restrict__ owner) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) atomicMax(&owner[slots[i]], static_cast<int32_t>(i));
}
__global__ void chain_scatter_kernel(const int32_t* __restrict__ slots, const double* __restrict__ vals, int64_t n,
                                     const int32_t* __restrict__ owner, double* __restrict__ g) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && owner[slots[i]] == static_cast<int32_t>(i)) g[slots[i]] = vals[i];
}
__global__ void chain_gather_kernel(const int32_t* __restrict__ slots, int64_t n, const double* __restrict__ g,
                                    double* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = g[slots[i]];
}

int chain_issue_upload(adept_cuda_ctx* c, ChainBlock& b) {
  Chain& ch = c->chain;
  Shard& s = c->sh[0];
  std::string* errp = &s.err;
  const int k = b.slot;
  const size_t bytes[6] = {size_t(b.n_stmt) * 8, size_t(b.n_ops) * 8, size_t(b.n_ops) * 4, size_t(b.n_seed) * 4,
                           b.seed_from_handoff ? 0 : size_t(b.n_seed) * 8, size_t(b.n_out) * 4};
  Buf* dst[6] = {&ch.dtape[k][0], &ch.dtape[k][1], &ch.dtape[k][2], &ch.dseed_slots[k], &ch.dseed_vals[k], &ch.dout_slots[k]};
  for (int a = 0; a < 6; ++a) {
    RET(ensure(errp, *dst[a], bytes[a]));
    if (bytes[a]) CK(cudaMemcpyAsync(dst[a]->p, ch.hs[k].p[a], bytes[a], cudaMemcpyHostToDevice, ch.copy));
  }
  CK(cudaEventRecord(ch.h2d_done[k], ch.copy));
  b.uploaded = true;
  return 0;
}

// one block on the worker thread: (upload) -> schedule -> seed -> sweep -> gather; the next block's upload is issued
// before this block's sweep is waited for
int chain_run_block(adept_cuda_ctx* c, ChainBlock b) {
  Chain& ch = c->chain;
  Shard& s = c->sh[0];
  std::string* errp = &s.err;
  CK(cudaSetDevice(s.dev));
  auto t0 = std::chrono::steady_clock::now();
  if (!b.uploaded) RET(chain_issue_upload(c, b));
  CK(cudaStreamWaitEvent(s.stream, ch.h2d_done[b.slot], 0));
  // the shard's tape buffers and this slot's trade places: the previous block's raw arrays are dead (its sweep only
  // reads the streams derived from them, and those are rebuilt below)
  std::swap(s.stmt, ch.dtape[b.slot][0]);
  std::swap(s.mult, ch.dtape[b.slot][1]);
  std::swap(s.idx, ch.dtape[b.slot][2]);
  s.n_stmt = b.n_stmt;
  s.n_ops = b.n_ops;
  s.max_gradient = b.max_gradient;
  s.have_tape = true;
  s.have_schedule = false;
  auto ta = std::chrono::steady_clock::now();
  RET(build_schedule(c, s));
  ch.analysis_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ta).count();
  const int64_t G = b.max_gradient;
  RET(ensure(errp, s.g1, size_t(G) * 8));
  RET(ensure(errp, ch.owner, size_t(G) * 4));
  double* g = s.g1.as<double>();
  CK(cudaMemsetAsync(g, 0, size_t(G) * 8, s.stream));                 // initialize_gradients: all zero (adept/Stack.cpp:515-531)
  CK(cudaMemsetAsync(ch.owner.p, 0xFF, size_t(G) * 4, s.stream));
  if (b.n_seed > 0) {
    const unsigned gs = unsigned((b.n_seed + 255) / 256);
    const int32_t* slots = ch.dseed_slots[b.slot].as<int32_t>();
    const double* vals = b.seed_from_handoff ? ch.handoff.as<double>() : ch.dseed_vals[b.slot].as<double>();
    chain_owner_kernel<<<gs, 256, 0, s.stream>>>(slots, b.n_seed, ch.owner.as<int32_t>());
    chain_scatter_kernel<<<gs, 256, 0, s.stream>>>(slots, vals, b.n_seed, ch.owner.as<int32_t>(), g);
    s.launches += 2;
  }
  CK(cudaEventRecord(s.ev[0], s.stream));
  RET(enqueue_sweep1(c, s, 1, g, 0, nullptr));
  CK(cudaEventRecord(s.ev[1], s.stream));
  chain_gather_kernel<<<unsigned((b.n_out + 255) / 256), 256, 0, s.stream>>>(ch.dout_slots[b.slot].as<int32_t>(), b.n_out, g,
                                                                            ch.handoff.as<double>());
  s.launches += 1;
  CK(cudaGetLastError());
  {  // the next block, if it is already waiting: start its upload now, behind nothing but the copy engine
    std::unique_lock<std::mutex> lk(ch.mu);
    if (!ch.q.empty() && !ch.q.front().uploaded) {
      ChainBlock& nb = ch.q.front();
      lk.unlock();          // only this thread pops; the caller only appends
      RET(chain_issue_upload(c, nb));
    }
  }
  CK(cudaStreamSynchronize(s.stream));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]));
  ch.sweep_ms += ms;
  ch.busy_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  ch.blocks += 1;
  return 0;
}

void chain_worker(adept_cuda_ctx* c) {
  Chain& ch = c->chain;
  for (;;) {
    ChainBlock b;
    {
      std::unique_lock<std::mutex> lk(ch.mu);
      ch.cv.wait(lk, [&] { return ch.stop || !ch.q.empty(); });
      if (ch.q.empty()) return;
      b = ch.q.front();
      ch.q.pop_front();
    }
    int rc = 0;
    {
      std::unique_lock<std::mutex> lk(ch.mu);
      rc = ch.rc;
    }
    if (!rc) {
      rc = chain_run_block(c, b);
      if (rc) {
        std::unique_lock<std::mutex> lk(ch.mu);
        ch.rc = rc;
        ch.err = c->sh[0].err;
      }
    }
    {
      std::unique_lock<std::mutex> lk(ch.mu);
      ch.hs[b.slot].busy = false;
    }
    ch.cv.notify_all();
  }
}

void chain_shutdown(adept_cuda_ctx* c) {
  Chain& ch = c->chain;
  if (ch.worker.joinable()) {
    {
      std::unique_lock<std::mutex> lk(ch.mu);
      ch.stop = true;
    }
    ch.cv.notify_all();
    ch.worker.join();
  }
  ch.stop = false;
  ch.active = false;
}

void chain_free(adept_cuda_ctx* c) {
  chain_shutdown(c);
  Chain& ch = c->chain;
  cudaSetDevice(c->sh[0].dev);
  for (int k = 0; k < 2; ++k) {
    for (int a = 0; a < 3; ++a) release(ch.dtape[k][a]);
    release(ch.dseed_slots[k]);
    release(ch.dseed_vals[k]);
    release(ch.dout_slots[k]);
    for (int a = 0; a < 6; ++a)
      if (ch.hs[k].p[a]) { cudaFreeHost(ch.hs[k].p[a]); ch.hs[k].p[a] = nullptr; ch.hs[k].cap[a] = 0; }
    if (ch.h2d_done[k]) { cudaEventDestroy(ch.h2d_done[k]); ch.h2d_done[k] = nullptr; }
  }
  release(ch.handoff);
  release(ch.owner);
  if (ch.copy) { cudaStreamDestroy(ch.copy); ch.copy = nullptr; }
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char* adept_cuda_version(void) { return "adept-2_b200 0.1 (sm_100a)"; }

const char* adept_cuda_last_error(const adept_cuda_ctx* c) { return c ? c->err.c_str() : "null context"; }

int adept_cuda_create(adept_cuda_ctx** out, const int* devices, int n_devices) {
  if (!out || n_devices < 1) return ADEPT_CUDA_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count < 1) return ADEPT_CUDA_ERR_CUDA;  // no CPU fallback by design
  adept_cuda_ctx* c = new adept_cuda_ctx();
  c->sh.resize(n_devices);
  for (int k = 0; k < n_devices; ++k) {
    const int dev = devices ? devices[k] : k;
    if (dev < 0 || dev >= count) { delete c; return ADEPT_CUDA_ERR_ARG; }
    int rc = init_shard(c->sh[k], dev);
    if (rc) {
      fprintf(stderr, "adept_cuda_create: %s\n", c->sh[k].err.c_str());
      for (Shard& s : c->sh) free_shard(s);
      delete c;
      return rc;
    }
  }
  c->world = n_devices;
  c->rank0 = 0;
  if (n_devices > 1) {
    std::vector<ncclComm_t> comms(n_devices);
    std::vector<int> devs(n_devices);
    for (int k = 0; k < n_devices; ++k) devs[k] = c->sh[k].dev;
    ncclResult_t r = ncclCommInitAll(comms.data(), n_devices, devs.data());
    if (r != ncclSuccess) {
      fprintf(stderr, "adept_cuda_create: ncclCommInitAll failed: %s\n", ncclGetErrorString(r));
      for (Shard& s : c->sh) free_shard(s);
      delete c;
      return ADEPT_CUDA_ERR_NCCL;
    }
    for (int k = 0; k < n_devices; ++k) c->sh[k].comm = comms[k];
  }
  // $ADEPT_CUDA_OPTIONS = "key=value,key=value": option defaults for callers that only see the Stack API (the drop-in)
  if (const char* env = getenv("ADEPT_CUDA_OPTIONS")) {
    std::string all(env);
    size_t pos = 0;
    while (pos < all.size()) {
      size_t end = all.find(',', pos);
      if (end == std::string::npos) end = all.size();
      const std::string kv = all.substr(pos, end - pos);
      pos = end + 1;
      const size_t eq = kv.find('=');
      if (kv.empty()) continue;
      int rc = ADEPT_CUDA_ERR_ARG;
      if (eq != std::string::npos && eq > 0 && eq + 1 < kv.size()) {
        char* endp = nullptr;
        const long long v = strtoll(kv.c_str() + eq + 1, &endp, 10);
        if (endp && *endp == 0) rc = adept_cuda_set_option(c, kv.substr(0, eq).c_str(), int64_t(v));
      }
      if (rc) {
        fprintf(stderr, "adept_cuda_create: bad entry '%s' in $ADEPT_CUDA_OPTIONS: %s\n", kv.c_str(), c->err.c_str());
        adept_cuda_destroy(c);
        return rc;
      }
    }
  }
  *out = c;
  return 0;
}

void adept_cuda_destroy(adept_cuda_ctx* c) {
  if (!c) return;
  chain_free(c);
  for (Shard& s : c->sh) free_shard(s);
  delete c;
}

int adept_cuda_unique_id(void* id128) {
  if (!id128) return ADEPT_CUDA_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return ADEPT_CUDA_ERR_NCCL;
  std::memcpy(id128, &id, 128);
  return 0;
}

int adept_cuda_join(adept_cuda_ctx* c, const void* id128, int rank, int world_size) {
  if (!c || !id128 || rank < 0 || rank >= world_size) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  if (c->sh.size() != 1) return fail(errp, ADEPT_CUDA_ERR_ARG, "adept_cuda_join needs a one-device context");
  Shard& s = c->sh[0];
  CK(cudaSetDevice(s.dev));
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  if (s.comm) { ncclCommDestroy(s.comm); s.comm = nullptr; }
  NK(ncclCommInitRank(&s.comm, world_size, id, rank));
  c->world = world_size;
  c->rank0 = rank;
  c->multi_process = true;
  return 0;
}

int adept_cuda_upload_tape(adept_cuda_ctx* c, const void* statements, int64_t n_stmt, const double* multiplier,
                           const int32_t* index, int64_t n_ops, int64_t max_gradient, uint64_t epoch) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  if (!statements || n_stmt < 1 || n_ops < 0 || max_gradient < 1 || (n_ops > 0 && (!multiplier || !index)))
    return fail(errp, ADEPT_CUDA_ERR_ARG, "bad tape arguments");
  if (max_gradient >= (int64_t(1) << 31)) return fail(errp, ADEPT_CUDA_ERR_LIMIT, "max_gradient >= 2^31");
  if (c->chain.active) return fail(errp, ADEPT_CUDA_ERR_ARG, "a checkpoint chain is open on this context (adept_cuda_chain_end first)");
  Shard& s = c->sh[0];
  CK(cudaSetDevice(s.dev));
  auto t0 = std::chrono::steady_clock::now();
  for (Shard& d : c->sh) { d.have_tape = false; d.have_schedule = false; }
  RET(ensure(errp, s.stmt, size_t(n_stmt) * 8));
  RET(ensure(errp, s.mult, size_t(n_ops) * 8));
  RET(ensure(errp, s.idx, size_t(n_ops) * 4));
  // The level analysis reads only statement_ and index_: when the multipliers sit in page-locked memory (true DMA)
  // and this context analyses right here, they travel on a stream of their own WHILE the analysis runs; the
  // shard's stream waits for them just before the first kernel that reads them (settle_multipliers).
  bool overlap = false;
  if (n_ops > 0 && !c->multi_process && c->sh.size() == 1 && !c->opt_remove_null && c->opt_overlap_upload) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, multiplier) == cudaSuccess) overlap = (at.type == cudaMemoryTypeHost);
    else cudaGetLastError();
  }
  s.mult_pending = false;
  {
    int rc = upload(s, s.stmt.p, statements, size_t(n_stmt) * 8);
    if (!rc) rc = upload(s, s.idx.p, index, size_t(n_ops) * 4);
    if (!rc && !overlap) rc = upload(s, s.mult.p, multiplier, size_t(n_ops) * 8);
    if (rc) { c->err = s.err; return rc; }
  }
  if (overlap) {
    if (!s.mult_stream) {
      CK(cudaStreamCreateWithFlags(&s.mult_stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&s.mult_ev, cudaEventDisableTiming));
    }
    CK(cudaMemcpyAsync(s.mult.p, multiplier, size_t(n_ops) * 8, cudaMemcpyHostToDevice, s.mult_stream));
    CK(cudaEventRecord(s.mult_ev, s.mult_stream));
    s.mult_pending = true;
  }
  c->stat["upload_overlapped"] = overlap ? 1.0 : 0.0;
  CK(cudaStreamSynchronize(s.stream));
  s.n_stmt = n_stmt;
  s.n_ops = n_ops;
  s.n_ops_recorded = n_ops;
  s.max_gradient = max_gradient;
  s.have_tape = true;
  c->epoch = epoch;
  auto t1 = std::chrono::steady_clock::now();
  c->stat["upload_ms"] = std::chrono::duration<double, std::milli>(t1 - t0).count();
  // single-process multi-GPU: broadcast right away
  if (!c->multi_process && c->sh.size() > 1) {
    for (size_t k = 1; k < c->sh.size(); ++k) {
      Shard& d = c->sh[k];
      CK(cudaSetDevice(d.dev));
      RET(ensure(errp, d.stmt, size_t(n_stmt) * 8));
      RET(ensure(errp, d.mult, size_t(n_ops) * 8));
      RET(ensure(errp, d.idx, size_t(n_ops) * 4));
      d.n_stmt = n_stmt; d.n_ops = n_ops; d.max_gradient = max_gradient;
    }
    ncclResult_t nr = ncclSuccess;
    NKG(nr, ncclGroupStart());
    for (Shard& d : c->sh) {
      NKG(nr, ncclBroadcast(d.stmt.p, d.stmt.p, size_t(n_stmt) * 8, ncclChar, 0, d.comm, d.stream));
      NKG(nr, ncclBroadcast(d.mult.p, d.mult.p, size_t(n_ops) * 8, ncclChar, 0, d.comm, d.stream));
      NKG(nr, ncclBroadcast(d.idx.p, d.idx.p, size_t(n_ops) * 4, ncclChar, 0, d.comm, d.stream));
    }
    NKG(nr, ncclGroupEnd());
    if (nr != ncclSuccess) return fail(errp, ADEPT_CUDA_ERR_NCCL, "tape broadcast failed: %s", ncclGetErrorString(nr));
    for (Shard& d : c->sh) { CK(cudaSetDevice(d.dev)); CK(cudaStreamSynchronize(d.stream)); d.have_tape = true; }
    c->stat["bcast_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
  }
  // One-process-per-GPU job: the schedule is built by adept_cuda_share_tape on all ranks at once (or on first use).
  // Otherwise: now, every local shard on its own host thread.
  if (!c->multi_process) {
    auto t2 = std::chrono::steady_clock::now();
    const int brc = over_shards(c, [&](size_t k) -> int { return ensure_schedule(c, c->sh[k]); });
    if (s.mult_pending) {   // (only when the build failed before reading them) the caller's array must not be in use on return
      cudaStreamSynchronize(s.mult_stream);
      s.mult_pending = false;
    }
    RET(brc);
    c->stat["analysis_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count();
    publish_schedule_stats(c);
    c->stat["analysis_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count();
  } else {
    c->stat["analysis_ms"] = 0.0;
  }
  c->stat["upload_total_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

int adept_cuda_share_tape(adept_cuda_ctx* c, int root_rank) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  if (!c->multi_process) return 0;  // single-process contexts broadcast inside upload_tape
  if (root_rank < 0 || root_rank >= c->world) return fail(errp, ADEPT_CUDA_ERR_ARG, "root_rank out of range");
  Shard& s = c->sh[0];
  CK(cudaSetDevice(s.dev));
  auto t0 = std::chrono::steady_clock::now();
  const bool is_root = (c->rank0 == root_rank);
  // header first: {status, n_stmt, n_ops, max_gradient, epoch}.  A root without a tape says so in the header,
  // so that its peers (already inside the broadcast) leave with the same error instead of waiting for arrays.
  RET(ensure(errp, s.scratch, 64));
  int64_t hdr[5] = {is_root && !s.have_tape ? int64_t(ADEPT_CUDA_ERR_NO_TAPE) : 0, s.n_stmt, s.n_ops, s.max_gradient,
                    int64_t(c->epoch)};
  if (is_root) CK(cudaMemcpyAsync(s.scratch.p, hdr, sizeof hdr, cudaMemcpyHostToDevice, s.stream));
  NK(ncclBroadcast(s.scratch.p, s.scratch.p, sizeof hdr, ncclChar, root_rank, s.comm, s.stream));
  CK(cudaMemcpyAsync(hdr, s.scratch.p, sizeof hdr, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  if (hdr[0]) return fail(errp, int(hdr[0]), "rank %d has no tape to share", root_rank);
  int rc = 0;
  if (!is_root) {
    s.have_tape = false;
    s.have_schedule = false;
    s.n_stmt = hdr[1]; s.n_ops = hdr[2]; s.max_gradient = hdr[3]; c->epoch = uint64_t(hdr[4]);
    rc = ensure(errp, s.stmt, size_t(s.n_stmt) * 8);
    if (!rc) rc = ensure(errp, s.mult, size_t(s.n_ops) * 8);
    if (!rc) rc = ensure(errp, s.idx, size_t(s.n_ops) * 4);
  }
  RET(agree(c, rc));   // an allocation that failed on one rank must not leave the others in the broadcast
  ncclResult_t nr = ncclSuccess;
  NKG(nr, ncclGroupStart());
  NKG(nr, ncclBroadcast(s.stmt.p, s.stmt.p, size_t(s.n_stmt) * 8, ncclChar, root_rank, s.comm, s.stream));
  NKG(nr, ncclBroadcast(s.mult.p, s.mult.p, size_t(s.n_ops) * 8, ncclChar, root_rank, s.comm, s.stream));
  NKG(nr, ncclBroadcast(s.idx.p, s.idx.p, size_t(s.n_ops) * 4, ncclChar, root_rank, s.comm, s.stream));
  NKG(nr, ncclGroupEnd());
  if (nr != ncclSuccess) return fail(errp, ADEPT_CUDA_ERR_NCCL, "tape broadcast failed: %s", ncclGetErrorString(nr));
  CK(cudaStreamSynchronize(s.stream));
  s.have_tape = true;
  c->stat["bcast_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  // every rank analyses its copy NOW, at the same time (the root did not do it in upload_tape)
  auto t2 = std::chrono::steady_clock::now();
  rc = ensure_schedule(c, s);
  if (rc) c->err = s.err;
  RET(agree(c, rc));   // e.g. a malformed tape: refused by every rank alike
  publish_schedule_stats(c);
  c->stat["analysis_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count();
  return 0;
}

int adept_cuda_tape_epoch(const adept_cuda_ctx* c, uint64_t* epoch, int64_t* n_stmt, int64_t* n_ops) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  if (!c->sh[0].have_tape) return ADEPT_CUDA_ERR_NO_TAPE;
  if (epoch) *epoch = c->epoch;
  if (n_stmt) *n_stmt = c->sh[0].n_stmt;
  if (n_ops) *n_ops = c->sh[0].n_ops_recorded;
  return 0;
}

int adept_cuda_jacobian(adept_cuda_ctx* c, int mode, const int32_t* indep, int64_t n_indep, const int32_t* dep,
                        int64_t n_dep, double* jac_out_host, int64_t dep_offset, int64_t indep_offset) {
  return do_jacobian(c, mode, indep, n_indep, dep, n_dep, jac_out_host, false, dep_offset, indep_offset);
}

int adept_cuda_jacobian_device(adept_cuda_ctx* c, int mode, const int32_t* indep, int64_t n_indep, const int32_t* dep,
                               int64_t n_dep, double* jac_out_device, int64_t dep_offset, int64_t indep_offset) {
  return do_jacobian(c, mode, indep, n_indep, dep, n_dep, jac_out_device, true, dep_offset, indep_offset);
}

int adept_cuda_tangent_linear(adept_cuda_ctx* c, double* gradient_host, int initialized, int flags) {
  return do_sweep1(c, 0, gradient_host, initialized, flags);
}

int adept_cuda_adjoint(adept_cuda_ctx* c, double* gradient_host, int initialized, int flags) {
  return do_sweep1(c, 1, gradient_host, initialized, flags);
}

int adept_cuda_tangent_linear_device(adept_cuda_ctx* c, double* gradient_device, int flags) {
  return do_sweep1(c, 0, gradient_device, 1, flags, true);
}

int adept_cuda_adjoint_device(adept_cuda_ctx* c, double* gradient_device, int flags) {
  return do_sweep1(c, 1, gradient_device, 1, flags, true);
}

int adept_cuda_jacobian_batch(adept_cuda_ctx* c, int mode, const void* statements, int64_t n_stmt,
                              const int32_t* index, int64_t n_ops, int64_t max_gradient, const double* multipliers,
                              int64_t n_members, const int32_t* indep, int64_t n_indep, const int32_t* dep,
                              int64_t n_dep, double* jac_out_host) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  if (mode != ADEPT_CUDA_FORWARD && mode != ADEPT_CUDA_REVERSE) return fail(errp, ADEPT_CUDA_ERR_ARG, "mode must be 0 or 1");
  if (n_indep <= 0 || n_dep <= 0)
    return fail(errp, ADEPT_CUDA_ERR_NOT_IDENTIFIED, "dependent or independent variables not identified");
  if (!statements || n_stmt < 1 || n_ops < 0 || max_gradient < 1 || n_members < 1 || !indep || !dep || !jac_out_host ||
      (n_ops > 0 && (!index || !multipliers)))
    return fail(errp, ADEPT_CUDA_ERR_ARG, "bad ensemble arguments");
  for (int64_t i = 0; i < n_indep; ++i)
    if (indep[i] < 0 || indep[i] >= max_gradient) return fail(errp, ADEPT_CUDA_ERR_RANGE, "independent slot out of range");
  for (int64_t i = 0; i < n_dep; ++i)
    if (dep[i] < 0 || dep[i] >= max_gradient) return fail(errp, ADEPT_CUDA_ERR_RANGE, "dependent slot out of range");
  const int64_t D = (mode == 0) ? n_indep : n_dep;
  const int64_t n_out = (mode == 0) ? n_dep : n_indep;
  const int32_t* seeds = (mode == 0) ? indep : dep;
  const int32_t* outs = (mode == 0) ? dep : indep;
  const int P = c->opt_ref_block;
  const int Dp = int(((D + P - 1) / P) * P);  // the reference's P-lane blocks stay inside one member
  const int n_sh = int(c->sh.size());
  const int64_t per = (n_members + n_sh - 1) / n_sh;
  auto t_begin = std::chrono::steady_clock::now();
  RET(over_shards(c, [&](size_t k) -> int {
    Shard& s = c->sh[k];
    std::string* errp = &s.err;
    const int64_t b_lo = std::min<int64_t>(n_members, int64_t(k) * per), b_hi = std::min<int64_t>(n_members, b_lo + per);
    if (b_hi <= b_lo) return 0;
    CK(cudaSetDevice(s.dev));
    // the shared structure replaces whatever tape this context held
    s.have_tape = false;
    s.have_schedule = false;
    s.have_levels = false;
    s.have_tape_streams = false;
    RET(ensure(errp, s.stmt, size_t(n_stmt) * 8));
    RET(ensure(errp, s.idx, size_t(n_ops) * 4));
    RET(upload(s, s.stmt.p, statements, size_t(n_stmt) * 8));
    RET(upload(s, s.idx.p, index, size_t(n_ops) * 4));
    s.n_stmt = n_stmt;
    s.n_ops = n_ops;
    s.max_gradient = max_gradient;
    const int64_t S = n_stmt - 1, R = n_ops + S;
    s.R = R;
    RET(ensure(errp, s.scratch, 64));
    CK(cudaMemsetAsync(s.scratch.p, 0, 64, s.stream));
    validate_kernel<<<grid_for(std::max(n_ops, n_stmt), 256, s.sm_count), 256, 0, s.stream>>>(
        s.stmt.as<int2>(), n_stmt, s.idx.as<int32_t>(), nullptr, n_ops, max_gradient, s.scratch.as<int>());
    s.launches += 1;
    int flags = 0;
    CK(cudaMemcpyAsync(&flags, s.scratch.p, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    if (flags) return fail(errp, ADEPT_CUDA_ERR_RANGE, "malformed tape structure");
    const int64_t one_chunk_host[2] = {0, R};
    RET(ensure(errp, s.one_chunk, sizeof one_chunk_host));
    CK(cudaMemcpyAsync(s.one_chunk.p, one_chunk_host, sizeof one_chunk_host, cudaMemcpyHostToDevice, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    RET(ensure(errp, s.fwd_t, size_t(std::max<int64_t>(R, 1)) * sizeof(Rec)));
    RET(ensure(errp, s.rev_t, size_t(std::max<int64_t>(R, 1)) * sizeof(Rec)));
    if (R > 0) {
      build_streams_kernel<<<grid_for(std::max(n_ops, S), 256, s.sm_count), 256, 0, s.stream>>>(
          s.stmt.as<int2>(), n_stmt, nullptr, s.idx.as<int32_t>(), n_ops, s.fwd_t.as<Rec>(), s.rev_t.as<Rec>(), nullptr);
      s.launches += 1;
    }
    RET(ensure(errp, s.seed_slots, size_t(D) * 4));
    RET(ensure(errp, s.out_slots, size_t(n_out) * 4));
    CK(cudaMemcpyAsync(s.seed_slots.p, seeds, size_t(D) * 4, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.out_slots.p, outs, size_t(n_out) * 4, cudaMemcpyHostToDevice, s.stream));
    // Members are processed in chunks through a two-deep pipeline: while chunk k is swept on the shard's stream, the
    // multipliers of chunk k+1 are copied in on the copy stream and the Jacobians of chunk k-1 are copied out on the
    // third one (two buffer sets).  With pinned caller memory (adept_cuda_host_alloc) the three overlap; pageable
    // memory still works, the copies then stage through the driver.
    // chunk size from the workspace budget: per member max_gradient*Dp gradients + n_ops multipliers + outputs, x2 sets
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    size_t held = s.mult.bytes;
    for (int q = 0; q < 2; ++q) held += s.bg[q].bytes + s.bmult[q].bytes + s.bout[q].bytes;
    const size_t budget = c->opt_workspace_bytes > 0 ? size_t(c->opt_workspace_bytes) : size_t(double(free_b + held) * 0.6);
    const size_t per_member = 2 * (size_t(max_gradient) * Dp * 8 + size_t(n_ops) * 8 + size_t(n_dep) * n_indep * 8);
    const int64_t n_mine = b_hi - b_lo;
    int64_t nb_chunk = std::max<int64_t>(1, int64_t(budget / per_member));
    nb_chunk = std::min<int64_t>(nb_chunk, (int64_t(65535) * 8 * 32) / Dp);  // gridDim.y limit
    const int64_t want_chunks = std::max<int64_t>(1, std::min<int64_t>(c->opt_batch_chunks, n_mine / 256));   // >= 256 members per chunk
    nb_chunk = std::min<int64_t>(nb_chunk, (n_mine + want_chunks - 1) / want_chunks);
    nb_chunk = std::max<int64_t>(nb_chunk, 1);
    if (!s.bcopy) {
      CK(cudaStreamCreateWithFlags(&s.bcopy, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&s.bback, cudaStreamNonBlocking));
      for (int q = 0; q < 2; ++q) {
        CK(cudaEventCreateWithFlags(&s.b_in[q], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.b_done[q], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.b_out[q], cudaEventDisableTiming));
      }
    }
    CK(cudaStreamSynchronize(s.stream));   // structure, streams and slot lists are in place before the copy stream starts
    int64_t ck = 0;
    for (int64_t b0 = b_lo; b0 < b_hi; b0 += nb_chunk, ++ck) {
      const int q = int(ck & 1);
      const int64_t nb = std::min(nb_chunk, b_hi - b0);
      const int64_t lanes = nb * Dp;
      const int64_t K = ((lanes + 31) / 32) * 32;
      if (ck >= 2) {   // this buffer set was used by chunk ck-2: its sweep and its copy-out must be over
        CK(cudaEventSynchronize(s.b_out[q]));
      }
      RET(ensure(errp, s.bg[q], size_t(max_gradient) * size_t(K) * 8));
      RET(ensure(errp, s.bmult[q], size_t(std::max<int64_t>(n_ops, 1)) * size_t(nb) * 8));
      RET(ensure(errp, s.bout[q], size_t(nb) * n_dep * n_indep * 8));
      if (n_ops > 0)
        CK(cudaMemcpy2DAsync(s.bmult[q].p, size_t(nb) * 8, multipliers + b0, size_t(n_members) * 8, size_t(nb) * 8,
                             size_t(n_ops), cudaMemcpyHostToDevice, s.bcopy));
      CK(cudaEventRecord(s.b_in[q], s.bcopy));
      CK(cudaMemsetAsync(s.bg[q].p, 0, size_t(max_gradient) * size_t(K) * 8, s.stream));
      batch_seed_kernel<<<unsigned((nb * D + 255) / 256), 256, 0, s.stream>>>(s.bg[q].as<double>(), K, s.seed_slots.as<int32_t>(),
                                                                            int(D), Dp, nb);
      s.launches += 1;
      CK(cudaStreamWaitEvent(s.stream, s.b_in[q], 0));
      if (R > 0) {
        ReplayArgs A{};
        A.recs = (mode == 0 ? s.fwd_t : s.rev_t).as<Rec>();
        A.chunk_begin = s.one_chunk.as<int64_t>();
        A.first_chunk = 0;
        A.g = s.bg[q].as<double>();
        A.K = K;
        A.n_colblocks = int(K / 32);
        A.n_dirs = lanes;
        A.ref_block = P;
        A.bmult = s.bmult[q].as<double>();
        A.n_members = nb;
        A.dirs_padded = Dp;
        const int warps = 8;
        const unsigned gy = unsigned((A.n_colblocks + warps - 1) / warps);
        if (mode == 0) replay_forward_kernel<true, true><<<dim3(1, gy), warps * 32, 0, s.stream>>>(A);
        else replay_reverse_kernel<true><<<dim3(1, gy), warps * 32, 0, s.stream>>>(A);
        s.launches += 1;
      }
      const int64_t total = nb * n_out * D;
      batch_extract_kernel<<<unsigned((total + 255) / 256), 256, 0, s.stream>>>(
          s.bg[q].as<double>(), K, s.out_slots.as<int32_t>(), int(n_out), int(D), Dp, nb, mode, n_dep, n_indep,
          s.bout[q].as<double>());
      s.launches += 1;
      CK(cudaGetLastError());
      CK(cudaEventRecord(s.b_done[q], s.stream));
      CK(cudaStreamWaitEvent(s.bback, s.b_done[q], 0));
      CK(cudaMemcpyAsync(jac_out_host + b0 * n_dep * n_indep, s.bout[q].p, size_t(nb) * n_dep * n_indep * 8,
                         cudaMemcpyDeviceToHost, s.bback));
      CK(cudaEventRecord(s.b_out[q], s.bback));
    }
    CK(cudaStreamSynchronize(s.bback));
    CK(cudaStreamSynchronize(s.stream));
    s.stat["batch_chunks"] = double(ck);
    return 0;
  }));
  c->stat["batch_chunks"] = c->sh[0].stat["batch_chunks"];
  c->stat["batch_wall_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  c->stat["kernel_launches"] = c->total_launches();
  return 0;
}

void* adept_cuda_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void adept_cuda_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int adept_cuda_chain_begin(adept_cuda_ctx* c, int64_t n_handoff, const double* handoff_host) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  if (n_handoff < 1) return fail(errp, ADEPT_CUDA_ERR_ARG, "chain: the hand-off vector needs at least one element");
  Chain& ch = c->chain;
  if (ch.active) return fail(errp, ADEPT_CUDA_ERR_ARG, "chain: already open (call adept_cuda_chain_end first)");
  Shard& s = c->sh[0];
  CK(cudaSetDevice(s.dev));
  if (!ch.copy) CK(cudaStreamCreateWithFlags(&ch.copy, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k)
    if (!ch.h2d_done[k]) CK(cudaEventCreateWithFlags(&ch.h2d_done[k], cudaEventDisableTiming));
  RET(ensure(errp, ch.handoff, size_t(n_handoff) * 8));
  if (handoff_host) CK(cudaMemcpy(ch.handoff.p, handoff_host, size_t(n_handoff) * 8, cudaMemcpyHostToDevice));
  else CK(cudaMemset(ch.handoff.p, 0, size_t(n_handoff) * 8));
  ch.n_handoff = n_handoff;
  ch.rc = 0;
  ch.err.clear();
  ch.blocks = 0;
  ch.busy_ms = ch.sweep_ms = ch.analysis_ms = 0;
  ch.q.clear();
  ch.hs[0].busy = ch.hs[1].busy = false;
  ch.stop = false;
  ch.active = true;
  ch.worker = std::thread(chain_worker, c);
  return 0;
}

int adept_cuda_chain_push(adept_cuda_ctx* c, const void* statements, int64_t n_stmt, const double* multiplier,
                          const int32_t* index, int64_t n_ops, int64_t max_gradient, const int32_t* seed_slots,
                          int64_t n_seed, const double* seed_values, const int32_t* out_slots, int64_t n_out) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  Chain& ch = c->chain;
  if (!ch.active) return fail(errp, ADEPT_CUDA_ERR_ARG, "chain: not open (call adept_cuda_chain_begin first)");
  if (!statements || n_stmt < 1 || n_ops < 0 || max_gradient < 1 || (n_ops > 0 && (!multiplier || !index)))
    return fail(errp, ADEPT_CUDA_ERR_ARG, "bad tape arguments");
  if (max_gradient >= (int64_t(1) << 31)) return fail(errp, ADEPT_CUDA_ERR_LIMIT, "max_gradient >= 2^31");
  if (n_seed < 0 || (n_seed > 0 && !seed_slots) || !out_slots) return fail(errp, ADEPT_CUDA_ERR_ARG, "chain: null slot list");
  if (n_out != ch.n_handoff)
    return fail(errp, ADEPT_CUDA_ERR_ARG, "chain: a block hands %lld gradients on, the chain was opened for %lld", (long long)n_out,
                (long long)ch.n_handoff);
  if (!seed_values && n_seed != ch.n_handoff)
    return fail(errp, ADEPT_CUDA_ERR_ARG, "chain: seeding from the hand-off vector needs n_seed == n_handoff");
  for (int64_t i = 0; i < n_seed; ++i)
    if (seed_slots[i] < 0 || seed_slots[i] >= max_gradient) return fail(errp, ADEPT_CUDA_ERR_RANGE, "chain: seed slot out of range");
  for (int64_t i = 0; i < n_out; ++i)
    if (out_slots[i] < 0 || out_slots[i] >= max_gradient) return fail(errp, ADEPT_CUDA_ERR_RANGE, "chain: output slot out of range");
  // a free pinned slot (at most two blocks are in flight: one being swept, one waiting or uploading)
  int k = -1;
  {
    std::unique_lock<std::mutex> lk(ch.mu);
    ch.cv.wait(lk, [&] { return ch.rc != 0 || !ch.hs[0].busy || !ch.hs[1].busy; });
    if (ch.rc) { c->err = ch.err; return ch.rc; }
    k = !ch.hs[0].busy ? 0 : 1;
    ch.hs[k].busy = true;
  }
  CK(cudaSetDevice(c->sh[0].dev));
  const void* src[6] = {statements, multiplier, index, seed_slots, seed_values, out_slots};
  const size_t bytes[6] = {size_t(n_stmt) * 8, size_t(n_ops) * 8, size_t(n_ops) * 4, size_t(n_seed) * 4,
                           seed_values ? size_t(n_seed) * 8 : 0, size_t(n_out) * 4};
  for (int a = 0; a < 6; ++a) {
    if (!bytes[a]) continue;
    ChainHostSlot& h = ch.hs[k];
    if (h.cap[a] < bytes[a]) {
      if (h.p[a]) CK(cudaFreeHost(h.p[a]));
      h.p[a] = nullptr;
      h.cap[a] = 0;
      const size_t want = bytes[a] + bytes[a] / 4 + 4096;
      cudaError_t e = cudaMallocHost(&h.p[a], want);
      if (e != cudaSuccess) {
        std::unique_lock<std::mutex> lk(ch.mu);
        h.busy = false;
        return fail(errp, ADEPT_CUDA_ERR_CUDA, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
      }
      h.cap[a] = want;
    }
    // host copy into the pinned slot, split over a few threads for long tapes (one core moves ~10 GB/s)
    const int nt = bytes[a] >= (size_t(16) << 20) ? 4 : 1;
    const size_t part = (bytes[a] + nt - 1) / nt;
    char* d = static_cast<char*>(h.p[a]);
    const char* sp = static_cast<const char*>(src[a]);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) {
      const size_t o = size_t(t) * part;
      if (o < bytes[a]) th.emplace_back([=]() { std::memcpy(d + o, sp + o, std::min(part, bytes[a] - o)); });
    }
    std::memcpy(d, sp, std::min(part, bytes[a]));
    for (std::thread& t : th) t.join();
  }
  ChainBlock b;
  b.slot = k;
  b.n_stmt = n_stmt;
  b.n_ops = n_ops;
  b.max_gradient = max_gradient;
  b.n_seed = n_seed;
  b.n_out = n_out;
  b.seed_from_handoff = (seed_values == nullptr);
  {
    std::unique_lock<std::mutex> lk(ch.mu);
    ch.q.push_back(b);
  }
  ch.cv.notify_all();
  return 0;
}

int adept_cuda_chain_end(adept_cuda_ctx* c, double* handoff_host) {
  if (!c) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  Chain& ch = c->chain;
  if (!ch.active) return fail(errp, ADEPT_CUDA_ERR_ARG, "chain: not open");
  {  // drain
    std::unique_lock<std::mutex> lk(ch.mu);
    ch.cv.wait(lk, [&] { return ch.q.empty() && !ch.hs[0].busy && !ch.hs[1].busy; });
  }
  chain_shutdown(c);
  c->stat["chain_blocks"] = double(ch.blocks);
  c->stat["chain_busy_ms"] = ch.busy_ms;
  c->stat["chain_sweep_ms"] = ch.sweep_ms;
  c->stat["chain_analysis_ms"] = ch.analysis_ms;
  c->stat["kernel_launches"] = c->total_launches();
  c->epoch = 0;
  if (ch.rc) { c->err = ch.err; return ch.rc; }
  CK(cudaSetDevice(c->sh[0].dev));
  if (handoff_host) CK(cudaMemcpy(handoff_host, ch.handoff.p, size_t(ch.n_handoff) * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int adept_cuda_set_option(adept_cuda_ctx* c, const char* key, int64_t value) {
  if (!c || !key) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  const std::string k(key);
  if (k == "schedule") {
    if (value < 0 || value > 2) return fail(errp, ADEPT_CUDA_ERR_ARG, "schedule must be 0, 1 or 2");
    c->opt_schedule = int(value);
  } else if (k == "ref_block") {
    if (value < 1 || value > 32 || (value & (value - 1))) return fail(errp, ADEPT_CUDA_ERR_ARG, "ref_block must be a power of two <= 32");
    c->opt_ref_block = int(value);
  } else if (k == "pass_dirs") {
    if (value < 0) return fail(errp, ADEPT_CUDA_ERR_ARG, "pass_dirs must be >= 0");
    c->opt_pass_dirs = value;
  } else if (k == "workspace_bytes") {
    if (value < 0) return fail(errp, ADEPT_CUDA_ERR_ARG, "workspace_bytes must be >= 0");
    c->opt_workspace_bytes = value;
  } else if (k == "pdl") {
    c->opt_pdl = value ? 1 : 0;
  } else if (k == "streams") {
    if (value < 1 || value > 8) return fail(errp, ADEPT_CUDA_ERR_ARG, "streams must be in 1..8");
    c->opt_streams = int(value);
  } else if (k == "smallg") {
    c->opt_smallg = value ? 1 : 0;
  } else if (k == "fused") {
    c->opt_fused = value ? 1 : 0;
  } else if (k == "fused_persistent") {
    c->opt_fused_persistent = value > 0 ? 1 : (value < 0 ? -1 : 0);
  } else if (k == "auto_split") {
    c->opt_auto_split = value ? 1 : 0;
  } else if (k == "split") {
    if (value < 0 || value > 16) return fail(errp, ADEPT_CUDA_ERR_ARG, "split must be in 0..16");
    c->opt_split = int(value);
  } else if (k == "sub") {
    if (value != 0 && (value < 4 || value > 256)) return fail(errp, ADEPT_CUDA_ERR_ARG, "sub must be 0 or in 4..256");
    c->opt_sub = int(value);
  } else if (k == "spin_limit") {
    if (value < 64 || value > (int64_t(1) << 31) - 1) return fail(errp, ADEPT_CUDA_ERR_ARG, "spin_limit must be in 64 .. 2^31-1");
    c->opt_spin_limit = value;
  } else if (k == "tail") {
    c->opt_tail = value ? 1 : 0;
  } else if (k == "fwd_v4") {
    c->opt_fwd_v4 = value ? 1 : 0;
  } else if (k == "overlap_upload") {
    c->opt_overlap_upload = value ? 1 : 0;
  } else if (k == "coop_pack") {
    c->opt_coop_pack = value ? 1 : 0;
  } else if (k == "batch_chunks") {
    if (value < 1 || value > 64) return fail(errp, ADEPT_CUDA_ERR_ARG, "batch_chunks must be in 1..64");
    c->opt_batch_chunks = int(value);
  } else if (k == "wide_streams") {
    if (value < 0 || value > 8) return fail(errp, ADEPT_CUDA_ERR_ARG, "wide_streams must be in 0..8");
    c->opt_wide_streams = int(value);
  } else if (k == "fwd_balanced") {
    c->opt_fwd_balanced = value ? 1 : 0;
  } else if (k == "remove_null") {
    c->opt_remove_null = value ? 1 : 0;
  } else if (k == "profile_every") {
    if (value < 0) return fail(errp, ADEPT_CUDA_ERR_ARG, "profile_every must be >= 0");
    c->opt_profile_every = int(value);
  } else if (k == "cbs") {
    if (value < 0 || value > kFusedCbs) return fail(errp, ADEPT_CUDA_ERR_ARG, "cbs must be in 0..32");
    c->opt_cbs = int(value);
  } else if (k == "chunk_recs") {
    if (value < 0 || value > 256) return fail(errp, ADEPT_CUDA_ERR_ARG, "chunk_recs must be in 0..256");
    c->opt_chunk_recs = value;
  } else if (k == "persistent") {
    c->opt_persistent = value ? 1 : 0;
  } else if (k == "time_kernels") {
    c->opt_time_kernels = value ? 1 : 0;
  } else if (k == "window") {
    if (value < 0 || (value > 0 && value < 32)) return fail(errp, ADEPT_CUDA_ERR_ARG, "window must be 0 (default) or >= 32");
    c->opt_window = value;
  } else if (k == "warps_per_cta") {
    if (value < 1 || value > 8) return fail(errp, ADEPT_CUDA_ERR_ARG, "warps_per_cta must be in 1..8");
    c->opt_warps = int(value);
  } else {
    return fail(errp, ADEPT_CUDA_ERR_ARG, "unknown option '%s'", key);
  }
  return 0;
}

int adept_cuda_get_stat(const adept_cuda_ctx* c, const char* key, double* value) {
  if (!c || !key || !value) return ADEPT_CUDA_ERR_ARG;
  auto it = c->stat.find(key);
  if (it == c->stat.end()) return ADEPT_CUDA_ERR_ARG;
  *value = it->second;
  return 0;
}

int adept_cuda_get_levels(adept_cuda_ctx* c, int32_t* level_out, int64_t n_stmt) {
  if (!c || !level_out) return ADEPT_CUDA_ERR_ARG;
  std::string* errp = &c->err;
  Shard& s = c->sh[0];
  if (!s.have_tape) return fail(errp, ADEPT_CUDA_ERR_NO_TAPE, "no tape uploaded");
  if (int rc = ensure_schedule(c, s)) { c->err = s.err; return rc; }
  if (!s.have_levels) return fail(errp, ADEPT_CUDA_ERR_ARG, "no level schedule was built for this tape");
  if (n_stmt != s.n_stmt) return fail(errp, ADEPT_CUDA_ERR_ARG, "n_stmt mismatch");
  CK(cudaSetDevice(s.dev));
  CK(cudaMemcpy(level_out, s.level.p, size_t(n_stmt) * 4, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
