/* adept_cuda.h -- C-ABI of libadept_cuda: B200-native replay of an Adept-2 derivative tape.
 *
 * This library replaces the REPLAY half of adept::Stack -- the bodies of
 *   Stack::compute_adjoint          (reference adept/Stack.cpp:139-164)
 *   Stack::compute_tangent_linear   (reference adept/Stack.cpp:170-190)
 *   Stack::jacobian_forward[_openmp](reference adept/jacobian.cpp:127-315)
 *   Stack::jacobian_reverse[_openmp](reference adept/jacobian.cpp:331-647)
 * -- with hand-written sm_100a kernels.  Recording stays on the host, unchanged.
 * The reference has no FFI of its own: the seam is the set of out-of-line Stack
 * members above; the C++ shim that forwards them to these entry points is
 * adept-2_b200/dropin/stack_replay_cuda.cpp.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or
 * an ADEPT_CUDA_ERR_* code and never throws; adept_cuda_last_error() gives the text.
 * All entry points are synchronous (results are complete on return), like the
 * reference's.  A context may be used by one host thread at a time; use one
 * context per adept::Stack / per host thread (reference threading model:
 * include/adept/Stack.h:52-61).  There is NO CPU fallback: without a usable
 * CUDA device adept_cuda_create fails.
 */
#ifndef ADEPT_CUDA_H
#define ADEPT_CUDA_H 1

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct adept_cuda_ctx adept_cuda_ctx;

/* status codes */
#define ADEPT_CUDA_OK                   0
#define ADEPT_CUDA_ERR_ARG              1   /* bad argument                                         */
#define ADEPT_CUDA_ERR_CUDA             2   /* CUDA runtime failure (text in last_error)            */
#define ADEPT_CUDA_ERR_NCCL             3   /* NCCL failure                                         */
#define ADEPT_CUDA_ERR_NO_TAPE          4   /* replay requested before adept_cuda_upload_tape       */
#define ADEPT_CUDA_ERR_RANGE            5   /* slot index outside [0,max_gradient) (reference:
                                               gradient_out_of_range, include/adept/Stack.h:296-298)*/
#define ADEPT_CUDA_ERR_LIMIT            6   /* tape exceeds an engine limit (documented in DESIGN.md)*/
#define ADEPT_CUDA_ERR_NOT_IDENTIFIED  10   /* reference: dependents_or_independents_not_identified
                                               (adept/jacobian.cpp:208-210, :466-468)               */
#define ADEPT_CUDA_ERR_GRAD_NOT_INIT   11   /* reference: gradients_not_initialized
                                               (adept/Stack.cpp:161-163, :187-189)                  */

/* Jacobian sweep direction */
#define ADEPT_CUDA_FORWARD 0   /* Stack::jacobian_forward: one direction per independent */
#define ADEPT_CUDA_REVERSE 1   /* Stack::jacobian_reverse: one direction per dependent   */

/* adept_cuda_adjoint / adept_cuda_tangent_linear flags */
#define ADEPT_CUDA_SWEEP_ORDERED  1  /* bit0: strictly tape-ordered kernel (bit-exact, slow for long tapes) */
/* default (flags==0): level-scheduled kernel built from the device-side dependency analysis. */

/* ---- lifetime ---------------------------------------------------------------------- */

/* Create a context on the given CUDA devices of THIS process (n_devices >= 1; devices==NULL
   means device 0..n_devices-1).  With n_devices > 1 the Jacobian directions are partitioned
   over the devices: the tape is ncclBroadcast from devices[0] and result blocks are gathered
   back on devices[0] (one process, ncclCommInitAll). */
int  adept_cuda_create(adept_cuda_ctx** ctx, const int* devices, int n_devices);
void adept_cuda_destroy(adept_cuda_ctx* ctx);

/* One-process-per-GPU deployments (torchrun): rank 0 calls adept_cuda_unique_id, ships the
   128 opaque bytes to every rank by any means, then every rank calls adept_cuda_join on a
   context created with ONE device.  After that upload_tape on rank 0 + adept_cuda_share_tape
   on all ranks broadcast the tape; adept_cuda_jacobian* is then collective: every rank
   computes its contiguous block of directions and rank 0 receives the whole matrix. */
int  adept_cuda_unique_id(void* id128);
int  adept_cuda_join(adept_cuda_ctx* ctx, const void* id128, int rank, int world_size);
int  adept_cuda_share_tape(adept_cuda_ctx* ctx, int root_rank);
/* Collective calls (share_tape, jacobian*) agree on their status before any data-path collective: an error on one rank
   (a root without a tape, a slot out of range, a malformed tape, a failed allocation) is returned by EVERY rank; no
   rank is left waiting inside a broadcast or a gather. */

/* ---- the tape ------------------------------------------------------------------------ */

/* Upload one recording.  `statements` is the reference's internal::Statement array
   (include/adept/Statement.h:24-30: {int32 index; int32 end_plus_one}) of n_stmt entries
   INCLUDING the {-1,0} sentinel pushed by new_recording (include/adept/Stack.h:528-543);
   `multiplier`/`index` are the operation stack (include/adept/StackStorageOrig.h:155-158),
   n_ops entries; max_gradient is Stack::max_gradient_.  The arrays are only read during the
   call (staged through pinned memory, cudaMemcpyAsync); pinned caller memory is DMA'd
   directly.  `epoch` is an opaque caller tag (e.g. a recording counter) returned by
   adept_cuda_tape_epoch so a shim can tell whether its host tape changed.
   Also validates the tape and builds the level schedule and the forward record stream on the device (the reverse streams
   are built by the first reverse sweep).  Multipliers in page-locked memory (adept_cuda_host_alloc, cudaHostRegister)
   are uploaded on a stream of their own WHILE the analysis runs -- it reads only statements and indices -- and are
   complete on return like everything else.  In a one-process-per-GPU job (adept_cuda_join) the schedule is NOT built
   here but by adept_cuda_share_tape, on all ranks at the same time (or by the first replay call). */
int  adept_cuda_upload_tape(adept_cuda_ctx* ctx, const void* statements, int64_t n_stmt,
                            const double* multiplier, const int32_t* index, int64_t n_ops,
                            int64_t max_gradient, uint64_t epoch);
int  adept_cuda_tape_epoch(const adept_cuda_ctx* ctx, uint64_t* epoch, int64_t* n_stmt, int64_t* n_ops);

/* ---- replay -------------------------------------------------------------------------- */

/* Jacobian of the uploaded tape.  indep/dep are Stack::independent_index_/dependent_index_
   (include/adept/Stack.h:788-789).  Output element (i = dependent, j = independent) is
   written to jac_out[i*dep_offset + j*indep_offset]; an offset <= 0 means "size of the other
   dimension" exactly as adept/jacobian.cpp:215-220.  Bit-identical to the reference built
   without FMA contraction.  jac_out is HOST memory in adept_cuda_jacobian and DEVICE memory
   (on devices[0]) in adept_cuda_jacobian_device. */
int  adept_cuda_jacobian(adept_cuda_ctx* ctx, int mode,
                         const int32_t* indep, int64_t n_indep,
                         const int32_t* dep, int64_t n_dep,
                         double* jac_out_host, int64_t dep_offset, int64_t indep_offset);
int  adept_cuda_jacobian_device(adept_cuda_ctx* ctx, int mode,
                                const int32_t* indep, int64_t n_indep,
                                const int32_t* dep, int64_t n_dep,
                                double* jac_out_device, int64_t dep_offset, int64_t indep_offset);

/* Single-direction sweeps on the caller's gradient vector (Stack::gradient_, max_gradient
   doubles, in/out).  initialized==0 reproduces the reference's pre-check and returns
   ADEPT_CUDA_ERR_GRAD_NOT_INIT without touching anything. */
int  adept_cuda_tangent_linear(adept_cuda_ctx* ctx, double* gradient_host, int initialized, int flags);
int  adept_cuda_adjoint(adept_cuda_ctx* ctx, double* gradient_host, int initialized, int flags);

/* The same sweeps IN PLACE on a gradient vector that already lives on devices[0] (max_gradient doubles): no
   host<->device copies.  For optimiser loops that keep the gradient on the GPU between iterations (the callers of
   BASELINE config 5: adept/minimize_limited_memory_bfgs.cpp:115-139, adept/line_search.cpp:45) and for
   checkpoint-chained replay (test/test_checkpoint.cpp:209-225).  Synchronous: the sweep has finished on return. */
int  adept_cuda_tangent_linear_device(adept_cuda_ctx* ctx, double* gradient_device, int flags);
int  adept_cuda_adjoint_device(adept_cuda_ctx* ctx, double* gradient_device, int flags);

/* ---- checkpoint-chained replay ----------------------------------------------------------------
   reference test/test_checkpoint.cpp:166-225: a long simulation is differentiated block by block, last block first.
   Per block the user re-records the block (new_recording ... ), seeds the block's OUTPUT variables with the gradients
   the previous block produced at ITS inputs (adept::set_gradients), calls Stack::reverse() and reads the gradients at
   the block's inputs (adept::get_gradients): many short tapes, each swept exactly once.  Done call by call through
   adept_cuda_upload_tape + adept_cuda_adjoint every block pays a blocking upload and two max_gradient-long PCIe copies.

   The chain keeps the hand-off vector (n_handoff doubles) ON THE DEVICE and pipelines the blocks:
     adept_cuda_chain_begin(ctx, n, h0)   opens a chain; the hand-off vector starts as h0 (NULL: zeros)
     adept_cuda_chain_push(ctx, tape..., seed_slots, n_seed, seed_values, out_slots, n_out)
         one block: gradient vector = 0; g[seed_slots[i]] = seed_values[i] (seed_values == NULL: the hand-off vector,
         n_seed == n_handoff; a slot named twice keeps the LAST value, as set_gradients would); reverse sweep
         (Stack::compute_adjoint, adept/Stack.cpp:139-164, level-scheduled, same bits); hand-off[i] = g[out_slots[i]]
         (n_out == n_handoff).  Returns as soon as the block has been copied to a pinned staging slot: the caller may
         re-record into the same arrays at once.  A worker thread owns the device side: block k+1 is uploaded (own copy
         stream, second device tape slot) while block k is being swept.  Errors of earlier blocks surface at the next
         push or at chain_end.
     adept_cuda_chain_end(ctx, h)         waits for the last block and copies the hand-off vector to h (NULL: discard)
   No other call may be made on the context while a chain is open. */
int  adept_cuda_chain_begin(adept_cuda_ctx* ctx, int64_t n_handoff, const double* handoff_host);
int  adept_cuda_chain_push(adept_cuda_ctx* ctx, const void* statements, int64_t n_stmt,
                           const double* multiplier, const int32_t* index, int64_t n_ops, int64_t max_gradient,
                           const int32_t* seed_slots, int64_t n_seed, const double* seed_values,
                           const int32_t* out_slots, int64_t n_out);
int  adept_cuda_chain_end(adept_cuda_ctx* ctx, double* handoff_host);

/* Ensemble of B tapes sharing ONE structure (statements, index, indep, dep) with per-member
   multipliers laid out [op][member] (BASELINE config 3: simulate_radiances ensemble).  Output is
   member-major: jac_out[b*n_dep*n_indep + i + j*n_dep] (the reference's default column-major
   layout per member).  No counterpart in the Stack API (one Stack holds one tape). */
int  adept_cuda_jacobian_batch(adept_cuda_ctx* ctx, int mode,
                               const void* statements, int64_t n_stmt,
                               const int32_t* index, int64_t n_ops, int64_t max_gradient,
                               const double* multipliers, int64_t n_members,
                               const int32_t* indep, int64_t n_indep,
                               const int32_t* dep, int64_t n_dep,
                               double* jac_out_host);

/* Page-locked host memory for buffers handed to the entry points above (tape arrays, multiplier matrices, Jacobians):
   copies from/to it are true DMA and overlap with kernels; NULL when the allocation fails.  Plain malloc'ed memory
   is accepted everywhere too (it is staged). */
void* adept_cuda_host_alloc(size_t bytes);
void  adept_cuda_host_free(void* p);

/* ---- knobs and introspection ------------------------------------------------------- */

/* Options (key, value):
     "schedule"      0 = auto (default), 1 = always the tape-ordered kernel, 2 = always level-scheduled
     "ref_block"     P of the reference build being matched (1,2,4,8,16,32; default 2): granularity of
                     the reverse-mode "skip the statement if all P lanes are zero" test
                     (adept/jacobian.cpp:392-407).  Only visible with non-finite multipliers.
     "pass_dirs"     max directions resident per pass (0 = as many as fit in workspace_bytes)
     "workspace_bytes" cap for the direction-major gradient workspace (0 = 60% of free memory)
     "warps_per_cta" replay CTA size in warps (1..8, default 8)
     "window"        statements per dependency-analysis window (0 = default 65536)
     "streams"       CUDA streams (1..8, default 4) over which independent direction ranges of a level-scheduled
                     sweep are spread, each fed by its own host thread; hides the launch gap and tail of one
                     range's step behind another range's kernels
     "pdl"           1 (default) = the per-step kernels of a level sweep are chained with programmatic dependent
                     launch: a step's kernel starts (and TMA-fetches its records) while the previous one drains
     "smallg"        1 (default) = Jacobians of tapes whose gradient list fits shared memory (max_gradient*32*8 bytes
                     plus the reverse snapshot <= 200 KB) run in the shared-memory-resident kernel; 0 = never
     "chunk_recs"    records per level chunk = one CTA's share of a step (0 = default 64, max 256); takes effect at
                     the next adept_cuda_upload_tape
     "fused"         1 (default) = the level-scheduled reverse sweep runs ONE kernel per step over versioned rows
                     (two workspace rows per slot, no snapshot buffer); 0 = snapshot + gather, two kernels per step
     "cbs"           fused reverse kernel and balanced forward kernel: 64-direction column blocks per CTA (1..32; 0 = automatic:
                     32 reverse, 16 forward, fewer when a device holds few directions)
     "tail"          fused kernel: 1 (default) = steps of a chunk or two are run by the last CTA of the previous step's
                     launch instead of a launch of their own
     "split"         fused reverse kernels, work distribution: every (chunk, column group) goes to `split` CTAs (1..16), each
                     owning a range of the chunk's sub-tasks, whose warps take (sub-task, active column block) pairs in turn.
                     0 (default) = automatic: one CTA walks the whole chunk, or see "auto_split".  Stat "fused_split"
     "sub"           ... records per sub-task (4..256; sub-tasks are cut at target-group heads).  0 (default) = automatic: the
                     chunk length divided by "split", at least 8.  Stat "fused_sub"
     "auto_split"    1 (default) = when the device holds few directions (<= 1024, e.g. one rank of eight), the steps are narrow and no
                     chunk is longer than a staged piece, a chunk is shared by TWO CTAs: a Jacobian's activity is banded, so a
                     step's work then sits in a few CTAs on a mostly idle machine (BASELINE config 2, 512 directions: 41 ->
                     33 ms per Jacobian); 0 = never
   Option defaults can also be given in the environment, for callers that only see the Stack API:
     ADEPT_CUDA_OPTIONS="key=value,key=value" is applied by adept_cuda_create (a bad entry makes it fail).
     "fused_persistent" the fused reverse sweep as ONE resident kernel for the whole pass: CTAs take the (chunk, column
                     group) items of every step round robin, count themselves in per step and wait only for the participants
                     of the step above; no launch, drain and refill per step.  1 = always, -1 = when the steps are narrow (a
                     step's items would not fill the resident grid twice), 0 (default) = never: bit-identical, but measured
                     slower than one launch per step on this hardware.  Stat "used_fused_persistent"
     "spin_limit"    polls (64 .. 2^31-1, default 2^24) after which a waiting CTA of that kernel gives up: the call then
                     returns ADEPT_CUDA_ERR_CUDA instead of hanging the device
     "persistent"    1 = the level-scheduled reverse sweep runs as ONE cooperative kernel with grid-wide barriers
                     between steps (implies the two-kernel data layout); 0 (default): per-step launches, measured faster
     "time_kernels"  1 = bracket every launch of the dominant replay kernel with CUDA events
                     (stats "dominant_ms", "dominant_launches"); off by default
     "fwd_balanced"  1 (default) = the level-scheduled forward sweep uses forward_balanced_kernel (a CTA owns a chunk and
                     up to 32 column blocks, its warps take the active ones in turn and walk a compacted list) for every step
                     whose chunks fit one staged piece; 0 = forward_level_kernel (warp bound to a column block) throughout
     "fwd_v4"        1 = the balanced forward kernel keeps FOUR adjacent directions per lane (1024-byte row segments per warp,
                     128-direction column blocks) on tapes without over-long steps; 0 (default) = two
     "wide_streams"  N > 0: spread the forward level sweep over N streams even when its steps are wide (default 0: two
                     streams on wide steps, "streams" on narrow ones)
     "overlap_upload" 1 (default) = when the multiplier array is in page-locked memory it is uploaded on a stream of its own
                     while the level analysis (which reads only statements and indices) already runs; 0 = one after the other
     "coop_pack"     1 (default) = the window-packing pass of the dependency analysis runs as ONE cooperative kernel with two
                     grid-wide barriers per window; 0 = three dependent launches per window from the host
     "batch_chunks"  adept_cuda_jacobian_batch: chunks per device through its copy-in / sweep / copy-out pipeline (1..64,
                     default 4; a chunk holds at least 256 members)
     "remove_null"   1 = tape compaction on upload: operations whose multiplier is zero are dropped, as a reference built
                     with -DADEPT_REMOVE_NULL_STATEMENTS does while recording (include/adept/StackStorageOrig.h:46-65);
                     results then equal THAT build's bit for bit (and the default build's whenever no gradient is inf/NaN:
                     0*inf is the only thing a dropped term could have contributed).  Default 0.  Takes effect at the next
                     upload.  Stats: "ops_removed", and for every tape "frac_mult_zero" / "frac_mult_one" /
                     "frac_mult_minus_one" (Stack::fraction_multipliers_equal_to, include/adept/Stack.h:707-715)
     "profile_every" N > 0: cudaProfilerStart/Stop around every N-th launch of a level sweep (for
                     `ncu --profile-from-start off`); the bracketed launches are listed in $ADEPT_CUDA_PROFILE_LOG
*/
int  adept_cuda_set_option(adept_cuda_ctx* ctx, const char* key, int64_t value);

/* Statistics of the last call (key -> value): "n_levels", "n_chunks", "n_records",
   "kernel_launches" (cumulative count of this library's own kernel launches),
   "replay_ms" (CUDA-event time of the last replay: zero-fill+seed+sweep+extract(+gather)),
   "sweep_ms" (the replay kernels alone), "upload_ms", "analysis_ms", "d2h_ms",
   "used_schedule" (1 tape-ordered / 2 level / 3 level with shared-memory gradients), "used_fused", "pass_dirs",
   "n_passes", "an_window_ms" / "an_forward_ms" / "an_target_ms" / "an_fused_ms" (phases of the schedule build; the two
   reverse streams are built on first use), "host_issue_ms" (host time spent enqueueing the last reverse level sweep). */
int  adept_cuda_get_stat(const adept_cuda_ctx* ctx, const char* key, double* value);

/* Copies the per-statement level (1-based; level[0]=0 for the sentinel) of the last
   uploaded tape to the host: the device-side dependency analysis, exposed for tests. */
int  adept_cuda_get_levels(adept_cuda_ctx* ctx, int32_t* level_out, int64_t n_stmt);

const char* adept_cuda_last_error(const adept_cuda_ctx* ctx);
const char* adept_cuda_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ADEPT_CUDA_H */
