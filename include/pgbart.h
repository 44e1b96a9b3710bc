/* pgbart.h — C ABI of the B200-native Particle-Gibbs BART sweep (libpgbart_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of GStechschulte/bart-rs: the PGBART
 * step.  Every entry point names the reference interface it replaces (paths relative to the
 * reference tree).  Plain pointers and sizes only; every call returns 0 on success and a
 * non-zero status otherwise, with text from pgbart_last_error().  There is no CPU fallback:
 * pgbart_create fails when no CUDA device is usable.
 *
 * Reference surface being replaced (src/lib.rs):
 *   PyBartSettings.__new__   lib.rs:54-104   -> pgbart_settings
 *   PySampler.init           lib.rs:115-180  -> pgbart_create (+ pgbart_set_likelihood)
 *   PySampler.step           lib.rs:182-202  -> pgbart_step
 *   LogpFunc callback        lib.rs:33, weight.rs:10-30 -> pgbart_set_likelihood (device families)
 * and, for Rust callers of the rlib (benches/particle_gibbs.rs:105-127,257-282):
 *   BartKernel::init/step    kernel.rs:38-126 -> pgbart_create / pgbart_step_device
 */
#ifndef PGBART_H
#define PGBART_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgbart_sampler* pgbart_handle;

/* PyBartSettings (lib.rs:35-104), field for field.  response_rule / resampling_rule are accepted
 * and ignored exactly as the reference ignores them (always Gaussian leaf rule, systematic
 * resampling; lib.rs:47-48,166).  split_rules: "ContinuousSplit" or "OneHotSplit" per column, padded with
 * ContinuousSplit (lib.rs:134-144); an unknown name is an error as in lib.rs:129-132.  A OneHotSplit column
 * (splitting.rs:72-104) must hold integer codes spanning fewer than 64 consecutive values; its split value is one of
 * the node's distinct codes, uniformly, and the partition tests x < code as in the reference (particle.rs:128).  A
 * sampler with a OneHotSplit column runs on the generic control code (slower than the n_particles <= 32 path). */
typedef struct pgbart_settings {
    uint64_t n_trees;
    uint64_t n_particles;          /* includes the reference stump (smc.rs:42-52) */
    uint32_t max_depth;            /* u8 in the reference */
    double alpha, beta, sigma;     /* sigma = leaf standard deviation */
    const double* split_prior;     /* used as RAW weights like the reference (smc.rs:242-254) */
    uint64_t n_split_prior;        /* 0 = uniform integer draw (lib.rs:153-157) */
    const char* const* split_rules;
    uint64_t n_split_rules;
    const char* response_rule;
    const char* resampling_rule;
    double batch_tune, batch_post; /* defaults 0.1 in the reference */
    uint64_t seed;                 /* default 0 */
} pgbart_settings;

/* Likelihood families evaluated on the device in place of the host logp callback. */
enum { PGBART_GAUSSIAN = 0,        /* params = {sigma}: sum -0.5((y-mu)/sigma)^2 - log sigma - 0.5 log 2pi */
       PGBART_BERNOULLI_LOGIT = 1, /* no params:        sum y*mu - softplus(mu) */
       PGBART_GAUSSIAN_KERNEL = 2  /* no params:        sum -0.5 (mu-y)^2  (benches/particle_gibbs.rs:22-32) */ };

enum { PGBART_X_F64 = 0, PGBART_X_F32 = 1 };
/* OR into x_dtype: accept round-to-nearest fp32 storage of f64 x and y.  Without it pgbart_create (and pgbart_predict)
 * refuse, with status 1, data that is not exactly representable in fp32: the device computes on fp32 X / y
 * (north_star) while the reference computes on the f64 values (data.rs:31-35), and a silent quantisation could merge
 * distinct values, change ties and child counts, and so diverge from the reference. */
enum { PGBART_ALLOW_FP32_ROUNDING = 0x100 };
enum { PGBART_ORDER_C = 0, PGBART_ORDER_F = 1 };

/* PySampler.init (lib.rs:115-180).  x: n-by-p host matrix (f64 or f32, C or Fortran order), y: f64[n].
 * Copies both (as the reference does), stores X column-major fp32 and y fp32 in HBM, builds m stump
 * trees and predictions = mean(y) (kernel.rs:38-58).  device = CUDA ordinal.  f64 values must be exactly
 * representable in fp32 unless x_dtype carries PGBART_ALLOW_FP32_ROUNDING (checked before any device work). */
int pgbart_create(const void* x, int x_dtype, int x_order, uint64_t n, uint64_t p, const double* y,
                  const pgbart_settings* settings, int device, pgbart_handle* out);
int pgbart_destroy(pgbart_handle h);
/* Page-locked host buffers for out_pred (freed by the caller; error text via pgbart_last_error(NULL)). */
int pgbart_host_alloc(uint64_t bytes, void** out);
int pgbart_host_free(void* p);
const char* pgbart_last_error(pgbart_handle h);   /* h may be NULL: error of the last failed create */

/* Replaces the LogpFunc pointer (lib.rs:125-126): callable before every step, e.g. with the
 * current sigma, like CompiledPyMCModel.update_shared_arrays (pgbart.py:190). */
int pgbart_set_likelihood(pgbart_handle h, int family, const double* params, int nparams);

/* PySampler.step (lib.rs:182-202): tune = 1/0, or -1 to keep the current flag (None).  Writes the
 * new sum-of-trees, f64[n], to HOST memory out_pred (device->host copy inside the call): one copy at link speed
 * when out_pred is page-locked (pgbart_host_alloc / cudaHostRegister), a chunked copy through a pinned staging
 * buffer otherwise. */
int pgbart_step(pgbart_handle h, int tune, double* out_pred);
/* Same step, result left in HBM; asynchronous until pgbart_sync. */
int pgbart_step_device(pgbart_handle h, int tune);
int pgbart_sync(pgbart_handle h);
const double* pgbart_predictions_device(pgbart_handle h);
int pgbart_get_predictions(pgbart_handle h, double* out_pred);

/* BartInfo of the last step (state.rs:16-20; dropped at the PyO3 boundary, lib.rs:197). */
int pgbart_get_info(pgbart_handle h, double* log_likelihood, int64_t* acceptance_count,
                    uint8_t* tree_depths, int* n_depths /* in: capacity, out: count */);

/* Forest export: TreeArrays (tree.rs:9-23).  split_var = -1 for leaves (u32::MAX sentinel). */
int pgbart_tree_size(pgbart_handle h, uint64_t tree, int* size);
int pgbart_get_tree(pgbart_handle h, uint64_t tree, int32_t* split_var, double* split_val, double* leaf_val);
int pgbart_get_leaf_indices(pgbart_handle h, uint64_t tree, uint32_t* out /* n */);
int pgbart_get_state(pgbart_handle h, uint64_t* next_tree_idx, int* tune, uint64_t* n_updates);

/* Random source.  Default: Philox4x32-10 keyed by settings.seed.
 * Tape (replay mode): typed draws by coordinates, host arrays copied to the device:
 *   slot[n_upd][r_cap][n_particles-1][5] = u1,u2,u3,zL,zR ; round[n_upd][r_cap] = u6 ; upd[n_upd] = u7.
 * The tape's first record corresponds to the sampler's next tree update. */
int pgbart_set_rng_philox(pgbart_handle h, uint64_t seed);
int pgbart_set_rng_tape(pgbart_handle h, const double* slot, const double* round, const double* upd,
                        int64_t n_upd, int r_cap);

/* Decision trace for parity tests (same layout as the oracle's):
 *   slot[u_cap][r_cap][S][12]: status(0 idle,1 depth-prior reject,2 empty,3 min>=max,4 accepted), node, var,
 *        split, vL, vR, cntL, cntR, log-weight, min, max, sum-others-left
 *   round[u_cap][r_cap][2+2S]: used, n_mutated, normalised weights[S], ancestors[S]
 *   upd[u_cap][4+2P]: rounds, selected, tree, accepted, log W[P], normalised W[P] */
int pgbart_trace_enable(pgbart_handle h, int64_t u_cap, int r_cap);
int pgbart_trace_read(pgbart_handle h, double* slot, double* round, double* upd, int64_t* n_upd, int* overflow);

/* Options: "driver" = "persistent" (default) | "stepwise"; "pool_factor" = integer f >= 1 (segment pool of f * (P-1) * n
 * entries; the default is the bound no tree update can exceed, (P-1) * (max_depth+1) * n, whenever that is under a quarter
 * of the free device memory, so the pool cannot run out; a pool that does run out ends the step with status 2);
 * "weight_mode" = "reference" (default: the reference's arithmetic, smc.rs:53,89-102 — slot weights are normalised in
 * place and meet next round's log-likelihoods, the reference particle is a fresh stump) | "pg_correct" (NOT the
 * reference's behaviour: every slot keeps the log-likelihood of the particle it holds, permuted with the ancestors, and
 * particle 0 is the tree being replaced — conditional SMC; runs on the generic control code; set before the first step);
 * "root_pass" = "auto" (default: the ring pass when a pass block holds at most 2048 rows, else the direct pass) | "direct"
 * (register loads over warp sub-ranges of the block's rows) | "ring" (bulk-copy ring in shared memory, cp.async.bulk);
 * "issue_ahead" = 1 | 0 (the next update's root pass is published while the current pass streams); "early_stats" = 1 | 0
 * (the statistics pass of rounds 1-2 is queued behind the partition pass and derives its own thresholds);
 * "prefetch_next" = 0 | 1 (L2 prefetch of the next tree update's root columns after a root pass);
 * "prewarm_host" = 1: create now the page-locked staging buffer that pgbart_step uses for pageable out_pred (observation-
 * sharded samplers call it before their first step: nothing may be allocated while the peers' kernels wait for this rank). */
int pgbart_set_option(pgbart_handle h, const char* key, const char* value);

/* Counters since creation: [0] compulsory bytes of this implementation's passes, [1] bytes by the
 * SURVEY.md §8(d) contract formula, [2] grid passes, [3] tree updates, [4] kernel launches,
 * [5] GPU time of the last step in microseconds (CUDA events on the sampler's stream),
 * [6] grid blocks, [7] dynamic shared memory per block. */
int pgbart_get_counters(pgbart_handle h, double* out8);

/* Speculation / early-commit counters since creation: [0] root passes issued ahead of the decision
 * they depend on, [1] tree updates whose outcome was predicted after round 0, [2] predictions the final selection
 * contradicted (the pass issued ahead was drained and re-issued), [3] passes issued ahead that were awaited and discarded. */
int pgbart_get_spec_counters(pgbart_handle h, uint64_t* out4);
/* Early statistics pass (statistics of rounds 1-2 queued behind the partition pass): out2 = passes issued since creation,
 * rounds that nevertheless had to run a regular statistics pass (the cached threshold did not match). */
int pgbart_get_early_stats_counters(pgbart_handle h, uint64_t* out2);

/* In-kernel stopwatch, SM cycles since creation: out32[0..7] pass time by phase as seen by block 0,
 * [8..15] pass end -> barrier exit (wait for the slowest block + serial section), [16..23] serial section alone,
 * [24..31] passes by phase (phases: 1 ROOT, 2 MINMAX, 3 STATS, 4 BERN, 5 PART, 6 FINAL); [32..39] serial sub-step
 * cycles (reduce, after_stats, after_minmax, finish_round, prep_part, finalize, begin_update, draw_round), [40..47] calls, [48] root passes issued speculatively,
 * [49..64] control block: root sections with / without a speculative issue (cycles, count, cycles, count), then cycles spent
 * waiting for a ROOT / STATS / PART / other pass to complete; [57..68] signalling diagnostics, filled only when the library
 * is built with -DPG_DIAG: for ROOT, STATS, PART: ns issue -> last block saw the command, issue -> last arrival,
 * issue -> control block saw completion, and cycles of the longest pass body, all summed over passes. */
int pgbart_get_profile(pgbart_handle h, double* out69);

/* Out-of-sample prediction: out[i] = sum over the forest's trees, in tree order, of the leaf value row i of x reaches
 * (reference TreeArrays::predict_batch_test, src/tree.rs:169-192, applied to every tree; comparison `x < split_val` in
 * f64).  x: n_new x p, dtype / order as in pgbart_create.  Training saw fp32 x and compared it with thr32 = the smallest
 * float >= split_val, which for fp32-representable x IS the reference's f64 comparison; new rows are treated the same way
 * (rounded to fp32 when the sampler was created with PGBART_ALLOW_FP32_ROUNDING, refused with status 1 otherwise if not
 * representable), so predict(X_train) always equals the in-sample predictions.  Under observation sharding every rank
 * holds the whole forest. */
int pgbart_predict(pgbart_handle h, const void* x, int x_dtype, int x_order, uint64_t n_new, double* out);

/* Split counts per variable over the current forest (reference BartState.variable_inclusion, src/state.rs:10 — declared
 * and zero-initialised at src/kernel.rs:49, never updated; python/pymc_bart/pgbart.py:56,196 reports a placeholder). */
int pgbart_variable_inclusion(pgbart_handle h, uint32_t* out_p);

/* Observation sharding: the rows of one chain are split across `world` GPUs of one node, one process and
 * one handle per GPU, each created with its own rows.  Every rank runs the same control code on the same draws; the
 * reductions (root totals, per-proposal counts and sums, min/max, Bernoulli log-likelihoods) are exchanged inside the
 * persistent kernel through peer mailboxes over NVLink and combined in rank order, so all ranks take bit-identical
 * decisions.  Replaces nothing in the reference (it is single-process CPU); north_star asks for it.
 *   1. pgbart_shard_config: before the first step.  n_global, the mean of all y and the global per-column min / max
 *      (the reference's root split candidates, splitting.rs:43-49) replace the values create() derived from the local rows.
 *   2. pgbart_shard_mailbox: 64-byte CUDA IPC handle of this rank's mailbox — exchange the handles between ranks
 *      (torch.distributed all_gather; bart_rs_b200/sharded.py).
 *   3. pgbart_shard_connect: handles = world x 64 bytes, in rank order.
 * pgbart_step* then returns this rank's rows of the predictions.  n_particles <= 32, persistent driver only. */
int pgbart_shard_config(pgbart_handle h, int rank, int world, uint64_t n_global, double y_mean_global,
                        const float* col_min_global, const float* col_max_global);
int pgbart_shard_mailbox(pgbart_handle h, unsigned char* handle64);
int pgbart_shard_connect(pgbart_handle h, const unsigned char* handles);

/* Diagnostics (-DPG_DIAG builds): out[b*8 + phase] = SM cycles pass block b has spent in pass bodies of that phase;
 * out[G*8 + b*8 + i] = cycles between checkpoints i of the staged root pass (G = pass blocks; cap >= 16 G). */
int pgbart_debug_block_cycles(pgbart_handle h, double* out, int cap);
/* Where the device state machine stands (after a step, or after a step that failed gracefully — e.g. an exchange that timed
 * out in an observation-sharded run): out8 = phase, tree update in the batch, round, tree updates since creation,
 * exchanges between ranks since creation, grid passes since creation, device error code, tree being updated. */
int pgbart_debug_state(pgbart_handle h, double* out8);

/* CUDA-event stopwatch on the sampler's own stream (the stream the sweep kernels are launched on):
 * start records an event, stop records a second one, synchronises and returns the milliseconds between. */
int pgbart_timer_start(pgbart_handle h);
int pgbart_timer_stop(pgbart_handle h, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif
