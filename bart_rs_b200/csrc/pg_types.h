// pg_types.h — plain data layout of the device-resident PGBART sampler.
//
// Everything the sweep touches lives in HBM behind the pointers of PgState:
//   X      column-major fp32 [p][ld]   (reference: data.rs OwnedData.x, f64 F-order)
//   y      fp32 [ld]                   (data.rs OwnedData.y)
//   A      fp64 [ld]  running array: holds `predictions` between steps and
//                     `residuals` (= sum of all OTHER trees, kernel.rs:84-87)
//                     while a tree is being updated
//   forest heap-indexed SoA tree arrays [m][maxn]   (tree.rs:9-23, state.rs:8)
//   slots  per-particle tree + BFS queue + per-node segment/statistics tables
//          [2][S][maxn] (double-buffered for resampling)   (particle.rs:48-52)
//   pool   u32 segment pool: per-leaf observation-index lists, appended by the
//          stable partition (particle.rs:13-23 LeafSamplesFlat)
// The header is shared by the CUDA build and by the host-side simulation used
// in tests (tests/cpu_sim), so it contains no CUDA-only constructs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#define PG_HDN __host__ __device__
#else
#define PG_HD inline
#define PG_HDN
#endif

enum PgPhase : int {
    PH_BEGIN = 0,   // start of a step: serial section prepares the first tree update
    PH_ROOT = 1,    // fused pass: A += tree_add(x); A -= tree_sub(x); root totals; round-1 child statistics
    PH_MINMAX = 2,  // per-job min/max of the split column over a node's index segment
    PH_STATS = 3,   // per-job child statistics over a node's index segment
    PH_BERN = 4,    // Bernoulli-logit: per-job child log-likelihood pass
    PH_PART = 5,    // stable partition of surviving particles' freshly split nodes into the pool
    PH_FINAL = 6,   // last pass of a step: A += tree_add(x) (A becomes `predictions` again)
    PH_DONE = 7
};

enum PgFamily : int { PG_GAUSSIAN = 0, PG_BERNOULLI_LOGIT = 1, PG_GAUSSIAN_KERNEL = 2 };
enum PgRngMode : int { PG_RNG_TAPE = 0, PG_RNG_PHILOX = 1 };

enum PgDraw : int { PG_D_GROW = 0, PG_D_VAR = 1, PG_D_SPLIT = 2, PG_D_ZLEFT = 3, PG_D_ZRIGHT = 4,
                    PG_D_RESAMPLE = 5, PG_D_SELECT = 6 };

enum PgError : int {
    PG_OK = 0,
    PG_ERR_DEPTH = 1,      // split past max_depth: the reference panics here (tree.rs:98-102)
    PG_ERR_POOL = 2,       // segment pool exhausted
    PG_ERR_TAPE = 3,       // RNG tape exhausted / split draw not below max
    PG_ERR_WATCHDOG = 4,   // grid barrier timed out
    PG_ERR_WEIGHTS = 5,    // non-finite weights (reference: WeightedIndex::new(..).unwrap() panics)
    PG_ERR_INTERNAL = 6,
    PG_ERR_PRUNE = 7       // prune_dead_proposals: the rounding slack of the weight bound is not negligible for this data
};

#define PG_SEG_IDENTITY 0xFFFFFFFFu  // node's samples are 0..n (root): no index list is read
#define PG_SEG_NONE 0xFFFFFFFEu      // children not materialised yet
#define PG_TS_FIELDS 12
#ifndef PG_THREADS
#define PG_THREADS 512
#endif
#define PG_WARPS (PG_THREADS / 32)

#define PG_MAX_RANKS 8
#define PG_MAIL_DOUBLES 128                    // payload of one exchange
// mailbox of one rank: [2 parities][PG_MAX_RANKS senders][PG_MAIL_DOUBLES] entries of 16 bytes — the two halves of a double,
// each in an 8-byte word that carries the exchange's sequence number in its upper half (flag-in-data: fw_exchange)
#define PG_MAIL_BYTES ((size_t)2 * PG_MAX_RANKS * PG_MAIL_DOUBLES * 16)

// A pass command: what a grid pass needs to know, = the first 16 ints of PgCtrl.  The P <= 32 control block publishes
// commands into PgCtrl.cmd[generation & 1] (two may be outstanding: the pass in progress and a root pass issued ahead);
// every pass block copies the command it is about to execute into its shared memory (PgSmem.cmd) and the passes read it there.
struct PgCmd {
    int phase, k, batch, t_cur, t_add, t_sub, round, n_jobs, n_groups, n_pjobs, cur, error, snext, max_size, t_add_internal, flags;
};

struct PgCtrl {
    // ---- state machine (written by the serial section, read by all CTAs after the barrier)
    int phase;
    int k;            // tree update index within this step's batch
    int batch;
    int t_cur;        // tree being updated
    int t_add;        // tree whose (new) values the next ROOT/FINAL pass adds to A, -1 = none, -2 = a stump (constant init_leaf)
    int t_sub;        // tree whose (old) values the next ROOT pass subtracts, -1 = none
    int round;
    int n_jobs;
    int n_groups;
    int n_pjobs;
    int cur;          // which half of the slot tables is live
    int error;
    int snext;        // next serial sub-step (PgSerialNext), SN_NONE = run the grid pass `phase`
    int max_size;     // largest slot tree size (bounds the resampling copy)
    int t_add_internal;   // internal nodes of tree t_add (bytes accounting)
    int flags;        // pass command flags: bit0 A source buffer, bit1 A destination buffer, bit2 job/partial buffer parity
    unsigned long long upd;       // tree-update counter since creation (RNG coordinate)
    unsigned long long pool_top;
    int next_tree_idx;
    int tune;
    // ---- per-update scalars
    double T_o, T_r, T_rr, T_ll0;  // root totals over all rows: sum others, sum (y-others), sum (y-others)^2, stump ll (Bernoulli)
    double C0;                      // particle-independent part of the log-likelihood
    double Cconst, inv_s2;          // per step: -n (log sigma + 0.5 log 2 pi) and 1/sigma^2 (Gaussian), else 0 and 1
    double g_stump;
    int acc_update;                 // accepted grows in this update
    int any_queue;
    // ---- BartInfo (state.rs:16-20) of the current step
    double info_loglik;
    long long info_acc;
    int info_n;
    // ---- roofline counters (bytes the passes must move)
    unsigned long long bytes_model;     // this implementation's compulsory bytes
    unsigned long long bytes_contract;  // SURVEY.md §8(d) formula
    unsigned long long n_phases;        // grid passes executed
    unsigned long long n_updates;       // tree updates executed
    unsigned long long n_spec;          // root passes issued speculatively (pg_serial_fast.cuh)
    unsigned long long n_pred;          // early commit: tree updates whose outcome was predicted after round 0 (fw_predict)
    unsigned long long n_pred_miss;     // ... of which the final selection disagreed (the root pass issued ahead was drained)
    unsigned long long n_drain;         // speculative / predicted root passes that were awaited and discarded
    unsigned long long trace_n;         // tree updates recorded in the trace
    int trace_overflow;
    int a_cur;                      // which A buffer holds the running array of the update in progress
    unsigned long long xseq;        // observation sharding: exchanges between ranks done so far (same on every rank)
    // serial-section breakdown (cycles / calls): 0 reduce, 1 after_stats, 2 after_minmax, 3 finish_round, 4 prep_part,
    // 5 finalize, 6 begin_update, 7 draw_round; [8..15] call counts
    unsigned long long prof2[16];
    unsigned long long prof_ser[8];     // serial section cycles by phase (kept by the control block)
    unsigned long long prof3[16];       // control block: [0,1] root sections with speculation (cycles, count), [2,3] without,
                                        // [4..7] cycles spent waiting for a ROOT / STATS / PART / other pass to complete
    // ---- grid barrier
    unsigned int bar_count;
    unsigned int bar_pad[31];
    unsigned int bar_gen;
    unsigned int bar_pad2[31];
    // ---- in-kernel stopwatch (SM cycles), kept by the LAST block (outside the serial block's working copy):
    // [0..7] pass time by phase, [8..15] pass end -> barrier exit (wait for the slowest block + serial section),
    // [16..23] serial section alone (written by the serial block), [24..31] number of passes by phase
    unsigned long long prof[32];
    // ---- signalling diagnostics (globaltimer ns): [0] time of the last issue, [1] latest command observation,
    // [2] latest arrival, [3] longest pass body (SM cycles) of the pass in flight; [4..] per phase class c = 0 ROOT,
    // 1 STATS, 2 PART: [4+4c] sum issue->last observed, [5+4c] sum issue->last arrival, [6+4c] sum issue->completion seen
    // by the control block, [7+4c] sum of the longest pass body (cycles)
    unsigned long long dbg[16];
    // ---- hint for the pass blocks: number of root jobs staged ahead in the job arrays of each buffer parity
    // (written by the control block when it pre-stages the next tree update; read after a root pass to prefetch)
    int pre_J[2];
    // ---- set by pg_prepare_step when an earlier step left a device error behind: steps queued after it must not run on
    // the broken state (pgbart_step_device can be queued several times before pgbart_sync reports the error)
    int poisoned;
    int pad_cmd[21];
    unsigned long long hostdiag;    // device address of a host-mapped buffer (8 x u64) that a block fills before it traps on a watchdog: the
                                    // context is lost after the trap, the host memory is not (pgbart_last_error then says who waited for what)
    double ll_old;                  // weight_mode pg_correct: log-likelihood of the current predictions (= of the tree being replaced)
    unsigned long long n_es, n_es_miss;   // early statistics passes issued / rounds that had to run a regular statistics pass after one (written in place by the control block)
    // ---- P <= 32 persistent kernels: command buffers by generation parity, and completion counters by generation parity
    // (zeroed with bar_gen by pg_prepare_step); counter q holds 147 x (generations of parity q completed)
    int cmd[2][16];
    unsigned int cnt[2][32];
};

// Logarithm table of the Bernoulli pass (pg_log_ge1, pg_serial.h); filled on the host.
#include <math.h>
#define PG_LOGT_K 64
#define PG_LOGT_REP 8   // shared-memory copies on the device: the 8 lanes of a quarter warp read 8 different 16-byte entries
                        // (lane & 7 picks the copy), so the lookup never has a bank conflict
inline void pg_log_table_fill(double* tbl /* [PG_LOGT_K][2] */) {
    for (int k = 0; k < PG_LOGT_K; ++k) {
        const long double c = 1.0L + ((long double)k + 0.5L) / (long double)PG_LOGT_K;
        const double inv = (double)floorl(4096.0L / c + 0.5L) / 4096.0;
        tbl[2 * k + 0] = inv;
        tbl[2 * k + 1] = (double)(-logl((long double)inv));
    }
}

struct PgState {
    // dimensions
    long long n, ld;
    int p, m, P, S, max_depth, maxn;
    int grid, n_gwarps;          // CTAs in the sweep grid, global warps (= grid * PG_WARPS)
    int stage_bytes;             // shared-memory staging area per block behind the pass arrays (async-copy ring)
    // ---- observation sharding: this rank holds n of n_global rows; reductions are exchanged
    // between the ranks' control blocks through peer mailboxes (pg_serial_fast.cuh, fw_exchange)
    int rank, world;
    long long n_global;
    double* mail[PG_MAX_RANKS];  // mail[r] = rank r's mailbox (peer memory for r != rank)
    int prune;                   // option prune_dead_proposals: do not evaluate proposals whose particle provably gets weight 0
    int prefetch_next;           // pass blocks prefetch the next update's root columns into L2 after a root pass
    int root_ring;               // P <= 32 persistent kernels: 0 = direct root pass (register loads), 1 = ring root pass (pg_pass_root_ring),
                                 // 2 = ring when a pass block holds at most 2048 rows, else direct (default)
    int root_tile_groups;        // ring root pass: cap on the row groups (128 rows) per tile, 0 = as many as fit (<= 8)
    int root_stages;             // ring root pass: cap on the stages of the ring, 0 = as many as fit
    int pg_correct;              // SURVEY.md 8(f4) weight mode (NOT the reference's behaviour, off by default): slot log-weights are kept as
                                 // log-likelihoods and permuted with the ancestors, particle 0 is the tree being replaced (conditional SMC)
    int generic;                 // 1 = the generic control code (pg_serial.h) runs the step: P > 32, pg_correct, or a OneHotSplit column
    int heartbeat;               // diagnostics: the control block records where it is in the host-mapped watchdog buffer (PgCtrl.hostdiag)
    int early_stats;             // P <= 32 control block: 1 (default) = the statistics pass of rounds 1-2 is queued behind the partition pass and derives its own thresholds
    int issue_ahead;             // P <= 32 control block: 1 (default) = the next update's root pass is issued right behind the pass in progress
    int root_dbg;                // ring root pass, timing experiments only (results are wrong): 1 no proxy fence in the loop, 2 no job loop, 4 no row work at all
    // data
    const float* X;
    const float* y;
    double* A;                   // buffer 0: holds `predictions` between steps
    double* A2;                  // buffer 1: ping-pong partner for speculatively issued root passes
    const float* col_min;
    const float* col_max;
    const double* prior;         // split prior used as raw weights (Q3); null = uniform integer draw
    double prior_total;
    // ---- OneHotSplit columns (splitting.rs:72-104; generic control code only).  The reference casts the node's values to i32,
    // collects the distinct codes and picks one uniformly (its partition still tests x < code: SURVEY.md Q7).  Here a
    // column's codes must span fewer than 64 consecutive integers, so a node's distinct codes are a 64-bit presence mask.
    const unsigned char* onehot; // [p] 1 = OneHotSplit, or null when every column is ContinuousSplit
    const int* col_imin;         // [p] smallest code of the column (bit 0 of the masks)
    const unsigned long long* col_bits;   // [p] presence mask over all rows (the root's candidates)
    unsigned long long* job_bits;         // [S] presence mask of each job's node
    const double* log_tbl;       // [PG_LOGT_K][2] (inv_k, -log inv_k): logarithm table of the Bernoulli pass (pg_log_ge1)
    const double* thr_depth;     // [max_depth+2] 1 - alpha (1+d)^-beta (host pow, smc.rs:143)
    double leaf_sigma, m_f, init_leaf, y_mean;
    double batch_tune, batch_post;
    // likelihood (weight.rs)
    int family;
    double lik_sigma;
    // forest
    int* f_split_var; float* f_thr32; double* f_split_val; double* f_leaf_val; int* f_size;
    int* f_internal;             // [m] number of internal nodes per forest tree
    // particle slot tables, [2][S][maxn] / [2][S]
    int* pt_split_var; float* pt_thr32; double* pt_split_val; double* pt_leaf_val;
    unsigned int* pt_seg_off; unsigned int* pt_cnt;
    unsigned int* pt_len;        // rows of the node held by THIS rank (= pt_cnt on one GPU; pt_cnt is the global count)
    double* pt_So; double* pt_Sr; double* pt_g;
    int* pt_queue; int* pt_size; int* pt_qhead; int* pt_qtail; double* pt_G;
    double* w; int* mutated; int* anc; int* slot_job;   // [S]
    int* pend_node; int* pend_var;                       // [S] proposal being evaluated
    double* Wfinal;                                      // [2P] log / normalised final weights
    // jobs of the current grid pass, [S]
    int* job_slot; int* job_node; int* job_var; float* job_thr32; double* job_split;
    unsigned int* job_seg_off; unsigned int* job_len; float* job_lo; float* job_hi;
    double* job_vL; double* job_vR;
    int* job_order;              // [S] permutation of jobs: jobs sharing one index segment are adjacent
    int* grp_first;              // [S+1] group boundaries in job_order
    double* red_So; double* red_Sr; double* red_llL; double* red_llR;   // [S] reduced per-job sums (LEFT child; Bernoulli: both)
    unsigned long long* red_cnt;                      // [S] reduced LEFT-child counts per job
    // partition jobs, [S]
    int* pj_anc; int* pj_job; int* pj_var; float* pj_thr32;
    unsigned int* pj_src_off; unsigned int* pj_len; unsigned int* pj_dst_off; unsigned int* pj_cntL;
    unsigned int* pj_warp_left;  // [S][n_gwarps] exclusive prefix of left counts per global warp
    int* pj_of_anc;              // [S] partition job of an ancestor slot, -1 = none
    unsigned int* scan_tmp;      // [PG_THREADS + 1]
    // per-pass partial results
    double* part_sum;            // [grid][S][2]  (sum others, sum y-others) of the LEFT child
    unsigned int* part_cnt;      // [n_gwarps][S] left count per global warp (prefix for the partition)
    unsigned int* part_ccnt;     // [grid][S]     left count per block
    double* part_tot;            // [grid][4]
    unsigned int* part_mm;       // [grid][2S][2] ordered-uint keys of min / max (2S: the partition pass may carry look-ahead requests)
    double* mq_u3; float* es_thr;   // [2S] early statistics pass: the split draw of each request; the threshold the pass derived (NaN: none)
    int* mq_pj; int* mq_side; int* mq_var;   // [2S] look-ahead min/max requests of the partition pass: (partition job, child side, column)
    double* part_ll;             // [grid][S][2]  Bernoulli: log-likelihood of the left / right child
    double* part_old;            // [grid] per-block sum behind ll_old (root pass, pg_correct only)
    unsigned* blk_done;          // [grid] generation of the last command each pass block completed (fast kernels): the control
                                 // block's reduce loads a block's partials as soon as that block is done, not after the slowest one
    unsigned long long* diag_blocks;   // [grid][8] per pass block, by phase: cycles spent in pass bodies (-DPG_DIAG only)
    // segment pool
    unsigned int* pool; unsigned long long pool_cap;
    // rng
    int rng_mode; unsigned int philox_key0, philox_key1;
    const double* tape_slot; const double* tape_round; const double* tape_upd;
    long long tape_n_upd; int tape_r_cap; unsigned long long tape_base_upd;
    // trace (tests): same layout as the oracle's
    double* tr_slot; double* tr_round; double* tr_upd; long long tr_u_cap; int tr_r_cap;
    unsigned char* info_depths;  // [m]
    PgCtrl* ctrl;
};

// ---- small helpers ---------------------------------------------------------

PG_HD int pg_depth_of(int node) {  // tree.rs:54-57
    int d = 0;
    unsigned v = (unsigned)node + 1u;
    while (v > 1u) { v >>= 1; ++d; }
    return d;
}

// Smallest float >= s, so that for float x:  (double)x < s  <=>  x < pg_thr32(s).
PG_HD float pg_thr32(double s) {
    float f = (float)s;  // round to nearest
    if ((double)f < s) {
        // next float up
        union { float f; uint32_t u; } c; c.f = f;
        if (f == 0.0f) c.u = 1u;                 // smallest positive subnormal
        else if (f > 0.0f) c.u += 1u;
        else c.u -= 1u;
        f = c.f;
    }
    return f;
}

// Order-preserving map float -> uint32 (for integer min/max reductions).
PG_HD uint32_t pg_f2key(float f) {
    union { float f; uint32_t u; } c; c.f = f;
    return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}
PG_HD float pg_key2f(uint32_t k) {
    union { float f; uint32_t u; } c;
    c.u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return c.f;
}

// Contiguous range of segment entries owned by one global warp.  Identity
// segments use granule 4 (float4 loads), index segments granule 1.
// Rows of pass block `bid` in the block-unit passes (staged root pass, identity partitions): contiguous, a multiple of
// 4 rows (16 B of a column) so every tile is a legal bulk copy; the last range ends at n rounded up to 4 (<= ld).
PG_HD void pg_block_range(long long n, int bid, int grid, long long* b0, long long* b1) {
    const long long units = (n + 3) / 4;
    const long long per = (units + grid - 1) / grid;
    long long b = (long long)bid * per * 4, e = b + per * 4;
    const long long n4 = units * 4;
    if (b > n4) b = n4;
    if (e > n4) e = n4;
    *b0 = b; *b1 = e;
}

// Entries of an index segment of length L owned by pass block `bid` in the block-unit scheme: the union of the 16 warp
// ranges pg_warp_range(L, gw, n_gwarps, 1) of its warps, so per-block counts agree whichever statistics pass (warp-range
// or block-unit) produced them.
PG_HD void pg_seg_block_range(unsigned long long L, int bid, int n_gwarps, int warps_per_block,
                              unsigned long long* b0, unsigned long long* b1) {
    const unsigned long long per = (L + (unsigned long long)n_gwarps - 1) / (unsigned long long)n_gwarps;
    unsigned long long b = (unsigned long long)bid * (unsigned long long)warps_per_block * per;
    unsigned long long e = b + (unsigned long long)warps_per_block * per;
    if (b > L) b = L;
    if (e > L) e = L;
    *b0 = b; *b1 = e;
}

PG_HD void pg_warp_range(unsigned long long L, int gw, int n_gwarps, int granule,
                         unsigned long long* beg, unsigned long long* end) {
    unsigned long long units = (L + (unsigned long long)granule - 1) / (unsigned long long)granule;
    unsigned long long per = (units + (unsigned long long)n_gwarps - 1) / (unsigned long long)n_gwarps;
    unsigned long long b = (unsigned long long)gw * per * (unsigned long long)granule;
    unsigned long long e = b + per * (unsigned long long)granule;
    if (b > L) b = L;
    if (e > L) e = L;
    *beg = b; *end = e;
}
