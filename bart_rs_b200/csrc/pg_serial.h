// pg_serial.h — the scalar control logic of the sweep, run between grid passes
// by ONE thread block (the last block to arrive at the grid barrier).
//
// It is the device counterpart of the non-streaming parts of the reference:
//   kernel.rs:60-126   batch loop, tree replacement, BartInfo
//   smc.rs:26-130      round loop, stale-weight rule (Q1), resampling, final selection
//   smc.rs:133-187     propose_mutation: depth prior, split variable, split value
//   smc.rs:219-239     leaf values
//   smc.rs:242-268     sample_feature_from_probs, normalize_weights_inplace
//   resampling.rs:24-44 systematic resampling
//   particle.rs:91-148 bookkeeping half of apply_mutation (tree split, queue push)
// The streaming halves (sums, min/max, partition) are the grid passes in
// pg_passes.cuh; their per-block partial results are reduced here in a fixed
// order, so a step is reproducible bit-for-bit from run to run.
//
// Written against a "team" (rank, size, sync) so the same source runs as one
// CUDA block (rank = threadIdx.x) and, in tests/cpu_sim, as plain host code
// (rank 0 of 1).  No CPU path exists in the product library.
#pragma once
#include "pg_types.h"
#include "pg_rng.h"
#include <math.h>
#include <string.h>

struct PgTeam {
    int rank, size;
    PG_HD void sync() const {
#if defined(__CUDA_ARCH__)
        __syncthreads();
#endif
    }
};

enum PgSerialNext : int { SN_NONE = 0, SN_BEGIN_UPDATE, SN_DRAW_ROUND, SN_FINISH_ROUND, SN_PREP_PART, SN_FINALIZE,
                          SN_PART_ISSUED /* pg_serial_fast.cuh: the partition pass is already in flight */,
                          SN_WAIT_EARLY /* pg_serial_fast.cuh: the round waits for the early statistics pass in flight */ };

#define PG_FOR(i, N) for (int i = tm.rank; i < (N); i += tm.size)
#define PG_SINGLE if (tm.rank == 0)

// The order-sensitive loops of the reference (normalisation, cumulative weights, resampling, selection) stay sequential
// in one thread, but over a copy of the weights in the team's scratch memory (shared memory on the device), and everything
// elementwise around them runs on the whole team: with 255 slots a single thread walking global arrays needed ~1 ms per
// tree update (BASELINE C5), the staged version a few microseconds.
#define PG_SCR_N 512      // slots / particles the scratch arrays hold; larger samplers work on the global arrays directly
#if defined(__CUDA_ARCH__)
#define PG_SCRATCH(type, name, count) __shared__ type name[count]
PG_HD void pg_team_add(int* p, int v) { atomicAdd(p, v); }
PG_HD void pg_team_max(int* p, int v) { atomicMax(p, v); }
PG_HD void pg_team_add64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
#else
#define PG_SCRATCH(type, name, count) type name[count]
PG_HD void pg_team_add(int* p, int v) { *p += v; }
PG_HD void pg_team_max(int* p, int v) { if (v > *p) *p = v; }
PG_HD void pg_team_add64(unsigned long long* p, unsigned long long v) { *p += v; }
#endif
// Number of set flags before position i (flags in scratch memory): the ordered compaction of a few hundred entries is
// done by every member summing its own prefix from shared memory, a microsecond, instead of one thread walking global arrays.
PG_HD int pg_flags_before(const int* flag, int i) {
    int n = 0;
    for (int q = 0; q < i; ++q) n += flag[q];
    return n;
}

PG_HD double pg_nan() {
    union { unsigned long long u; double d; } c;
    c.u = 0x7ff8000000000000ull;
    return c.d;
}
PG_HD size_t pg_tix(const PgState& st, int buf, int s, int node) {
    return ((size_t)buf * (size_t)st.S + (size_t)s) * (size_t)st.maxn + (size_t)node;
}
PG_HD int pg_six(const PgState& st, int buf, int s) { return buf * st.S + s; }

PG_HD double* pg_trace_slot(const PgState& st, PgCtrl* c, int round, int s) {
    if (!st.tr_slot) return 0;
    if ((long long)c->trace_n >= st.tr_u_cap || round >= st.tr_r_cap) { c->trace_overflow = 1; return 0; }
    return st.tr_slot + (((size_t)c->trace_n * (size_t)st.tr_r_cap + (size_t)round) * (size_t)st.S + (size_t)s) * PG_TS_FIELDS;
}

PG_HD double* pg_trace_slot(const PgState& st, int round, int s) { return pg_trace_slot(st, st.ctrl, round, s); }

// Separately rounded a*b + c (the reference is Rust: no FMA contraction).
PG_HD double pg_mul_add(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dmul_rn(a, b), c);
#else
    volatile double prod = a * b;
    return prod + c;
#endif
}

// smc.rs:242-254 — the prior is consumed as raw weights (Q3).
PG_HD int pg_feature_from_probs(const PgState& st, double u) {
    double target = u * st.prior_total;
    for (int j = 0; j < st.p; ++j) {
        target -= st.prior[j];
        if (target <= 0.0) return j;
    }
    return st.p - 1;
}

// Fixed-order sum of one column of a [grid][stride] array of per-block partials.
PG_HD double pg_reduce_blocks(const double* base, int grid, size_t stride, size_t off) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int b = 0;
    for (; b + 3 < grid; b += 4) {
        a0 += base[(size_t)b * stride + off];
        a1 += base[(size_t)(b + 1) * stride + off];
        a2 += base[(size_t)(b + 2) * stride + off];
        a3 += base[(size_t)(b + 3) * stride + off];
    }
    for (; b < grid; ++b) a0 += base[(size_t)b * stride + off];
    return (a0 + a1) + (a2 + a3);
}

PG_HD double pg_softplus(double mu) { return fmax(mu, 0.0) + log1p(exp(-fabs(mu))); }

// ---- log(u) for finite u >= 1: table + polynomial (Bernoulli pass) -------------------------------------------------
// u = m 2^e with m in [1, 2).  Entry k (the top 6 mantissa bits of m) holds (inv_k, -log inv_k) with inv_k ~ 1 / (1 + (k +
// 0.5) / 64) rounded to 12 fractional bits, so r = fma(m, inv_k, -1) is rounded once and |r| <= 0.0081;
// log u = e ln 2 - log inv_k + log1p(r), log1p(r) = r - r^2/2 + ... + r^7/7 (truncation <= r^8/8 = 2.3e-18).  Against long
// double over 1e7 arguments: <= 1.1e-16 absolute for u in [1, 2], <= 1.3 ulp beyond.  Nine FMAs, no division, no exp.
// tbl: entry k of copy `rep` at tbl[(k * REP + rep) * 2]
template <int REP>
PG_HD double pg_log_ge1(double u, const double* tbl, int rep) {
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(u);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(u));
    const double2 t = *reinterpret_cast<const double2*>(tbl + (size_t)((((hi >> 14) & (PG_LOGT_K - 1)) * REP + rep) * 2));
    const double inv = t.x, lc = t.y;
#else
    unsigned long long bits;
    memcpy(&bits, &u, 8);
    const int hi = (int)(bits >> 32);
    const unsigned long long mb = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m;
    memcpy(&m, &mb, 8);
    const double* t = tbl + (size_t)((((hi >> 14) & (PG_LOGT_K - 1)) * REP + rep) * 2);
    const double inv = t[0], lc = t[1];
#endif
    const double e = (double)((hi >> 20) - 1023);
    const double r = fma(m, inv, -1.0);
    double p = 1.0 / 7.0;
    p = fma(p, r, -1.0 / 6.0); p = fma(p, r, 0.2); p = fma(p, r, -0.25); p = fma(p, r, 1.0 / 3.0); p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    return fma(p, r, fma(e, 0.6931471805599453, lc));
}

// exp(o) of a row / exp(v) of a leaf value for pg_bern_term; +inf marks "outside the range where the product form is
// safe" and sends the term to the direct formula
PG_HD double pg_bern_exp(double z) { return fabs(z) < 300.0 ? exp(z) : INFINITY; }

// y mu - softplus(mu) for mu = o + v with eo = pg_bern_exp(o) (once per row) and ev = pg_bern_exp(v) (once per job and
// child): softplus(mu) = log(1 + exp(o) exp(v)), one FMA and one table logarithm instead of an exp and a log1p per row-job.
template <int REP>
PG_HD double pg_bern_term(double y, double o, double v, double eo, double ev, const double* tbl, int rep) {
    const double mu = o + v;
    const double u = fma(eo, ev, 1.0);
    if (!(u < 1e300)) return y * mu - pg_softplus(mu);
    return y * mu - pg_log_ge1<REP>(u, tbl, rep);
}

// ---------------------------------------------------------------------------

PG_HDN inline void pg_s_begin_step(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    PG_SINGLE {
        const double frac = c->tune ? st.batch_tune : st.batch_post;       // kernel.rs:65-72
        long long b = (long long)floor(frac * (double)st.m + 0.5);          // f64::round for frac >= 0
        if (b < 1) b = 1;
        if (b > st.m) b = st.m;
        c->batch = (int)b;
        c->k = 0;
        c->t_add = -1;
        c->info_loglik = 0.0;
        c->info_acc = 0;
        c->info_n = 0;
        c->snext = SN_BEGIN_UPDATE;
    }
}

// smc.rs:42-53 — fresh particles for one tree update.
PG_HDN inline void pg_s_begin_update(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    PG_SINGLE {
        const int t = (c->next_tree_idx + c->k) % st.m;  // kernel.rs:82
        c->t_cur = t;
        c->t_sub = t;
        c->round = 0;
        c->acc_update = 0;
        c->pool_top = 0;
    }
    tm.sync();
    const int cur = c->cur;
    PG_FOR(s, st.S) {
        const size_t i0 = pg_tix(st, cur, s, 0);
        st.pt_split_var[i0] = -1;
        st.pt_split_val[i0] = pg_nan();
        st.pt_thr32[i0] = 0.0f;
        st.pt_leaf_val[i0] = st.init_leaf;
        st.pt_seg_off[i0] = PG_SEG_IDENTITY;
        st.pt_cnt[i0] = (unsigned int)st.n;
        st.pt_So[i0] = 0.0; st.pt_Sr[i0] = 0.0; st.pt_g[i0] = 0.0;   // filled after the root pass
        st.pt_queue[i0] = 0;
        const int ss = pg_six(st, cur, s);
        st.pt_size[ss] = 1; st.pt_qhead[ss] = 0; st.pt_qtail[ss] = 1; st.pt_G[ss] = 0.0;
        st.w[s] = 0.0;  // smc.rs:53: initialised once per update, never reset (Q1)
    }
    tm.sync();
    PG_SINGLE { c->snext = SN_DRAW_ROUND; }
}

// Split value for a job whose min/max are known (splitting.rs:39-56).
PG_HD void pg_decide_split(const PgState& st, int j) {
    PgCtrl* c = st.ctrl;
    const int s = st.job_slot[j];
    const int node = st.job_node[j];
    double* tr = pg_trace_slot(st, c->round, s);
    if (st.onehot && st.onehot[st.job_var[j]]) {
        // splitting.rs:127-137 -> :79-90: one of the node's distinct integer codes, uniformly (taken in ascending order;
        // the reference's HashSet order is process-random, Q7); random_range(0..k) = floor of the f64 range draw on [0, k)
        const unsigned long long bits = st.job_bits[j];
        const int imin = st.col_imin[st.job_var[j]];
        int nu = 0, first = -1, last = -1;
        for (int b = 0; b < 64; ++b) if ((bits >> b) & 1ull) { nu += 1; if (first < 0) first = b; last = b; }
        if (tr) { tr[9] = (double)(imin + first); tr[10] = (double)(imin + last); }
        if (nu <= 1) { if (tr) tr[0] = 3; st.job_node[j] = -1; return; }     // splitting.rs:85-87
        if (2 * node + 2 >= st.maxn) { c->error = PG_ERR_DEPTH; st.job_node[j] = -1; return; }
        int k = (int)pg_split_draw(st, c->upd, c->round, s, 0.0, (double)nu);
        if (k > nu - 1) k = nu - 1;
        int code = imin;
        for (int b = 0, seen = 0; b < 64; ++b) if ((bits >> b) & 1ull) { if (seen == k) { code = imin + b; break; } seen += 1; }
        st.job_split[j] = (double)code;
        st.job_thr32[j] = (float)code;            // |code| < 2^24 (checked at creation): exact, and x < (float)code <=> (double)x < code
        if (tr) tr[3] = (double)code;
        return;
    }
    const float lo = st.job_lo[j], hi = st.job_hi[j];
    if (tr) { tr[9] = (double)lo; tr[10] = (double)hi; }
    if (lo >= hi) {                 // splitting.rs:51-53
        if (tr) tr[0] = 3;
        st.job_node[j] = -1;        // job dropped: particle not mutated, no further draws
        return;
    }
    if (2 * node + 2 >= st.maxn) {  // tree.rs:98-102: the reference panics
        c->error = PG_ERR_DEPTH;
        st.job_node[j] = -1;
        return;
    }
    const double split = pg_split_draw(st, c->upd, c->round, s, (double)lo, (double)hi);
    st.job_split[j] = split;
    st.job_thr32[j] = pg_thr32(split);
    if (tr) tr[3] = split;
}

// Drop dead jobs (order preserved) and rebuild slot->job.
PG_HDN inline void pg_compact_jobs_single(const PgState& st) {
    PgCtrl* c = st.ctrl;
    int J = 0;
    for (int s = 0; s < st.S; ++s) st.slot_job[s] = -1;
    for (int j = 0; j < c->n_jobs; ++j) {
        if (st.job_node[j] < 0) continue;
        if (J != j) {
            st.job_slot[J] = st.job_slot[j]; st.job_node[J] = st.job_node[j]; st.job_var[J] = st.job_var[j];
            st.job_thr32[J] = st.job_thr32[j]; st.job_split[J] = st.job_split[j];
            st.job_seg_off[J] = st.job_seg_off[j]; st.job_len[J] = st.job_len[j];
            st.job_lo[J] = st.job_lo[j]; st.job_hi[J] = st.job_hi[j];
            if (st.onehot) st.job_bits[J] = st.job_bits[j];
        }
        st.slot_job[st.job_slot[J]] = J;
        ++J;
    }
    c->n_jobs = J;
}

// job_order: permutation with the jobs of one index segment adjacent (first-seen order);
// grp_first: group boundaries in job_order.  J is small (<= S).  Returns the number of groups.
PG_HDN inline int pg_group_jobs_arrays(const unsigned int* seg_off, const unsigned int* seg_len, int* job_order, int* grp_first, int J) {
    int G = 0, filled = 0;
    bool one = true;   // common case (round 1, or all slots copies of one ancestor): a single segment
    for (int j = 1; j < J; ++j)
        if (seg_off[j] != seg_off[0] || seg_len[j] != seg_len[0]) { one = false; break; }
    if (one) {
        for (int j = 0; j < J; ++j) job_order[j] = j;
        grp_first[0] = 0; grp_first[J > 0 ? 1 : 0] = J;
        return J > 0 ? 1 : 0;
    }
    for (int j = 0; j < J; ++j) job_order[j] = -1;
    for (int j = 0; j < J; ++j) {
        bool placed = false;
        for (int q = 0; q < filled; ++q) if (job_order[q] == j) { placed = true; break; }
        if (placed) continue;
        grp_first[G++] = filled;
        const unsigned int off = seg_off[j], len = seg_len[j];
        for (int q = j; q < J; ++q)
            if (seg_off[q] == off && seg_len[q] == len) job_order[filled++] = q;
    }
    grp_first[G] = filled;
    return G;
}
PG_HDN inline void pg_group_jobs_single(const PgState& st, PgCtrl* c) {
    c->n_groups = pg_group_jobs_arrays(st.job_seg_off, st.job_len, st.job_order, st.grp_first, c->n_jobs);
}

PG_HDN inline void pg_group_jobs_single(const PgState& st) { pg_group_jobs_single(st, st.ctrl); }

PG_HDN inline int pg_count_unique_vars(const PgState& st) {
    const int J = st.ctrl->n_jobs;
    int u = 0;
    for (int j = 0; j < J; ++j) {
        bool dup = false;
        for (int q = 0; q < j; ++q) if (st.job_var[q] == st.job_var[j]) { dup = true; break; }
        if (!dup) ++u;
    }
    return u;
}

PG_HDN inline int pg_tree_internal_nodes(const PgState& st, int t) {
    if (t < 0) return 0;
    int cnt = 0;
    const int sz = st.f_size[t];
    for (int i = 0; i < sz; ++i) if (st.f_split_var[(size_t)t * st.maxn + i] >= 0) ++cnt;
    return cnt;
}

// smc.rs:61-86 front half: pop the next expandable node of every slot, depth-prior test, split variable.
PG_HDN inline void pg_s_draw_round(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    const int cur = c->cur;
    const int r = c->round;
    const unsigned long long upd = c->upd;
    PG_FOR(s, st.S) {
        st.mutated[s] = 0;
        st.slot_job[s] = -1;
        st.pend_node[s] = -1;
        double* tr = pg_trace_slot(st, r, s);
        if (tr) { for (int f = 0; f < PG_TS_FIELDS; ++f) tr[f] = pg_nan(); tr[0] = 0; }
        const int ss = pg_six(st, cur, s);
        const int qh = st.pt_qhead[ss];
        if (qh >= st.pt_qtail[ss]) continue;                      // no expandable node
        const int node = st.pt_queue[pg_tix(st, cur, s, qh)];
        st.pt_qhead[ss] = qh + 1;                                  // popped on accept and on reject (smc.rs:77,83)
        if (tr) { tr[0] = 1; tr[1] = (double)node; tr[2] = -1; }
        const int depth = pg_depth_of(node);
        const double u1 = pg_uniform(st, PG_D_GROW, upd, r, s);
        if (st.thr_depth[depth] > u1) continue;                   // smc.rs:145-147
        if (st.pt_cnt[pg_tix(st, cur, s, node)] == 0u) {          // smc.rs:149-152
            if (tr) tr[0] = 2;
            continue;
        }
        const double u2 = pg_uniform(st, PG_D_VAR, upd, r, s);
        int var;
        if (st.prior) var = pg_feature_from_probs(st, u2);
        else { var = (int)(u2 * (double)st.p); if (var > st.p - 1) var = st.p - 1; }
        if (tr) tr[2] = (double)var;
        st.pend_node[s] = node;
        st.pend_var[s] = var;
    }
    tm.sync();
    PG_SCRATCH(int, s_flag, PG_SCR_N);
    PG_SCRATCH(int, s_var, PG_SCR_N);
    PG_SCRATCH(int, s_cnt, 2);
    const bool staged = st.S <= PG_SCR_N;
    if (staged) {
        // jobs in slot order: each proposing slot finds its job number from the flags before it
        PG_FOR(s, st.S) s_flag[s] = st.pend_node[s] >= 0 ? 1 : 0;
        tm.sync();
        PG_FOR(s, st.S) {
            const int J = pg_flags_before(s_flag, s);
            if (s_flag[s]) {
                const int node = st.pend_node[s];
                const size_t iN = pg_tix(st, cur, s, node);
                st.job_slot[J] = s; st.job_node[J] = node; st.job_var[J] = st.pend_var[s];
                st.job_seg_off[J] = st.pt_seg_off[iN]; st.job_len[J] = st.pt_cnt[iN];
                st.slot_job[s] = J;
            }
            if (s == st.S - 1) c->n_jobs = J + s_flag[s];
        }
    } else {
        PG_SINGLE {
            int J = 0;
            for (int s = 0; s < st.S; ++s) {
                if (st.pend_node[s] < 0) continue;
                const int node = st.pend_node[s];
                const size_t iN = pg_tix(st, cur, s, node);
                st.job_slot[J] = s; st.job_node[J] = node; st.job_var[J] = st.pend_var[s];
                st.job_seg_off[J] = st.pt_seg_off[iN]; st.job_len[J] = st.pt_cnt[iN];
                st.slot_job[s] = J;
                ++J;
            }
            c->n_jobs = J;
        }
    }
    tm.sync();
    const int J = c->n_jobs;
    if (r == 0) {
        // round 1: every node is the root, whose column min/max are per-column constants
        PG_FOR(j, J) {
            st.job_lo[j] = st.col_min[st.job_var[j]];
            st.job_hi[j] = st.col_max[st.job_var[j]];
            if (st.onehot) st.job_bits[j] = st.col_bits[st.job_var[j]];
            pg_decide_split(st, j);
        }
        tm.sync();
        // Dropped jobs (constant column) are rare: the team counts them and the distinct split variables; only when a
        // job was dropped does one thread compact the list.  All round-1 jobs share the root's identity segment: one group.
        int fast_cols = -1;
        if (staged) {
            PG_SINGLE { s_cnt[0] = 0; s_cnt[1] = 0; }
            PG_FOR(j, J) s_var[j] = st.job_node[j] >= 0 ? st.job_var[j] : -1;
            tm.sync();
            {
                int dropped = 0, uniq = 0;
                PG_FOR(j, J) {
                    const int v = s_var[j];
                    if (v < 0) { dropped += 1; continue; }
                    // position of job j when the jobs are ordered by split variable (stable): consecutive jobs of the
                    // segment passes then read the same column (one group: every round-1 job is on the root's rows)
                    int before = 0, same_before = 0;
                    for (int q = 0; q < J; ++q) {
                        const int vq = s_var[q];
                        before += (vq < v) ? 1 : 0;
                        same_before += (vq == v && q < j) ? 1 : 0;
                    }
                    if (same_before == 0) uniq += 1;
                    st.job_order[before + same_before] = j;
                }
                pg_team_add(&s_cnt[0], dropped);
                pg_team_add(&s_cnt[1], uniq);
            }
            tm.sync();
            if (s_cnt[0] == 0) fast_cols = s_cnt[1];
        }
        PG_SINGLE {
            const int J0 = c->n_jobs;
            int ucols;
            if (fast_cols >= 0) {
                st.grp_first[0] = 0; st.grp_first[J0 > 0 ? 1 : 0] = J0;
                c->n_groups = J0 > 0 ? 1 : 0;
                ucols = fast_cols;
            } else {
                pg_compact_jobs_single(st);
                pg_group_jobs_single(st);
                ucols = pg_count_unique_vars(st);
            }
            const unsigned long long n = (unsigned long long)st.n;
            const int cols = ucols + pg_tree_internal_nodes(st, c->t_add)
                           + pg_tree_internal_nodes(st, c->t_sub);
            c->bytes_model += n * (8 + 8 + 4) + 4ull * n * (unsigned long long)cols;
            c->bytes_contract += 20ull * n + 4ull * n * (unsigned long long)J0 + 16ull * n * (unsigned long long)c->n_jobs;
            c->phase = PH_ROOT;    // the root pass always runs (it also commits the previous tree and gives the totals)
            c->n_phases += 1;
            c->snext = SN_NONE;
        }
    } else {
        PG_SINGLE {
            if (J > 0) {
                pg_group_jobs_single(st);
                unsigned long long bm = 0, bc = 0;
                for (int g = 0; g < c->n_groups; ++g) {
                    const unsigned long long L = st.job_len[st.job_order[st.grp_first[g]]];
                    bm += 4ull * L + 4ull * L * (unsigned long long)(st.grp_first[g + 1] - st.grp_first[g]);
                }
                for (int j = 0; j < J; ++j) bc += 8ull * st.job_len[j];
                c->bytes_model += bm; c->bytes_contract += bc;
                c->phase = PH_MINMAX;
                c->n_phases += 1;
                c->snext = SN_NONE;
            } else {
                c->snext = SN_FINISH_ROUND;   // nobody proposes: the round still normalises and resamples
            }
        }
    }
}

// After PH_MINMAX: per-job min/max over blocks, then the split value.
PG_HDN inline void pg_s_after_minmax(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    const int J = c->n_jobs;
    PG_FOR(j, J) {
        uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
        for (int b = 0; b < st.grid; ++b) {
            const uint32_t a = st.part_mm[((size_t)b * 2 * st.S + j) * 2 + 0];
            const uint32_t z = st.part_mm[((size_t)b * 2 * st.S + j) * 2 + 1];
            kmin = a < kmin ? a : kmin;
            kmax = z > kmax ? z : kmax;
        }
        st.job_lo[j] = pg_key2f(kmin);
        st.job_hi[j] = pg_key2f(kmax);
        if (st.onehot && st.onehot[st.job_var[j]]) {       // the pass wrote the low / high words of the presence mask instead
            unsigned long long bits = 0ull;
            for (int b = 0; b < st.grid; ++b)
                bits |= (unsigned long long)st.part_mm[((size_t)b * 2 * st.S + j) * 2 + 0]
                      | ((unsigned long long)st.part_mm[((size_t)b * 2 * st.S + j) * 2 + 1] << 32);
            st.job_bits[j] = bits;
        }
        pg_decide_split(st, j);
    }
    tm.sync();
    PG_SINGLE {
        pg_compact_jobs_single(st);
        if (c->n_jobs > 0) {
            pg_group_jobs_single(st);
            unsigned long long bm = 0, bc = 0;
            for (int g = 0; g < c->n_groups; ++g) {
                const unsigned long long L = st.job_len[st.job_order[st.grp_first[g]]];
                bm += (4ull + 8ull + 4ull) * L + 4ull * L * (unsigned long long)(st.grp_first[g + 1] - st.grp_first[g]);
            }
            for (int j = 0; j < c->n_jobs; ++j) bc += 20ull * st.job_len[j];
            c->bytes_model += bm; c->bytes_contract += bc;
            c->phase = PH_STATS;
            c->n_phases += 1;
            c->snext = SN_NONE;
        } else {
            c->snext = SN_FINISH_ROUND;
        }
    }
}

// particle.rs:91-148 bookkeeping half + smc.rs:76-81: split the slot's node, record child statistics.
PG_HD void pg_apply_mutation(const PgState& st, int j, double vL, double vR, double gL, double gR,
                             unsigned int cL, unsigned int cR, double SoL, double SoR, double SrL, double SrR) {
    PgCtrl* c = st.ctrl;
    const int cur = c->cur;
    const int s = st.job_slot[j];
    const int node = st.job_node[j];
    const int L = 2 * node + 1, R = 2 * node + 2;
    const size_t iN = pg_tix(st, cur, s, node), iL = pg_tix(st, cur, s, L), iR = pg_tix(st, cur, s, R);
    const int ss = pg_six(st, cur, s);
    const int old_size = st.pt_size[ss];
    for (int q = old_size; q < R + 1; ++q) {          // tree.rs:105-110 filler nodes
        const size_t iq = pg_tix(st, cur, s, q);
        st.pt_split_var[iq] = -1; st.pt_split_val[iq] = pg_nan(); st.pt_thr32[iq] = 0.0f; st.pt_leaf_val[iq] = 0.0;
        st.pt_seg_off[iq] = PG_SEG_NONE; st.pt_cnt[iq] = 0u; st.pt_So[iq] = 0.0; st.pt_Sr[iq] = 0.0; st.pt_g[iq] = 0.0;
    }
    st.pt_split_var[iN] = st.job_var[j];
    st.pt_split_val[iN] = st.job_split[j];
    st.pt_thr32[iN] = st.job_thr32[j];
    st.pt_leaf_val[iN] = pg_nan();
    st.pt_split_var[iL] = -1; st.pt_split_val[iL] = pg_nan(); st.pt_thr32[iL] = 0.0f; st.pt_leaf_val[iL] = vL;
    st.pt_split_var[iR] = -1; st.pt_split_val[iR] = pg_nan(); st.pt_thr32[iR] = 0.0f; st.pt_leaf_val[iR] = vR;
    st.pt_seg_off[iL] = PG_SEG_NONE; st.pt_seg_off[iR] = PG_SEG_NONE;
    st.pt_cnt[iL] = cL; st.pt_cnt[iR] = cR;
    st.pt_So[iL] = SoL; st.pt_So[iR] = SoR; st.pt_Sr[iL] = SrL; st.pt_Sr[iR] = SrR;
    st.pt_g[iL] = gL; st.pt_g[iR] = gR;
    if (R + 1 > old_size) st.pt_size[ss] = R + 1;
    st.pt_G[ss] = (st.pt_G[ss] - st.pt_g[iN]) + (gL + gR);
    const int qt = st.pt_qtail[ss];
    st.pt_queue[pg_tix(st, cur, s, qt)] = L;           // particle.rs:146-147
    st.pt_queue[pg_tix(st, cur, s, qt + 1)] = R;
    st.pt_qtail[ss] = qt + 2;
    st.mutated[s] = 1;
    st.w[s] = c->C0 + st.pt_G[ss];                     // smc.rs:89-95: fresh log-likelihood of the mutated particle
    double* tr = pg_trace_slot(st, c->round, s);
    if (tr) {
        tr[0] = 4; tr[4] = vL; tr[5] = vR; tr[6] = (double)cL; tr[7] = (double)cR; tr[8] = st.w[s]; tr[11] = SoL;
    }
}

// Gaussian leaf term: sum_{i in leaf} -(r_i - v)^2 / (2 s^2) without the (particle-independent) r^2 part.
PG_HD double pg_gauss_g(double v, double Sr, unsigned int cnt, double inv_s2) {
    return (v * Sr - 0.5 * (double)cnt * v * v) * inv_s2;
}

// After PH_ROOT / PH_STATS: reduce partials, draw leaf values (smc.rs:219-239); Gaussian: weights.
PG_HDN inline void pg_s_after_stats(const PgState& st, const PgTeam& tm, bool root) {
    PgCtrl* c = st.ctrl;
    const int J = c->n_jobs;
    const int cur = c->cur;
    const bool bern = st.family == PG_BERNOULLI_LOGIT;
    if (root) {
        PG_FOR(q, 4) {
            const double v = pg_reduce_blocks(st.part_tot, st.grid, 4, (size_t)q);
            if (q == 0) c->T_o = v; else if (q == 1) c->T_r = v; else if (q == 2) c->T_rr = v; else c->T_ll0 = v;
        }
    }
    PG_FOR(idx, J * 3) {
        const int j = idx / 3, q = idx % 3;
        if (q == 0) st.red_So[j] = pg_reduce_blocks(st.part_sum, st.grid, (size_t)st.S * 2, (size_t)j * 2);
        else if (q == 1) st.red_Sr[j] = pg_reduce_blocks(st.part_sum, st.grid, (size_t)st.S * 2, (size_t)j * 2 + 1);
        else {
            unsigned long long cnt = 0;
            for (int b = 0; b < st.grid; ++b) cnt += st.part_ccnt[(size_t)b * st.S + j];
            st.red_cnt[j] = cnt;
        }
    }
    tm.sync();
    if (root) {
        PG_SINGLE {
            const double nn = (double)st.n;
            if (st.family == PG_GAUSSIAN) {
                const double s2 = st.lik_sigma * st.lik_sigma;
                c->C0 = -nn * (log(st.lik_sigma) + 0.5 * log(6.283185307179586476925286766559)) - 0.5 * c->T_rr / s2;
                c->g_stump = pg_gauss_g(st.init_leaf, c->T_r, (unsigned int)st.n, 1.0 / s2);
            } else if (st.family == PG_GAUSSIAN_KERNEL) {
                c->C0 = -0.5 * c->T_rr;
                c->g_stump = pg_gauss_g(st.init_leaf, c->T_r, (unsigned int)st.n, 1.0);
            } else {
                c->C0 = 0.0;
                c->g_stump = c->T_ll0;
            }
        }
        if (st.pg_correct) {
            // conditional reference particle: the log-likelihood of the current predictions (all trees, the old one included)
            PG_SINGLE {
                const double t_old = pg_reduce_blocks(st.part_old, st.grid, 1, 0);
                if (st.family == PG_GAUSSIAN) c->ll_old = -(double)st.n * (log(st.lik_sigma) + 0.5 * log(6.283185307179586476925286766559)) - 0.5 * t_old / (st.lik_sigma * st.lik_sigma);
                else if (st.family == PG_GAUSSIAN_KERNEL) c->ll_old = -0.5 * t_old;
                else c->ll_old = t_old;
            }
        }
        tm.sync();
        PG_FOR(s, st.S) {
            const size_t i0 = pg_tix(st, cur, s, 0);
            st.pt_So[i0] = c->T_o; st.pt_Sr[i0] = c->T_r; st.pt_g[i0] = c->g_stump;
            st.pt_G[pg_six(st, cur, s)] = c->g_stump;
            if (st.pg_correct) st.w[s] = c->C0 + c->g_stump;      // every slot starts as a stump: its log-likelihood, not 0.0 (Q1)
        }
        tm.sync();
    }
    const double inv_s2 = st.family == PG_GAUSSIAN ? 1.0 / (st.lik_sigma * st.lik_sigma) : 1.0;
    PG_FOR(j, J) {
        const int s = st.job_slot[j];
        const size_t iN = pg_tix(st, cur, s, st.job_node[j]);
        const unsigned int nc = st.pt_cnt[iN];
        const unsigned int cL = (unsigned int)st.red_cnt[j], cR = nc - cL;
        const double SoL = st.red_So[j], SoR = st.pt_So[iN] - SoL;
        const double zL = pg_normal(st, PG_D_ZLEFT, c->upd, c->round, s);     // drawn for both children, left first
        const double zR = pg_normal(st, PG_D_ZRIGHT, c->upd, c->round, s);
        const double vL = cL == 0u ? zL * st.leaf_sigma : pg_mul_add(zL, st.leaf_sigma, SoL / (double)cL / st.m_f);
        const double vR = cR == 0u ? zR * st.leaf_sigma : pg_mul_add(zR, st.leaf_sigma, SoR / (double)cR / st.m_f);
        st.job_vL[j] = vL; st.job_vR[j] = vR;
        if (!bern) {
            const double SrL = st.red_Sr[j], SrR = st.pt_Sr[iN] - SrL;
            pg_apply_mutation(st, j, vL, vR, pg_gauss_g(vL, SrL, cL, inv_s2), pg_gauss_g(vR, SrR, cR, inv_s2),
                              cL, cR, SoL, SoR, SrL, SrR);
        }
    }
    tm.sync();
    PG_SCRATCH(unsigned long long, s_bytes, 1);
    if (bern && J > 0) {
        PG_SINGLE { s_bytes[0] = 0ull; }
        tm.sync();
        unsigned long long bc = 0;
        PG_FOR(j, J) bc += 12ull * st.job_len[j];
        pg_team_add64(&s_bytes[0], bc);
        tm.sync();
    }
    PG_SINGLE {
        if (bern && J > 0) {
            unsigned long long bm = 0;
            const unsigned long long bc = s_bytes[0];
            for (int g = 0; g < c->n_groups; ++g) {
                const int j0 = st.job_order[st.grp_first[g]];
                const unsigned long long L = st.job_len[j0];
                const unsigned long long fixed = st.job_seg_off[j0] == PG_SEG_IDENTITY ? 12ull : 16ull;
                bm += fixed * L + 4ull * L * (unsigned long long)(st.grp_first[g + 1] - st.grp_first[g]);
            }
            c->bytes_model += bm; c->bytes_contract += bc;
            c->phase = PH_BERN;
            c->n_phases += 1;
            c->snext = SN_NONE;
        } else {
            c->snext = SN_FINISH_ROUND;
        }
    }
}

// After PH_BERN: child log-likelihoods are known; complete the mutations.
PG_HDN inline void pg_s_after_bern(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    const int J = c->n_jobs;
    const int cur = c->cur;
    PG_FOR(idx, J * 2) {
        const int j = idx / 2, q = idx % 2;
        const double v = pg_reduce_blocks(st.part_ll, st.grid, (size_t)st.S * 2, (size_t)j * 2 + (size_t)q);
        if (q == 0) st.red_llL[j] = v; else st.red_llR[j] = v;
    }
    tm.sync();
    PG_FOR(j, J) {
        const int s = st.job_slot[j];
        const size_t iN = pg_tix(st, cur, s, st.job_node[j]);
        const unsigned int nc = st.pt_cnt[iN];
        const unsigned int cL = (unsigned int)st.red_cnt[j], cR = nc - cL;
        const double SoL = st.red_So[j], SoR = st.pt_So[iN] - SoL;
        const double SrL = st.red_Sr[j], SrR = st.pt_Sr[iN] - SrL;
        pg_apply_mutation(st, j, st.job_vL[j], st.job_vR[j], st.red_llL[j], st.red_llR[j], cL, cR, SoL, SoR, SrL, SrR);
    }
    tm.sync();
    PG_SINGLE { c->snext = SN_FINISH_ROUND; }
}

// smc.rs:96-102: normalise (stale-weight rule, Q1), systematic resampling, clone survivors.
PG_HDN inline void pg_s_finish_round(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    const int S = st.S;
    const int src = c->cur, dst = 1 - c->cur;
    PG_SCRATCH(double, s_w, PG_SCR_N);
    PG_SCRATCH(int, s_anc, PG_SCR_N);
    PG_SCRATCH(double, s_red, 2);
    PG_SCRATCH(int, s_ired, 2);
    const bool staged = S <= PG_SCR_N;
    double* w = staged ? s_w : st.w;
    int* anc = staged ? s_anc : st.anc;
    PG_SCRATCH(double, s_lw, PG_SCR_N);        // weight_mode pg_correct: the slots' log-likelihoods, permuted with the ancestors below
    PG_SINGLE { s_ired[0] = 0; s_ired[1] = 1; }
    if (staged) { PG_FOR(s, S) { w[s] = st.w[s]; s_lw[s] = st.w[s]; } }
    tm.sync();
    {   // mutated slots and the largest slot tree: integer partial results, order-free
        int nm = 0, mz = 1;
        PG_FOR(s, S) { nm += st.mutated[s]; const int z = st.pt_size[pg_six(st, src, s)]; mz = z > mz ? z : mz; }
        pg_team_add(&s_ired[0], nm);
        pg_team_max(&s_ired[1], mz);
    }
    // smc.rs:257-268.  w[] keeps last round's normalised values for slots that did not mutate.
    PG_SINGLE {
        double mx = -INFINITY;
        for (int s = 0; s < S; ++s) mx = fmax(mx, w[s]);
        s_red[0] = mx;
    }
    tm.sync();
    { const double mx = s_red[0]; PG_FOR(s, S) w[s] = exp(w[s] - mx); }
    tm.sync();
    PG_SINGLE {
        double sum = 0.0;
        for (int s = 0; s < S; ++s) sum += w[s];      // sequential, in slot order, like the reference's iter().sum()
        s_red[1] = sum;
    }
    tm.sync();
    { const double sum = s_red[1]; PG_FOR(s, S) w[s] /= sum; }
    tm.sync();
    PG_SINGLE {
        const int n_mut = s_ired[0];
        c->acc_update += n_mut;
        // resampling.rs:24-44
        const double u6 = pg_uniform(st, PG_D_RESAMPLE, c->upd, c->round, 0);
        int cidx = 0;
        double cum = 0.0;
        for (int i = 0; i < S; ++i) {
            const double target = ((double)i + u6) / (double)S;
            while (cum < target && cidx < S) { cum += w[cidx]; cidx += 1; }
            anc[i] = cidx == 0 ? 0 : cidx - 1;
        }
        c->max_size = s_ired[1];
    }
    tm.sync();
    // Reference (Q1): the normalised weights stay in w[] and meet next round's fresh log-likelihoods.  pg_correct: slot s
    // continues with the log-likelihood of the particle it now holds.
    if (staged) { PG_FOR(s, S) { st.w[s] = st.pg_correct ? s_lw[anc[s]] : w[s]; st.anc[s] = anc[s]; } }
    if (st.tr_round) {
        if ((long long)c->trace_n < st.tr_u_cap && c->round < st.tr_r_cap) {
            double* tr = st.tr_round + ((size_t)c->trace_n * st.tr_r_cap + c->round) * (size_t)(2 + 2 * S);
            PG_SINGLE { tr[0] = 1; tr[1] = (double)s_ired[0]; }
            PG_FOR(s, S) { tr[2 + s] = w[s]; tr[2 + S + s] = (double)anc[s]; }
        } else { PG_SINGLE { c->trace_overflow = 1; } }
    }
    tm.sync();
    // smc.rs:99-102: slot i becomes a copy of slot anc[i] (tree, queue, per-node tables).
    const int mxs = c->max_size;
    PG_FOR(idx, S * mxs) {
        const int i = idx / mxs, node = idx % mxs;
        const int a = st.anc[i];
        if (node >= st.pt_size[pg_six(st, src, a)]) continue;
        const size_t f = pg_tix(st, src, a, node), t = pg_tix(st, dst, i, node);
        st.pt_split_var[t] = st.pt_split_var[f]; st.pt_thr32[t] = st.pt_thr32[f];
        st.pt_split_val[t] = st.pt_split_val[f]; st.pt_leaf_val[t] = st.pt_leaf_val[f];
        st.pt_seg_off[t] = st.pt_seg_off[f]; st.pt_cnt[t] = st.pt_cnt[f];
        st.pt_So[t] = st.pt_So[f]; st.pt_Sr[t] = st.pt_Sr[f]; st.pt_g[t] = st.pt_g[f];
        st.pt_queue[t] = st.pt_queue[f];   // queue entries live in [qhead, qtail) <= size
    }
    PG_FOR(i, S) {
        const int a = st.anc[i];
        const int fs = pg_six(st, src, a), ts = pg_six(st, dst, i);
        st.pt_size[ts] = st.pt_size[fs]; st.pt_qhead[ts] = st.pt_qhead[fs]; st.pt_qtail[ts] = st.pt_qtail[fs];
        st.pt_G[ts] = st.pt_G[fs];
        st.pj_of_anc[i] = -1;
    }
    tm.sync();
    {   // does any slot still have queued nodes / did any surviving ancestor grow?  (order-free: the whole team looks)
        PG_SINGLE { s_ired[0] = 0; s_ired[1] = 0; }
        tm.sync();
        int aq = 0, am = 0;
        PG_FOR(i, S) {
            const int ts = pg_six(st, dst, i);
            if (st.pt_qhead[ts] < st.pt_qtail[ts]) aq = 1;
            if (st.mutated[st.anc[i]]) am = 1;
        }
        pg_team_max(&s_ired[0], aq);
        pg_team_max(&s_ired[1], am);
        tm.sync();
    }
    PG_SINGLE {
        c->cur = dst;
        // Survivors that grew this round need their children's index lists: one stable partition per
        // distinct surviving ancestor (particles that died are never materialised), in order of first appearance.
        int K = 0;
        const int any_queue = s_ired[0];
        if (s_ired[1]) {
            for (int i = 0; i < S; ++i) {
                const int a = st.anc[i];
                if (!st.mutated[a] || st.pj_of_anc[a] >= 0) continue;
                const int j = st.slot_job[a];
                st.pj_of_anc[a] = K;
                st.pj_anc[K] = a; st.pj_job[K] = j; st.pj_var[K] = st.job_var[j]; st.pj_thr32[K] = st.job_thr32[j];
                st.pj_src_off[K] = st.job_seg_off[j]; st.pj_len[K] = st.job_len[j];
                st.pj_cntL[K] = (unsigned int)st.red_cnt[j];
                if (c->pool_top + st.job_len[j] > st.pool_cap) { c->error = PG_ERR_POOL; }
                st.pj_dst_off[K] = (unsigned int)c->pool_top;
                c->pool_top += st.job_len[j];
                ++K;
            }
        }
        c->n_pjobs = K;
        c->any_queue = any_queue;
        if (c->error != PG_OK) { c->phase = PH_DONE; c->snext = SN_NONE; }
        else if (K > 0) c->snext = SN_PREP_PART;
        else if (any_queue) { c->round += 1; c->snext = SN_DRAW_ROUND; }
        else c->snext = SN_FINALIZE;
    }
}

// Exclusive prefix, over global warps, of each partition job's per-warp left counts.
PG_HDN inline void pg_s_prep_part(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    const int K = c->n_pjobs;
    const int W = st.n_gwarps;
    const int chunk = (W + tm.size - 1) / tm.size;
    for (int k = 0; k < K; ++k) {
        const int j = st.pj_job[k];
        const int b = tm.rank * chunk, e = (b + chunk < W) ? b + chunk : W;
        unsigned int local = 0;
        for (int g = b; g < e; ++g) local += st.part_cnt[(size_t)g * st.S + j];
        st.scan_tmp[tm.rank] = local;
        tm.sync();
        PG_SINGLE {
            unsigned int run = 0;
            for (int q = 0; q < tm.size; ++q) { const unsigned int v = st.scan_tmp[q]; st.scan_tmp[q] = run; run += v; }
        }
        tm.sync();
        unsigned int run = st.scan_tmp[tm.rank];
        for (int g = b; g < e; ++g) {
            st.pj_warp_left[(size_t)k * W + g] = run;
            run += st.part_cnt[(size_t)g * st.S + j];
        }
        tm.sync();
    }
    PG_SINGLE {
        unsigned long long bm = 0;
        for (int k = 0; k < K; ++k) {
            const unsigned long long L = st.pj_len[k];
            bm += (st.pj_src_off[k] == PG_SEG_IDENTITY ? 0ull : 4ull * L) + 4ull * L + 4ull * L;
        }
        c->bytes_model += bm;
        c->n_groups = 0;      // no look-ahead requests ride on the partition pass here
        c->phase = PH_PART;
        c->n_phases += 1;
        c->snext = SN_NONE;
    }
}

// After PH_PART: every copy of a partitioned ancestor learns where its children's index lists are.
PG_HDN inline void pg_s_after_part(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    const int cur = c->cur;
    PG_FOR(i, st.S) {
        const int k = st.pj_of_anc[st.anc[i]];
        if (k < 0) continue;
        const int node = st.job_node[st.pj_job[k]];
        st.pt_seg_off[pg_tix(st, cur, i, 2 * node + 1)] = st.pj_dst_off[k];
        st.pt_seg_off[pg_tix(st, cur, i, 2 * node + 2)] = st.pj_dst_off[k] + st.pj_cntL[k];
    }
    tm.sync();
    PG_SINGLE { c->round += 1; c->snext = SN_DRAW_ROUND; }   // a survivor that grew always has queued children
}

// smc.rs:105-122 final selection + kernel.rs:99-117 tree replacement and diagnostics.
PG_HDN inline void pg_s_finalize(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    const int P = st.P, S = st.S, cur = c->cur, t = c->t_cur;
    PG_SCRATCH(double, s_W, 2 * PG_SCR_N);
    PG_SCRATCH(double, s_fr, 2);
    const bool staged = P <= PG_SCR_N;
    double* W = staged ? s_W : st.Wfinal;       // [0,P) log-likelihoods, [P,2P) normalised
    PG_SINGLE { W[0] = st.pg_correct ? c->ll_old : c->C0 + c->g_stump; }   // the reference particle is a fresh stump (Q2); pg_correct: the tree being replaced
    PG_FOR(s_, S) W[1 + s_] = c->C0 + st.pt_G[pg_six(st, cur, s_)];
    tm.sync();
    PG_SINGLE {
        double mx = -INFINITY;
        for (int i = 0; i < P; ++i) mx = fmax(mx, W[i]);
        s_fr[0] = mx;
    }
    tm.sync();
    { const double mx = s_fr[0]; PG_FOR(i, P) W[P + i] = exp(W[i] - mx); }
    tm.sync();
    PG_SINGLE {
        double sum = 0.0;
        for (int i = 0; i < P; ++i) sum += W[P + i];      // sequential, in particle order (smc.rs:257-268)
        s_fr[1] = sum;
    }
    tm.sync();
    { const double sum = s_fr[1]; PG_FOR(i, P) W[P + i] /= sum; }
    tm.sync();
    PG_SINGLE {
        // rand WeightedIndex: cumulative weights without the last, uniform in [0,total), count of cum <= chosen
        double total = 0.0;
        for (int i = 0; i < P; ++i) total += W[P + i];
        if (!(total > 0.0) || !(total < INFINITY)) c->error = PG_ERR_WEIGHTS;   // reference: unwrap() panics
        const double u7 = pg_uniform(st, PG_D_SELECT, c->upd, 0, 0);
        const double chosen = u7 * total;
        double cum = 0.0;
        int sel = 0;
        for (int i = 0; i + 1 < P; ++i) {
            cum += W[P + i];
            if (cum <= chosen) sel = i + 1; else break;
        }
        c->any_queue = sel;   // reused as "selected particle" for the copy below
        c->info_loglik += W[P + 0];
        c->info_acc += c->acc_update;
    }
    if (st.tr_upd) {
        tm.sync();
        if ((long long)c->trace_n < st.tr_u_cap) {
            double* tr = st.tr_upd + (size_t)c->trace_n * (size_t)(4 + 2 * P);
            PG_SINGLE { tr[0] = (double)(c->round + 1); tr[1] = (double)c->any_queue; tr[2] = (double)t; tr[3] = (double)c->acc_update; }
            PG_FOR(i, 2 * P) tr[4 + i] = W[i];
        } else { PG_SINGLE { c->trace_overflow = 1; } }
    }
    tm.sync();
    const int sel = c->any_queue;
    const size_t fb = (size_t)t * st.maxn;
    if (sel == 0 && st.pg_correct) {
        // conditional SMC kept the tree being replaced: the forest entry stays as it is
    } else if (sel == 0) {
        PG_SINGLE {
            st.f_split_var[fb] = -1; st.f_split_val[fb] = pg_nan(); st.f_thr32[fb] = 0.0f; st.f_leaf_val[fb] = st.init_leaf;
            st.f_size[t] = 1;
        }
    } else {
        const int s = sel - 1;
        const int sz = st.pt_size[pg_six(st, cur, s)];
        PG_FOR(node, sz) {
            const size_t f = pg_tix(st, cur, s, node);
            st.f_split_var[fb + node] = st.pt_split_var[f]; st.f_split_val[fb + node] = st.pt_split_val[f];
            st.f_thr32[fb + node] = st.pt_thr32[f]; st.f_leaf_val[fb + node] = st.pt_leaf_val[f];
        }
        PG_SINGLE { st.f_size[t] = sz; }
    }
    tm.sync();
    PG_SINGLE {
        int depth = 0;   // kernel.rs:107-112 (filler nodes count as leaves, as in the reference)
        const int sz = st.f_size[t];
        for (int i = 0; i < sz; ++i)
            if (st.f_split_var[fb + i] < 0) { const int d = pg_depth_of(i); depth = d > depth ? d : depth; }
        if (c->info_n < st.m) st.info_depths[c->info_n] = (unsigned char)depth;
        c->info_n += 1;
        st.f_internal[t] = pg_tree_internal_nodes(st, t);
        c->t_add_internal = st.f_internal[t];
        if (st.tr_upd && (long long)c->trace_n < st.tr_u_cap) c->trace_n += 1;
        c->upd += 1;
        c->n_updates += 1;
        c->k += 1;
        c->t_add = t;
        if (c->k < c->batch) c->snext = SN_BEGIN_UPDATE;
        else {
            c->t_sub = -1;
            c->bytes_model += (unsigned long long)st.n * (16ull + 4ull * (unsigned long long)pg_tree_internal_nodes(st, t));
            c->phase = PH_FINAL;
            c->n_phases += 1;
            c->snext = SN_NONE;
        }
    }
}

// Stopwatch of the sub-steps (device builds; read through pgbart_get_profile as serial_sub_*; slot 0 = after_bern here).
#if defined(__CUDA_ARCH__)
#define PG_SUB_T0() pg_sub_t0 = clock64()
#define PG_SUB_T1(i) do { tm.sync(); PG_SINGLE { c->prof2[(i)] += (unsigned long long)(clock64() - pg_sub_t0); c->prof2[8 + (i)] += 1ull; } } while (0)
#else
#define PG_SUB_T0() (void)pg_sub_t0
#define PG_SUB_T1(i) (void)0
#endif

// Entry point: called by one team after the grid pass `ctrl->phase` has completed (its partial
// results are visible).  Runs serial sub-steps until another grid pass is needed or the step is done.
PG_HDN inline void pg_serial(const PgState& st, const PgTeam& tm) {
    PgCtrl* c = st.ctrl;
    long long pg_sub_t0 = 0;
    tm.sync();
    const int phase = c->phase;
    tm.sync();
    PG_SUB_T0();
    switch (phase) {
        case PH_BEGIN: pg_s_begin_step(st, tm); break;
        case PH_ROOT: pg_s_after_stats(st, tm, true); PG_SUB_T1(1); break;
        case PH_MINMAX: pg_s_after_minmax(st, tm); PG_SUB_T1(2); break;
        case PH_STATS: pg_s_after_stats(st, tm, false); PG_SUB_T1(1); break;
        case PH_BERN: pg_s_after_bern(st, tm); PG_SUB_T1(0); break;
        case PH_PART: pg_s_after_part(st, tm); break;
        case PH_FINAL:
            PG_SINGLE {
                c->next_tree_idx = (c->next_tree_idx + c->batch) % st.m;   // kernel.rs:117
                c->t_add = -1;
                c->phase = PH_DONE;
                c->snext = SN_NONE;
            }
            break;
        default:
            PG_SINGLE { c->snext = SN_NONE; }
            break;
    }
    for (int guard = 0; guard < (1 << 20); ++guard) {
        tm.sync();
        const int sn = c->snext;
        const int err = c->error;
        tm.sync();
        if (err != PG_OK) { PG_SINGLE { c->phase = PH_DONE; c->snext = SN_NONE; } break; }
        if (sn == SN_NONE) break;
        PG_SUB_T0();
        switch (sn) {
            case SN_BEGIN_UPDATE: pg_s_begin_update(st, tm); PG_SUB_T1(6); break;
            case SN_DRAW_ROUND: pg_s_draw_round(st, tm); PG_SUB_T1(7); break;
            case SN_FINISH_ROUND: pg_s_finish_round(st, tm); PG_SUB_T1(3); break;
            case SN_PREP_PART: pg_s_prep_part(st, tm); PG_SUB_T1(4); break;
            case SN_FINALIZE: pg_s_finalize(st, tm); PG_SUB_T1(5); break;
            default: PG_SINGLE { c->error = PG_ERR_INTERNAL; } break;
        }
    }
    tm.sync();
}
