// pg_serial_fast.cuh — latency-organised serial section for the persistent kernel (P <= 32).
//
// Same state machine and the same arithmetic as pg_serial.h (which stays the reference
// implementation: host simulation, stepwise driver, P > 32).  What differs is WHERE the state lives
// while the section runs, because every dependent global load here costs ~0.35 us (an L2 round trip
// at 1.97 GHz) and nothing else is running on the GPU meanwhile:
//   * PgCtrl, the per-slot scalars (w, mutated, anc, slot_job, size, queue head/tail, G), the job list
//     of the pass that just ran, and the sampler constants (depth-prior table, split prior, column
//     min/max) are staged in shared memory: one batch of loads at entry, one batch of stores at exit.
//     The generic helpers of pg_serial.h run unchanged on a PgState copy whose pointers are redirected
//     to those shared arrays.
//   * per-block partials are reduced with lane = job, warp = block subset: every load of the
//     reduction is in flight at once (one L2 round trip), then a fixed-order combine.
//   * one lane per particle slot / per job; the order-sensitive loops of the reference
//       normalize_weights_inplace  smc.rs:257-268  (sequential sum)
//       systematic resampling      resampling.rs:24-44 (sequential cumulative walk)
//       WeightedIndex              smc.rs:113-114
//     keep the reference's summation order through warp-uniform loops over shuffle broadcasts.
#pragma once
#include <cstddef>
#include "pg_passes.cuh"

// Where the section's hot state lives while it runs (shared memory) — passed by value, stays in registers.
struct FwCx {
    PgFastSm* f;
    PgCtrl* c;              // working copy of the control block (f->lc)
    const double* thr;      // depth-prior thresholds
    const double* prior;    // split prior (null = uniform draw)
    const float* cmin; const float* cmax;
};

__device__ __forceinline__ double fw_sum_seq(double v, int n) {   // lanes 0..n-1 in order; uniform result
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += __shfl_sync(PG_FULL, v, k);
    return acc;
}
__device__ __forceinline__ double fw_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(PG_FULL, v, o));
    return v;
}
__device__ __forceinline__ int fw_imax(int v) { return __reduce_max_sync(PG_FULL, v); }

// Number of 8-byte words of PgCtrl that the serial section owns (everything before the barrier words).
__device__ __forceinline__ int fw_ctrl_words() { return (int)(offsetof(PgCtrl, bar_count) / 8); }

// ---- stage A: reduce the pass's per-block partials --------------------------------------------
// lane = job, warp = subset of blocks {warp, warp + PG_WARPS, ...}; rows of the [block][job] partial
// arrays are contiguous in job, so every load is coalesced and all of a warp's loads are in flight.
#define FW_RB 12   // blocks per warp per batch
__device__ void fw_reduce(const PgState& st, PgFastSm* f, int phase, int J) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = st.grid, S = st.S;
    const bool sums = phase == PH_ROOT || phase == PH_STATS || phase == PH_BERN;
    if (sums) {
        const double* src = phase == PH_BERN ? st.part_ll : st.part_sum;
        const bool cnts = phase != PH_BERN;
        double a = 0.0, bsum = 0.0;
        unsigned cc = 0u;
        if (lane < J) {
            for (int b0 = warp; b0 < G; b0 += PG_WARPS * FW_RB) {
                double2 v[FW_RB];
                unsigned k[FW_RB];
#pragma unroll
                for (int q = 0; q < FW_RB; ++q) {
                    const int b = b0 + q * PG_WARPS;
                    if (b < G) {
                        v[q] = __ldcg(reinterpret_cast<const double2*>(src + ((size_t)b * S + lane) * 2));
                        k[q] = cnts ? __ldcg(&st.part_ccnt[(size_t)b * S + lane]) : 0u;
                    } else { v[q] = make_double2(0.0, 0.0); k[q] = 0u; }
                }
#pragma unroll
                for (int q = 0; q < FW_RB; ++q) { a += v[q].x; bsum += v[q].y; cc += k[q]; }
            }
        }
        f->wsA[warp][lane] = a; f->wsB[warp][lane] = bsum; f->wsC[warp][lane] = cc;
    } else if (phase == PH_MINMAX) {
        unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
        if (lane < J)
            for (int b = warp; b < G; b += PG_WARPS) {
                const uint2 v = __ldcg(reinterpret_cast<const uint2*>(st.part_mm + ((size_t)b * S + lane) * 2));
                kmin = min(kmin, v.x); kmax = max(kmax, v.y);
            }
        f->wsC[warp][lane] = kmin;
        f->wsD[warp][lane] = kmax;
    }
    if (phase == PH_ROOT) {   // root totals: thread = block
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        for (int b = threadIdx.x; b < G; b += blockDim.x) {
            const double2 u = __ldcg(reinterpret_cast<const double2*>(st.part_tot + (size_t)b * 4));
            const double2 v = __ldcg(reinterpret_cast<const double2*>(st.part_tot + (size_t)b * 4 + 2));
            t0 += u.x; t1 += u.y; t2 += v.x; t3 += v.y;
        }
        for (int o = 16; o > 0; o >>= 1) {
            t0 += __shfl_xor_sync(PG_FULL, t0, o); t1 += __shfl_xor_sync(PG_FULL, t1, o);
            t2 += __shfl_xor_sync(PG_FULL, t2, o); t3 += __shfl_xor_sync(PG_FULL, t3, o);
        }
        if (lane == 0) { f->wsT[warp][0] = t0; f->wsT[warp][1] = t1; f->wsT[warp][2] = t2; f->wsT[warp][3] = t3; }
    }
    __syncthreads();
    if (warp == 0) {   // fixed-order combine over warps
        if (sums && lane < J) {
            double a = 0.0, b = 0.0;
            unsigned cc = 0u;
#pragma unroll
            for (int w = 0; w < PG_WARPS; ++w) { a += f->wsA[w][lane]; b += f->wsB[w][lane]; cc += f->wsC[w][lane]; }
            if (phase == PH_BERN) { f->rllL[lane] = a; f->rllR[lane] = b; }
            else { f->rSo[lane] = a; f->rSr[lane] = b; f->rCnt[lane] = cc; }
        }
        if (phase == PH_MINMAX && lane < J) {
            unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
            for (int w = 0; w < PG_WARPS; ++w) { kmin = min(kmin, f->wsC[w][lane]); kmax = max(kmax, f->wsD[w][lane]); }
            f->rMin[lane] = kmin; f->rMax[lane] = kmax;
        }
        if (phase == PH_ROOT && lane < 4) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < PG_WARPS; ++w) v += f->wsT[w][lane];
            f->tot[lane] = v;
        }
    }
}

// ---- per-lane job view ------------------------------------------------------------------------
struct FwJob {
    int slot, node, var;
    unsigned seg_off, len;
    float lo, hi, thr;
    double split;
    bool live;
};

// Compact live jobs (order preserved) into the GLOBAL job arrays (read by the next grid pass) and
// into the shared job cache; slot_job goes to the shared slot scalars.  Returns the new J.
__device__ __forceinline__ int fw_write_jobs(const PgState& st, const FwCx& x, const FwJob& jb, int lane, bool write_all) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    const unsigned live = __ballot_sync(PG_FULL, jb.live);
    const int J = __popc(live);
    const int pos = __popc(live & ((1u << lane) - 1u));
    if (lane < st.S) x.f->hsjob[lane] = -1;
    __syncwarp();
    if (jb.live) {
        st.job_slot[pos] = jb.slot; st.job_node[pos] = jb.node; st.job_var[pos] = jb.var;
        st.job_seg_off[pos] = jb.seg_off; st.job_len[pos] = jb.len;
        x.f->jslot[pos] = jb.slot; x.f->jnode[pos] = jb.node; x.f->jvar[pos] = jb.var;
        x.f->jseg[pos] = jb.seg_off; x.f->jlen[pos] = jb.len;
        if (write_all) {
            st.job_thr32[pos] = jb.thr; st.job_split[pos] = jb.split; st.job_lo[pos] = jb.lo; st.job_hi[pos] = jb.hi;
            x.f->jthr[pos] = jb.thr; x.f->jsplit[pos] = jb.split;
        }
        st.job_order[pos] = pos;
        x.f->hsjob[jb.slot] = pos;
    }
    const int first = __ffs(live) - 1;
    const unsigned off0 = __shfl_sync(PG_FULL, jb.seg_off, first < 0 ? 0 : first);
    const unsigned len0 = __shfl_sync(PG_FULL, jb.len, first < 0 ? 0 : first);
    const bool same = !jb.live || (jb.seg_off == off0 && jb.len == len0);
    const bool one = __all_sync(PG_FULL, same);
    if (lane == 0) c->n_jobs = J;
    if (one) {
        if (lane == 0) { st.grp_first[0] = 0; st.grp_first[J > 0 ? 1 : 0] = J; c->n_groups = J > 0 ? 1 : 0; }
    } else {
        __threadfence_block();
        __syncwarp();
        if (lane == 0) {          // rare (several distinct survivors): generic grouping on the global arrays
            pg_group_jobs_single(st, c);
        }
    }
    __syncwarp();
    return J;
}

__device__ __forceinline__ int fw_unique(int var, bool live, int lane) {
    const unsigned lv = __ballot_sync(PG_FULL, live);
    const unsigned m = __match_any_sync(PG_FULL, live ? var : -1 - lane);
    const bool leader = live && ((__ffs(m & lv) - 1) == lane);
    return __popc(__ballot_sync(PG_FULL, leader));
}

// smc.rs:242-254 with the prior in shared memory; 8 subtractions per trip, same order as the reference
__device__ __forceinline__ int fw_feature_from_probs(const PgState& st, const FwCx& x, double u) {
    double target = u * st.prior_total;
    const int p = st.p;
    int j = 0;
    for (; j + 8 <= p; j += 8) {
        double q[8];
#pragma unroll
        for (int z = 0; z < 8; ++z) q[z] = x.prior[j + z];
#pragma unroll
        for (int z = 0; z < 8; ++z) { target -= q[z]; if (target <= 0.0) return j + z; }
    }
    for (; j < p; ++j) { target -= x.prior[j]; if (target <= 0.0) return j; }
    return p - 1;
}

// splitting.rs:39-56 for one job lane (mirrors pg_decide_split)
__device__ __forceinline__ void fw_decide(const PgState& st, const FwCx& x, FwJob& jb) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    if (!jb.live) return;
    double* tr = pg_trace_slot(st, x.c, c->round, jb.slot);
    if (tr) { tr[9] = (double)jb.lo; tr[10] = (double)jb.hi; }
    if (jb.lo >= jb.hi) { if (tr) tr[0] = 3; jb.live = false; return; }
    if (2 * jb.node + 2 >= st.maxn) { c->error = PG_ERR_DEPTH; jb.live = false; return; }
    jb.split = pg_split_draw(st, x.c, c->upd, c->round, jb.slot, (double)jb.lo, (double)jb.hi);
    jb.thr = pg_thr32(jb.split);
    if (tr) tr[3] = jb.split;
}

// ---- sub-steps (warp 0; every lane calls) -------------------------------------------------------

__device__ __forceinline__ int fw_begin_step(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    if (lane == 0) {
        const double frac = c->tune ? st.batch_tune : st.batch_post;
        long long b = (long long)floor(frac * (double)st.m + 0.5);
        if (b < 1) b = 1;
        if (b > st.m) b = st.m;
        c->batch = (int)b; c->k = 0; c->t_add = -1; c->t_add_internal = 0;
        c->info_loglik = 0.0; c->info_acc = 0; c->info_n = 0;
    }
    __syncwarp();
    return SN_BEGIN_UPDATE;
}

__device__ __forceinline__ int fw_begin_update(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    const int cur = c->cur;
    if (lane == 0) {
        const int t = (c->next_tree_idx + c->k) % st.m;
        c->t_cur = t; c->t_sub = t; c->round = 0; c->acc_update = 0; c->pool_top = 0;
    }
    if (lane < st.S) {
        const int s = lane;
        const size_t i0 = pg_tix(st, cur, s, 0);
        st.pt_split_var[i0] = -1; st.pt_split_val[i0] = pg_nan(); st.pt_thr32[i0] = 0.0f; st.pt_leaf_val[i0] = st.init_leaf;
        st.pt_seg_off[i0] = PG_SEG_IDENTITY; st.pt_cnt[i0] = (unsigned)st.n;
        st.pt_So[i0] = 0.0; st.pt_Sr[i0] = 0.0; st.pt_g[i0] = 0.0; st.pt_queue[i0] = 0;
        const int ss = pg_six(st, cur, s);
        x.f->hsize[ss] = 1; x.f->hqh[ss] = 0; x.f->hqt[ss] = 1; x.f->hG[ss] = 0.0;
        x.f->hw[s] = 0.0;
    }
    __syncwarp();
    return SN_DRAW_ROUND;
}

// fresh = the slot tables were just initialised by fw_begin_update in this same serial section
__device__ __forceinline__ int fw_draw_round(const PgState& st, const FwCx& x, int lane, bool fresh) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    const int cur = c->cur, r = c->round, S = st.S;
    const unsigned long long upd = c->upd;
    FwJob jb;
    jb.live = false; jb.slot = lane; jb.node = -1; jb.var = 0; jb.seg_off = 0; jb.len = 0; jb.lo = jb.hi = jb.thr = 0.f; jb.split = 0.0;
    if (lane < S) {
        const int s = lane;
        x.f->hmut[s] = 0;
        double* tr = pg_trace_slot(st, x.c, r, s);
        if (tr) { for (int q = 0; q < PG_TS_FIELDS; ++q) tr[q] = pg_nan(); tr[0] = 0; }
        const int ss = pg_six(st, cur, s);
        const int qh = x.f->hqh[ss], qt = x.f->hqt[ss];
        if (qh < qt) {
            const int node = fresh ? 0 : st.pt_queue[pg_tix(st, cur, s, qh)];
            x.f->hqh[ss] = qh + 1;
            if (tr) { tr[0] = 1; tr[1] = (double)node; tr[2] = -1; }
            const double u1 = pg_uniform(st, x.c, PG_D_GROW, upd, r, s);
            if (!(x.thr[pg_depth_of(node)] > u1)) {
                const size_t iN = pg_tix(st, cur, s, node);
                const unsigned cnt = fresh ? (unsigned)st.n : st.pt_cnt[iN];
                if (cnt == 0u) { if (tr) tr[0] = 2; }
                else {
                    const double u2 = pg_uniform(st, x.c, PG_D_VAR, upd, r, s);
                    int var;
                    if (x.prior) var = fw_feature_from_probs(st, x, u2);
                    else { var = (int)(u2 * (double)st.p); if (var > st.p - 1) var = st.p - 1; }
                    if (tr) tr[2] = (double)var;
                    jb.live = true; jb.node = node; jb.var = var; jb.len = cnt;
                    jb.seg_off = fresh ? PG_SEG_IDENTITY : st.pt_seg_off[iN];
                }
            }
        }
    }
    if (r == 0) {
        if (jb.live) { jb.lo = x.cmin[jb.var]; jb.hi = x.cmax[jb.var]; }
        const int J0 = __popc(__ballot_sync(PG_FULL, jb.live));
        fw_decide(st, x, jb);
        const int J = fw_write_jobs(st, x, jb, lane, true);
        const int uniq = fw_unique(jb.var, jb.live, lane);
        if (lane == 0) {
            const unsigned long long n = (unsigned long long)st.n;
            const int cols = uniq + c->t_add_internal + (c->t_sub >= 0 ? st.f_internal[c->t_sub] : 0);
            c->bytes_model += n * (8 + 8 + 4) + 4ull * n * (unsigned long long)cols;
            c->bytes_contract += 20ull * n + 4ull * n * (unsigned long long)J0 + 16ull * n * (unsigned long long)J;
            c->phase = PH_ROOT; c->n_phases += 1;
        }
        __syncwarp();
        return SN_NONE;
    }
    const int J = fw_write_jobs(st, x, jb, lane, false);
    if (J == 0) return SN_FINISH_ROUND;
    if (lane == 0) {
        unsigned long long bm = 0, bc = 0;
        for (int g = 0; g < c->n_groups; ++g) {
            const unsigned long long L = st.job_len[st.job_order[st.grp_first[g]]];
            bm += 4ull * L + 4ull * L * (unsigned long long)(st.grp_first[g + 1] - st.grp_first[g]);
        }
        for (int j = 0; j < J; ++j) bc += 8ull * x.f->jlen[j];
        c->bytes_model += bm; c->bytes_contract += bc;
        c->phase = PH_MINMAX; c->n_phases += 1;
    }
    __syncwarp();
    return SN_NONE;
}

__device__ __forceinline__ FwJob fw_load_job(const FwCx& x, int lane, int J) {
    FwJob jb;
    jb.live = lane < J;
    jb.slot = jb.live ? x.f->jslot[lane] : 0;
    jb.node = jb.live ? x.f->jnode[lane] : 0;
    jb.var = jb.live ? x.f->jvar[lane] : 0;
    jb.seg_off = jb.live ? x.f->jseg[lane] : 0u;
    jb.len = jb.live ? x.f->jlen[lane] : 0u;
    jb.thr = jb.live ? x.f->jthr[lane] : 0.f;
    jb.split = jb.live ? x.f->jsplit[lane] : 0.0;
    jb.lo = jb.hi = 0.f;
    return jb;
}

__device__ __forceinline__ int fw_after_minmax(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    const int J0 = c->n_jobs;
    FwJob jb = fw_load_job(x, lane, J0);
    if (jb.live) { jb.lo = pg_key2f(f->rMin[lane]); jb.hi = pg_key2f(f->rMax[lane]); }
    fw_decide(st, x, jb);
    const int J = fw_write_jobs(st, x, jb, lane, true);
    if (J == 0) return SN_FINISH_ROUND;
    if (lane == 0) {
        unsigned long long bm = 0, bc = 0;
        for (int g = 0; g < c->n_groups; ++g) {
            const unsigned long long L = st.job_len[st.job_order[st.grp_first[g]]];
            bm += (4ull + 8ull + 4ull) * L + 4ull * L * (unsigned long long)(st.grp_first[g + 1] - st.grp_first[g]);
        }
        for (int j = 0; j < J; ++j) bc += 20ull * x.f->jlen[j];
        c->bytes_model += bm; c->bytes_contract += bc;
        c->phase = PH_STATS; c->n_phases += 1;
    }
    __syncwarp();
    return SN_NONE;
}

// particle.rs:91-148 bookkeeping half + smc.rs:76-81 (mirrors pg_apply_mutation on the shared slot scalars)
__device__ __forceinline__ void fw_apply(const PgState& st, const FwCx& x, int j, double vL, double vR, double gL, double gR,
                                         unsigned cL, unsigned cR, double SoL, double SoR, double SrL, double SrR) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int cur = c->cur;
    const int s = f->jslot[j];
    const int node = f->jnode[j];
    const int L = 2 * node + 1, R = 2 * node + 2;
    const size_t iN = pg_tix(st, cur, s, node), iL = pg_tix(st, cur, s, L), iR = pg_tix(st, cur, s, R);
    const int ss = pg_six(st, cur, s);
    const int old_size = f->hsize[ss];
    const double g_node = node == 0 ? c->g_stump : st.pt_g[iN];   // the root's leaf term is the stump's
    for (int q = old_size; q < R + 1; ++q) {          // tree.rs:105-110 filler nodes
        const size_t iq = pg_tix(st, cur, s, q);
        st.pt_split_var[iq] = -1; st.pt_split_val[iq] = pg_nan(); st.pt_thr32[iq] = 0.0f; st.pt_leaf_val[iq] = 0.0;
        st.pt_seg_off[iq] = PG_SEG_NONE; st.pt_cnt[iq] = 0u; st.pt_So[iq] = 0.0; st.pt_Sr[iq] = 0.0; st.pt_g[iq] = 0.0;
    }
    st.pt_split_var[iN] = f->jvar[j];
    st.pt_split_val[iN] = f->jsplit[j];
    st.pt_thr32[iN] = f->jthr[j];
    st.pt_leaf_val[iN] = pg_nan();
    st.pt_split_var[iL] = -1; st.pt_split_val[iL] = pg_nan(); st.pt_thr32[iL] = 0.0f; st.pt_leaf_val[iL] = vL;
    st.pt_split_var[iR] = -1; st.pt_split_val[iR] = pg_nan(); st.pt_thr32[iR] = 0.0f; st.pt_leaf_val[iR] = vR;
    st.pt_seg_off[iL] = PG_SEG_NONE; st.pt_seg_off[iR] = PG_SEG_NONE;
    st.pt_cnt[iL] = cL; st.pt_cnt[iR] = cR;
    st.pt_So[iL] = SoL; st.pt_So[iR] = SoR; st.pt_Sr[iL] = SrL; st.pt_Sr[iR] = SrR;
    st.pt_g[iL] = gL; st.pt_g[iR] = gR;
    if (R + 1 > old_size) f->hsize[ss] = R + 1;
    const double G = (f->hG[ss] - g_node) + (gL + gR);
    f->hG[ss] = G;
    const int qt = f->hqt[ss];
    st.pt_queue[pg_tix(st, cur, s, qt)] = L;           // particle.rs:146-147
    st.pt_queue[pg_tix(st, cur, s, qt + 1)] = R;
    f->hqt[ss] = qt + 2;
    f->hmut[s] = 1;
    const double w = c->C0 + G;                        // smc.rs:89-95
    f->hw[s] = w;
    double* tr = pg_trace_slot(st, c, c->round, s);
    if (tr) { tr[0] = 4; tr[4] = vL; tr[5] = vR; tr[6] = (double)cL; tr[7] = (double)cR; tr[8] = w; tr[11] = SoL; }
}

// After PH_ROOT / PH_STATS (and, for Bernoulli, completes after PH_BERN).
__device__ __forceinline__ int fw_after_stats(const PgState& st, const FwCx& x, int lane, int phase) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    const int J = c->n_jobs, cur = c->cur, S = st.S;
    const bool bern = st.family == PG_BERNOULLI_LOGIT;
    const bool root = phase == PH_ROOT;
    if (root) {
        const double T_o = f->tot[0], T_r = f->tot[1], T_rr = f->tot[2], T_ll0 = f->tot[3];
        const double nn = (double)st.n;
        double C0, g_stump;
        if (st.family == PG_GAUSSIAN) {
            const double s2 = st.lik_sigma * st.lik_sigma;
            C0 = -nn * (log(st.lik_sigma) + 0.5 * log(6.283185307179586476925286766559)) - 0.5 * T_rr / s2;
            g_stump = pg_gauss_g(st.init_leaf, T_r, (unsigned)st.n, 1.0 / s2);
        } else if (st.family == PG_GAUSSIAN_KERNEL) {
            C0 = -0.5 * T_rr;
            g_stump = pg_gauss_g(st.init_leaf, T_r, (unsigned)st.n, 1.0);
        } else { C0 = 0.0; g_stump = T_ll0; }
        if (lane == 0) { c->T_o = T_o; c->T_r = T_r; c->T_rr = T_rr; c->T_ll0 = T_ll0; c->C0 = C0; c->g_stump = g_stump; }
        if (lane < S) {
            const size_t i0 = pg_tix(st, cur, lane, 0);
            st.pt_So[i0] = T_o; st.pt_Sr[i0] = T_r; st.pt_g[i0] = g_stump;
            x.f->hG[pg_six(st, cur, lane)] = g_stump;
        }
        __syncwarp();
    }
    if (phase == PH_BERN) {
        if (lane < J) {
            const int s = x.f->jslot[lane];
            const size_t iN = pg_tix(st, cur, s, x.f->jnode[lane]);
            const unsigned nc = st.pt_cnt[iN];
            const unsigned cL = (unsigned)st.red_cnt[lane], cR = nc - cL;
            const double SoL = st.red_So[lane], SoR = st.pt_So[iN] - SoL;
            const double SrL = st.red_Sr[lane], SrR = st.pt_Sr[iN] - SrL;
            fw_apply(st, x, lane, st.job_vL[lane], st.job_vR[lane], f->rllL[lane], f->rllR[lane], cL, cR, SoL, SoR, SrL, SrR);
        }
        __syncwarp();
        return SN_FINISH_ROUND;
    }
    const double inv_s2 = st.family == PG_GAUSSIAN ? 1.0 / (st.lik_sigma * st.lik_sigma) : 1.0;
    if (lane < J) {
        const int s = x.f->jslot[lane];
        const int node = x.f->jnode[lane];
        const size_t iN = pg_tix(st, cur, s, node);
        // node totals: the root's are the pass totals just reduced (no table read needed)
        const unsigned nc = root ? (unsigned)st.n : st.pt_cnt[iN];
        const double nSo = root ? f->tot[0] : st.pt_So[iN];
        const double nSr = root ? f->tot[1] : st.pt_Sr[iN];
        const unsigned cL = f->rCnt[lane], cR = nc - cL;
        const double SoL = f->rSo[lane], SoR = nSo - SoL;
        const double zL = pg_normal(st, x.c, PG_D_ZLEFT, c->upd, c->round, s);
        const double zR = pg_normal(st, x.c, PG_D_ZRIGHT, c->upd, c->round, s);
        const double vL = cL == 0u ? zL * st.leaf_sigma : pg_mul_add(zL, st.leaf_sigma, SoL / (double)cL / st.m_f);
        const double vR = cR == 0u ? zR * st.leaf_sigma : pg_mul_add(zR, st.leaf_sigma, SoR / (double)cR / st.m_f);
        const double SrL = f->rSr[lane], SrR = nSr - SrL;
        st.red_cnt[lane] = cL;   // the partition (and the Bernoulli completion) may read these in a later section
        if (bern) {
            st.job_vL[lane] = vL; st.job_vR[lane] = vR; st.red_So[lane] = SoL; st.red_Sr[lane] = SrL;
        } else {
            fw_apply(st, x, lane, vL, vR, pg_gauss_g(vL, SrL, cL, inv_s2), pg_gauss_g(vR, SrR, cR, inv_s2),
                              cL, cR, SoL, SoR, SrL, SrR);
        }
    }
    __syncwarp();
    if (bern && J > 0) {
        if (lane == 0) {
            unsigned long long bm = 0, bc = 0;
            for (int g = 0; g < c->n_groups; ++g) {
                const int j0 = st.job_order[st.grp_first[g]];
                const unsigned long long L = x.f->jlen[j0];
                const unsigned long long fixed = x.f->jseg[j0] == PG_SEG_IDENTITY ? 12ull : 16ull;
                bm += fixed * L + 4ull * L * (unsigned long long)(st.grp_first[g + 1] - st.grp_first[g]);
            }
            for (int j = 0; j < J; ++j) bc += 12ull * x.f->jlen[j];
            c->bytes_model += bm; c->bytes_contract += bc;
            c->phase = PH_BERN; c->n_phases += 1;
        }
        __syncwarp();
        return SN_NONE;
    }
    return SN_FINISH_ROUND;
}

// smc.rs:96-102 + bookkeeping of the lazy partition
__device__ __forceinline__ int fw_finish_round(const PgState& st, const FwCx& x, int lane, bool cnt_in_smem) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    const int S = st.S;
    const int src = c->cur, dst = 1 - c->cur;
    const bool act = lane < S;
    const int mut = act ? x.f->hmut[lane] : 0;
    double w = act ? x.f->hw[lane] : -INFINITY;
    const int n_mut = __popc(__ballot_sync(PG_FULL, mut != 0));
    // smc.rs:257-268
    const double mx = fw_max(w);
    double e = act ? exp(w - mx) : 0.0;
    const double sum = fw_sum_seq(e, S);
    const double wn = e / sum;
    if (act) x.f->hw[lane] = wn;
    // resampling.rs:24-44
    const double u6 = pg_uniform(st, x.c, PG_D_RESAMPLE, c->upd, c->round, 0);
    const double target = ((double)lane + u6) / (double)S;
    int cidx = 0;
    double cum = 0.0;
    for (int k = 0; k < S; ++k) {          // cum == sum of the first k weights, in order, on every lane
        if (cum < target) cidx = k + 1;     // the reference's pointer passes k
        cum += __shfl_sync(PG_FULL, wn, k);
    }
    const int anc = act ? (cidx == 0 ? 0 : cidx - 1) : 0;
    if (act) x.f->hanc[lane] = anc;
    if (st.tr_round && lane == 0) {
        if (!((long long)c->trace_n < st.tr_u_cap && c->round < st.tr_r_cap)) c->trace_overflow = 1;
    }
    if (st.tr_round && (long long)c->trace_n < st.tr_u_cap && c->round < st.tr_r_cap) {
        double* tr = st.tr_round + ((size_t)c->trace_n * st.tr_r_cap + c->round) * (size_t)(2 + 2 * S);
        if (lane == 0) { tr[0] = 1; tr[1] = (double)n_mut; }
        if (act) { tr[2 + lane] = wn; tr[2 + S + lane] = (double)anc; }
    }
    // clone survivors: slot `lane` <- slot `anc`
    const int fs = pg_six(st, src, anc), ts = pg_six(st, dst, lane);
    const int sz = act ? x.f->hsize[fs] : 0;
    const int qh = act ? x.f->hqh[fs] : 0, qt = act ? x.f->hqt[fs] : 0;
    const double Gs = act ? x.f->hG[fs] : 0.0;
    const int mxs = fw_imax(sz);
    for (int node = 0; node < mxs; ++node) {
        if (node < sz) {
            const size_t a = pg_tix(st, src, anc, node), t = pg_tix(st, dst, lane, node);
            st.pt_split_var[t] = st.pt_split_var[a]; st.pt_thr32[t] = st.pt_thr32[a];
            st.pt_split_val[t] = st.pt_split_val[a]; st.pt_leaf_val[t] = st.pt_leaf_val[a];
            st.pt_seg_off[t] = st.pt_seg_off[a]; st.pt_cnt[t] = st.pt_cnt[a];
            st.pt_So[t] = st.pt_So[a]; st.pt_Sr[t] = st.pt_Sr[a]; st.pt_g[t] = st.pt_g[a];
            st.pt_queue[t] = st.pt_queue[a];
        }
    }
    __syncwarp();
    if (act) { x.f->hsize[ts] = sz; x.f->hqh[ts] = qh; x.f->hqt[ts] = qt; x.f->hG[ts] = Gs; st.pj_of_anc[lane] = -1; }
    // one stable partition per distinct surviving ancestor that grew this round
    const int mut_a = __shfl_sync(PG_FULL, mut, anc);
    const bool need = act && mut_a != 0;
    const unsigned needm = __ballot_sync(PG_FULL, need);
    const unsigned same = __match_any_sync(PG_FULL, need ? anc : -1 - lane);
    const bool leader = need && ((__ffs(same & needm) - 1) == lane);
    const unsigned leadm = __ballot_sync(PG_FULL, leader);
    const int K = __popc(leadm);
    const int k = __popc(leadm & ((1u << lane) - 1u));
    int j = 0, node = 0;
    unsigned len = 0u, cntL = 0u;
    if (need) {
        j = x.f->hsjob[anc];
        node = x.f->jnode[j];
        len = x.f->jlen[j];
        cntL = cnt_in_smem ? f->rCnt[j] : (unsigned)st.red_cnt[j];
    }
    unsigned long long off = 0;
    {
        const unsigned long long mylen = leader ? (unsigned long long)len : 0ull;
        unsigned long long run = c->pool_top;
        for (int q = 0; q < 32; ++q) {
            if (!((leadm >> q) & 1u)) continue;       // leadm is warp-uniform
            const unsigned long long v = __shfl_sync(PG_FULL, mylen, q);
            if (q == lane) off = run;
            run += v;
        }
        if (lane == 0) {
            if (run > st.pool_cap) c->error = PG_ERR_POOL;
            c->pool_top = run;
        }
    }
    if (leader) {
        st.pj_of_anc[anc] = k;
        st.pj_anc[k] = anc; st.pj_job[k] = j; st.pj_var[k] = x.f->jvar[j]; st.pj_thr32[k] = x.f->jthr[j];
        st.pj_src_off[k] = x.f->jseg[j]; st.pj_len[k] = len; st.pj_cntL[k] = cntL; st.pj_dst_off[k] = (unsigned)off;
    }
    const int lead_lane = need ? (__ffs(same & needm) - 1) : 0;
    const unsigned long long off_l = __shfl_sync(PG_FULL, off, lead_lane);
    if (need) {
        st.pt_seg_off[pg_tix(st, dst, lane, 2 * node + 1)] = (unsigned)off_l;
        st.pt_seg_off[pg_tix(st, dst, lane, 2 * node + 2)] = (unsigned)off_l + cntL;
    }
    const int any_queue = __any_sync(PG_FULL, act && qh < qt);
    if (lane == 0) {
        c->acc_update += n_mut;
        c->cur = dst;
        c->n_pjobs = K;
        c->any_queue = any_queue;
    }
    __syncwarp();
    if (c->error != PG_OK) return SN_NONE;
    if (K > 0) return SN_PREP_PART;
    if (any_queue) { if (lane == 0) c->round += 1; __syncwarp(); return SN_DRAW_ROUND; }
    return SN_FINALIZE;
}

// block-wide: exclusive prefix, over global warps, of each partition job's per-warp left counts
#define FB_CH 8
__device__ __forceinline__ void fb_prep_part(const PgState& st, const FwCx& x) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    const int K = c->n_pjobs, W = st.n_gwarps, S = st.S;
    const int tid = threadIdx.x, T = blockDim.x;
    const int chunk = (W + T - 1) / T;
    for (int k = 0; k < K; ++k) {
        const int j = st.pj_job[k];
        const int b = tid * chunk;
        unsigned local = 0;
        if (chunk <= FB_CH) {
            unsigned v[FB_CH];
#pragma unroll
            for (int q = 0; q < FB_CH; ++q) v[q] = (q < chunk && b + q < W) ? __ldcg(&st.part_cnt[(size_t)(b + q) * S + j]) : 0u;
#pragma unroll
            for (int q = 0; q < FB_CH; ++q) local += v[q];
            f->scan[tid] = local;
            __syncthreads();
            if (tid < 32) {
                unsigned carry = 0;
                for (int base = 0; base < T; base += 32) {
                    const unsigned x0 = f->scan[base + tid];
                    unsigned x = x0;
                    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(PG_FULL, x, o); if (tid >= o) x += y; }
                    f->scan[base + tid] = carry + x - x0;
                    carry += __shfl_sync(PG_FULL, x, 31);
                }
            }
            __syncthreads();
            unsigned run = f->scan[tid];
#pragma unroll
            for (int q = 0; q < FB_CH; ++q)
                if (q < chunk && b + q < W) { st.pj_warp_left[(size_t)k * W + b + q] = run; run += v[q]; }
        } else {
            const int e = min(b + chunk, W);
            for (int g = b; g < e; ++g) local += __ldcg(&st.part_cnt[(size_t)g * S + j]);
            f->scan[tid] = local;
            __syncthreads();
            if (tid == 0) { unsigned run = 0; for (int q = 0; q < T; ++q) { const unsigned v = f->scan[q]; f->scan[q] = run; run += v; } }
            __syncthreads();
            unsigned run = f->scan[tid];
            for (int g = b; g < e; ++g) { st.pj_warp_left[(size_t)k * W + g] = run; run += __ldcg(&st.part_cnt[(size_t)g * S + j]); }
        }
        __syncthreads();
    }
    if (tid == 0) {
        unsigned long long bm = 0;
        for (int k = 0; k < K; ++k) {
            const unsigned long long L = st.pj_len[k];
            bm += (st.pj_src_off[k] == PG_SEG_IDENTITY ? 0ull : 4ull * L) + 4ull * L + 4ull * L;
        }
        c->bytes_model += bm;
        c->phase = PH_PART; c->n_phases += 1;
    }
}

// smc.rs:105-122 + kernel.rs:99-117
__device__ __forceinline__ int fw_finalize(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    (void)f;
    const int P = st.P, cur = c->cur, t = c->t_cur;
    const bool act = lane < P;
    double Wl = -INFINITY;
    if (act) Wl = c->C0 + (lane == 0 ? c->g_stump : x.f->hG[pg_six(st, cur, lane == 0 ? 0 : lane - 1)]);
    const double mx = fw_max(Wl);
    const double e = act ? exp(Wl - mx) : 0.0;
    const double sum = fw_sum_seq(e, P);
    const double Wn = e / sum;
    const double total = fw_sum_seq(act ? Wn : 0.0, P);
    const double u7 = pg_uniform(st, x.c, PG_D_SELECT, c->upd, 0, 0);
    const double chosen = u7 * total;
    int sel = 0;
    {
        double cum = 0.0;
        bool open = true;
        for (int i = 0; i + 1 < P; ++i) {
            cum += __shfl_sync(PG_FULL, Wn, i);
            if (open && cum <= chosen) sel = i + 1; else open = false;
        }
    }
    if (lane == 0) {
        if (!(total > 0.0) || !(total < INFINITY)) c->error = PG_ERR_WEIGHTS;
        c->info_loglik += Wn;      // lane 0 holds the reference stump's normalised weight
        c->info_acc += c->acc_update;
    }
    if (st.tr_upd) {
        if ((long long)c->trace_n < st.tr_u_cap) {
            double* tr = st.tr_upd + (size_t)c->trace_n * (size_t)(4 + 2 * P);
            if (lane == 0) { tr[0] = (double)(c->round + 1); tr[1] = (double)sel; tr[2] = (double)t; tr[3] = (double)c->acc_update; }
            if (act) { tr[4 + lane] = Wl; tr[4 + P + lane] = Wn; }
        } else if (lane == 0) c->trace_overflow = 1;
    }
    const size_t fb = (size_t)t * st.maxn;
    int depth = 0, internal = 0, sz = 1;
    if (sel == 0) {
        if (lane == 0) {
            st.f_split_var[fb] = -1; st.f_split_val[fb] = pg_nan(); st.f_thr32[fb] = 0.0f; st.f_leaf_val[fb] = st.init_leaf;
            st.f_size[t] = 1;
        }
    } else {
        const int s = sel - 1;
        sz = x.f->hsize[pg_six(st, cur, s)];
        for (int node = lane; node < sz; node += 32) {
            const size_t a = pg_tix(st, cur, s, node);
            const int sv = st.pt_split_var[a];
            st.f_split_var[fb + node] = sv; st.f_split_val[fb + node] = st.pt_split_val[a];
            st.f_thr32[fb + node] = st.pt_thr32[a]; st.f_leaf_val[fb + node] = st.pt_leaf_val[a];
            if (sv < 0) { const int d = pg_depth_of(node); depth = d > depth ? d : depth; } else internal += 1;
        }
        depth = fw_imax(depth);
        internal = __reduce_add_sync(PG_FULL, internal);
        if (lane == 0) st.f_size[t] = sz;
    }
    int next = SN_NONE;
    if (lane == 0) {
        st.f_internal[t] = internal;
        c->t_add_internal = internal;
        if (c->info_n < st.m) st.info_depths[c->info_n] = (unsigned char)depth;
        c->info_n += 1;
        if (st.tr_upd && (long long)c->trace_n < st.tr_u_cap) c->trace_n += 1;
        c->upd += 1; c->n_updates += 1; c->k += 1; c->t_add = t;
        if (c->k < c->batch) next = SN_BEGIN_UPDATE;
        else {
            c->t_sub = -1;
            c->bytes_model += (unsigned long long)st.n * (16ull + 4ull * (unsigned long long)internal);
            c->phase = PH_FINAL; c->n_phases += 1;
            next = SN_NONE;
        }
    }
    __syncwarp();
    next = __shfl_sync(PG_FULL, next, 0);
    return next;
}

// ---- entry: the whole serial block (block 0) -------------------------------------------------------
// phase / J: the pass that just completed (every block knows them: no load needed).
__device__ void pg_serial_fast(const PgState& st, const PgSmemView& sm, int phase, int J) {
    PgCtrl* gc = st.ctrl;
    PgFastSm* f = &sm.base->fast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = st.S;
    const long long t_entry = clock64();

    // ---- prologue: ONE batch of loads. Warp 1: PgCtrl; warp 2: slot scalars; warp 3: job list; then every warp
    // takes its share of the partial reduction (fw_reduce) — all in the same L2 round trip.
    if (warp == 1) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gc);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&f->lc);
        const int nw = fw_ctrl_words();
        for (int i = lane; i < nw; i += 32) dst[i] = __ldcg(&src[i]);
    } else if (warp == 2) {
        if (lane < S) {   // both halves of the double-buffered scalars: no dependent load on ctrl->cur
            const int s0 = lane, s1 = S + lane;
            const double w = __ldcg(&st.w[lane]); const int mu = __ldcg(&st.mutated[lane]);
            const int an = __ldcg(&st.anc[lane]); const int sj = __ldcg(&st.slot_job[lane]);
            const int sz0 = __ldcg(&st.pt_size[s0]), qh0 = __ldcg(&st.pt_qhead[s0]), qt0 = __ldcg(&st.pt_qtail[s0]);
            const int sz1 = __ldcg(&st.pt_size[s1]), qh1 = __ldcg(&st.pt_qhead[s1]), qt1 = __ldcg(&st.pt_qtail[s1]);
            const double G0 = __ldcg(&st.pt_G[s0]), G1 = __ldcg(&st.pt_G[s1]);
            f->hw[lane] = w; f->hmut[lane] = mu; f->hanc[lane] = an; f->hsjob[lane] = sj;
            f->hsize[s0] = sz0; f->hqh[s0] = qh0; f->hqt[s0] = qt0; f->hG[s0] = G0;
            f->hsize[s1] = sz1; f->hqh[s1] = qh1; f->hqt[s1] = qt1; f->hG[s1] = G1;
        }
    } else if (warp == 3) {
        if (lane < J) {
            const int a = __ldcg(&st.job_slot[lane]), b = __ldcg(&st.job_node[lane]), v = __ldcg(&st.job_var[lane]);
            const unsigned so = __ldcg(&st.job_seg_off[lane]), ln = __ldcg(&st.job_len[lane]);
            const float th = __ldcg(&st.job_thr32[lane]); const double sp = __ldcg(&st.job_split[lane]);
            f->jslot[lane] = a; f->jnode[lane] = b; f->jvar[lane] = v; f->jseg[lane] = so; f->jlen[lane] = ln;
            f->jthr[lane] = th; f->jsplit[lane] = sp;
        }
    }
    fw_reduce(st, f, phase, J);   // contains a __syncthreads before its combine step
    __syncthreads();

    FwCx x;
    x.f = f;
    x.c = &f->lc;
    const bool staged = sm.base->cstaged != 0;
    x.thr = staged ? sm.base->cthr : st.thr_depth;
    x.prior = st.prior ? (staged ? sm.cprior : st.prior) : nullptr;
    x.cmin = staged ? sm.cmin : st.col_min;
    x.cmax = staged ? sm.cmax : st.col_max;
    PgCtrl* c = x.c;

    long long tq = clock64();
#define FW_TICK(slot) do { if (threadIdx.x == 0) { const long long now_ = clock64(); c->prof2[slot] += (unsigned long long)(now_ - tq); c->prof2[8 + slot] += 1ull; tq = now_; } } while (0)
    if (threadIdx.x == 0) { c->prof2[0] += (unsigned long long)(tq - t_entry); c->prof2[8] += 1ull; }

    bool fresh = false;
    bool cnt_in_smem = (phase == PH_ROOT || phase == PH_STATS);   // f->rCnt holds this round's left counts
    if (warp == 0) {
        int sn = SN_NONE;
        switch (phase) {
            case PH_BEGIN: sn = fw_begin_step(st, x, lane); break;
            case PH_ROOT: case PH_STATS: case PH_BERN: sn = fw_after_stats(st, x, lane, phase); FW_TICK(1); break;
            case PH_MINMAX: sn = fw_after_minmax(st, x, lane); FW_TICK(2); break;
            case PH_PART:
                if (lane == 0) c->round += 1;
                __syncwarp();
                sn = SN_DRAW_ROUND;
                break;
            case PH_FINAL:
                if (lane == 0) { c->next_tree_idx = (c->next_tree_idx + c->batch) % st.m; c->t_add = -1; c->phase = PH_DONE; }
                break;
            default: break;
        }
        if (lane == 0) f->sn = sn;
    }
    for (int guard = 0; guard < (1 << 20); ++guard) {
        __syncthreads();
        const int sn = f->sn;
        __syncthreads();
        if (sn == SN_NONE) break;
        if (sn == SN_PREP_PART) {
            tq = clock64();
            fb_prep_part(st, x);
            FW_TICK(4);
            if (threadIdx.x == 0) f->sn = SN_NONE;
            continue;
        }
        if (warp == 0) {
            int nx = SN_NONE;
            int cur_sn = sn;
            tq = clock64();
            for (;;) {   // warp-only sub-steps chain without block barriers
                if (c->error != PG_OK) { nx = SN_NONE; break; }
                if (cur_sn == SN_BEGIN_UPDATE) { nx = fw_begin_update(st, x, lane); fresh = true; FW_TICK(6); }
                else if (cur_sn == SN_DRAW_ROUND) { nx = fw_draw_round(st, x, lane, fresh); fresh = false; FW_TICK(7); }
                else if (cur_sn == SN_FINISH_ROUND) { nx = fw_finish_round(st, x, lane, cnt_in_smem); cnt_in_smem = false; FW_TICK(3); }
                else if (cur_sn == SN_FINALIZE) { nx = fw_finalize(st, x, lane); FW_TICK(5); }
                else { if (lane == 0) c->error = PG_ERR_INTERNAL; nx = SN_NONE; }
                if (nx == SN_NONE || nx == SN_PREP_PART) break;
                cur_sn = nx;
            }
            if (lane == 0) { if (c->error != PG_OK) c->phase = PH_DONE; f->sn = nx; }
        }
    }
    __syncthreads();
    // ---- epilogue: publish PgCtrl (warp 1) and the slot scalars (warp 2)
    if (threadIdx.x == 0) c->prof[16 + (phase & 7)] += (unsigned long long)(clock64() - t_entry);
    __syncthreads();
    if (warp == 1) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(gc);
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&f->lc);
        const int nw = fw_ctrl_words();
        for (int i = lane; i < nw; i += 32) dst[i] = src[i];
    } else if (warp == 2) {
        const int cur = f->lc.cur;
        if (lane < S) {
            st.w[lane] = f->hw[lane]; st.mutated[lane] = f->hmut[lane]; st.anc[lane] = f->hanc[lane];
            st.slot_job[lane] = f->hsjob[lane];
            const int ss = cur * S + lane;
            st.pt_size[ss] = f->hsize[ss]; st.pt_qhead[ss] = f->hqh[ss]; st.pt_qtail[ss] = f->hqt[ss]; st.pt_G[ss] = f->hG[ss];
        }
    }
    __syncthreads();
}
