// pg_sim.cpp — HOST SIMULATION of the device sweep.  TEST INFRASTRUCTURE ONLY.
//
// It compiles the product's serial control logic (bart_rs_b200/csrc/pg_serial.h, pg_rng.h,
// pg_types.h) for the host and replaces the CUDA grid passes of pg_passes.cuh by plain loops with
// the same partial-result layout (per "block", per "global warp").  Purpose: exercise the state
// machine, the stale-weight rule, resampling, slot-table copies, job grouping, partition offsets
// and the trace against the CPU oracle in this GPU-less container, before GPU time is spent.
// It is never linked into libpgbart_b200.so and is not a fallback of the product.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../bart_rs_b200/csrc/pg_serial.h"
#include "../../include/pgbart.h"

namespace {

struct Sim {
    PgState st{};
    std::vector<void*> allocs;
    std::string err;
    std::vector<double> tape_slot, tape_round, tape_upd;
    bool dead = false;
};

template <typename T>
T* halloc(Sim* s, size_t count) {
    void* p = std::calloc(count ? count : 1, sizeof(T));
    s->allocs.push_back(p);
    return (T*)p;
}

double unrolled_mean(const double* xs, size_t n) {
    if (n == 0) return 0.0;
    double p[8] = {0, 0, 0, 0, 0, 0, 0, 0}, acc = 0.0;
    size_t len = n;
    while (len >= 8) { for (int j = 0; j < 8; ++j) p[j] += xs[j]; xs += 8; len -= 8; }
    acc += p[0] + p[4]; acc += p[1] + p[5]; acc += p[2] + p[6]; acc += p[3] + p[7];
    for (size_t i = 0; i < len && i < 7; ++i) acc += xs[i];
    return acc / (double)n;
}

double tree_eval(const PgState& st, int t, long long row) {
    const size_t fb = (size_t)t * st.maxn;
    int node = 0, sv;
    while ((sv = st.f_split_var[fb + node]) >= 0) {
        const float x = st.X[(size_t)sv * st.ld + row];
        node = 2 * node + 1 + (x < st.f_thr32[fb + node] ? 0 : 1);
    }
    return st.f_leaf_val[fb + node];
}

void pass_root(const PgState& st) {
    PgCtrl* c = st.ctrl;
    const bool is_root = c->phase == PH_ROOT;
    const int J = is_root ? c->n_jobs : 0;
    const int t_add = c->t_add, t_sub = is_root ? c->t_sub : -1;
    const int S = st.S;
    for (int b = 0; b < st.grid; ++b) {
        double tot[4] = {0, 0, 0, 0}, tot_old = 0.0;
        std::vector<double> so(J, 0.0), sr(J, 0.0);
        std::vector<unsigned> cc(J, 0u);
        for (int w = 0; w < PG_WARPS; ++w) {
            const int gw = b * PG_WARPS + w;
            unsigned long long beg, end;
            pg_warp_range((unsigned long long)st.n, gw, st.n_gwarps, 4, &beg, &end);
            std::vector<unsigned> wc(J, 0u);
            for (unsigned long long row = beg; row < end; ++row) {
                double o = st.A[row];
                if (t_add >= 0) o = o + tree_eval(st, t_add, (long long)row);
                if (st.pg_correct && is_root) {     // log-likelihood of the current predictions (conditional reference particle)
                    if (st.family == PG_BERNOULLI_LOGIT) tot_old += (double)st.y[row] * o - pg_softplus(o);
                    else { const double d = (double)st.y[row] - o; tot_old += d * d; }
                }
                if (t_sub >= 0) o = o - tree_eval(st, t_sub, (long long)row);
                st.A[row] = o;
                if (!is_root) continue;
                const double r = (double)st.y[row] - o;
                tot[0] += o; tot[1] += r; tot[2] += r * r;
                if (st.family == PG_BERNOULLI_LOGIT) {
                    const double mu = o + st.init_leaf;
                    tot[3] += (double)st.y[row] * mu - pg_softplus(mu);
                }
                for (int j = 0; j < J; ++j)
                    if (st.X[(size_t)st.job_var[j] * st.ld + row] < st.job_thr32[j]) { so[j] += o; sr[j] += r; cc[j] += 1; wc[j] += 1; }
            }
            for (int j = 0; j < J; ++j) st.part_cnt[(size_t)gw * S + j] = wc[j];
        }
        for (int q = 0; q < 4; ++q) st.part_tot[(size_t)b * 4 + q] = tot[q];
        if (st.pg_correct && is_root) st.part_old[b] = tot_old;
        for (int j = 0; j < J; ++j) {
            st.part_sum[((size_t)b * S + j) * 2 + 0] = so[j];
            st.part_sum[((size_t)b * S + j) * 2 + 1] = sr[j];
            st.part_ccnt[(size_t)b * S + j] = cc[j];
        }
    }
}

void pass_minmax(const PgState& st) {
    PgCtrl* c = st.ctrl;
    const int J = c->n_jobs, S = st.S;
    for (int b = 0; b < st.grid; ++b)
        for (int j = 0; j < J; ++j) {
            uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
            const unsigned off = st.job_seg_off[j];
            const unsigned long long L = st.job_len[j];
            if (st.onehot && st.onehot[st.job_var[j]]) {     // presence mask of the node's integer codes (low / high words)
                unsigned long long bits = 0ull;
                for (int w = 0; w < PG_WARPS; ++w) {
                    unsigned long long beg, end;
                    pg_warp_range(L, b * PG_WARPS + w, st.n_gwarps, 1, &beg, &end);
                    for (unsigned long long e = beg; e < end; ++e)
                        bits |= 1ull << ((int)st.X[(size_t)st.job_var[j] * st.ld + st.pool[(size_t)off + e]] - st.col_imin[st.job_var[j]]);
                }
                st.part_mm[((size_t)b * 2 * S + j) * 2 + 0] = (unsigned)(bits & 0xFFFFFFFFull);
                st.part_mm[((size_t)b * 2 * S + j) * 2 + 1] = (unsigned)(bits >> 32);
                continue;
            }
            for (int w = 0; w < PG_WARPS; ++w) {
                unsigned long long beg, end;
                pg_warp_range(L, b * PG_WARPS + w, st.n_gwarps, 1, &beg, &end);
                for (unsigned long long e = beg; e < end; ++e) {
                    const float x = st.X[(size_t)st.job_var[j] * st.ld + st.pool[(size_t)off + e]];
                    const uint32_t k = pg_f2key(x);
                    if (x == x) { kmin = k < kmin ? k : kmin; kmax = k > kmax ? k : kmax; }
                }
            }
            st.part_mm[((size_t)b * 2 * S + j) * 2 + 0] = kmin;
            st.part_mm[((size_t)b * 2 * S + j) * 2 + 1] = kmax;
        }
}

void pass_seg(const PgState& st, int mode) {
    PgCtrl* c = st.ctrl;
    const int J = c->n_jobs, S = st.S;
    for (int b = 0; b < st.grid; ++b)
        for (int j = 0; j < J; ++j) {
            const unsigned off = st.job_seg_off[j];
            const bool ident = off == PG_SEG_IDENTITY;
            const unsigned long long L = st.job_len[j];
            double a = 0.0, bb = 0.0;
            unsigned cc = 0;
            for (int w = 0; w < PG_WARPS; ++w) {
                const int gw = b * PG_WARPS + w;
                unsigned long long beg, end;
                pg_warp_range(L, gw, st.n_gwarps, ident ? 4 : 1, &beg, &end);
                unsigned wc = 0;
                for (unsigned long long e = beg; e < end; ++e) {
                    const unsigned idx = ident ? (unsigned)e : st.pool[(size_t)off + e];
                    const double o = st.A[idx];
                    const double yv = (double)st.y[idx];
                    const bool left = st.X[(size_t)st.job_var[j] * st.ld + idx] < st.job_thr32[j];
                    if (mode == 0) {
                        if (left) { a += o; bb += yv - o; cc += 1; wc += 1; }
                    } else {
                        // the device pass's arithmetic (pg_bern_term): exp(o) per row, exp(v) per job and child
                        const double v = left ? st.job_vL[j] : st.job_vR[j];
                        const double f = pg_bern_term<1>(yv, o, v, pg_bern_exp(o), pg_bern_exp(v), st.log_tbl, 0);
                        if (left) a += f; else bb += f;
                    }
                }
                if (mode == 0) st.part_cnt[(size_t)gw * S + j] = wc;
            }
            double* dst = mode == 0 ? st.part_sum : st.part_ll;
            dst[((size_t)b * S + j) * 2 + 0] = a;
            dst[((size_t)b * S + j) * 2 + 1] = bb;
            if (mode == 0) st.part_ccnt[(size_t)b * S + j] = cc;
        }
}

void pass_part(const PgState& st) {
    PgCtrl* c = st.ctrl;
    const int K = c->n_pjobs;
    for (int k = 0; k < K; ++k) {
        const unsigned src = st.pj_src_off[k];
        const bool ident = src == PG_SEG_IDENTITY;
        const unsigned long long L = st.pj_len[k];
        unsigned* dstL = st.pool + st.pj_dst_off[k];
        unsigned* dstR = dstL + st.pj_cntL[k];
        for (int gw = 0; gw < st.n_gwarps; ++gw) {
            unsigned long long beg, end;
            pg_warp_range(L, gw, st.n_gwarps, ident ? 4 : 1, &beg, &end);
            unsigned left_run = st.pj_warp_left[(size_t)k * st.n_gwarps + gw];
            unsigned right_run = (unsigned)beg - left_run;
            const unsigned left0 = left_run;
            for (unsigned long long e = beg; e < end; ++e) {
                const unsigned idx = ident ? (unsigned)e : st.pool[(size_t)src + e];
                if (st.X[(size_t)st.pj_var[k] * st.ld + idx] < st.pj_thr32[k]) dstL[left_run++] = idx;
                else dstR[right_run++] = idx;
            }
            if (left_run - left0 != st.part_cnt[(size_t)gw * st.S + st.pj_job[k]]) c->error = PG_ERR_INTERNAL;
        }
    }
}

const char* err_text(int code) {
    switch (code) {
        case PG_ERR_DEPTH: return "Tree mutation would exceed maximum capacity";
        case PG_ERR_POOL: return "segment pool exhausted";
        case PG_ERR_TAPE: return "RNG tape exhausted";
        case PG_ERR_WEIGHTS: return "weights not finite";
        default: return "internal error";
    }
}

}  // namespace

extern "C" {

void* pgsim_create(const double* x_f, uint64_t n, uint64_t p, const double* y, const pgbart_settings* cfg, int grid) {
    Sim* s = new Sim();
    PgState& st = s->st;
    st.n = (long long)n; st.ld = (long long)((n + 31) / 32 * 32);
    st.p = (int)p; st.m = (int)cfg->n_trees; st.P = (int)cfg->n_particles; st.S = st.P - 1;
    st.max_depth = (int)cfg->max_depth; st.maxn = (1 << (st.max_depth + 1)) - 1;
    st.grid = grid; st.n_gwarps = grid * PG_WARPS;
    st.leaf_sigma = cfg->sigma; st.m_f = (double)st.m;
    st.batch_tune = cfg->batch_tune; st.batch_post = cfg->batch_post;
    st.family = PG_GAUSSIAN; st.lik_sigma = 1.0;
    st.y_mean = unrolled_mean(y, n);
    st.init_leaf = st.y_mean / (double)st.m;
    float* X = halloc<float>(s, (size_t)st.ld * p);
    float* cmin = halloc<float>(s, p);
    float* cmax = halloc<float>(s, p);
    for (uint64_t j = 0; j < p; ++j) {
        float mn = std::numeric_limits<float>::infinity(), mx = -mn;
        for (uint64_t i = 0; i < n; ++i) {
            const float v = (float)x_f[j * n + i];
            X[j * st.ld + i] = v;
            mn = std::fmin(mn, v); mx = std::fmax(mx, v);
        }
        cmin[j] = mn; cmax[j] = mx;
    }
    st.X = X; st.col_min = cmin; st.col_max = cmax;
    {   // OneHotSplit columns (same preparation as pgbart_create)
        unsigned char* oh = halloc<unsigned char>(s, p);
        int* imin = halloc<int>(s, p);
        unsigned long long* cb = halloc<unsigned long long>(s, p);
        bool any = false;
        for (uint64_t j = 0; j < cfg->n_split_rules && j < p; ++j) {
            if (!cfg->split_rules || std::strcmp(cfg->split_rules[j], "OneHotSplit") != 0) continue;
            oh[j] = 1; any = true;
            int lo = INT32_MAX;
            for (uint64_t i = 0; i < n; ++i) { const int c = (int)X[j * st.ld + i]; lo = c < lo ? c : lo; }
            imin[j] = lo;
            for (uint64_t i = 0; i < n; ++i) cb[j] |= 1ull << ((int)X[j * st.ld + i] - lo);
        }
        if (any) { st.onehot = oh; st.col_imin = imin; st.col_bits = cb; st.job_bits = halloc<unsigned long long>(s, (size_t)st.S); }
    }
    float* yy = halloc<float>(s, (size_t)st.ld);
    for (uint64_t i = 0; i < n; ++i) yy[i] = (float)y[i];
    st.y = yy;
    if (cfg->n_split_prior) {
        double* pr = halloc<double>(s, p);
        double tot = 0.0;
        for (uint64_t j = 0; j < p; ++j) { pr[j] = cfg->split_prior[j]; tot += pr[j]; }
        st.prior = pr; st.prior_total = tot;
    }
    double* thr = halloc<double>(s, (size_t)st.max_depth + 2);
    for (int d = 0; d < st.max_depth + 2; ++d) thr[d] = 1.0 - (cfg->alpha * std::pow(1.0 + (double)d, -cfg->beta));
    st.thr_depth = thr;
    double* ltb = halloc<double>(s, 2 * PG_LOGT_K);
    pg_log_table_fill(ltb);
    st.log_tbl = ltb;
    const size_t S = st.S, MN = st.maxn, M = st.m, G = st.grid, W = st.n_gwarps;
    st.A = halloc<double>(s, (size_t)st.ld); st.A2 = halloc<double>(s, (size_t)st.ld);
    st.f_split_var = halloc<int>(s, M * MN); st.f_thr32 = halloc<float>(s, M * MN); st.f_split_val = halloc<double>(s, M * MN);
    st.f_leaf_val = halloc<double>(s, M * MN); st.f_size = halloc<int>(s, M); st.f_internal = halloc<int>(s, M);
    st.pt_split_var = halloc<int>(s, 2 * S * MN); st.pt_thr32 = halloc<float>(s, 2 * S * MN); st.pt_split_val = halloc<double>(s, 2 * S * MN);
    st.pt_leaf_val = halloc<double>(s, 2 * S * MN); st.pt_seg_off = halloc<unsigned>(s, 2 * S * MN); st.pt_cnt = halloc<unsigned>(s, 2 * S * MN);
    st.pt_So = halloc<double>(s, 2 * S * MN); st.pt_Sr = halloc<double>(s, 2 * S * MN); st.pt_g = halloc<double>(s, 2 * S * MN);
    st.pt_queue = halloc<int>(s, 2 * S * MN); st.pt_size = halloc<int>(s, 2 * S); st.pt_qhead = halloc<int>(s, 2 * S);
    st.pt_qtail = halloc<int>(s, 2 * S); st.pt_G = halloc<double>(s, 2 * S);
    st.w = halloc<double>(s, S); st.mutated = halloc<int>(s, S); st.anc = halloc<int>(s, S); st.slot_job = halloc<int>(s, S);
    st.pend_node = halloc<int>(s, S); st.pend_var = halloc<int>(s, S); st.Wfinal = halloc<double>(s, 2 * (S + 1));
    st.job_slot = halloc<int>(s, S); st.job_node = halloc<int>(s, S); st.job_var = halloc<int>(s, S); st.job_thr32 = halloc<float>(s, S);
    st.job_split = halloc<double>(s, S); st.job_seg_off = halloc<unsigned>(s, S); st.job_len = halloc<unsigned>(s, S);
    st.job_lo = halloc<float>(s, S); st.job_hi = halloc<float>(s, S); st.job_vL = halloc<double>(s, S); st.job_vR = halloc<double>(s, S);
    st.job_order = halloc<int>(s, S); st.grp_first = halloc<int>(s, S + 1);
    st.red_So = halloc<double>(s, S); st.red_Sr = halloc<double>(s, S); st.red_llL = halloc<double>(s, S); st.red_llR = halloc<double>(s, S);
    st.red_cnt = halloc<unsigned long long>(s, S);
    st.pj_anc = halloc<int>(s, S); st.pj_job = halloc<int>(s, S); st.pj_var = halloc<int>(s, S); st.pj_thr32 = halloc<float>(s, S);
    st.pj_src_off = halloc<unsigned>(s, S); st.pj_len = halloc<unsigned>(s, S); st.pj_dst_off = halloc<unsigned>(s, S);
    st.pj_cntL = halloc<unsigned>(s, S); st.pj_warp_left = halloc<unsigned>(s, S * W); st.pj_of_anc = halloc<int>(s, S);
    st.scan_tmp = halloc<unsigned>(s, PG_THREADS + 1);
    st.part_sum = halloc<double>(s, G * S * 2); st.part_cnt = halloc<unsigned>(s, W * S); st.part_ccnt = halloc<unsigned>(s, G * S);
    st.part_tot = halloc<double>(s, G * 4); st.part_mm = halloc<unsigned>(s, G * 2 * S * 2); st.mq_pj = halloc<int>(s, 2 * S); st.mq_side = halloc<int>(s, 2 * S); st.mq_var = halloc<int>(s, 2 * S); st.part_ll = halloc<double>(s, G * S * 2); st.part_old = halloc<double>(s, G);
    st.info_depths = halloc<unsigned char>(s, M);
    st.ctrl = halloc<PgCtrl>(s, 1);
    st.pool_cap = (unsigned long long)std::max<size_t>(2 * S, (size_t)st.max_depth + 2) * n;
    st.pool = halloc<unsigned>(s, (size_t)st.pool_cap);
    st.rng_mode = PG_RNG_PHILOX;
    st.philox_key0 = (unsigned)cfg->seed; st.philox_key1 = (unsigned)(cfg->seed >> 32);
    st.ctrl->phase = PH_DONE; st.ctrl->tune = 1; st.ctrl->t_add = -1; st.ctrl->t_sub = -1;
    for (uint64_t i = 0; i < n; ++i) st.A[i] = st.y_mean;
    for (size_t t = 0; t < M; ++t) {
        st.f_split_var[t * MN] = -1; st.f_split_val[t * MN] = pg_nan(); st.f_leaf_val[t * MN] = st.init_leaf; st.f_size[t] = 1;
    }
    return s;
}

void pgsim_destroy(void* h) {
    Sim* s = (Sim*)h;
    for (void* p : s->allocs) std::free(p);
    delete s;
}

const char* pgsim_last_error(void* h) { return ((Sim*)h)->err.c_str(); }

int pgsim_set_likelihood(void* h, int family, const double* params, int nparams) {
    Sim* s = (Sim*)h;
    s->st.family = family;
    if (family == PG_GAUSSIAN && nparams >= 1) s->st.lik_sigma = params[0];
    return 0;
}

// SURVEY.md 8(f4) weight mode (pgbart_set_option "weight_mode")
int pgsim_set_pg_correct(void* h, int on) {
    Sim* s = (Sim*)h;
    s->st.pg_correct = on != 0;
    return 0;
}

int pgsim_set_rng_philox(void* h, uint64_t seed) {
    Sim* s = (Sim*)h;
    s->st.rng_mode = PG_RNG_PHILOX; s->st.philox_key0 = (unsigned)seed; s->st.philox_key1 = (unsigned)(seed >> 32);
    return 0;
}

int pgsim_set_rng_tape(void* h, const double* slot, const double* round, const double* upd, int64_t n_upd, int r_cap) {
    Sim* s = (Sim*)h;
    s->tape_slot.assign(slot, slot + (size_t)n_upd * r_cap * s->st.S * 5);
    s->tape_round.assign(round, round + (size_t)n_upd * r_cap);
    s->tape_upd.assign(upd, upd + (size_t)n_upd);
    s->st.rng_mode = PG_RNG_TAPE;
    s->st.tape_slot = s->tape_slot.data(); s->st.tape_round = s->tape_round.data(); s->st.tape_upd = s->tape_upd.data();
    s->st.tape_n_upd = n_upd; s->st.tape_r_cap = r_cap; s->st.tape_base_upd = s->st.ctrl->upd;
    return 0;
}

// trace buffers are caller-owned host arrays
int pgsim_trace_attach(void* h, double* slot, double* round, double* upd, int64_t u_cap, int r_cap) {
    Sim* s = (Sim*)h;
    s->st.tr_slot = slot; s->st.tr_round = round; s->st.tr_upd = upd; s->st.tr_u_cap = u_cap; s->st.tr_r_cap = r_cap;
    s->st.ctrl->trace_n = 0; s->st.ctrl->trace_overflow = 0;
    return 0;
}
int64_t pgsim_trace_count(void* h) { return (int64_t)((Sim*)h)->st.ctrl->trace_n; }
int pgsim_trace_overflow(void* h) { return ((Sim*)h)->st.ctrl->trace_overflow; }

int pgsim_step(void* h, int tune, double* out_pred) {
    Sim* s = (Sim*)h;
    if (s->dead) return 2;
    const PgState& st = s->st;
    PgCtrl* c = st.ctrl;
    if (tune >= 0) c->tune = tune;
    c->phase = PH_BEGIN; c->snext = SN_NONE;
    PgTeam tm{0, 1};
    pg_serial(st, tm);
    while (c->phase != PH_DONE) {
        switch (c->phase) {
            case PH_ROOT: case PH_FINAL: pass_root(st); break;
            case PH_MINMAX: pass_minmax(st); break;
            case PH_STATS: pass_seg(st, 0); break;
            case PH_BERN: pass_seg(st, 1); break;
            case PH_PART: pass_part(st); break;
            default: c->error = PG_ERR_INTERNAL; break;
        }
        pg_serial(st, tm);
    }
    if (c->error != PG_OK) { s->err = err_text(c->error); s->dead = true; return 3; }
    if (out_pred) std::memcpy(out_pred, st.A, (size_t)st.n * 8);
    return 0;
}

int pgsim_get_info(void* h, double* ll, int64_t* acc, uint8_t* depths, int* n_depths) {
    Sim* s = (Sim*)h;
    *ll = s->st.ctrl->info_loglik; *acc = s->st.ctrl->info_acc;
    const int cap = *n_depths;
    *n_depths = s->st.ctrl->info_n;
    for (int i = 0; i < cap && i < s->st.ctrl->info_n; ++i) depths[i] = s->st.info_depths[i];
    return 0;
}

int pgsim_tree_size(void* h, uint64_t t) { return ((Sim*)h)->st.f_size[t]; }
int pgsim_get_tree(void* h, uint64_t t, int32_t* sv, double* sp, double* lv) {
    const PgState& st = ((Sim*)h)->st;
    const size_t fb = (size_t)t * st.maxn;
    for (int i = 0; i < st.f_size[t]; ++i) { sv[i] = st.f_split_var[fb + i]; sp[i] = st.f_split_val[fb + i]; lv[i] = st.f_leaf_val[fb + i]; }
    return 0;
}
int pgsim_get_leaf_indices(void* h, uint64_t t, uint32_t* out) {
    const PgState& st = ((Sim*)h)->st;
    const size_t fb = (size_t)t * st.maxn;
    for (long long i = 0; i < st.n; ++i) {
        int node = 0, sv;
        while ((sv = st.f_split_var[fb + node]) >= 0) node = 2 * node + 1 + (st.X[(size_t)sv * st.ld + i] < st.f_thr32[fb + node] ? 0 : 1);
        out[i] = (uint32_t)node;
    }
    return 0;
}
void pgsim_get_counters(void* h, double* out4) {
    PgCtrl* c = ((Sim*)h)->st.ctrl;
    out4[0] = (double)c->bytes_model; out4[1] = (double)c->bytes_contract; out4[2] = (double)c->n_phases; out4[3] = (double)c->n_updates;
}
float pgsim_thr32(double s) { return pg_thr32(s); }
uint32_t pgsim_f2key(float f) { return pg_f2key(f); }
float pgsim_key2f(uint32_t k) { return pg_key2f(k); }
void pgsim_warp_range(uint64_t L, int gw, int W, int gran, uint64_t* b, uint64_t* e) {
    unsigned long long bb, ee;
    pg_warp_range(L, gw, W, gran, &bb, &ee);
    *b = bb; *e = ee;
}

// y mu - softplus(mu) through the product's table logarithm (pg_bern_term), elementwise: accuracy test
void pgsim_bern_terms(const double* y, const double* o, const double* v, double* out, long long n) {
    double tbl[2 * PG_LOGT_K];
    pg_log_table_fill(tbl);
    for (long long i = 0; i < n; ++i) out[i] = pg_bern_term<1>(y[i], o[i], v[i], pg_bern_exp(o[i]), pg_bern_exp(v[i]), tbl, 0);
}
void pgsim_log_ge1(const double* u, double* out, long long n) {
    double tbl[2 * PG_LOGT_K];
    pg_log_table_fill(tbl);
    for (long long i = 0; i < n; ++i) out[i] = pg_log_ge1<1>(u[i], tbl, 0);
}
}  // extern "C"
