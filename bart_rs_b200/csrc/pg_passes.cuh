// pg_passes.cuh — the streaming (grid-wide) halves of the sweep, sm_100a.
//
// Every pass partitions a segment of observation indices into one contiguous
// range per GLOBAL WARP (pg_warp_range): the same ranges are used by the
// statistics pass and, later, by the partition pass, so the per-warp LEFT
// counts of the former are the scatter offsets of the latter.
//
//   pg_pass_root    PH_ROOT / PH_FINAL. One pass over all n rows:
//                     A += tree_add(x)          kernel.rs:99-102  (commit previous tree)
//                     A -= tree_sub(x)          kernel.rs:84-87   (open this tree: A = "residuals")
//                     totals  sum A, sum (y-A), sum (y-A)^2       (stump weight, smc.rs:105-110)
//                     per job (slot): LEFT count, sum A, sum (y-A) for x[var] < split
//                                               smc.rs:201-217 fused over all slots of round 1
//                   A, y are read once per row and reused from registers by every slot; only the
//                   slots' split columns are streamed per slot.
//   pg_pass_minmax  PH_MINMAX. splitting.rs:43-49 over a node's index segment.
//   pg_pass_stats   PH_STATS.  smc.rs:201-217 over a node's index segment (gather).
//   pg_pass_bern    PH_BERN.   Bernoulli-logit child log-likelihoods (weight.rs via smc.rs:89-95).
//   pg_pass_part    PH_PART.   particle.rs:117-138 as a STABLE partition (warp ballot + prefix
//                   compaction) appended to the segment pool; only for surviving particles.
#pragma once
#include <cuda_runtime.h>
#include "pg_types.h"
#include "pg_serial.h"

#define PG_TREE_SM 63      // tree nodes staged in shared memory (depth <= 5); larger trees read global
#define PG_FULL 0xFFFFFFFFu

// Index of this block among the PASS blocks (st.grid of them).  In the persistent kernel block 0 is the serial
// block and streams nothing (its SM keeps the control code in its instruction caches), so pass block = blockIdx - 1.
#define PG_BID (sm.base->pass_bid)

// Buffers a pass works on, selected by the command flags (PgCtrl.flags): which of the two A buffers it reads /
// writes, and which parity of the job / partial-result buffers it uses.  The generic control code always
// issues flags = 0 (in-place A, parity 0); the speculative root passes of pg_serial_fast.cuh use both parities.
struct PgBuf {
    const double* A_src; double* A_dst;
    double* part_sum; unsigned* part_cnt; unsigned* part_ccnt; double* part_tot;
    const int* job_var; const float* job_thr32; const int* job_order; const int* grp_first;
    const unsigned* job_seg_off; const unsigned* job_len; const double* job_vL; const double* job_vR;
};
__host__ __device__ inline size_t pg_off_part_sum(const PgState& st, int pb) { return (size_t)pb * st.grid * st.S * 2; }
__host__ __device__ inline size_t pg_off_part_cnt(const PgState& st, int pb) { return (size_t)pb * st.n_gwarps * st.S; }
__host__ __device__ inline size_t pg_off_part_ccnt(const PgState& st, int pb) { return (size_t)pb * st.grid * st.S; }
__host__ __device__ inline size_t pg_off_part_tot(const PgState& st, int pb) { return (size_t)pb * st.grid * 4; }
__device__ __forceinline__ PgBuf pg_make_buf(const PgState& st, int flags) {
    PgBuf b;
    const int pb = (flags >> 2) & 1;
    b.A_src = (flags & 1) ? st.A2 : st.A;
    b.A_dst = (flags & 2) ? st.A2 : st.A;
    b.part_sum = st.part_sum + pg_off_part_sum(st, pb);
    b.part_cnt = st.part_cnt + pg_off_part_cnt(st, pb);
    b.part_ccnt = st.part_ccnt + pg_off_part_ccnt(st, pb);
    b.part_tot = st.part_tot + pg_off_part_tot(st, pb);
    const size_t jo = (size_t)pb * st.S;
    b.job_var = st.job_var + jo; b.job_thr32 = st.job_thr32 + jo; b.job_order = st.job_order + jo;
    b.grp_first = st.grp_first + (size_t)pb * (st.S + 1);
    b.job_seg_off = st.job_seg_off + jo; b.job_len = st.job_len + jo; b.job_vL = st.job_vL + jo; b.job_vR = st.job_vR + jo;
    return b;
}

struct PgTreeSm {
    int size;
    int from_global;       // tree too large for shared memory
    int split_var[PG_TREE_SM];
    float thr32[PG_TREE_SM];
    double leaf_val[PG_TREE_SM];
};

#define FW_DR 4            // SMC rounds whose data-independent draws are precomputed per tree update
#define FW_GW (PG_WARPS - FW_DR - 1)          // warps of the control group: warp 0 + warps FW_DR+2 .. PG_WARPS-1
#define FW_GT (FW_GW * 32)                   // threads of the control group (named barrier 1)

struct PgFastSm {          // state of the latency-organised serial section (pg_serial_fast.cuh), P <= 32.
    // Lives in block 0's shared memory for the whole kernel: loaded at launch, stored at exit.
    PgCtrl lc;                                   // working copy of the control block
    double hw[32], hG[64];                       // slot scalars: weights; [2][S] particle log-lik sums
    int hmut[32], hanc[32], hsjob[32], hsize[64], hqh[64], hqt[64];
    int jslot[32], jnode[32], jvar[32];          // job list of the current round
    unsigned jseg[32], jlen[32];
    float jthr[32];
    double jsplit[32];
    // round-0 proposals' results, kept in registers/shared memory until resampling decides who survives
    double jr_vL[32], jr_vR[32], jr_gL[32], jr_gR[32], jr_SoL[32], jr_SrL[32];
    unsigned jr_cL[32];
    double rSo[32], rSr[32], rllL[32], rllR[32]; // reduced per-job partials
    double tot[4];
    unsigned rCnt[32], rMin[32], rMax[32];
    double wsA[PG_WARPS][32], wsB[PG_WARPS][32], wsT[PG_WARPS][4];   // per-warp stage of the reduction
    unsigned wsC[PG_WARPS][32], wsD[PG_WARPS][32];
    // draws precomputed by idle warps: [update mod 3][round][slot] (the update in progress and the next two)
    double d_u1[3][FW_DR][32], d_u3[3][FW_DR][32], d_zL[3][FW_DR][32], d_zR[3][FW_DR][32];
    int d_var[3][FW_DR][32];
    double d_u6[3][FW_DR], d_u7[3];
    int d_ok[3][FW_DR + 1];                      // draw source in range for (round) / for u7
    int d0_status[3][32];                        // complete round-0 proposal: 1 depth reject, 3 min>=max, 5 live, 6 tape error, 7 depth error
    float d0_lo[3][32], d0_hi[3][32], d0_thr[3][32];
    double d0_split[3][32];
    unsigned long long d_upd[3];                 // which tree update each buffer holds
    volatile unsigned long long d_flag[3][FW_DR + 1];   // per helper warp: token (= update + 1) when its part is written
    int sn;
    // control loop of the serial block: the pass that is in flight
    unsigned issued;                             // commands issued so far (= value of bar_gen)
    int out_phase, out_J, out_pb;                // phase / job count / buffer parity of the outstanding pass
    int quit;
    // look-ahead caches of the update in progress (rounds 1 and 2), indexed [round-1][slot]
    int mm_valid[2][32]; unsigned mm_seg[2][32]; int mm_var[2][32]; float mm_lo[2][32], mm_hi[2][32];
    int sc_valid[32]; unsigned sc_seg[32]; int sc_var[32]; float sc_thr[32]; unsigned sc_cnt[32]; double sc_So[32], sc_Sr[32];
    int rq_r[64], rq_s[64]; unsigned rq_seg[64];  // request -> (round, slot, child segment)
    int n_mq, pf_n, pf_slot[32];
    unsigned long long pre_upd;                  // update whose root jobs are already staged in the global job arrays
    int pre_J, pre_G, spec_now;
    volatile unsigned long long want_base, want_top;   // draw helpers: updates [want_base, want_top] are wanted
    volatile int helpers_quit;
    double stump_total, stump_cum[32];           // final selection when every particle is a stump: sequential sums of 1/P
    unsigned scan[PG_THREADS + 1];
};

struct PgSmem {
    PgTreeSm tadd, tsub;
    double tot[PG_WARPS][4];
    PgFastSm fast;
    double cthr[16];       // depth-prior thresholds (staged once per kernel)
    int cstaged;           // split prior / column min-max are staged in the dynamic area
    int pass_bid;          // this block's index among the pass blocks (-1: serial-only block)
    int is_last;
    // followed by dynamic arrays, see pg_smem_* accessors
};

#define FW_MAXP 512   // split prior / column min-max staged in shared memory up to this many columns

// dynamic shared memory layout after PgSmem (all sized by S):
//   int   jvar[S]; float jthr[S]; int jorder[S]; int grp[S+1]; double jvL[S]; double jvR[S];
//   double accA[PG_WARPS][S]; double accB[PG_WARPS][S]; unsigned accC[PG_WARPS][S];
//   unsigned mmin[S]; unsigned mmax[S];
__host__ __device__ inline size_t pg_smem_bytes(int S) {
    size_t b = sizeof(PgSmem);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)S * (4 + 4 + 4) + (size_t)(S + 1) * 4;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)S * 16;                       // jvL, jvR
    b += (size_t)PG_WARPS * S * (8 + 8 + 4);   // accA, accB, accC
    b += (size_t)S * 16;                       // mmin, mmax (2S each)
    b = (b + 15) & ~(size_t)15;
    b += (size_t)FW_MAXP * (8 + 4 + 4);        // cprior, cmin, cmax
    return b + 64;
}

struct PgSmemView {
    PgSmem* base;
    int* jvar; float* jthr; int* jorder; int* grp; double* jvL; double* jvR;
    double* accA; double* accB; unsigned* accC; unsigned* mmin; unsigned* mmax;
    double* cprior; float* cmin; float* cmax;
};

__device__ __forceinline__ PgSmemView pg_smem_view(unsigned char* raw, int S) {
    PgSmemView v;
    v.base = reinterpret_cast<PgSmem*>(raw);
    size_t off = (sizeof(PgSmem) + 15) & ~(size_t)15;
    v.jvL = reinterpret_cast<double*>(raw + off); off += (size_t)S * 8;
    v.jvR = reinterpret_cast<double*>(raw + off); off += (size_t)S * 8;
    v.accA = reinterpret_cast<double*>(raw + off); off += (size_t)PG_WARPS * S * 8;
    v.accB = reinterpret_cast<double*>(raw + off); off += (size_t)PG_WARPS * S * 8;
    v.accC = reinterpret_cast<unsigned*>(raw + off); off += (size_t)PG_WARPS * S * 4;
    v.mmin = reinterpret_cast<unsigned*>(raw + off); off += (size_t)S * 8;
    v.mmax = reinterpret_cast<unsigned*>(raw + off); off += (size_t)S * 8;
    v.jvar = reinterpret_cast<int*>(raw + off); off += (size_t)S * 4;
    v.jthr = reinterpret_cast<float*>(raw + off); off += (size_t)S * 4;
    v.jorder = reinterpret_cast<int*>(raw + off); off += (size_t)S * 4;
    v.grp = reinterpret_cast<int*>(raw + off); off += (size_t)(S + 1) * 4;
    off = (off + 15) & ~(size_t)15;
    v.cprior = reinterpret_cast<double*>(raw + off); off += (size_t)FW_MAXP * 8;
    v.cmin = reinterpret_cast<float*>(raw + off); off += (size_t)FW_MAXP * 4;
    v.cmax = reinterpret_cast<float*>(raw + off);
    return v;
}

// ---- loads ------------------------------------------------------------------
// X and y never change: read-only path.  A and the pool are rewritten between passes by other
// SMs: read through L2 (ld.global.cg) so a stale L1 line can never be observed.
// X columns are streamed once per pass and are far larger than L2: load them with the streaming (evict-first) policy so
// they do not push out what IS reused every pass — A, y, and the control block's code and tables.
__device__ __forceinline__ float4 pg_ld_x4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ double2 pg_ld_a2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

// ---- warp-transposed reduction of 4 jobs' partial sums ----------------------
// Each lane holds 4 values (one per job).  After the call every lane holds the warp total of job
// ((lane & 1) << 1) | ((lane >> 1) & 1): 6 shuffles instead of 20.  Fixed pattern => deterministic.
template <typename T>
__device__ __forceinline__ T pg_xreduce4(T v0, T v1, T v2, T v3, int lane) {
    const bool b0 = lane & 1, b1 = lane & 2;
    T s0 = b0 ? v0 : v2, k0 = b0 ? v2 : v0;
    T s1 = b0 ? v1 : v3, k1 = b0 ? v3 : v1;
    T u0 = k0 + __shfl_xor_sync(PG_FULL, s0, 1);
    T u1 = k1 + __shfl_xor_sync(PG_FULL, s1, 1);
    T s = b1 ? u0 : u1, k = b1 ? u1 : u0;
    T t = k + __shfl_xor_sync(PG_FULL, s, 2);
    t = t + __shfl_xor_sync(PG_FULL, t, 4);
    t = t + __shfl_xor_sync(PG_FULL, t, 8);
    t = t + __shfl_xor_sync(PG_FULL, t, 16);
    return t;
}
__device__ __forceinline__ int pg_xreduce4_job(int lane) { return ((lane & 1) << 1) | ((lane >> 1) & 1); }

__device__ __forceinline__ double pg_warp_sum(double v) {
    v += __shfl_xor_sync(PG_FULL, v, 16);
    v += __shfl_xor_sync(PG_FULL, v, 8);
    v += __shfl_xor_sync(PG_FULL, v, 4);
    v += __shfl_xor_sync(PG_FULL, v, 2);
    v += __shfl_xor_sync(PG_FULL, v, 1);
    return v;
}

// ---- tree staging / evaluation (tree.rs:169-192 traversal on training rows) ------------------
__device__ __forceinline__ void pg_stage_tree(const PgState& st, int t, PgTreeSm& sm) {
    if (t < 0) { if (threadIdx.x == 0) { sm.size = 0; sm.from_global = 0; } return; }
    const int sz = __ldcg(&st.f_size[t]);
    if (threadIdx.x == 0) { sm.size = sz; sm.from_global = sz > PG_TREE_SM; }
    if (sz <= PG_TREE_SM)
        for (int i = threadIdx.x; i < sz; i += blockDim.x) {
            sm.split_var[i] = __ldcg(&st.f_split_var[(size_t)t * st.maxn + i]);
            sm.thr32[i] = __ldcg(&st.f_thr32[(size_t)t * st.maxn + i]);
            sm.leaf_val[i] = __ldcg(&st.f_leaf_val[(size_t)t * st.maxn + i]);
        }
}

__device__ __forceinline__ double pg_tree_eval(const PgState& st, int t, const PgTreeSm& sm, long long row) {
    int node = 0;
    if (!sm.from_global) {
        int sv;
        while ((sv = sm.split_var[node]) >= 0) {
            const float x = __ldg(&st.X[(size_t)sv * st.ld + row]);
            node = 2 * node + 1 + (x < sm.thr32[node] ? 0 : 1);
        }
        return sm.leaf_val[node];
    }
    const size_t fb = (size_t)t * st.maxn;
    int sv;
    while ((sv = __ldcg(&st.f_split_var[fb + node])) >= 0) {
        const float x = __ldg(&st.X[(size_t)sv * st.ld + row]);
        node = 2 * node + 1 + (x < __ldcg(&st.f_thr32[fb + node]) ? 0 : 1);
    }
    return __ldcg(&st.f_leaf_val[fb + node]);
}

// Stage the job list of the current pass into shared memory.
__device__ __forceinline__ void pg_stage_jobs(const PgState& st, const PgBuf& bf, const PgSmemView& sm, int J, int G, bool with_values) {
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        sm.jvar[j] = __ldcg(&bf.job_var[j]);
        sm.jthr[j] = __ldcg(&bf.job_thr32[j]);
        sm.jorder[j] = __ldcg(&bf.job_order[j]);
        if (with_values) { sm.jvL[j] = __ldcg(&bf.job_vL[j]); sm.jvR[j] = __ldcg(&bf.job_vR[j]); }
    }
    for (int g = threadIdx.x; g <= G; g += blockDim.x) sm.grp[g] = __ldcg(&bf.grp_first[g]);
}

__device__ __forceinline__ void pg_zero_acc(const PgSmemView& sm, int S) {
    for (int i = threadIdx.x; i < PG_WARPS * S; i += blockDim.x) { sm.accA[i] = 0.0; sm.accB[i] = 0.0; sm.accC[i] = 0u; }
}

// Write the block's per-job results: fixed warp order => deterministic.
__device__ __forceinline__ void pg_flush_acc(const PgState& st, const PgBuf& bf, const PgSmemView& sm, int J, bool counts, bool as_ll) {
    const int S = st.S;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        double a = 0.0, b = 0.0;
        unsigned cc = 0u;
        for (int w = 0; w < PG_WARPS; ++w) { a += sm.accA[w * S + j]; b += sm.accB[w * S + j]; cc += sm.accC[w * S + j]; }
        double* dst = as_ll ? st.part_ll : bf.part_sum;
        dst[((size_t)PG_BID * S + j) * 2 + 0] = a;
        dst[((size_t)PG_BID * S + j) * 2 + 1] = b;
        if (counts) {
            bf.part_ccnt[(size_t)PG_BID * S + j] = cc;
            for (int w = 0; w < PG_WARPS; ++w)
                bf.part_cnt[((size_t)PG_BID * PG_WARPS + w) * S + j] = sm.accC[w * S + j];
        }
    }
}

// =============================================================================
// PH_ROOT / PH_FINAL
// =============================================================================
// Rows per lane per iteration: 8 (two float4 groups) => 256 rows per warp-iteration.
template <bool BERN>
__device__ void pg_pass_root(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    PgCtrl* c = st.ctrl;
    const int J = __ldcg(&c->n_jobs) * (__ldcg(&c->phase) == PH_ROOT ? 1 : 0);
    const int t_add = __ldcg(&c->t_add), t_sub = __ldcg(&c->phase) == PH_ROOT ? __ldcg(&c->t_sub) : -1;
    const int S = st.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = PG_BID * PG_WARPS + warp;

    pg_stage_tree(st, t_add, sm.base->tadd);   // t_add == -2: a stump (constant init_leaf), no forest read
    pg_stage_tree(st, t_sub, sm.base->tsub);
    pg_stage_jobs(st, bf, sm, J, 0, false);
    pg_zero_acc(sm, S);
    __syncthreads();
    const bool add_nontrivial = t_add >= 0 && sm.base->tadd.size > 1;
    const bool sub_nontrivial = t_sub >= 0 && sm.base->tsub.size > 1;
    const double add_c = t_add == -2 ? st.init_leaf : ((t_add >= 0 && !add_nontrivial) ? sm.base->tadd.leaf_val[0] : 0.0);
    const double sub_c = (t_sub >= 0 && !sub_nontrivial) ? sm.base->tsub.leaf_val[0] : 0.0;

    unsigned long long beg, end;
    pg_warp_range((unsigned long long)st.n, gw, st.n_gwarps, 4, &beg, &end);
    const long long n = st.n;
    const size_t ld = (size_t)st.ld;
    double tot_o = 0.0, tot_r = 0.0, tot_rr = 0.0, tot_ll = 0.0;

    for (unsigned long long base = beg; base < end; base += 256) {
        const long long r0 = (long long)base + 4 * lane, r1 = r0 + 128;
        const bool g0 = r0 < (long long)end, g1 = r1 < (long long)end;   // group start inside the warp's range
        double o[8];
        float yv[8];
        {
            double2 a;
            float4 yy;
            if (g0) {
                a = pg_ld_a2(bf.A_src + r0); o[0] = a.x; o[1] = a.y;
                a = pg_ld_a2(bf.A_src + r0 + 2); o[2] = a.x; o[3] = a.y;
                yy = pg_ld_x4(st.y + r0); yv[0] = yy.x; yv[1] = yy.y; yv[2] = yy.z; yv[3] = yy.w;
            } else { o[0] = o[1] = o[2] = o[3] = 0.0; yv[0] = yv[1] = yv[2] = yv[3] = 0.0f; }
            if (g1) {
                a = pg_ld_a2(bf.A_src + r1); o[4] = a.x; o[5] = a.y;
                a = pg_ld_a2(bf.A_src + r1 + 2); o[6] = a.x; o[7] = a.y;
                yy = pg_ld_x4(st.y + r1); yv[4] = yy.x; yv[5] = yy.y; yv[6] = yy.z; yv[7] = yy.w;
            } else { o[4] = o[5] = o[6] = o[7] = 0.0; yv[4] = yv[5] = yv[6] = yv[7] = 0.0f; }
        }
        bool valid[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) valid[q] = ((q < 4 ? r0 : r1 - 4) + q) < n && (q < 4 ? g0 : g1);
        // kernel.rs:99-102 then :84-87, each a separately rounded f64 operation
        if (t_add >= 0 || t_add == -2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double v = add_nontrivial ? (valid[q] ? pg_tree_eval(st, t_add, sm.base->tadd, (q < 4 ? r0 : r1 - 4) + q) : 0.0) : add_c;
                o[q] = o[q] + v;
            }
        }
        if (t_sub >= 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double v = sub_nontrivial ? (valid[q] ? pg_tree_eval(st, t_sub, sm.base->tsub, (q < 4 ? r0 : r1 - 4) + q) : 0.0) : sub_c;
                o[q] = o[q] - v;
            }
        }
        if (g0) {
            __stcg(reinterpret_cast<double2*>(bf.A_dst + r0), make_double2(o[0], o[1]));
            __stcg(reinterpret_cast<double2*>(bf.A_dst + r0 + 2), make_double2(o[2], o[3]));
        }
        if (g1) {
            __stcg(reinterpret_cast<double2*>(bf.A_dst + r1), make_double2(o[4], o[5]));
            __stcg(reinterpret_cast<double2*>(bf.A_dst + r1 + 2), make_double2(o[6], o[7]));
        }
        if (J == 0 && t_sub < 0) continue;   // PH_FINAL: commit only
        double r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (!valid[q]) { o[q] = 0.0; r[q] = 0.0; }
            else r[q] = (double)yv[q] - o[q];
            tot_o += o[q];
            tot_r += r[q];
            tot_rr += r[q] * r[q];
            if (BERN && valid[q]) {
                const double mu = o[q] + st.init_leaf;
                tot_ll += (double)yv[q] * mu - pg_softplus(mu);
            }
        }
        // smc.rs:201-217 for every slot of round 1, four slots at a time
        for (int jb = 0; jb < J; jb += 4) {
            double so[4], sr[4];
            int cn[4];
            float4 xa[4], xb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool live = jb + q < J;
                const float* col = st.X + (size_t)(live ? sm.jvar[jb + q] : 0) * ld;
                xa[q] = (live && g0) ? pg_ld_x4(col + r0) : make_float4(0.f, 0.f, 0.f, 0.f);
                xb[q] = (live && g1) ? pg_ld_x4(col + r1) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool live = jb + q < J;
                const float thr = live ? sm.jthr[jb + q] : 0.0f;
                const float xs[8] = {xa[q].x, xa[q].y, xa[q].z, xa[q].w, xb[q].x, xb[q].y, xb[q].z, xb[q].w};
                double a = 0.0, b = 0.0;
                int cc = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const bool left = live && valid[e] && (xs[e] < thr);
                    a += left ? o[e] : 0.0;
                    b += left ? r[e] : 0.0;
                    cc += left ? 1 : 0;
                }
                so[q] = a; sr[q] = b; cn[q] = cc;
            }
            const double ta = pg_xreduce4(so[0], so[1], so[2], so[3], lane);
            const double tb = pg_xreduce4(sr[0], sr[1], sr[2], sr[3], lane);
            const int tc = pg_xreduce4(cn[0], cn[1], cn[2], cn[3], lane);
            const int jj = jb + pg_xreduce4_job(lane);
            if (lane < 4 && jj < J) {
                sm.accA[warp * S + jj] += ta;
                sm.accB[warp * S + jj] += tb;
                sm.accC[warp * S + jj] += (unsigned)tc;
            }
        }
    }
    // block totals
    tot_o = pg_warp_sum(tot_o); tot_r = pg_warp_sum(tot_r); tot_rr = pg_warp_sum(tot_rr); tot_ll = pg_warp_sum(tot_ll);
    if (lane == 0) { sm.base->tot[warp][0] = tot_o; sm.base->tot[warp][1] = tot_r; sm.base->tot[warp][2] = tot_rr; sm.base->tot[warp][3] = tot_ll; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0.0;
        for (int w = 0; w < PG_WARPS; ++w) v += sm.base->tot[w][threadIdx.x];
        bf.part_tot[(size_t)PG_BID * 4 + threadIdx.x] = v;
    }
    pg_flush_acc(st, bf, sm, J, true, false);
}

// =============================================================================
// PH_MINMAX — splitting.rs:43-49 over index segments
// =============================================================================
__device__ void pg_pass_minmax(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    PgCtrl* c = st.ctrl;
    const int J = __ldcg(&c->n_jobs);
    const int S = st.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = PG_BID * PG_WARPS + warp;
    for (int j = threadIdx.x; j < J; j += blockDim.x) { sm.mmin[j] = 0xFFFFFFFFu; sm.mmax[j] = 0u; sm.jvar[j] = __ldcg(&bf.job_var[j]); }
    __syncthreads();
    const size_t ld = (size_t)st.ld;
    for (int j = 0; j < J; ++j) {
        const unsigned off = __ldcg(&bf.job_seg_off[j]);
        const unsigned long long L = __ldcg(&bf.job_len[j]);
        const float* col = st.X + (size_t)sm.jvar[j] * ld;
        unsigned long long beg, end;
        pg_warp_range(L, gw, st.n_gwarps, 1, &beg, &end);
        float mn = INFINITY, mx = -INFINITY;
        bool any = false;
        for (unsigned long long e = beg + lane; e < end; e += 32) {
            const unsigned idx = __ldcg(&st.pool[(size_t)off + e]);
            const float x = __ldg(&col[idx]);
            mn = fminf(mn, x); mx = fmaxf(mx, x);   // f64::min/max ignore NaN like fminf/fmaxf
            any = true;
        }
        unsigned kmn = any ? pg_f2key(mn) : 0xFFFFFFFFu, kmx = any ? pg_f2key(mx) : 0u;
        kmn = __reduce_min_sync(PG_FULL, kmn);
        kmx = __reduce_max_sync(PG_FULL, kmx);
        if (lane == 0) { atomicMin(&sm.mmin[j], kmn); atomicMax(&sm.mmax[j], kmx); }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        st.part_mm[((size_t)PG_BID * 2 * S + j) * 2 + 0] = sm.mmin[j];
        st.part_mm[((size_t)PG_BID * 2 * S + j) * 2 + 1] = sm.mmax[j];
    }
}

// =============================================================================
// PH_STATS / PH_BERN — over index segments (gather); jobs sharing a segment share idx/A/y loads
// =============================================================================
// MODE 0: statistics (LEFT count, sum A, sum y-A).  MODE 1: Bernoulli child log-likelihoods
// (A = left-child sum, B = right-child sum); identity segments allowed.
template <int MODE>
__device__ void pg_pass_seg(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    PgCtrl* c = st.ctrl;
    const int J = __ldcg(&c->n_jobs), G = __ldcg(&c->n_groups);
    const int S = st.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = PG_BID * PG_WARPS + warp;
    pg_stage_jobs(st, bf, sm, J, G, MODE == 1);
    pg_zero_acc(sm, S);
    __syncthreads();
    const size_t ld = (size_t)st.ld;
    constexpr int R = 4;   // entries per lane per iteration
    for (int g = 0; g < G; ++g) {
        const int jb0 = sm.grp[g], jb1 = sm.grp[g + 1];
        const int j0 = sm.jorder[jb0];
        const unsigned off = __ldcg(&bf.job_seg_off[j0]);
        const bool ident = off == PG_SEG_IDENTITY;
        const unsigned long long L = __ldcg(&bf.job_len[j0]);
        unsigned long long beg, end;
        pg_warp_range(L, gw, st.n_gwarps, ident ? 4 : 1, &beg, &end);
        for (unsigned long long base = beg; base < end; base += 32 * R) {
            unsigned idx[R];
            double o[R], r[R];
            float yv[R];
            bool valid[R];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const unsigned long long e = base + lane + 32ull * k;
                valid[k] = e < end;
                idx[k] = valid[k] ? (ident ? (unsigned)e : __ldcg(&st.pool[(size_t)off + e])) : 0u;
            }
#pragma unroll
            for (int k = 0; k < R; ++k) {
                o[k] = valid[k] ? __ldcg(&bf.A_src[idx[k]]) : 0.0;
                yv[k] = valid[k] ? __ldg(&st.y[idx[k]]) : 0.0f;
                r[k] = valid[k] ? (double)yv[k] - o[k] : 0.0;
            }
            for (int jb = jb0; jb < jb1; jb += 4) {
                double va[4], vb[4];
                int cn[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool live = jb + q < jb1;
                    const int j = live ? sm.jorder[jb + q] : j0;
                    const float* col = st.X + (size_t)sm.jvar[j] * ld;
                    const float thr = sm.jthr[j];
                    double a = 0.0, b = 0.0;
                    int cc = 0;
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        const bool ok = live && valid[k];
                        const float x = ok ? __ldg(&col[idx[k]]) : 0.0f;
                        const bool left = ok && (x < thr);
                        if (MODE == 0) {
                            a += left ? o[k] : 0.0;
                            b += left ? r[k] : 0.0;
                            cc += left ? 1 : 0;
                        } else if (ok) {
                            const double mu = o[k] + (left ? sm.jvL[j] : sm.jvR[j]);
                            const double f = (double)yv[k] * mu - pg_softplus(mu);
                            if (left) a += f; else b += f;
                        }
                    }
                    va[q] = a; vb[q] = b; cn[q] = cc;
                }
                const double ta = pg_xreduce4(va[0], va[1], va[2], va[3], lane);
                const double tb = pg_xreduce4(vb[0], vb[1], vb[2], vb[3], lane);
                const int tc = MODE == 0 ? pg_xreduce4(cn[0], cn[1], cn[2], cn[3], lane) : 0;
                const int q = pg_xreduce4_job(lane);
                if (lane < 4 && jb + q < jb1) {
                    const int j = sm.jorder[jb + q];
                    sm.accA[warp * S + j] += ta;
                    sm.accB[warp * S + j] += tb;
                    sm.accC[warp * S + j] += (unsigned)tc;
                }
            }
        }
    }
    __syncthreads();
    pg_flush_acc(st, bf, sm, J, MODE == 0, MODE == 1);
}

// =============================================================================
// PH_PART — stable partition of a segment by x[var] < split into the pool
// =============================================================================
__device__ void pg_pass_part(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    PgCtrl* c = st.ctrl;
    const int K = __ldcg(&c->n_pjobs);
    const int Q = __ldcg(&c->n_groups);     // look-ahead min/max requests riding on this pass (0 in the generic control code)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = PG_BID * PG_WARPS + warp;
    const unsigned lt_mask = (1u << lane) - 1u;
    const size_t ld = (size_t)st.ld;
    if (Q > 0) {
        for (int q = threadIdx.x; q < Q; q += blockDim.x) { sm.mmin[q] = 0xFFFFFFFFu; sm.mmax[q] = 0u; }
        __syncthreads();
    }
    for (int k = 0; k < K; ++k) {
        const unsigned src = __ldcg(&st.pj_src_off[k]);
        const bool ident = src == PG_SEG_IDENTITY;
        const unsigned long long L = __ldcg(&st.pj_len[k]);
        const float* col = st.X + (size_t)__ldcg(&st.pj_var[k]) * ld;
        const float thr = __ldcg(&st.pj_thr32[k]);
        unsigned long long beg, end;
        pg_warp_range(L, gw, st.n_gwarps, ident ? 4 : 1, &beg, &end);
        unsigned left_run = __ldcg(&st.pj_warp_left[(size_t)k * st.n_gwarps + gw]);
        unsigned right_run = (unsigned)beg - left_run;
        const unsigned left0 = left_run;
        unsigned* dstL = st.pool + (size_t)__ldcg(&st.pj_dst_off[k]);
        unsigned* dstR = dstL + __ldcg(&st.pj_cntL[k]);
        for (unsigned long long base = beg; base < end; base += 32) {
            const unsigned long long e = base + lane;
            const bool valid = e < end;
            const unsigned idx = valid ? (ident ? (unsigned)e : __ldcg(&st.pool[(size_t)src + e])) : 0u;
            const float x = valid ? __ldg(&col[idx]) : 0.0f;
            const bool left = valid && (x < thr);
            const unsigned bl = __ballot_sync(PG_FULL, left);
            const unsigned bv = __ballot_sync(PG_FULL, valid);
            const unsigned br = bv & ~bl;
            if (valid) {
                if (left) __stcg(&dstL[left_run + __popc(bl & lt_mask)], idx);
                else __stcg(&dstR[right_run + __popc(br & lt_mask)], idx);
            }
            left_run += __popc(bl);
            right_run += __popc(br);
        }
        // the statistics pass counted the same predicate over the same range
        const int j = __ldcg(&st.pj_job[k]);
        if (lane == 0 && (left_run - left0) != __ldcg(&bf.part_cnt[(size_t)gw * st.S + j])) c->error = PG_ERR_INTERNAL;
        // look-ahead: splitting.rs:43-49 for proposals the control block expects on the two children (their rows are
        // streaming through this SM's caches right now: no gather pass later)
        for (int q = 0; q < Q; ++q) {
            if (__ldcg(&st.mq_pj[q]) != k) continue;
            const bool want_left = __ldcg(&st.mq_side[q]) == 0;
            const float* col2 = st.X + (size_t)__ldcg(&st.mq_var[q]) * ld;
            float mn = INFINITY, mx = -INFINITY;
            bool any = false;
            for (unsigned long long e = beg + lane; e < end; e += 32) {
                const unsigned idx = ident ? (unsigned)e : __ldcg(&st.pool[(size_t)src + e]);
                const bool left = __ldg(&col[idx]) < thr;
                if (left == want_left) { const float v = __ldg(&col2[idx]); mn = fminf(mn, v); mx = fmaxf(mx, v); any = true; }
            }
            unsigned kmn = any ? pg_f2key(mn) : 0xFFFFFFFFu, kmx = any ? pg_f2key(mx) : 0u;
            kmn = __reduce_min_sync(PG_FULL, kmn);
            kmx = __reduce_max_sync(PG_FULL, kmx);
            if (lane == 0) { atomicMin(&sm.mmin[q], kmn); atomicMax(&sm.mmax[q], kmx); }
        }
    }
    if (Q > 0) {
        __syncthreads();
        for (int q = threadIdx.x; q < Q; q += blockDim.x) {
            st.part_mm[((size_t)PG_BID * 2 * st.S + q) * 2 + 0] = sm.mmin[q];
            st.part_mm[((size_t)PG_BID * 2 * st.S + q) * 2 + 1] = sm.mmax[q];
        }
    }
}

// Dispatch one grid pass.
__device__ __forceinline__ void pg_run_pass(const PgState& st, const PgSmemView& sm, int phase) {
    const PgBuf bf = pg_make_buf(st, __ldcg(&st.ctrl->flags));
    switch (phase) {
        case PH_ROOT:
        case PH_FINAL:
            if (st.family == PG_BERNOULLI_LOGIT) pg_pass_root<true>(st, bf, sm); else pg_pass_root<false>(st, bf, sm);
            break;
        case PH_MINMAX: pg_pass_minmax(st, bf, sm); break;
        case PH_STATS: pg_pass_seg<0>(st, bf, sm); break;
        case PH_BERN: pg_pass_seg<1>(st, bf, sm); break;
        case PH_PART: pg_pass_part(st, bf, sm); break;
        default: break;
    }
}

// ---- grid-level synchronisation primitives -------------------------------------------------------
// Polling load: relaxed at gpu scope (served by L2, does not touch this SM's L1); the acquire is one
// fence after the flag has flipped, so a spinning block does not keep invalidating the L1 that the
// serial section running on the same SM depends on.
__device__ __forceinline__ unsigned pg_ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pg_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void pg_spin_until(const unsigned* p, unsigned target, PgCtrl* c) {
    const long long t0 = clock64();
    unsigned ns = 20;
    while (pg_ld_relaxed(p) < target) {
        __nanosleep(ns);
        if (ns < 160) ns <<= 1;
        if (clock64() - t0 > 40000000000ll) {   // ~20 s at 2 GHz: a lost block becomes an error, not a hang
            c->error = PG_ERR_WATCHDOG;
            __threadfence();
            __trap();
        }
    }
    __threadfence();   // acquire side of the fence/atomic/fence pattern
}

