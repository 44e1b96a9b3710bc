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
#include <type_traits>
#include "pg_types.h"
#include "pg_serial.h"

#define PG_TREE_SM 63      // tree nodes staged in shared memory (depth <= 5); larger trees read global
#define PGR_RG 128         // ring root pass: rows per row group (32 lanes x 4 consecutive rows)
#define PGR_MAXG 8         //   row groups per tile at most (1024 rows: 4 KB per column copy)
#define PGR_MAXST 8        //   stages of the ring at most
#define PGR_MAXU 40        //   distinct columns per tile (S <= 31 job columns + tree columns)
#define PGR_JW 8           //   jobs per warp: 4 job groups x 8 >= S
#define PG_FULL 0xFFFFFFFFu

// Index of this block among the PASS blocks (st.grid of them).  In the persistent kernel block 0 is the serial
// block and streams nothing (its SM keeps the control code in its instruction caches), so pass block = blockIdx - 1.
#define PG_BID (sm.base->pass_bid)

__device__ __forceinline__ void pg_spin_until(const unsigned* p, unsigned target, PgCtrl* c);   // defined with the other grid-level primitives below

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

struct PgFastSm {          // state of the latency-organised serial section (pg_serial_fast.cuh), P <= 32.
    // Lives in block 0's shared memory for the whole kernel: loaded at launch, stored at exit.
    PgCtrl lc;                                   // working copy of the control block
    double hw[32], hG[64];                       // slot scalars: weights; [2][S] particle log-lik sums
    int hmut[32], hanc[32], hsjob[32], hsize[64], hqh[64], hqt[64];
    int hlin[32];                                // lineage id of each slot: equal ids <=> identical slot tables
    double hGp[32];                              // rounds >= 1: log-lik sum a slot would have with its pending (deferred) mutation
    int jslot[32], jnode[32], jvar[32];          // job list of the current round
    unsigned jseg[32], jlen[32];
    float jthr[32];
    double jsplit[32];
    // round-0 proposals' results, kept in registers/shared memory until resampling decides who survives
    double jr_vL[32], jr_vR[32], jr_gL[32], jr_gR[32], jr_SoL[32], jr_SrL[32];
    unsigned jr_cL[32];
    unsigned jr_cLl[32];                         // left count among THIS rank's rows (= jr_cL on one GPU)
    double rSo[32], rSr[32], rllL[32], rllR[32]; // reduced per-job partials
    double tot[4];
    unsigned rCnt[32], rMin[32], rMax[32];
    unsigned rCntL[32];                          // rCnt before the exchange between ranks: this rank's rows only
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
    // control loop of the serial block: the passes in flight (at most two: the pass in progress and a root pass issued ahead)
    unsigned issued;                             // commands issued so far (= value of bar_gen)
    unsigned cur_gen;                            // generation of the pass whose completion the current section digests
    int out_phase[2], out_J[2], out_pb[2];       // by generation parity: phase / job count / buffer parity of an issued pass
    int quit;
    int stat_spec;                               // profiling: a root pass was in flight ahead during this section
    int pred_tried;                              // early commit: the predicted root pass of this update has been issued once
    // look-ahead caches of the update in progress (rounds 1 and 2), indexed [round-1][slot]
    int mm_valid[2][32]; unsigned mm_seg[2][32]; int mm_var[2][32]; float mm_lo[2][32], mm_hi[2][32];
    // statistics known ahead of their round, indexed [round-1][slot]: prefetched by the round-1 statistics pass (round 2) or
    // computed by the early statistics pass queued behind the partition pass (rounds 1 and 2)
    int sc_valid[2][32]; unsigned sc_seg[2][32]; int sc_var[2][32]; float sc_thr[2][32]; unsigned sc_cnt[2][32], sc_cntL[2][32]; double sc_So[2][32], sc_Sr[2][32];
    int sc_job[2][32], from_cache;                  // job index the prefetched statistics had in their pass (its per-block counts live there)
    int rq_r[64], rq_s[64]; unsigned rq_seg[64];  // request -> (round, slot, child segment)
    int n_mq, pf_n, pf_slot[32];
    unsigned long long pre_upd;                  // update whose root jobs are already staged in the global job arrays
    int pre_J, pre_G, spec_now;
    // early statistics pass: queued behind the partition pass (early_now), then the pass in progress whose
    // results the round logic waits for (es_pending); es_resume = what to do when they arrive (SN_DRAW_ROUND: finish the
    // round whose proposals are parked in wj_*; SN_FINALIZE: nothing needed them any more)
    int early_now, es_pending, es_resume, es_n, es_mask;   // es_mask: rounds (bit round-1) of this update an early statistics pass covered
    int wj_live[32], wj_node[32], wj_var[32]; unsigned wj_seg[32], wj_len[32]; float wj_lo[32], wj_hi[32], wj_thr[32]; double wj_split[32];
    int sl_valid, sl_lin; unsigned sl_cnt[2], sl_len[2], sl_seg[2]; double sl_So[2], sl_Sr[2], sl_g[2];   // single lineage after round 0: its two leaves (mirror of the slot tables)
    int pq_job[32], pq_ident[32]; unsigned pq_len[32];   // partition jobs of the round (mirrors of pj_job / pj_src_off / pj_len)
    int pred_valid, pred_sel, pred_lin, pred_ok;  // early commit: the tree this update will commit, predicted after round 0
    int lazy_J;                                  // pruned round 0: accepted root proposals that were not evaluated (they count as grows)
    double stump_total, stump_cum[32];           // final selection when every particle is a stump: sequential sums of 1/P
    unsigned scan[PG_THREADS + 1];
};

struct PgSmem {
    PgCmd cmd;             // the command of the pass being executed (pg_fetch_cmd)
    PgTreeSm tadd, tsub;
    double tot[PG_WARPS][4];
    double cthr[16];       // depth-prior thresholds (staged once per kernel)
    int cstaged;           // split prior / column min-max are staged in the dynamic area
    int pass_bid;          // this block's index among the pass blocks (-1: serial-only block)
    int is_last;
    unsigned pass_gen;     // generation of the command being executed (P <= 32 kernels)
    int es_groups;         // early statistics pass: number of job groups (child segments) it found
    // ---- ring root pass (pg_pass_root_ring): bulk-copy ring in the staging area behind the dynamic arrays
    unsigned long long mbar_full[PGR_MAXST];   // per stage: the tile's bytes have landed
    unsigned r_done[PGR_MAXST];            // per stage: warps that have finished with the tile it holds (the last one refills it)
    int r_ucol[PGR_MAXU];                  // distinct columns staged per tile
    int r_jslot[32];                       // job -> staged column
    short r_add_slot[PG_TREE_SM], r_sub_slot[PG_TREE_SM];   // tree node -> staged column, -1 = gather from global
    int r_U;
    unsigned long long r_src[2 + PGR_MAXU];   // per staged array (A, y, columns...): global address of the array's row 0
    double r_acc[4][32][2];                // [row quarter][job]: sum A, sum y-A over LEFT rows
    unsigned r_cnt[4][32];
    unsigned p_scan[2][PG_WARPS];          // block-unit identity partition: warp totals of a chunk, double-buffered
    int p_rq_idx[64], p_rq_side[64], p_rq_var[64], p_nreq;   // look-ahead requests of the partition job in progress
    unsigned s_seg[32], s_len[32];         // block-unit statistics pass: the jobs' segments, staged with the job list
    // block-unit partition pass: every job's parameters and the look-ahead requests, staged in one round trip
    unsigned pp_src[32], pp_len[32], pp_left[32], pp_dst[32], pp_cntL[32], pp_exp[32];
    int pp_var[32], pp_mq_pj[64], pp_mq_side[64], pp_mq_var[64];
    float pp_thr[32];
    // followed by dynamic arrays, see pg_smem_* accessors
};

#define FW_MAXP 512   // split prior / column min-max staged in shared memory up to this many columns

// Dynamic shared memory of every kernel: [PgSmem][pass arrays sized by S][union area].  The union area holds, in the
// control block (block 0 of the P <= 32 persistent kernels), its kernel-lifetime state PgFastSm and the staged split prior
// / column min-max; in the pass blocks, which never touch that state, the same bytes are the staging ring of the root pass.
//   pass arrays: double jvL[S], jvR[S], jexp[2][S] (exp(vL), exp(vR): Bernoulli pass); double accA[PG_WARPS][S], accB[PG_WARPS][S]; unsigned accC[PG_WARPS][S];
//                unsigned mmin[2S], mmax[2S]; int jvar[S]; float jthr[S]; int jorder[S]; int grp[S+1]
__host__ __device__ inline size_t pg_smem_pass_bytes(int S) {
    size_t b = (sizeof(PgSmem) + 15) & ~(size_t)15;
    b += (size_t)S * 32;                       // jvL, jvR, jexp[2]
    b += (size_t)PG_WARPS * S * (8 + 8 + 4);   // accA, accB, accC
    b += (size_t)S * 16;                       // mmin, mmax (2S each)
    b += (size_t)S * (4 + 4 + 4) + (size_t)(S + 1) * 4;
    return ((b + 127) & ~(size_t)127) + 128;   // the union area starts 128-byte aligned behind this
}
__host__ __device__ inline size_t pg_smem_ctrl_bytes() {   // the control block's part of the union area
    return ((sizeof(PgFastSm) + 15) & ~(size_t)15) + (size_t)FW_MAXP * (8 + 4 + 4);
}
__host__ __device__ inline size_t pg_smem_bytes(int S) { return pg_smem_pass_bytes(S) + pg_smem_ctrl_bytes(); }   // without a ring

struct PgSmemView {
    PgSmem* base;
    int* jvar; float* jthr; int* jorder; int* grp; double* jvL; double* jvR; double* jexp;
    double* accA; double* accB; unsigned* accC; unsigned* mmin; unsigned* mmax;
    PgFastSm* fast;            // control block only
    double* cprior; float* cmin; float* cmax;   // control block only
    unsigned char* stage;      // pass blocks only: staging ring (st.stage_bytes), 128-byte aligned; aliases fast / cprior / ...
};

__device__ __forceinline__ PgSmemView pg_smem_view(unsigned char* raw, int S) {
    PgSmemView v;
    v.base = reinterpret_cast<PgSmem*>(raw);
    size_t off = (sizeof(PgSmem) + 15) & ~(size_t)15;
    v.jvL = reinterpret_cast<double*>(raw + off); off += (size_t)S * 8;
    v.jvR = reinterpret_cast<double*>(raw + off); off += (size_t)S * 8;
    v.jexp = reinterpret_cast<double*>(raw + off); off += (size_t)S * 16;
    v.accA = reinterpret_cast<double*>(raw + off); off += (size_t)PG_WARPS * S * 8;
    v.accB = reinterpret_cast<double*>(raw + off); off += (size_t)PG_WARPS * S * 8;
    v.accC = reinterpret_cast<unsigned*>(raw + off); off += (size_t)PG_WARPS * S * 4;
    v.mmin = reinterpret_cast<unsigned*>(raw + off); off += (size_t)S * 8;
    v.mmax = reinterpret_cast<unsigned*>(raw + off); off += (size_t)S * 8;
    v.jvar = reinterpret_cast<int*>(raw + off); off += (size_t)S * 4;
    v.jthr = reinterpret_cast<float*>(raw + off); off += (size_t)S * 4;
    v.jorder = reinterpret_cast<int*>(raw + off); off += (size_t)S * 4;
    v.grp = reinterpret_cast<int*>(raw + off); off += (size_t)(S + 1) * 4;
    v.stage = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw + off) + 127) & ~(uintptr_t)127);
    unsigned char* u = v.stage;
    v.fast = reinterpret_cast<PgFastSm*>(u); u += (sizeof(PgFastSm) + 15) & ~(size_t)15;
    v.cprior = reinterpret_cast<double*>(u); u += (size_t)FW_MAXP * 8;
    v.cmin = reinterpret_cast<float*>(u); u += (size_t)FW_MAXP * 4;
    v.cmax = reinterpret_cast<float*>(u);
    return v;
}

// Copy a 16-int pass command from global memory into this block's shared memory; the caller synchronises the block.
__device__ __forceinline__ void pg_fetch_cmd(const int* src, PgSmem* b) {
    if (threadIdx.x < 16) reinterpret_cast<int*>(&b->cmd)[threadIdx.x] = __ldcg(&src[threadIdx.x]);
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
    // the size and the first PG_TREE_SM nodes are requested together (the node arrays have maxn entries per tree, so the
    // addresses are valid whatever the size turns out to be): one memory round trip instead of two
    const int i = threadIdx.x;
    const bool in = i < PG_TREE_SM && i < st.maxn;
    const int sz = __ldcg(&st.f_size[t]);
    const int sv = in ? __ldcg(&st.f_split_var[(size_t)t * st.maxn + i]) : -1;
    const float th = in ? __ldcg(&st.f_thr32[(size_t)t * st.maxn + i]) : 0.0f;
    const double lv = in ? __ldcg(&st.f_leaf_val[(size_t)t * st.maxn + i]) : 0.0;
    if (i == 0) { sm.size = sz; sm.from_global = sz > PG_TREE_SM; }
    if (in && i < sz && sz <= PG_TREE_SM) { sm.split_var[i] = sv; sm.thr32[i] = th; sm.leaf_val[i] = lv; }
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
        if (with_values) {
            const double vL = __ldcg(&bf.job_vL[j]), vR = __ldcg(&bf.job_vR[j]);
            sm.jvL[j] = vL; sm.jvR[j] = vR;
            sm.jexp[2 * j + 0] = pg_bern_exp(vL); sm.jexp[2 * j + 1] = pg_bern_exp(vR);
        }
    }
    for (int g = threadIdx.x; g <= G; g += blockDim.x) sm.grp[g] = __ldcg(&bf.grp_first[g]);
}

__device__ __forceinline__ void pg_zero_acc(const PgSmemView& sm, int S) {
    for (int i = threadIdx.x; i < PG_WARPS * S; i += blockDim.x) { sm.accA[i] = 0.0; sm.accB[i] = 0.0; sm.accC[i] = 0u; }
}

// Write the block's per-job results: fixed warp order => deterministic.
__device__ __forceinline__ void pg_flush_acc(const PgState& st, const PgBuf& bf, const PgSmemView& sm, int J, bool counts, bool as_ll,
                                             bool warp_counts = true) {
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
            if (warp_counts)                                 // only the warp-range partitions of the generic kernels read these
                for (int w = 0; w < PG_WARPS; ++w)
                    bf.part_cnt[((size_t)PG_BID * PG_WARPS + w) * S + j] = sm.accC[w * S + j];
        }
    }
}

// one row of one job: if (x < thr) { a += o; b += r; c += 1; } (smc.rs:201-217) without predication: the 0/1 mask
// times a finite value is exact, so a = fma(m, o, a) is the same sum (ptxas turns predicated f64 adds into add + two
// selects, twice the instructions)
__device__ __forceinline__ void pg_acc_left(double& a, double& b, unsigned& c, float x, float thr, double o, double r) {
    const bool left = x < thr;
    const double m = left ? 1.0 : 0.0;
    a = fma(m, o, a);
    b = fma(m, r, b);
    c += left ? 1u : 0u;
}

// =============================================================================
// PH_ROOT / PH_FINAL
// =============================================================================
// Rows per lane per iteration: 8 (two float4 groups) => 256 rows per warp-iteration.
// BLOCKR: the block's rows are pg_block_range (the block-unit scheme of the P <= 32 persistent kernel), split among its
// warps; otherwise every global warp owns one pg_warp_range (generic control code, stepwise driver).
template <bool BERN, bool BLOCKR>
__device__ __forceinline__ void pg_pass_root(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    const PgCmd& cm = sm.base->cmd;
    const int J = cm.n_jobs * (cm.phase == PH_ROOT ? 1 : 0);
    const int t_add = cm.t_add, t_sub = cm.phase == PH_ROOT ? cm.t_sub : -1;
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
    if (BLOCKR) {
        long long B0, B1;
        pg_block_range(st.n, PG_BID, st.grid, &B0, &B1);
        pg_warp_range((unsigned long long)(B1 - B0), warp, PG_WARPS, 4, &beg, &end);
        beg += (unsigned long long)B0; end += (unsigned long long)B0;
    } else pg_warp_range((unsigned long long)st.n, gw, st.n_gwarps, 4, &beg, &end);
    const long long n = st.n;
    const size_t ld = (size_t)st.ld;
    double tot_o = 0.0, tot_r = 0.0, tot_rr = 0.0, tot_ll = 0.0, tot_old = 0.0;

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
        if (!BLOCKR && st.pg_correct && cm.phase == PH_ROOT) {
            // weight_mode pg_correct: o is the current sum of ALL trees here; its log-likelihood is the weight of the
            // conditional reference particle (the tree being replaced)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (!valid[q]) continue;
                if (BERN) tot_old += (double)yv[q] * o[q] - pg_softplus(o[q]);
                else { const double d = (double)yv[q] - o[q]; tot_old += d * d; }
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
                // rows outside this warp's range read as +inf, like the padding rows of X: never LEFT
                xa[q] = (live && g0) ? pg_ld_x4(col + r0) : make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
                xb[q] = (live && g1) ? pg_ld_x4(col + r1) : make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool live = jb + q < J;
                const float thr = live ? sm.jthr[jb + q] : -INFINITY;
                const float xs[8] = {xa[q].x, xa[q].y, xa[q].z, xa[q].w, xb[q].x, xb[q].y, xb[q].z, xb[q].w};
                double a = 0.0, b = 0.0;
                unsigned cc = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) pg_acc_left(a, b, cc, xs[e], thr, o[e], r[e]);   // o = r = 0 on rows >= n
                so[q] = a; sr[q] = b; cn[q] = (int)cc;
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
    if (!BLOCKR && st.pg_correct && cm.phase == PH_ROOT) {
        __syncthreads();
        tot_old = pg_warp_sum(tot_old);
        if (lane == 0) sm.base->tot[warp][0] = tot_old;
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int w = 0; w < PG_WARPS; ++w) v += sm.base->tot[w][0];
            st.part_old[PG_BID] = v;
        }
    }
    pg_flush_acc(st, bf, sm, J, true, false, !BLOCKR);
}

// =============================================================================
// PH_ROOT / PH_FINAL, ring variant (P <= 32 persistent kernels; root_pass = "ring")
// =============================================================================
// The direct pass is latency bound: a block owns only ~6.8 K rows at BASELINE C3, so each warp runs two iterations of
// five dependent load -> reduce phases and the pass takes 24 us where the bytes need 12.  Here the loads are decoupled
// from the arithmetic: the block's rows are cut into tiles of G row groups (128 rows each, G <= 8) and a ring of tiles
// is kept in flight with bulk asynchronous copies (cp.async.bulk -> UBLKCP, completion on an mbarrier, evict-first for
// X): per tile ONE copy per array — A (8 B/row), y, and every DISTINCT column the pass needs (the slots' split variables
// plus the split variables of the trees being committed / opened, so tree evaluation reads shared memory too) — 4 KB per
// column at G = 8, because the copy engine accepts one copy per ~90 cycles and 2 KB copies cannot feed an SM its share of
// HBM bandwidth (the first staged version, 512-row tiles, lost to the direct pass for that reason).
// Warp w = (row quarter w / 4, job group w % 4): the tile's row groups are dealt round-robin to the quarters (rotating
// with the tile index so that every quarter gets the odd group in turn); a lane owns 4 consecutive rows of a row group
// and accumulates its job group's jobs in registers across ALL tiles of the block, so the cross-lane reduction happens
// once per pass.  There is no producer warp and no block barrier in the loop.  Lane 0 of warp w issues the copies of
// arrays w, w + 16, w + 32 (measured, tools/microbench_bulk.cu: one thread sustains a bulk copy per ~300-570 cycles whatever
// its size, so an SM needs several issuing threads to reach its share of HBM bandwidth with copies of a few KB; 147 blocks
// x 16 issuers x 2 KB copies stream at 6.9 TB/s).  A warp that has finished tile t counts itself out of the tile's stage
// and then refills the stage of tile t - 1 — free once all 16 warps have counted out of it, which they normally have by
// then — with tile t - 1 + NST: warps may drift apart by a tile, and NST - 1 tiles are in flight ahead of the consumers.
// (A first version let the LAST warp out of a tile issue all of the next tile's copies: that warp fell 1.2 us behind per
// tile, became the last one out again, and the pass took 31 us against 24 for the direct pass.)
// Fixed summation order => bit-reproducible.  Emits per-BLOCK results only (part_tot, part_sum, part_ccnt).
__device__ __forceinline__ unsigned pg_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long pg_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 pg_lds_f4(unsigned addr) {     // explicit shared-space loads: the staging area is reached through a
    float4 v;                                                      // generic pointer, which would otherwise compile to generic LD
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float pg_lds_f1(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 pg_lds_d2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void pg_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pg_mbar_inval(unsigned long long* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(pg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pg_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {      // expect `bytes`, then arrive
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pg_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pg_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = pg_smem_u32(bar);
    unsigned done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");      // (a suspend-time hint made waits last the whole hint: 40 % slower)
        if (!done && ++spins > (1u << 24)) __trap();     // a lost copy becomes an error, not a hang
    } while (!done);
}
__device__ __forceinline__ void pg_bulk_g2s_u32(unsigned dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(pg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pg_bulk_g2s_hint_u32(unsigned dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(pg_smem_u32(bar)), "l"(pol) : "memory");
}

// tree value of a row whose columns are (mostly) staged: tile-local row `lr` (stride `T` rows per staged column), global row `row`
__device__ __forceinline__ double pg_tree_eval_staged(const PgState& st, int t, const PgTreeSm& tr, const short* slot,
                                                       unsigned cols_u32, int T, int lr, long long row) {
    if (tr.from_global) return pg_tree_eval(st, t, tr, row);
    int node = 0, sv;
    while ((sv = tr.split_var[node]) >= 0) {
        const int sl = slot[node];
        const float x = sl >= 0 ? pg_lds_f1(cols_u32 + (unsigned)(sl * T + lr) * 4u) : __ldg(&st.X[(size_t)sv * st.ld + row]);
        node = 2 * node + 1 + (x < tr.thr32[node] ? 0 : 1);
    }
    return tr.leaf_val[node];
}

// Returns false (block-uniformly, before touching anything) when the staging area cannot hold two tiles of one row group
// for this pass's distinct columns: the caller then runs the direct pass.
template <bool BERN, int JW>
__device__ __forceinline__ bool pg_pass_root_ring(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    PgSmem* b = sm.base;
    const PgCmd& cm = sm.base->cmd;
    const bool is_root = cm.phase == PH_ROOT;
    const int J = is_root ? cm.n_jobs : 0;
    const int t_add = cm.t_add, t_sub = is_root ? cm.t_sub : -1;
    const int S = st.S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qd = warp >> 2, grp = warp & 3;

    pg_stage_tree(st, t_add, b->tadd);
    pg_stage_tree(st, t_sub, b->tsub);
    for (int j = tid; j < J; j += blockDim.x) { sm.jvar[j] = __ldcg(&bf.job_var[j]); sm.jthr[j] = __ldcg(&bf.job_thr32[j]); }
    __syncthreads();
    // ---- the distinct columns of this pass (warp 0)
    if (warp == 0) {
        const bool live = lane < J;
        const int var = live ? sm.jvar[lane] : -1 - lane;
        const unsigned same = __match_any_sync(PG_FULL, var);
        const bool leader = live && (__ffs(same) - 1) == lane;
        const unsigned leadm = __ballot_sync(PG_FULL, leader);
        const int slot_l = __popc(leadm & ((1u << lane) - 1u));
        if (leader) b->r_ucol[slot_l] = var;
        const int slot = __shfl_sync(PG_FULL, slot_l, live ? (__ffs(same) - 1) : 0);
        if (live) b->r_jslot[lane] = slot;
        int U = __popc(leadm);
        const int ucap = min(PGR_MAXU, (st.stage_bytes / (2 * PGR_RG) - 12) / 4);   // two tiles of one row group must fit
        __syncwarp();
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const PgTreeSm& tr = which == 0 ? b->tadd : b->tsub;
            short* slots = which == 0 ? b->r_add_slot : b->r_sub_slot;
            const int sz = (which == 0 ? t_add : t_sub) >= 0 && !tr.from_global ? tr.size : 0;
#pragma unroll 1
            for (int node = 0; node < sz; ++node) {
                const int sv = tr.split_var[node];
                if (sv < 0) { if (lane == 0) slots[node] = -1; continue; }
                const bool hit0 = lane < U && b->r_ucol[lane] == sv;
                const bool hit1 = lane + 32 < U && b->r_ucol[lane + 32] == sv;
                const unsigned m0 = __ballot_sync(PG_FULL, hit0), m1 = __ballot_sync(PG_FULL, hit1);
                int sl = m0 ? (__ffs(m0) - 1) : (m1 ? 32 + (__ffs(m1) - 1) : -1);
                if (sl < 0 && U < ucap) { sl = U; if (lane == 0) b->r_ucol[U] = sv; U += 1; }
                if (lane == 0) slots[node] = (short)sl;
                __syncwarp();
            }
        }
        if (lane == 0) b->r_U = U <= ucap ? U : -1;          // more job columns than the staging area holds: direct pass
    }
    __syncthreads();
    const int U = b->r_U;
    if (U < 0) { __syncthreads(); return false; }
    const size_t ld = (size_t)st.ld;
    const int n_arr = 2 + U;
    // ---- ring geometry: G row groups per tile, NST stages.  G = 4 (one row group per row quarter per tile: the quarters stay
    // balanced) when three such stages fit, else 2 or 1; at least two stages or the direct pass takes over.
    const unsigned bpg = (unsigned)PGR_RG * (12u + 4u * (unsigned)U);            // bytes of one row group of every staged array
    int G = 4;
    while (G > 1 && (unsigned)st.stage_bytes < 3u * (unsigned)G * bpg) G >>= 1;
    if (st.root_tile_groups > 0) G = min(st.root_tile_groups, PGR_MAXG);
    while (G > 1 && (unsigned)st.stage_bytes < 2u * (unsigned)G * bpg) G >>= 1;
    const int T = G * PGR_RG;
    const unsigned stage_bytes = (unsigned)G * bpg;
    int NST = (int)((unsigned)st.stage_bytes / stage_bytes);
    if (NST > PGR_MAXST) NST = PGR_MAXST;
    if (st.root_stages >= 2 && NST > st.root_stages) NST = st.root_stages;
    if (tid < PGR_MAXST) { pg_mbar_init(&b->mbar_full[tid], PG_WARPS); b->r_done[tid] = 0u; }
    for (int a = tid; a < n_arr; a += blockDim.x)
        b->r_src[a] = a == 0 ? (unsigned long long)bf.A_src : a == 1 ? (unsigned long long)st.y
                                                                    : (unsigned long long)(st.X + (size_t)b->r_ucol[a - 2] * ld);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const bool add_nontrivial = t_add >= 0 && b->tadd.size > 1;
    const bool sub_nontrivial = t_sub >= 0 && b->tsub.size > 1;
    const double add_c = t_add == -2 ? st.init_leaf : ((t_add >= 0 && !add_nontrivial) ? b->tadd.leaf_val[0] : 0.0);
    const double sub_c = (t_sub >= 0 && !sub_nontrivial) ? b->tsub.leaf_val[0] : 0.0;
    const bool do_add = t_add >= 0 || t_add == -2, do_sub = t_sub >= 0;
    const bool commit_only = J == 0 && t_sub < 0;      // PH_FINAL

    long long B0, B1;
    pg_block_range(st.n, PG_BID, st.grid, &B0, &B1);
    const int n_tiles = (int)((B1 - B0 + T - 1) / T);
    const unsigned long long pol = pg_policy_evict_first();
    const unsigned stage0 = pg_smem_u32(sm.stage);
    // One tile = one bulk copy per staged array: a = 0: A (8 B per row), 1: y, 2 + u: column u.  Lane 0 of warp w issues
    // arrays w, w + 16, w + 32 and arrives once per tile on the stage's "landed" barrier with the bytes it asked for.
    auto issue = [&](int t) {                                // lane 0 of every warp
        const int s_ = t % NST;
        const long long r0 = B0 + (long long)t * T;
        const unsigned rows = (unsigned)min((long long)T, B1 - r0);
        const unsigned sbase = stage0 + (unsigned)s_ * stage_bytes;
        unsigned bytes = 0;
#pragma unroll 1
        for (int a = warp; a < n_arr; a += PG_WARPS) bytes += rows * (a == 0 ? 8u : 4u);
        if (bytes) pg_mbar_expect_tx(&b->mbar_full[s_], bytes); else pg_mbar_arrive(&b->mbar_full[s_]);
#pragma unroll 1
        for (int a = warp; a < n_arr; a += PG_WARPS) {
            const unsigned es = a == 0 ? 8u : 4u;
            const unsigned dst = sbase + (a == 0 ? 0u : a == 1 ? (unsigned)T * 8u : (unsigned)T * 12u + (unsigned)(a - 2) * ((unsigned)T * 4u));
            const char* src = reinterpret_cast<const char*>(b->r_src[a]) + r0 * es;
            if (a < 2) pg_bulk_g2s_u32(dst, src, rows * es, &b->mbar_full[s_]);
            else pg_bulk_g2s_hint_u32(dst, src, rows * es, &b->mbar_full[s_], pol);
        }
    };
    if (lane == 0) {
        // generic-proxy writes this thread has observed (A, written by the previous pass, through the command's release /
        // acquire chain) are ordered before its bulk copies
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int t = 0; t < NST && t < n_tiles; ++t) issue(t);
    }
    __syncwarp();

    // my jobs: group g takes jobs j with (j + 1) % 4 == g, so group 0 (which also writes A and keeps the totals) gets the fewest
    const int jfirst = (grp + 3) & 3;
    const int myJ = J > jfirst ? (J - jfirst + 3) / 4 : 0;
    unsigned coff[JW];
    float jthr[JW];
    double accA[JW], accB[JW];
    unsigned accC[JW];
#pragma unroll
    for (int q = 0; q < JW; ++q) {
        const int j = jfirst + 4 * q;
        const bool lv = j < J;
        coff[q] = lv ? (unsigned)b->r_jslot[j] * ((unsigned)T * 4u) : 0u;
        jthr[q] = lv ? sm.jthr[j] : 0.0f;
        accA[q] = 0.0; accB[q] = 0.0; accC[q] = 0u;
    }
    double tot_o = 0.0, tot_r = 0.0, tot_rr = 0.0, tot_ll = 0.0;
    int stg = 0;

#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
        pg_mbar_wait(&b->mbar_full[stg], (unsigned)((t / NST) & 1));
        const unsigned sbase = stage0 + (unsigned)stg * stage_bytes;
        const long long r0 = B0 + (long long)t * T;
        const int rows = (int)min((long long)T, B1 - r0);
        const unsigned cols = sbase + (unsigned)T * 12u;          // shared-space address of the staged columns
        // this warp's row groups of the tile: g = (qd + t) mod 4, + 4, ...  (the odd group of a tile visits every quarter in turn)
#pragma unroll 1
        for (int g = (qd + t) & 3; g < G; g += 4) {
            const int lr = g * PGR_RG + lane * 4;               // tile-local first row of this lane
            if (lr >= rows) continue;                           // (a short last tile leaves some lanes without rows)
            if (st.root_dbg & 4) continue;
            const long long row = r0 + lr;
            const int nv = (int)min(4ll, st.n - row);           // valid rows of this lane's four (>= 1 except past n)
            double o0, o1, o2, o3;
            float4 yy;
            {
                const double2 a01 = pg_lds_d2(sbase + (unsigned)lr * 8u);
                const double2 a23 = pg_lds_d2(sbase + (unsigned)lr * 8u + 16u);
                yy = pg_lds_f4(sbase + (unsigned)T * 8u + (unsigned)lr * 4u);
                o0 = a01.x; o1 = a01.y; o2 = a23.x; o3 = a23.y;
            }
            // kernel.rs:99-102 then :84-87, each a separately rounded f64 operation
            if (do_add) {
                if (add_nontrivial) {
                    o0 = o0 + (nv > 0 ? pg_tree_eval_staged(st, t_add, b->tadd, b->r_add_slot, cols, T, lr + 0, row + 0) : 0.0);
                    o1 = o1 + (nv > 1 ? pg_tree_eval_staged(st, t_add, b->tadd, b->r_add_slot, cols, T, lr + 1, row + 1) : 0.0);
                    o2 = o2 + (nv > 2 ? pg_tree_eval_staged(st, t_add, b->tadd, b->r_add_slot, cols, T, lr + 2, row + 2) : 0.0);
                    o3 = o3 + (nv > 3 ? pg_tree_eval_staged(st, t_add, b->tadd, b->r_add_slot, cols, T, lr + 3, row + 3) : 0.0);
                } else { o0 = o0 + add_c; o1 = o1 + add_c; o2 = o2 + add_c; o3 = o3 + add_c; }
            }
            if (do_sub) {
                if (sub_nontrivial) {
                    o0 = o0 - (nv > 0 ? pg_tree_eval_staged(st, t_sub, b->tsub, b->r_sub_slot, cols, T, lr + 0, row + 0) : 0.0);
                    o1 = o1 - (nv > 1 ? pg_tree_eval_staged(st, t_sub, b->tsub, b->r_sub_slot, cols, T, lr + 1, row + 1) : 0.0);
                    o2 = o2 - (nv > 2 ? pg_tree_eval_staged(st, t_sub, b->tsub, b->r_sub_slot, cols, T, lr + 2, row + 2) : 0.0);
                    o3 = o3 - (nv > 3 ? pg_tree_eval_staged(st, t_sub, b->tsub, b->r_sub_slot, cols, T, lr + 3, row + 3) : 0.0);
                } else { o0 = o0 - sub_c; o1 = o1 - sub_c; o2 = o2 - sub_c; o3 = o3 - sub_c; }
            }
            if (grp == 0) {
                __stcg(reinterpret_cast<double2*>(bf.A_dst + row), make_double2(o0, o1));
                __stcg(reinterpret_cast<double2*>(bf.A_dst + row + 2), make_double2(o2, o3));
            }
            if (commit_only || (st.root_dbg & 2)) continue;
            const double r0v = (double)yy.x - o0, r1v = (double)yy.y - o1, r2v = (double)yy.z - o2, r3v = (double)yy.w - o3;
            if (grp == 0) {
                if (nv >= 4) {
                    tot_o += o0; tot_o += o1; tot_o += o2; tot_o += o3;
                    tot_r += r0v; tot_r += r1v; tot_r += r2v; tot_r += r3v;
                    tot_rr += r0v * r0v; tot_rr += r1v * r1v; tot_rr += r2v * r2v; tot_rr += r3v * r3v;
                    if (BERN) {
                        const double m0 = o0 + st.init_leaf, m1 = o1 + st.init_leaf, m2 = o2 + st.init_leaf, m3 = o3 + st.init_leaf;
                        tot_ll += (double)yy.x * m0 - pg_softplus(m0); tot_ll += (double)yy.y * m1 - pg_softplus(m1);
                        tot_ll += (double)yy.z * m2 - pg_softplus(m2); tot_ll += (double)yy.w * m3 - pg_softplus(m3);
                    }
                } else {                                        // the lane that straddles n: rows row .. row + nv - 1 are real
                    const double ov[4] = {o0, o1, o2, o3}, rv[4] = {r0v, r1v, r2v, r3v};
                    const float yv[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll 1
                    for (int e = 0; e < nv; ++e) {
                        tot_o += ov[e]; tot_r += rv[e]; tot_rr += rv[e] * rv[e];
                        if (BERN) { const double mu = ov[e] + st.init_leaf; tot_ll += (double)yv[e] * mu - pg_softplus(mu); }
                    }
                }
            }
            // smc.rs:201-217 for this warp's slots of round 1 (padding rows hold x = +inf: never LEFT).  Dispatch on the
            // warp-uniform job count so that the unrolled bodies carry no per-job branches.
            const unsigned xb = cols + (unsigned)lr * 4u;
            auto jobs = [&](auto njc) {
                constexpr int NJ = decltype(njc)::value;
#pragma unroll
                for (int q = 0; q < NJ; ++q) {
                    const float4 xv = pg_lds_f4(xb + coff[q]);
                    pg_acc_left(accA[q], accB[q], accC[q], xv.x, jthr[q], o0, r0v);
                    pg_acc_left(accA[q], accB[q], accC[q], xv.y, jthr[q], o1, r1v);
                    pg_acc_left(accA[q], accB[q], accC[q], xv.z, jthr[q], o2, r2v);
                    pg_acc_left(accA[q], accB[q], accC[q], xv.w, jthr[q], o3, r3v);
                }
            };
            switch (myJ) {
                case 0: break;
                case 1: jobs(std::integral_constant<int, 1>{}); break;
                case 2: jobs(std::integral_constant<int, 2>{}); break;
                case 3: jobs(std::integral_constant<int, 3>{}); break;
                case 4: jobs(std::integral_constant<int, 4>{}); break;
                case 5: jobs(std::integral_constant<int, 5>{}); break;
                default:
                    if constexpr (JW > 5) {
                        if (myJ == 6) jobs(std::integral_constant<int, 6>{});
                        else if (myJ == 7) jobs(std::integral_constant<int, 7>{});
                        else jobs(std::integral_constant<int, JW>{});
                    }
                    break;
            }
        }
        // done with this tile: count this warp out of its stage, then refill the stage tile t - 1 occupied (free once all
        // warps have counted out of it) with this warp's arrays of tile t - 1 + NST
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();                              // this warp's reads of the stage are done before the count moves
            atomicAdd(&b->r_done[stg], 1u);
            if (t >= 1 && t - 1 + NST < n_tiles) {
                const int sp = stg == 0 ? NST - 1 : stg - 1;
                const unsigned want = (unsigned)PG_WARPS * (unsigned)((t - 1) / NST + 1);
                unsigned spins = 0;
                while (*(volatile unsigned*)&b->r_done[sp] < want) { if (++spins > (1u << 26)) __trap(); }
                if (!(st.root_dbg & 1)) asm volatile("fence.proxy.async;" ::: "memory");   // the other warps' reads of that stage precede the copies into it
                issue(t - 1 + NST);
            }
        }
        __syncwarp();
        stg = stg + 1 == NST ? 0 : stg + 1;
    }

    // ---- block results
    if (!commit_only) {
#pragma unroll
        for (int q = 0; q < JW; ++q) {
            if (q < myJ) {                                  // warp-uniform
                const double a = pg_warp_sum(accA[q]), bb = pg_warp_sum(accB[q]);
                const unsigned cc = __reduce_add_sync(PG_FULL, accC[q]);
                if (lane == 0) { const int j = jfirst + 4 * q; b->r_acc[qd][j][0] = a; b->r_acc[qd][j][1] = bb; b->r_cnt[qd][j] = cc; }
            }
        }
        if (grp == 0) {
            tot_o = pg_warp_sum(tot_o); tot_r = pg_warp_sum(tot_r); tot_rr = pg_warp_sum(tot_rr); tot_ll = pg_warp_sum(tot_ll);
            if (lane == 0) { b->tot[qd][0] = tot_o; b->tot[qd][1] = tot_r; b->tot[qd][2] = tot_rr; b->tot[qd][3] = tot_ll; }
        }
    }
    __syncthreads();
    if (tid < PGR_MAXST) pg_mbar_inval(&b->mbar_full[tid]);
    if (commit_only) return true;
    if (tid < 4) {
        double v = 0.0;
        for (int q = 0; q < 4; ++q) v += b->tot[q][tid];
        bf.part_tot[(size_t)PG_BID * 4 + tid] = v;
    }
    for (int j = tid; j < J; j += blockDim.x) {
        double a = 0.0, bb = 0.0;
        unsigned cc = 0u;
        for (int q = 0; q < 4; ++q) { a += b->r_acc[q][j][0]; bb += b->r_acc[q][j][1]; cc += b->r_cnt[q][j]; }
        bf.part_sum[((size_t)PG_BID * S + j) * 2 + 0] = a;
        bf.part_sum[((size_t)PG_BID * S + j) * 2 + 1] = bb;
        bf.part_ccnt[(size_t)PG_BID * S + j] = cc;
    }
    return true;
}

// =============================================================================
// PH_MINMAX — splitting.rs:43-49 over index segments
// =============================================================================
template <bool ONEHOT>   // ONEHOT: the generic kernels, where OneHotSplit columns ask for a presence mask instead of min / max
__device__ __forceinline__ void pg_pass_minmax(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    const int J = sm.base->cmd.n_jobs;
    const int S = st.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = PG_BID * PG_WARPS + warp;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        const int var = __ldcg(&bf.job_var[j]);
        sm.jvar[j] = var;
        const bool oh = ONEHOT && st.onehot && st.onehot[var];
        sm.mmin[j] = oh ? 0u : 0xFFFFFFFFu; sm.mmax[j] = 0u;
    }
    __syncthreads();
    const size_t ld = (size_t)st.ld;
    for (int j = 0; j < J; ++j) {
        const unsigned off = __ldcg(&bf.job_seg_off[j]);
        const unsigned long long L = __ldcg(&bf.job_len[j]);
        const float* col = st.X + (size_t)sm.jvar[j] * ld;
        unsigned long long beg, end;
        pg_warp_range(L, gw, st.n_gwarps, 1, &beg, &end);
        if (ONEHOT && st.onehot && st.onehot[sm.jvar[j]]) {
            // OneHotSplit column (splitting.rs:79-90): which integer codes occur in the node — a 64-bit presence mask,
            // its low / high words in the slots the min / max keys use
            const int imin = st.col_imin[sm.jvar[j]];
            unsigned lo = 0u, hi = 0u;
            for (unsigned long long e = beg + lane; e < end; e += 32) {
                const unsigned idx = __ldcg(&st.pool[(size_t)off + e]);
                const int b = (int)__ldg(&col[idx]) - imin;      // `as i32` truncates toward zero, like the conversion
                if (b < 32) lo |= 1u << b; else hi |= 1u << (b - 32);
            }
            lo = __reduce_or_sync(PG_FULL, lo);
            hi = __reduce_or_sync(PG_FULL, hi);
            if (lane == 0) { atomicOr(&sm.mmin[j], lo); atomicOr(&sm.mmax[j], hi); }
            continue;
        }
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
// PH_STATS — over index segments (gather); jobs sharing a segment share idx/A/y loads
// =============================================================================
// LEFT count, sum A, sum y-A per job; identity segments allowed.
__device__ __forceinline__ void pg_pass_seg(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    const int J = sm.base->cmd.n_jobs, G = sm.base->cmd.n_groups;
    const int S = st.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = PG_BID * PG_WARPS + warp;
    pg_stage_jobs(st, bf, sm, J, G, false);
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
                r[k] = valid[k] ? (double)__ldg(&st.y[idx[k]]) - o[k] : 0.0;
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
                        const double m = left ? 1.0 : 0.0;      // exact 0/1 mask (see pg_acc_left)
                        a = fma(m, o[k], a);
                        b = fma(m, r[k], b);
                        cc += left ? 1 : 0;
                    }
                    va[q] = a; vb[q] = b; cn[q] = cc;
                }
                const double ta = pg_xreduce4(va[0], va[1], va[2], va[3], lane);
                const double tb = pg_xreduce4(vb[0], vb[1], vb[2], vb[3], lane);
                const int tc = pg_xreduce4(cn[0], cn[1], cn[2], cn[3], lane);
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
    pg_flush_acc(st, bf, sm, J, true, false);
}

// =============================================================================
// PH_BERN — Bernoulli-logit child log-likelihoods over index segments (weight.rs:26-30 via smc.rs:89-95)
// =============================================================================
// Per job: A = sum over the LEFT rows of y mu - softplus(mu) with mu = o + vL, B = the same over the RIGHT rows with vR.
// This pass is arithmetic bound (BASELINE C5: 255 jobs x 1M rows = 2.6e8 softplus evaluations per tree update), so the
// row-job is made as short as the formula allows: exp(o) once per row, exp(v) once per job and child, softplus(mu) =
// log(1 + exp(o) exp(v)) through the table logarithm of pg_serial.h read from shared memory without bank conflicts — 17
// f64 operations, no exp, no division — and kept free of branches inside the 4 rows x 4 jobs block so that sixteen
// independent dependency chains are in flight per lane: whether any row of the chunk or the job's leaf values are out
// of the product form's range is decided once per chunk / job, warp-uniformly, and such chunks take the direct formula.
__device__ __forceinline__ double pg_log_ge1_sm(double u, unsigned tab_addr) {       // pg_log_ge1 with the table in shared memory
    const int hi = __double2hiint(u);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(u));
    const double2 t = pg_lds_d2(tab_addr + (unsigned)(((hi >> 14) & (PG_LOGT_K - 1)) * (PG_LOGT_REP * 16)));
    const double e = (double)((hi >> 20) - 1023);
    const double r = fma(m, t.x, -1.0);
    double p = 1.0 / 7.0;
    p = fma(p, r, -1.0 / 6.0); p = fma(p, r, 0.2); p = fma(p, r, -0.25); p = fma(p, r, 1.0 / 3.0); p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    return fma(p, r, fma(e, 0.6931471805599453, t.y));
}

// One block of NQ jobs x 4 rows per lane.  FULL: all 128 rows of the chunk exist and all NQ = 4 jobs are live, so the hot
// path carries no validity logic at all (the rows past a warp's range and the last, partial job group take FULL = false).
template <bool FULL>
__device__ __forceinline__ void pg_bern_block(const PgState& st, const PgSmemView& sm, const unsigned (&idx)[4], const double (&o)[4],
                                              const double (&yd)[4], const double (&eo)[4], unsigned vmask, bool rows_wide,
                                              int jb, int jb1, int j0, unsigned tab_addr, int lane, int warp) {
    constexpr int R = 4;
    const size_t ld = (size_t)st.ld;
    const int S = st.S;
    int jq[4];
    float x[4][R];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const bool live = FULL || jb + q < jb1;
        jq[q] = live ? sm.jorder[jb + q] : j0;
        const float* col = st.X + (size_t)sm.jvar[jq[q]] * ld;
#pragma unroll
        for (int k = 0; k < R; ++k) x[q][k] = __ldg(&col[idx[k]]);      // rows past the end read row 0 (idx = 0) and are masked below
    }
    double va[4], vb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = jq[q];
        const bool live = FULL || jb + q < jb1;
        const float thr = sm.jthr[j];
        const double vL = sm.jvL[j], vR = sm.jvR[j], evL = sm.jexp[2 * j], evR = sm.jexp[2 * j + 1];
        double a = 0.0, b = 0.0;
        if (!(rows_wide || !(evL < INFINITY) || !(evR < INFINITY))) {          // warp-uniform
            double sall = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const bool left = x[q][k] < thr;                                   // a NaN in X goes RIGHT, as in the reference
                const double mu = o[k] + (left ? vL : vR);
                double f = yd[k] * mu - pg_log_ge1_sm(fma(eo[k], left ? evL : evR, 1.0), tab_addr);
                if (!FULL) f = (live && ((vmask >> k) & 1u)) ? f : 0.0;
                a = fma(left ? 1.0 : 0.0, f, a);
                sall += f;
            }
            b = sall - a;      // four terms per lane: the subtraction costs an ulp of a number of order 1
        } else {
#pragma unroll 1
            for (int k = 0; k < R; ++k) {
                if (!(live && ((vmask >> k) & 1u))) continue;
                const bool left = x[q][k] < thr;
                const double mu = o[k] + (left ? vL : vR);
                const double f = yd[k] * mu - pg_softplus(mu);
                if (left) a += f; else b += f;
            }
        }
        va[q] = a; vb[q] = b;
    }
    const double ta = pg_xreduce4(va[0], va[1], va[2], va[3], lane);
    const double tb = pg_xreduce4(vb[0], vb[1], vb[2], vb[3], lane);
    const int q = pg_xreduce4_job(lane);
    if (lane < 4 && (FULL || jb + q < jb1)) {
        const int j = sm.jorder[jb + q];
        sm.accA[warp * S + j] += ta;
        sm.accB[warp * S + j] += tb;
    }
}

__device__ __forceinline__ void pg_pass_bern(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    const int J = sm.base->cmd.n_jobs, G = sm.base->cmd.n_groups;
    const int S = st.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pg_stage_jobs(st, bf, sm, J, G, true);
    pg_zero_acc(sm, S);
    // the logarithm table, PG_LOGT_REP copies, in the union area (idle outside the root pass)
    double* ltab = reinterpret_cast<double*>(sm.stage);
    for (int i = threadIdx.x; i < PG_LOGT_K * PG_LOGT_REP; i += blockDim.x) {
        const int k = i / PG_LOGT_REP;
        ltab[2 * i + 0] = __ldg(&st.log_tbl[2 * k + 0]);
        ltab[2 * i + 1] = __ldg(&st.log_tbl[2 * k + 1]);
    }
    __syncthreads();
    const unsigned tab_addr = pg_smem_u32(ltab) + (unsigned)((lane & (PG_LOGT_REP - 1)) * 16);
    constexpr int R = 4;   // entries per lane per chunk
    // Work units: (chunk of 128 segment entries) x (slice of the group's jobs), dealt round-robin to the block's warps.
    // A block takes whole chunks (ceil(chunks / grid) of them), so only the segment's last chunk is partial, and slicing
    // the jobs keeps the warps level when a block has few chunks per warp (BASELINE C5: 54 chunks x 8 slices of 32 jobs =
    // 27 units per warp; one row range per warp would leave each warp a quarter-filled fourth chunk).
    for (int g = 0; g < G; ++g) {
        const int jb0 = sm.grp[g], jb1 = sm.grp[g + 1];
        const int j0 = sm.jorder[jb0];
        const unsigned off = __ldcg(&bf.job_seg_off[j0]);
        const bool ident = off == PG_SEG_IDENTITY;
        const unsigned long long L = __ldcg(&bf.job_len[j0]);
        const unsigned long long chunks = (L + 32ull * R - 1) / (32ull * R);
        const unsigned long long cper = (chunks + (unsigned long long)st.grid - 1) / (unsigned long long)st.grid;
        unsigned long long c0 = (unsigned long long)PG_BID * cper, c1 = c0 + cper;
        if (c0 > chunks) c0 = chunks;
        if (c1 > chunks) c1 = chunks;
        const int slice = (jb1 - jb0 >= 64) ? 32 : 8;                       // jobs per slice, a multiple of 4
        const int nsl = (jb1 - jb0 + slice - 1) / slice;
        const unsigned long long units = (c1 - c0) * (unsigned long long)nsl;
        for (unsigned long long u = (unsigned long long)warp; u < units; u += PG_WARPS) {
            const unsigned long long base = (c0 + u / (unsigned long long)nsl) * (32ull * R);
            const int sj0 = jb0 + (int)(u % (unsigned long long)nsl) * slice;
            const int sj1 = sj0 + slice < jb1 ? sj0 + slice : jb1;
            const unsigned rem = (unsigned)(L - base < 32ull * R ? L - base : 32ull * R);
            unsigned idx[R];
            double o[R], yd[R], eo[R];
            unsigned vmask = 0u;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const unsigned e = (unsigned)lane + 32u * k;
                const bool valid = e < rem;
                vmask |= valid ? (1u << k) : 0u;
                idx[k] = valid ? (ident ? (unsigned)(base + e) : __ldcg(&st.pool[(size_t)off + base + e])) : 0u;
            }
            bool rows_wide = false;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                o[k] = __ldcg(&bf.A_src[idx[k]]);
                yd[k] = (double)__ldg(&st.y[idx[k]]);
                eo[k] = pg_bern_exp(o[k]);
                rows_wide |= !(eo[k] < INFINITY);
            }
            rows_wide = __any_sync(PG_FULL, rows_wide);
            int jb = sj0;
            if (rem == 32u * R)
                for (; jb + 4 <= sj1; jb += 4) pg_bern_block<true>(st, sm, idx, o, yd, eo, vmask, rows_wide, jb, sj1, j0, tab_addr, lane, warp);
            for (; jb < sj1; jb += 4) pg_bern_block<false>(st, sm, idx, o, yd, eo, vmask, rows_wide, jb, sj1, j0, tab_addr, lane, warp);
        }
    }
    __syncthreads();
    pg_flush_acc(st, bf, sm, J, false, true);
}

// =============================================================================
// PH_PART — stable partition of a segment by x[var] < split into the pool
// =============================================================================
// =============================================================================
// PH_STATS, block-unit variant (P <= 32 persistent kernels, Gaussian families)
// =============================================================================
// Same sums as pg_pass_seg<0>, but a block takes its contiguous entries of the segment (pg_seg_block_range) 8 per
// thread at once: 8 index loads in flight, then the gathers of A, y and up to four jobs' columns for all 8 entries
// together — three memory round trips per 4096 entries instead of three per 128 per warp.  Emits per-block results.
__device__ __forceinline__ void pg_pass_seg_block(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    const int J = sm.base->cmd.n_jobs;
    int G = sm.base->cmd.n_groups;
    const int S = st.S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (sm.base->cmd.flags & 8) {
        // ---- early statistics pass (queued behind the partition pass; pg_serial_fast.cuh, fw_r0_finish): job q = look-ahead
        // request q of that pass.  Nobody has digested its results yet, so (i) wait until every pass block has finished it (its
        // completion counter: the index lists and the min / max partials come from all of them), (ii) reduce the requests'
        // min / max here — one warp per request, lanes over the blocks — and (iii) derive the split value from the recorded
        // draw with the arithmetic of fw_decide / pg_split_draw.  A request without a split (min >= max, or a draw that
        // would have to be repeated) keeps no job: its threshold is reported as NaN and the control block falls back.
        if (tid == 0) {
            const unsigned gp = sm.base->pass_gen - 1u;
            pg_spin_until(&st.ctrl->cnt[gp & 1u][0], ((gp + 1u) >> 1) * (gridDim.x - 1u), st.ctrl);
        }
        __syncthreads();
        const unsigned dst0 = __ldcg(&st.pj_dst_off[0]), cntL0 = __ldcg(&st.pj_cntL[0]), len0 = __ldcg(&st.pj_len[0]);
        int n_left = 0;
        for (int q = warp; q < J; q += PG_WARPS) {
            unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
            for (int b = lane; b < st.grid; b += 32) {
                const uint2 v = __ldcg(reinterpret_cast<const uint2*>(st.part_mm + ((size_t)b * 2 * S + q) * 2));
                kmin = min(kmin, v.x); kmax = max(kmax, v.y);
            }
            kmin = __reduce_min_sync(PG_FULL, kmin);
            kmax = __reduce_max_sync(PG_FULL, kmax);
            if (lane == 0) {
                const float lo = pg_key2f(kmin), hi = pg_key2f(kmax);
                const int side = __ldcg(&st.mq_side[q]);
                float thr = -INFINITY;
                bool ok = lo < hi;
                if (ok) {
                    const double split = __dadd_rn(__dmul_rn(__ldcg(&st.mq_u3[q]), (double)hi - (double)lo), (double)lo);
                    ok = split < (double)hi;
                    if (ok) thr = pg_thr32(split);
                }
                sm.jvar[q] = __ldcg(&st.mq_var[q]);
                sm.jthr[q] = thr;                       // -inf: no row is LEFT, the job's sums stay zero
                sm.jorder[q] = q;
                sm.base->s_seg[q] = dst0 + (side ? cntL0 : 0u);
                sm.base->s_len[q] = side ? len0 - cntL0 : cntL0;
                if (PG_BID == 0) st.es_thr[q] = ok ? thr : __int_as_float(0x7fc00000);
            }
        }
        __syncthreads();
        // requests come ordered: the left child's (round 1) first, then the right child's (round 2): at most two groups
        if (tid == 0) {
            for (int q = 0; q < J; ++q) n_left += sm.base->s_seg[q] == dst0 ? 1 : 0;
            int g = 0;
            sm.grp[0] = 0;
            if (n_left > 0) sm.grp[++g] = n_left;
            if (J - n_left > 0) sm.grp[++g] = J;
            sm.base->es_groups = g;
        }
        pg_zero_acc(sm, S);
        __syncthreads();
    } else {
    pg_stage_jobs(st, bf, sm, J, G, false);
    for (int j = tid; j < J && j < 32; j += blockDim.x) { sm.base->s_seg[j] = __ldcg(&bf.job_seg_off[j]); sm.base->s_len[j] = __ldcg(&bf.job_len[j]); }
    pg_zero_acc(sm, S);
    __syncthreads();
    }
    if (sm.base->cmd.flags & 8) G = sm.base->es_groups;
    const size_t ld = (size_t)st.ld;
    constexpr int E = 8;
    for (int g = 0; g < G; ++g) {
        const int jb0 = sm.grp[g], jb1 = sm.grp[g + 1];
        const int j0 = sm.jorder[jb0];
        const unsigned off = j0 < 32 ? sm.base->s_seg[j0] : __ldcg(&bf.job_seg_off[j0]);
        const bool ident = off == PG_SEG_IDENTITY;
        const unsigned long long L = j0 < 32 ? sm.base->s_len[j0] : __ldcg(&bf.job_len[j0]);
        unsigned long long S0, S1;
        pg_seg_block_range(L, PG_BID, st.n_gwarps, PG_WARPS, &S0, &S1);
#pragma unroll 1
        for (unsigned long long cb = S0; cb < S1; cb += (unsigned long long)E * PG_THREADS) {
            unsigned idx[E];
            bool valid[E];
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const unsigned long long e = cb + (unsigned long long)tid + (unsigned long long)k * PG_THREADS;
                valid[k] = e < S1;
                idx[k] = valid[k] ? (ident ? (unsigned)e : __ldcg(&st.pool[(size_t)off + e])) : 0u;
            }
            double o[E], r[E];
#pragma unroll
            for (int k = 0; k < E; ++k) {
                o[k] = valid[k] ? __ldcg(&bf.A_src[idx[k]]) : 0.0;
                r[k] = valid[k] ? (double)__ldg(&st.y[idx[k]]) : 0.0;
            }
            float x0[E][4];                                  // the first four jobs' columns ride with the A / y gathers
            auto gather = [&](int jb) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool live = jb + q < jb1;
                    const float* col = st.X + (size_t)sm.jvar[live ? sm.jorder[jb + q] : j0] * ld;
#pragma unroll
                    for (int k = 0; k < E; ++k) x0[k][q] = (live && valid[k]) ? __ldg(&col[idx[k]]) : INFINITY;
                }
            };
            gather(jb0);
#pragma unroll
            for (int k = 0; k < E; ++k) r[k] = valid[k] ? r[k] - o[k] : 0.0;
#pragma unroll 1
            for (int jb = jb0; jb < jb1; jb += 4) {
                if (jb != jb0) gather(jb);
                double va[4], vb[4];
                int cn[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool live = jb + q < jb1;
                    const float thr = live ? sm.jthr[sm.jorder[jb + q]] : -INFINITY;
                    double a = 0.0, b2 = 0.0;
                    unsigned cc = 0u;
#pragma unroll
                    for (int k = 0; k < E; ++k) pg_acc_left(a, b2, cc, x0[k][q], thr, o[k], r[k]);
                    va[q] = a; vb[q] = b2; cn[q] = (int)cc;
                }
                const double ta = pg_xreduce4(va[0], va[1], va[2], va[3], lane);
                const double tb = pg_xreduce4(vb[0], vb[1], vb[2], vb[3], lane);
                const int tc = pg_xreduce4(cn[0], cn[1], cn[2], cn[3], lane);
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
    pg_flush_acc(st, bf, sm, J, true, false, false);
}

// Identity segment (the root's rows), block-unit scheme of the P <= 32 persistent kernel: block b compacts its own
// contiguous rows (pg_block_range, the same ranges as the staged root pass, whose per-block LEFT counts give the
// block's output offsets).  Rows are taken 4 per thread with coalesced float4 loads, 4 chunks in flight; order is
// kept by a warp scan + a 16-entry cross-warp prefix per chunk.  Look-ahead min/max requests stream the two columns
// the same way (no gathers: the segment is the identity).
#define PGP_CH (PG_THREADS * 4)      // rows per chunk
__device__ __forceinline__ void pg_part_ident(const PgState& st, const PgBuf& bf, const PgSmemView& sm, int k, int Q,
                              const float* col, float thr, unsigned long long L) {
    PgCtrl* c = st.ctrl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const size_t ld = (size_t)st.ld;
    long long B0, B1;
    pg_block_range((long long)L, PG_BID, st.grid, &B0, &B1);
    const long long n = (long long)L;
    unsigned* scanw = &sm.base->p_scan[0][0];                    // [2][PG_WARPS] warp totals, double-buffered by chunk parity
    unsigned left_run = sm.base->pp_left[k];                     // block-unit: this block's offset among the LEFT rows
    const unsigned left0 = left_run;
    unsigned right_run = (unsigned)min(B0, n) - left_run;
    unsigned* dstL = st.pool + (size_t)sm.base->pp_dst[k];
    unsigned* dstR = dstL + sm.base->pp_cntL[k];
    int par = 0;
#pragma unroll 1
    for (long long cb = B0; cb < B1; cb += 4 * PGP_CH) {
        float4 xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long row = cb + (long long)u * PGP_CH + tid * 4;
            xv[u] = row < B1 ? __ldg(reinterpret_cast<const float4*>(col + row)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long cs = cb + (long long)u * PGP_CH;      // chunk start (block-uniform)
            if (cs >= B1) break;
            const long long row = cs + tid * 4;
            const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
            bool vl[4], lf[4];
            int cl = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) { vl[e] = row + e < n && row < B1; lf[e] = vl[e] && xs[e] < thr; cl += lf[e] ? 1 : 0; }
            int incl = cl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(PG_FULL, incl, o); if (lane >= o) incl += y; }
            if (lane == 31) scanw[par * PG_WARPS + warp] = (unsigned)incl;
            __syncthreads();
            unsigned wpre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < PG_WARPS; ++w) { const unsigned v = scanw[par * PG_WARPS + w]; tot += v; wpre += w < warp ? v : 0u; }
            const unsigned before_left = wpre + (unsigned)(incl - cl);
            unsigned wl = left_run + before_left;
            unsigned wr = right_run + ((unsigned)(tid * 4) - before_left);      // rows before mine in the chunk are all valid if mine are
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (vl[e]) {
                    if (lf[e]) __stcg(&dstL[wl++], (unsigned)(row + e));
                    else __stcg(&dstR[wr++], (unsigned)(row + e));
                }
            }
            const long long ce = min(min(cs + PGP_CH, B1), n);
            const unsigned nvalid = ce > cs ? (unsigned)(ce - cs) : 0u;
            left_run += tot;
            right_run += nvalid - tot;
            par ^= 1;
        }
    }
    // the root pass counted the same predicate over the same rows
    if (tid == 0 && (left_run - left0) != sm.base->pp_exp[k]) c->error = PG_ERR_INTERNAL;
    // look-ahead: splitting.rs:43-49 for the proposals expected on the two children.  The requests of this job are staged
    // in shared memory once, then taken three at a time with four chunks in flight each (16 independent float4 loads per
    // thread), so the pass pays a memory round trip per group of requests instead of several per request.
    int* rq_idx = sm.base->p_rq_idx; int* rq_side = sm.base->p_rq_side; int* rq_var = sm.base->p_rq_var;
    __syncthreads();
    if (warp == 0) {
        int nreq = 0;
        for (int q0 = 0; q0 < Q; q0 += 32) {
            const int q = q0 + lane;
            const bool mine = q < Q && sm.base->pp_mq_pj[q] == k;
            const unsigned m = __ballot_sync(PG_FULL, mine);
            if (mine) {
                const int pos = nreq + __popc(m & ((1u << lane) - 1u));
                rq_idx[pos] = q; rq_side[pos] = sm.base->pp_mq_side[q]; rq_var[pos] = sm.base->pp_mq_var[q];
            }
            nreq += __popc(m);
        }
        if (lane == 0) sm.base->p_nreq = nreq;
    }
    __syncthreads();
    const int nreq = sm.base->p_nreq;
#pragma unroll 1
    for (int g0 = 0; g0 < nreq; g0 += 3) {
        const float* c2[3];
        bool wl[3];
        float mn[3], mx[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int gi = min(g0 + u, nreq - 1);            // a short last group repeats its last request (result discarded)
            c2[u] = st.X + (size_t)rq_var[gi] * ld;
            wl[u] = rq_side[gi] == 0;
            mn[u] = INFINITY; mx[u] = -INFINITY;
        }
#pragma unroll 1
        for (long long cb = B0; cb < B1; cb += 4 * PGP_CH) {
            float4 xv[4], zv[3][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long row = cb + (long long)u * PGP_CH + tid * 4;
                const bool in = row < B1;
                xv[u] = in ? __ldg(reinterpret_cast<const float4*>(col + row)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int w = 0; w < 3; ++w) zv[w][u] = in ? __ldg(reinterpret_cast<const float4*>(c2[w] + row)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long row = cb + (long long)u * PGP_CH + tid * 4;
                const int nv = row < B1 ? (int)min(4ll, n - row) : 0;
                const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool valid = e < nv, left = xs[e] < thr;
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        const float z = e == 0 ? zv[w][u].x : e == 1 ? zv[w][u].y : e == 2 ? zv[w][u].z : zv[w][u].w;
                        if (valid && left == wl[w]) { mn[w] = fminf(mn[w], z); mx[w] = fmaxf(mx[w], z); }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            if (g0 + u < nreq) {                             // block-uniform
                // an untouched (+inf, -inf) pair maps to the identity keys of min / max
                unsigned kmn = mn[u] <= mx[u] ? pg_f2key(mn[u]) : 0xFFFFFFFFu, kmx = mn[u] <= mx[u] ? pg_f2key(mx[u]) : 0u;
                kmn = __reduce_min_sync(PG_FULL, kmn);
                kmx = __reduce_max_sync(PG_FULL, kmx);
                if (lane == 0) { atomicMin(&sm.mmin[rq_idx[g0 + u]], kmn); atomicMax(&sm.mmax[rq_idx[g0 + u]], kmx); }
            }
        }
    }
    __syncthreads();      // scanw is reused by the next job
}

// Pool segment (a non-root node's rows), block-unit scheme: block b compacts its own contiguous entries
// (pg_seg_block_range; the statistics pass's per-block LEFT counts give the offsets) — 4 entries per thread, 2 chunks in
// flight, same scan as pg_part_ident; the split column is gathered through the index list.
__device__ __forceinline__ void pg_part_seg(const PgState& st, const PgBuf& bf, const PgSmemView& sm, int k, unsigned src,
                            const float* col, float thr, unsigned long long L) {
    PgCtrl* c = st.ctrl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    unsigned long long S0, S1;
    pg_seg_block_range(L, PG_BID, st.n_gwarps, PG_WARPS, &S0, &S1);
    unsigned* scanw = &sm.base->p_scan[0][0];
    unsigned left_run = sm.base->pp_left[k];
    const unsigned left0 = left_run;
    unsigned right_run = (unsigned)S0 - left_run;
    unsigned* dstL = st.pool + (size_t)sm.base->pp_dst[k];
    unsigned* dstR = dstL + sm.base->pp_cntL[k];
    const unsigned* seg = st.pool + (size_t)src;
    int par = 0;
#pragma unroll 1
    for (unsigned long long cb = S0; cb < S1; cb += 2 * PGP_CH) {
        unsigned idx[2][4];
        float xs[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned long long en = cb + (unsigned long long)u * PGP_CH + (unsigned long long)tid * 4 + e;
                idx[u][e] = en < S1 ? __ldcg(&seg[en]) : 0u;
            }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned long long en = cb + (unsigned long long)u * PGP_CH + (unsigned long long)tid * 4 + e;
                xs[u][e] = en < S1 ? __ldg(&col[idx[u][e]]) : INFINITY;
            }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const unsigned long long cs = cb + (unsigned long long)u * PGP_CH;
            if (cs >= S1) break;
            const unsigned long long e0 = cs + (unsigned long long)tid * 4;
            bool vl[4], lf[4];
            int cl = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) { vl[e] = e0 + e < S1; lf[e] = vl[e] && xs[u][e] < thr; cl += lf[e] ? 1 : 0; }
            int incl = cl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(PG_FULL, incl, o); if (lane >= o) incl += y; }
            if (lane == 31) scanw[par * PG_WARPS + warp] = (unsigned)incl;
            __syncthreads();
            unsigned wpre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < PG_WARPS; ++w) { const unsigned v = scanw[par * PG_WARPS + w]; tot += v; wpre += w < warp ? v : 0u; }
            const unsigned before_left = wpre + (unsigned)(incl - cl);
            unsigned wl = left_run + before_left;
            unsigned wr = right_run + ((unsigned)(tid * 4) - before_left);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (vl[e]) {
                    if (lf[e]) __stcg(&dstL[wl++], idx[u][e]);
                    else __stcg(&dstR[wr++], idx[u][e]);
                }
            }
            const unsigned long long ce = cs + PGP_CH < S1 ? cs + PGP_CH : S1;
            left_run += tot;
            right_run += (unsigned)(ce - cs) - tot;
            par ^= 1;
        }
    }
    if (tid == 0 && (left_run - left0) != sm.base->pp_exp[k]) c->error = PG_ERR_INTERNAL;
    __syncthreads();      // scanw is reused by the next job
}

template <bool FASTP>
__device__ __forceinline__ void pg_pass_part(const PgState& st, const PgBuf& bf, const PgSmemView& sm) {
    PgCtrl* c = st.ctrl;
    const int K = sm.base->cmd.n_pjobs;
    const int Q = sm.base->cmd.n_groups;     // look-ahead min/max requests riding on this pass (0 in the generic control code)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = PG_BID * PG_WARPS + warp;
    const unsigned lt_mask = (1u << lane) - 1u;
    const size_t ld = (size_t)st.ld;
    if (Q > 0) {
        for (int q = threadIdx.x; q < Q; q += blockDim.x) { sm.mmin[q] = 0xFFFFFFFFu; sm.mmax[q] = 0u; }
    }
    if (FASTP) {   // every job's parameters and the look-ahead requests in one round trip (K <= S <= 31, Q <= 64)
        PgSmem* b = sm.base;
        const int t = threadIdx.x;
        if (t < K && t < 32) {
            b->pp_src[t] = __ldcg(&st.pj_src_off[t]); b->pp_len[t] = __ldcg(&st.pj_len[t]); b->pp_var[t] = __ldcg(&st.pj_var[t]);
            b->pp_thr[t] = __ldcg(&st.pj_thr32[t]);
            b->pp_dst[t] = __ldcg(&st.pj_dst_off[t]); b->pp_cntL[t] = __ldcg(&st.pj_cntL[t]);
        }
        // this block's offset among a job's LEFT entries = the per-block left counts of the blocks before it (the pass that
        // computed the job's statistics left them in part_ccnt): every block adds them up itself, one warp per job, instead
        // of the control block running a 147-entry prefix per job before it can issue this pass
        {
            const int w = t >> 5, l = t & 31;
            for (int k = w; k < K && k < 32; k += PG_WARPS) {
                const int j = __ldcg(&st.pj_job[k]);
                unsigned acc = 0u, mine = 0u;
                for (int bq = l; bq <= PG_BID; bq += 32) {
                    const unsigned v = __ldcg(&bf.part_ccnt[(size_t)bq * st.S + j]);
                    if (bq < PG_BID) acc += v; else mine = v;
                }
                acc = __reduce_add_sync(PG_FULL, acc);
                mine = __reduce_add_sync(PG_FULL, mine);
                if (l == 0) { b->pp_left[k] = acc; b->pp_exp[k] = mine; }
            }
        }
        if (t >= 64 && t - 64 < Q && t - 64 < 64) {
            const int q = t - 64;
            b->pp_mq_pj[q] = __ldcg(&st.mq_pj[q]); b->pp_mq_side[q] = __ldcg(&st.mq_side[q]); b->pp_mq_var[q] = __ldcg(&st.mq_var[q]);
        }
    }
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const unsigned src = FASTP ? sm.base->pp_src[k] : __ldcg(&st.pj_src_off[k]);
        const bool ident = src == PG_SEG_IDENTITY;
        const unsigned long long L = FASTP ? sm.base->pp_len[k] : __ldcg(&st.pj_len[k]);
        const float* col = st.X + (size_t)(FASTP ? sm.base->pp_var[k] : __ldcg(&st.pj_var[k])) * ld;
        const float thr = FASTP ? sm.base->pp_thr[k] : __ldcg(&st.pj_thr32[k]);
        if (FASTP) {                                         // block-unit scheme for every segment
            if (ident) pg_part_ident(st, bf, sm, k, Q, col, thr, L); else pg_part_seg(st, bf, sm, k, src, col, thr, L);
            continue;
        }
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

// After a root pass: the control block has usually staged the NEXT tree update's root jobs already (pg_serial_fast.cuh,
// fw_prestage), so this block asks L2 for its own rows of those columns now — one bulk prefetch per distinct column,
// issued after the block has reported completion, i.e. off the critical path.  The next root pass then streams from
// L2 instead of waiting on HBM round trips.  Purely a hint: stale job arrays only cost a useless prefetch.
__device__ __forceinline__ void pg_prefetch_next_root(const PgState& st, int flags) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const int pb_next = 1 - ((flags >> 2) & 1);
    const int J = min(__ldcg(&st.ctrl->pre_J[pb_next]), 32);
    const bool live = lane < J;
    const int var = live ? __ldcg(&st.job_var[(size_t)pb_next * st.S + lane]) : -1 - lane;
    const unsigned same = __match_any_sync(PG_FULL, var);
    if (!live || (__ffs(same) - 1) != lane || var < 0 || var >= st.p) return;
    long long B0, B1;
    pg_block_range(st.n, (int)blockIdx.x - 1, st.grid, &B0, &B1);
    if (B1 <= B0) return;
    const float* src = st.X + (size_t)var * st.ld + B0;
    const unsigned bytes = (unsigned)(B1 - B0) * 4u;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// Dispatch one grid pass.
// Pass variants.  Every persistent kernel instantiates exactly one, so that the register allocation and code size of
// one variant never disturb another (the kernels are monoliths: control block + every pass inlined):
//   PGV_GENERIC   warp-range passes, both likelihood families at run time: generic control code (P > 32), stepwise driver
//   PGV_DIRECT    P <= 32, Gaussian families: block-unit scheme, root pass = register loads over warp sub-ranges
//   PGV_RING      P <= 32, Gaussian families: block-unit scheme, root pass = bulk-copy ring in shared memory (default)
//   PGV_BERN      P <= 32, Bernoulli-logit: block-unit scheme, ring root pass with the stump log-likelihood
enum { PGV_GENERIC = 0, PGV_DIRECT = 1, PGV_RING = 3, PGV_BERN = 4 };

// Ring root pass: always when asked for (root_ring = 1); by default (2) only while a block's rows fit the ring a few tiles
// deep, where every copy of the pass is in flight at once — measured 8.3 us against 11.4 for the direct pass at n = 100k,
// but 34 us against 24 at n = 1M, where the refills are paced by the consumers.
__host__ __device__ inline bool pg_use_ring(const PgState& st) {
    return st.stage_bytes > 0 && (st.root_ring == 1 || (st.root_ring == 2 && st.n <= (long long)st.grid * 2048));
}

template <int VAR>
__device__ __forceinline__ void pg_run_pass(const PgState& st, const PgSmemView& sm, int phase) {
    const PgBuf bf = pg_make_buf(st, sm.base->cmd.flags);
    constexpr bool FASTP = VAR != PGV_GENERIC;
    switch (phase) {
        case PH_ROOT:
        case PH_FINAL:
            if (VAR == PGV_GENERIC) { if (st.family == PG_BERNOULLI_LOGIT) pg_pass_root<true, false>(st, bf, sm); else pg_pass_root<false, false>(st, bf, sm); }
            else if (VAR == PGV_BERN) { if (!pg_use_ring(st) || !pg_pass_root_ring<true, PGR_JW>(st, bf, sm)) pg_pass_root<true, true>(st, bf, sm); }
            else if (VAR == PGV_RING) {
                const bool done = st.S <= 20 ? pg_pass_root_ring<false, 5>(st, bf, sm) : pg_pass_root_ring<false, PGR_JW>(st, bf, sm);
                if (!done) pg_pass_root<false, true>(st, bf, sm);      // more distinct columns than the staging area holds
            }
            else pg_pass_root<false, true>(st, bf, sm);
            break;
        case PH_MINMAX: pg_pass_minmax<VAR == PGV_GENERIC>(st, bf, sm); break;
        case PH_STATS: if (FASTP && VAR != PGV_BERN) pg_pass_seg_block(st, bf, sm); else pg_pass_seg(st, bf, sm); break;
        case PH_BERN: if (VAR == PGV_GENERIC || VAR == PGV_BERN) pg_pass_bern(st, bf, sm); break;
        case PH_PART: pg_pass_part<FASTP>(st, bf, sm); break;
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
// arrival: add with release semantics at gpu scope (the block's results, ordered before by __syncthreads, become visible
// to whoever acquires the counter) — one instruction instead of a full fence followed by a relaxed atomic
__device__ __forceinline__ void pg_red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// acquire / release fence at gpu scope (MEMBAR.ALL.GPU + L1 invalidate).  __threadfence() is the sequentially consistent
// one (MEMBAR.SC.GPU), which none of the hand-offs here needs: they are all message passing (data, then flag).
__device__ __forceinline__ void pg_fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void pg_red_relaxed_add(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void pg_st_relaxed(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned pg_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pg_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long pg_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void pg_spin_until(const unsigned* p, unsigned target, PgCtrl* c) {
    const long long t0 = clock64();
    unsigned ns = 20;
    while (pg_ld_relaxed(p) < target) {
        __nanosleep(ns);
        if (ns < 160) ns <<= 1;
        if (clock64() - t0 > 80000000000ll) {   // ~40 s at 2 GHz (twice the exchange time-out of a sharded run): a lost block becomes an error, not a hang
            c->error = PG_ERR_WATCHDOG;
            unsigned long long* hd = reinterpret_cast<unsigned long long*>(c->hostdiag);
            if (hd && atomicCAS(&hd[0], 0ull, 0xD1A60000ull + blockIdx.x) == 0ull) {     // who waited for what (read by the host after the trap)
                hd[1] = (unsigned long long)(reinterpret_cast<const char*>(p) - reinterpret_cast<const char*>(c));
                hd[2] = target; hd[3] = pg_ld_relaxed(p);
                hd[4] = pg_ld_relaxed(&c->bar_gen); hd[5] = pg_ld_relaxed(&c->cnt[0][0]); hd[6] = pg_ld_relaxed(&c->cnt[1][0]);
                hd[7] = ((unsigned long long)(unsigned)c->cmd[0][0] << 32) | (unsigned)c->cmd[1][0];
                __threadfence_system();
            }
            __threadfence();
            __trap();
        }
    }
    pg_fence_acq_rel();   // acquire side of the fence/atomic/fence pattern
}

