// pg_serial_fast.cuh — latency-organised serial section of the persistent kernel (P <= 32).
//
// Same state machine and arithmetic as pg_serial.h (the generic implementation used by the host
// simulation, the stepwise driver and P > 32).  ncu on the isolated section showed it is bound by
// INSTRUCTION FETCH (~7 stall cycles per instruction on `no_instruction`: ~9k cold, straight-line
// instructions executed once) and by dependent L2 round trips (~0.35 us each), so this version is
// organised around executing few instructions and touching global memory rarely:
//   * Block 0 is always the serial block and keeps PgCtrl, the per-slot scalars and the job list in
//     its shared memory for the whole kernel (loaded at launch, stored at exit); after each section
//     it publishes only the few words the grid passes read.
//   * Everything that does not depend on the data — the Philox/tape draws of the next tree update
//     (depth-prior uniforms, split variables, split fractions, leaf normals, resampling and
//     selection uniforms) and, for round 0, the complete proposals (the root's column min/max are
//     constants) — is computed by otherwise idle warps (1..FW_DR+1) while warp 0 runs the logic.
//   * Round 0 is table-free: proposals' results stay in shared memory; per-slot node tables are only
//     written for particles that survive resampling after growing (the reference's stale-weight rule,
//     Q1, kills grown particles in 62 % of updates, which then end as a stump without touching a table).
//   * One lane per particle slot / per job; the order-sensitive loops of the reference
//       normalize_weights_inplace smc.rs:257-268, systematic resampling resampling.rs:24-44,
//       WeightedIndex smc.rs:113-114
//     keep the reference's summation order through warp-uniform loops over shuffle broadcasts.
//   * Loops are kept rolled (#pragma unroll 1): code bytes, not instruction count, are the cost.
#pragma once
#include <cstddef>
#include "pg_passes.cuh"

struct FwCx {               // passed by value: stays in registers
    PgFastSm* f;
    PgCtrl* c;              // working copy of the control block (f->lc)
    const double* thr;      // depth-prior thresholds
    const double* prior;    // split prior (null = uniform draw)
    const float* cmin; const float* cmax;
};

__device__ __forceinline__ double fw_sum_seq(double v, int n) {   // lanes 0..n-1 in order; uniform result
    double acc = 0.0;
#pragma unroll 8            // the shuffles of a batch are in flight together; the chain of additions keeps its order
    for (int k = 0; k < n; ++k) acc += __shfl_sync(PG_FULL, v, k);
    return acc;
}
__device__ __forceinline__ double fw_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(PG_FULL, v, o));
    return v;
}
__device__ __forceinline__ int fw_imax(int v) { return __reduce_max_sync(PG_FULL, v); }

// 8-byte words of PgCtrl owned by the serial block (everything before the barrier words).
__device__ __forceinline__ int fw_ctrl_words() { return (int)(offsetof(PgCtrl, bar_count) / 8); }

// ---- kernel-lifetime state of block 0 ----------------------------------------------------------
__device__ __forceinline__ void fw_load_state(const PgState& st, PgFastSm* f) {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(st.ctrl);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&f->lc);
    for (int i = threadIdx.x; i < fw_ctrl_words(); i += blockDim.x) dst[i] = __ldcg(&src[i]);
    const int S = st.S;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        f->hw[s] = __ldcg(&st.w[s]); f->hmut[s] = __ldcg(&st.mutated[s]);
        f->hanc[s] = __ldcg(&st.anc[s]); f->hsjob[s] = __ldcg(&st.slot_job[s]);
    }
    for (int s = threadIdx.x; s < 2 * S; s += blockDim.x) {
        f->hsize[s] = __ldcg(&st.pt_size[s]); f->hqh[s] = __ldcg(&st.pt_qhead[s]);
        f->hqt[s] = __ldcg(&st.pt_qtail[s]); f->hG[s] = __ldcg(&st.pt_G[s]);
    }
    if (threadIdx.x < 3) f->d_upd[threadIdx.x] = ~0ull;
    if (threadIdx.x < 3 * (FW_DR + 1)) f->d_flag[threadIdx.x / (FW_DR + 1)][threadIdx.x % (FW_DR + 1)] = 0ull;
    if (threadIdx.x == 0) { f->issued = 0u; f->cur_gen = 0u; f->out_phase[0] = f->out_phase[1] = PH_BEGIN; f->out_J[0] = f->out_J[1] = 0; f->out_pb[0] = f->out_pb[1] = 0;
                            f->quit = 0; f->pre_upd = ~0ull; f->spec_now = 0; f->stat_spec = 0; f->pred_tried = 0; f->lazy_J = 0;
                            f->early_now = 0; f->es_pending = 0; f->es_resume = SN_NONE; f->es_n = 0; f->es_mask = 0; }
}

__device__ __forceinline__ void fw_store_state(const PgState& st, PgFastSm* f) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(st.ctrl);
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&f->lc);
    for (int i = threadIdx.x; i < fw_ctrl_words(); i += blockDim.x) dst[i] = src[i];
    const int S = st.S;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        st.w[s] = f->hw[s]; st.mutated[s] = f->hmut[s]; st.anc[s] = f->hanc[s]; st.slot_job[s] = f->hsjob[s];
    }
    for (int s = threadIdx.x; s < 2 * S; s += blockDim.x) {
        st.pt_size[s] = f->hsize[s]; st.pt_qhead[s] = f->hqh[s]; st.pt_qtail[s] = f->hqt[s]; st.pt_G[s] = f->hG[s];
    }
}

// what the grid passes read: the first 16 ints of PgCtrl (phase .. max_size)
__device__ __forceinline__ void fw_publish(const PgState& st, PgFastSm* f, int lane) {
    if (lane < 16) reinterpret_cast<int*>(st.ctrl)[lane] = reinterpret_cast<const int*>(&f->lc)[lane];
}

// ---- stage A: reduce the pass's per-block partials --------------------------------------------
// lane = job, warp = subset of blocks {warp, warp + PG_WARPS, ...}; rows of the [block][job] partial
// arrays are contiguous in job: coalesced, and all of a warp's loads are in flight at once.
#define FW_RW (PG_WARPS - 1)   // loading warps: 1 .. PG_WARPS - 1 (warp 0 waits for the completion count and issues ahead meanwhile)
#define FW_RB 10               // blocks per loading warp per batch (147 pass blocks / 15 warps = 9.8)
// A loading warp waits for ITS blocks only (per-block completion flags, PgState::blk_done) and fetches their partials at
// once: the fetch of ~56 KB through one SM's L2 port (1.4 us when it started after the slowest block) overlaps the pass's
// tail, and what remains after the last arrival is one warp's share.
__device__ __forceinline__ void fw_wait_blocks(const PgState& st, int warp, int lane, unsigned g) {
    bool seen;
    do {
        seen = true;
        for (int b = warp - 1 + lane * FW_RW; b < st.grid; b += 32 * FW_RW)
            seen = seen && (int)(pg_ld_acquire(&st.blk_done[b]) - g) >= 0;
        seen = __all_sync(PG_FULL, seen);
        if (!seen) __nanosleep(20);
    } while (!seen);        // a block that never arrives: the watchdog of thread 0's wait on the count ends the kernel
    __syncwarp();
}
__device__ __forceinline__ void fw_reduce(const PgState& st, PgFastSm* f, int phase, int J, int pb, unsigned g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = st.grid, S = st.S;
    const double* p_sum = st.part_sum + pg_off_part_sum(st, pb);
    const unsigned* p_ccnt = st.part_ccnt + pg_off_part_ccnt(st, pb);
    const double* p_tot = st.part_tot + pg_off_part_tot(st, pb);
    const bool sums = phase == PH_ROOT || phase == PH_STATS || phase == PH_BERN;
    // Root pass: the block's four totals ride along as two more "jobs" (lanes J and J + 1 take them as two double2), so they
    // cost no load stage, no shuffle tree and no combine loop of their own.  J = 31 leaves no room: the separate stage below.
    const bool tot_lanes = phase == PH_ROOT && J + 2 <= 32;
    if (sums) {
        const double* src = phase == PH_BERN ? st.part_ll : p_sum;
        const bool isjob = lane < J;
        const bool cnts = phase != PH_BERN && isjob;
        double a = 0.0, bsum = 0.0;
        unsigned cc = 0u;
        if (warp > 0) fw_wait_blocks(st, warp, lane, g);
        if (warp > 0 && (isjob || (tot_lanes && lane < J + 2))) {
            const double* base = isjob ? src + (size_t)lane * 2 : p_tot + (size_t)(lane - J) * 2;
            const size_t stride = isjob ? (size_t)S * 2 : (size_t)4;
#pragma unroll 1
            for (int b0 = warp - 1; b0 < G; b0 += FW_RW * FW_RB) {
                double2 v[FW_RB];
                unsigned k[FW_RB];
#pragma unroll
                for (int q = 0; q < FW_RB; ++q) {
                    const int b = b0 + q * FW_RW;
                    if (b < G) {
                        v[q] = __ldcg(reinterpret_cast<const double2*>(base + (size_t)b * stride));
                        k[q] = cnts ? __ldcg(&p_ccnt[(size_t)b * S + lane]) : 0u;
                    } else { v[q] = make_double2(0.0, 0.0); k[q] = 0u; }
                }
#pragma unroll
                for (int q = 0; q < FW_RB; ++q) { a += v[q].x; bsum += v[q].y; cc += k[q]; }
            }
        }
        f->wsA[warp][lane] = a; f->wsB[warp][lane] = bsum; f->wsC[warp][lane] = cc;
    } else if (phase == PH_MINMAX || phase == PH_PART) {
        unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
        if (warp > 0) fw_wait_blocks(st, warp, lane, g);
        if (warp > 0 && lane < J) {
#pragma unroll 1
            for (int b = warp - 1; b < G; b += FW_RW) {
                const uint2 v = __ldcg(reinterpret_cast<const uint2*>(st.part_mm + ((size_t)b * 2 * S + lane) * 2));
                kmin = min(kmin, v.x); kmax = max(kmax, v.y);
            }
        }
        f->wsC[warp][lane] = kmin;
        f->wsD[warp][lane] = kmax;
    }
    if (phase == PH_ROOT && !tot_lanes) {   // root totals: thread = block
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        for (int b = threadIdx.x; b < G; b += blockDim.x) {
            while ((int)(pg_ld_acquire(&st.blk_done[b]) - g) < 0) __nanosleep(20);
            const double2 u = __ldcg(reinterpret_cast<const double2*>(p_tot + (size_t)b * 4));
            const double2 v = __ldcg(reinterpret_cast<const double2*>(p_tot + (size_t)b * 4 + 2));
            t0 += u.x; t1 += u.y; t2 += v.x; t3 += v.y;
        }
#pragma unroll 1
        for (int o = 16; o > 0; o >>= 1) {
            t0 += __shfl_xor_sync(PG_FULL, t0, o); t1 += __shfl_xor_sync(PG_FULL, t1, o);
            t2 += __shfl_xor_sync(PG_FULL, t2, o); t3 += __shfl_xor_sync(PG_FULL, t3, o);
        }
        if (lane == 0) { f->wsT[warp][0] = t0; f->wsT[warp][1] = t1; f->wsT[warp][2] = t2; f->wsT[warp][3] = t3; }
    }
        if (st.heartbeat && (threadIdx.x & 31) == 0) { volatile unsigned long long* hw_ = reinterpret_cast<volatile unsigned long long*>(st.ctrl->hostdiag); if (hw_) hw_[16 + (threadIdx.x >> 5)] = ((unsigned long long)f->issued << 32) | 2ull; }
    __syncthreads();
    if (warp == 0) {   // fixed-order combine over warps
        if (sums && (lane < J || (tot_lanes && lane < J + 2))) {
            double a = 0.0, b = 0.0;
            unsigned cc = 0u;
#pragma unroll 4
            for (int w = 0; w < PG_WARPS; ++w) { a += f->wsA[w][lane]; b += f->wsB[w][lane]; cc += f->wsC[w][lane]; }
            if (lane >= J) { f->tot[2 * (lane - J)] = a; f->tot[2 * (lane - J) + 1] = b; }
            else if (phase == PH_BERN) { f->rllL[lane] = a; f->rllR[lane] = b; }
            else { f->rSo[lane] = a; f->rSr[lane] = b; f->rCnt[lane] = cc; f->rCntL[lane] = cc; }
        }
        if ((phase == PH_MINMAX || phase == PH_PART) && lane < J) {
            unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll 4
            for (int w = 0; w < PG_WARPS; ++w) { kmin = min(kmin, f->wsC[w][lane]); kmax = max(kmax, f->wsD[w][lane]); }
            f->rMin[lane] = kmin; f->rMax[lane] = kmax;
        }
        if (phase == PH_ROOT && !tot_lanes && lane < 4) {
            double v = 0.0;
#pragma unroll 1
            for (int w = 0; w < PG_WARPS; ++w) v += f->wsT[w][lane];
            f->tot[lane] = v;
        }
        __syncwarp();
    }
}

// ---- stage A': observation sharding — exchange the reduced values between the ranks' control blocks ------------
// Every rank holds a contiguous block of rows and runs the identical control code on identical draws; the only
// data-dependent inputs of that code are the values fw_reduce produces.  Each rank writes its values into every peer's
// mailbox (stores into peer memory over NVLink), waits until every peer's values have arrived in its own, and combines all
// ranks' contributions IN RANK ORDER, so every rank computes bit-identical totals and takes bit-identical decisions; no
// host round trip, no collective launch.
// Flag-in-data: a value travels as two 8-byte words (low half | seq << 32, high half | seq << 32), each written by one
// store, which the hardware performs as a unit; the receiver polls the words themselves until both carry the sequence
// number of this exchange.  No release fence on the sender (a system-scope release waits for the acknowledgement of every
// earlier store to the peers: one more NVLink round trip per exchange) and no separate flag.  Mailbox slots alternate with
// the parity of the sequence number: a rank can be at most one exchange ahead of a peer that still reads the previous
// slot, and a stale word carries seq - 2.
__device__ __forceinline__ void fw_st_mail(unsigned long long* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void fw_ld_mail(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void fw_exchange(const PgState& st, PgFastSm* f, int phase, int J, int lane) {
    const int W = st.world, me = st.rank;
    const bool sums = phase == PH_ROOT || phase == PH_STATS || phase == PH_BERN;
    const bool mm = phase == PH_MINMAX || phase == PH_PART;
    if (!sums && !mm) return;
    // payload: [0..3] root totals, then per job three values
    const int n_val = 4 + 3 * J;
    const unsigned long long seq = f->lc.xseq + 1ull;
    const unsigned long long tag = (seq & 0xFFFFFFFFull) << 32;
    const int par = (int)(seq & 1ull);
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {                            // this lane's values: indices lane, lane + 32, ...
        const int i = lane + 32 * q;
        if (i >= n_val) break;
        double v = 0.0;
        if (i < 4) v = phase == PH_ROOT ? f->tot[i] : 0.0;
        else {
            const int j = (i - 4) / 3, k = (i - 4) % 3;
            if (phase == PH_BERN) v = k == 0 ? f->rllL[j] : k == 1 ? f->rllR[j] : 0.0;
            else if (sums) v = k == 0 ? f->rSo[j] : k == 1 ? f->rSr[j] : (double)f->rCnt[j];
            else v = k == 0 ? (double)f->rMin[j] : k == 1 ? (double)f->rMax[j] : 0.0;      // ordered keys: exact in a double
        }
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        const unsigned long long w0 = (bits & 0xFFFFFFFFull) | tag, w1 = (bits >> 32) | tag;
#pragma unroll
        for (int r = 0; r < PG_MAX_RANKS; ++r)               // my slot in every rank's mailbox (my own included)
            if (r < W) fw_st_mail(reinterpret_cast<unsigned long long*>(st.mail[r]) + (((size_t)par * PG_MAX_RANKS + me) * PG_MAIL_DOUBLES + i) * 2, w0, w1);
    }
    const unsigned long long* box = reinterpret_cast<const unsigned long long*>(st.mail[me]) + (size_t)par * PG_MAX_RANKS * PG_MAIL_DOUBLES * 2;
    const long long t0 = clock64();
    bool lost = false;
#pragma unroll 1
    for (int q = 0; q < 4 && !lost; ++q) {
        const int i = lane + 32 * q;
        if (i >= n_val) break;
        unsigned long long a[PG_MAX_RANKS], b[PG_MAX_RANKS];
        for (;;) {                                           // every rank's words requested at once, until all carry this exchange's tag
            bool ok = true;
#pragma unroll
            for (int r = 0; r < PG_MAX_RANKS; ++r)
                if (r < W) fw_ld_mail(box + ((size_t)r * PG_MAIL_DOUBLES + i) * 2, a[r], b[r]);
#pragma unroll
            for (int r = 0; r < PG_MAX_RANKS; ++r)
                if (r < W) ok = ok && (a[r] >> 32) == (tag >> 32) && (b[r] >> 32) == (tag >> 32);
            if (ok) break;
            if (clock64() - t0 > 40000000000ll) { lost = true; break; }   // a lost peer becomes an error (half the pass blocks' watchdog)
            __nanosleep(40);
        }
        if (lost) break;
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < PG_MAX_RANKS; ++r) {             // combined in rank order
            if (r >= W) continue;
            const double v = __longlong_as_double((long long)((a[r] & 0xFFFFFFFFull) | (b[r] << 32)));
            if (sums) acc += v;
            else { const int k = (i - 4) % 3; acc = r == 0 ? v : (k == 0 ? fmin(acc, v) : fmax(acc, v)); }
        }
        if (i < 4) { if (phase == PH_ROOT) f->tot[i] = acc; }
        else {
            const int j = (i - 4) / 3, k = (i - 4) % 3;
            if (phase == PH_BERN) { if (k == 0) f->rllL[j] = acc; else if (k == 1) f->rllR[j] = acc; }
            else if (sums) { if (k == 0) f->rSo[j] = acc; else if (k == 1) f->rSr[j] = acc; else f->rCnt[j] = (unsigned)acc; }
            else { if (k == 0) f->rMin[j] = (unsigned)acc; else if (k == 1) f->rMax[j] = (unsigned)acc; }
        }
    }
    if (__any_sync(PG_FULL, lost) && lane == 0) f->lc.error = PG_ERR_WATCHDOG;
    if (lane == 0) f->lc.xseq = seq;
    __syncwarp();
}

// smc.rs:242-254 (prior in shared memory when staged), same subtraction order as the reference
__device__ __forceinline__ int fw_feature_from_probs(const PgState& st, const double* prior, double u) {
    double target = u * st.prior_total;
    const int p = st.p;
#pragma unroll 1
    for (int j = 0; j < p; ++j) { target -= prior[j]; if (target <= 0.0) return j; }
    return p - 1;
}

// ---- cold draws -----------------------------------------------------------------------------------------
// Warp 0 normally finds its draws precomputed by the helper warps; the generator itself (ten Philox rounds, log / cos
// for the normals, the tape bounds checks) is needed only for rounds beyond FW_DR.  Inlined at every call site it
// would sit in the middle of the control block's hot code, which is instruction-fetch bound — so the fallback lives in
// one out-of-line function that takes the generator's few fields by value (taking the address of the kernel parameter
// `st` would force the whole struct into local memory).
struct FwRngView {
    int rng_mode, S, tape_r_cap;
    unsigned k0, k1;
    const double *tape_slot, *tape_round, *tape_upd;
    long long tape_n_upd;
    unsigned long long tape_base_upd;
};
__device__ __forceinline__ FwRngView fw_rng_view(const PgState& st) {
    FwRngView v;
    v.rng_mode = st.rng_mode; v.S = st.S; v.tape_r_cap = st.tape_r_cap; v.k0 = st.philox_key0; v.k1 = st.philox_key1;
    v.tape_slot = st.tape_slot; v.tape_round = st.tape_round; v.tape_upd = st.tape_upd;
    v.tape_n_upd = st.tape_n_upd; v.tape_base_upd = st.tape_base_upd;
    return v;
}
// what = 0 uniform, 1 normal, 2 split value in [lo, hi) (pg_rng.h: pg_uniform / pg_normal / pg_split_draw)
__device__ __noinline__ double fw_cold_draw(FwRngView v, PgCtrl* c, int what, int kind, unsigned long long upd, int round, int slot,
                                            double lo, double hi) {
    if (v.rng_mode == PG_RNG_TAPE) {
        const double* p = nullptr;
        if (upd >= v.tape_base_upd) {
            const unsigned long long u = upd - v.tape_base_upd;
            if ((long long)u < v.tape_n_upd && round < v.tape_r_cap) {
                if (kind <= PG_D_ZRIGHT)
                    p = v.tape_slot + (((u * (unsigned long long)v.tape_r_cap + (unsigned long long)round) * (unsigned long long)v.S
                                        + (unsigned long long)slot) * 5ull + (unsigned long long)kind);
                else if (kind == PG_D_RESAMPLE) p = v.tape_round + (u * (unsigned long long)v.tape_r_cap + (unsigned long long)round);
                else p = v.tape_upd + u;
            }
        }
        if (!p) { c->error = PG_ERR_TAPE; return 0.0; }
        const double u = *p;
        if (what != 2) return u;
        const double r = __dadd_rn(__dmul_rn(u, hi - lo), lo);
        if (!(r < hi)) c->error = PG_ERR_TAPE;
        return r;
    }
    uint32_t o[4];
    if (what != 2) {
        pg_philox4x32_10((uint32_t)upd, (uint32_t)(upd >> 32), (uint32_t)round, ((uint32_t)slot << 16) | ((uint32_t)kind << 8), v.k0, v.k1, o);
        const double ua = pg_bits53(o[0], o[1]);
        if (what == 0) return ua;
        return sqrt(-2.0 * log(1.0 - ua)) * cos(6.283185307179586476925286766559 * pg_bits53(o[2], o[3]));
    }
    for (int attempt = 0; attempt < 64; ++attempt) {
        pg_philox4x32_10((uint32_t)upd, (uint32_t)(upd >> 32), (uint32_t)round,
                         ((uint32_t)slot << 16) | ((uint32_t)PG_D_SPLIT << 8) | (uint32_t)(attempt & 0xff), v.k0, v.k1, o);
        const double r = __dadd_rn(__dmul_rn(pg_bits53(o[0], o[1]), hi - lo), lo);
        if (r < hi) return r;
    }
    return lo;
}

// ---- helper warps: draws of one tree update (data-independent) --------------------------------------
__device__ __forceinline__ bool fw_tape_ok(const PgState& st, unsigned long long upd, int round) {
    if (st.rng_mode != PG_RNG_TAPE) return true;
    if (upd < st.tape_base_upd) return false;
    return (long long)(upd - st.tape_base_upd) < st.tape_n_upd && round < st.tape_r_cap;
}

// warp hw (1..FW_DR): round hw-1; warp FW_DR+1: resampling / selection uniforms.  `scr` is a scratch control
// block: helpers must not raise errors for draws that may never be consumed.
__device__ __forceinline__ void fw_prefetch_draws(const PgState& st, const FwCx& x, unsigned long long upd, int hw, int lane, PgCtrl* scr) {
    PgFastSm* f = x.f;
    const int buf = (int)(upd % 3ull);
    const int S = st.S;
    if (hw <= FW_DR) {
        const int r = hw - 1;
        const bool ok = fw_tape_ok(st, upd, r);
        if (lane < S && ok) {
            const int s = lane;
            const double u1 = pg_uniform(st, scr, PG_D_GROW, upd, r, s);
            const double u2 = pg_uniform(st, scr, PG_D_VAR, upd, r, s);
            int var;
            if (x.prior) var = fw_feature_from_probs(st, x.prior, u2);
            else { var = (int)(u2 * (double)st.p); if (var > st.p - 1) var = st.p - 1; }
            double u3;
            if (st.rng_mode == PG_RNG_TAPE) u3 = *pg_tape_at(st, scr, PG_D_SPLIT, upd, r, s);
            else { double ub; pg_philox_pair(st, PG_D_SPLIT, upd, r, s, 0, &u3, &ub); }
            f->d_u1[buf][r][s] = u1; f->d_var[buf][r][s] = var; f->d_u3[buf][r][s] = u3;
            f->d_zL[buf][r][s] = pg_normal(st, scr, PG_D_ZLEFT, upd, r, s);
            f->d_zR[buf][r][s] = pg_normal(st, scr, PG_D_ZRIGHT, upd, r, s);
            if (r == 0) {   // complete root proposal: splitting.rs:39-56 with the column's constant min/max
                int status = 5;
                float lo = 0.f, hi = 0.f, thr = 0.f;
                double split = 0.0;
                if (x.thr[0] > u1) status = 1;
                else {
                    lo = x.cmin[var]; hi = x.cmax[var];
                    if (lo >= hi) status = 3;
                    else if (2 >= st.maxn) status = 7;
                    else {
                        scr->error = PG_OK;
                        split = pg_split_draw(st, scr, upd, 0, s, (double)lo, (double)hi);
                        if (scr->error != PG_OK) status = 6;
                        thr = pg_thr32(split);
                    }
                }
                f->d0_status[buf][s] = status; f->d0_lo[buf][s] = lo; f->d0_hi[buf][s] = hi;
                f->d0_thr[buf][s] = thr; f->d0_split[buf][s] = split;
            }
        }
        if (lane == 0) f->d_ok[buf][r] = ok ? 1 : 0;
    } else {
        if (lane < FW_DR) {
            const bool ok = fw_tape_ok(st, upd, lane);
            f->d_u6[buf][lane] = ok ? pg_uniform(st, scr, PG_D_RESAMPLE, upd, lane, 0) : 0.0;
        } else if (lane == FW_DR) {
            const bool ok = fw_tape_ok(st, upd, 0);
            f->d_u7[buf] = ok ? pg_uniform(st, scr, PG_D_SELECT, upd, 0, 0) : 0.0;
            f->d_ok[buf][FW_DR] = ok ? 1 : 0;
            f->d_upd[buf] = upd;
        }
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) f->d_flag[buf][hw - 1] = upd + 1ull;
}

__device__ __forceinline__ void fw_wait_draws(const PgState& st, PgFastSm* f, unsigned long long upd) {
    const int buf = (int)(upd % 3ull);
#pragma unroll 1
    for (int q = 0; q <= FW_DR; ++q) {
        const long long t0 = clock64();
        while (f->d_flag[buf][q] != upd + 1ull) {
            __nanosleep(20);
            if (clock64() - t0 > 40000000000ll) {       // the helper warps never produced these draws: report it instead of hanging
                unsigned long long* hd = reinterpret_cast<unsigned long long*>(st.ctrl->hostdiag);
                if (hd && atomicCAS(&hd[0], 0ull, 0xD2A30000ull + (unsigned)q) == 0ull) {
                    hd[1] = upd; hd[2] = f->d_flag[buf][q]; hd[3] = f->d_upd[buf]; hd[4] = f->d_upd[0]; hd[5] = f->d_upd[1]; hd[6] = f->d_upd[2]; hd[7] = f->issued;
                    __threadfence_system();
                }
                f->lc.error = PG_ERR_WATCHDOG;
                break;
            }
        }
    }
    __threadfence_block();
}

// draws as seen by warp 0: precomputed when in range, else the generic (error-raising) path
__device__ __forceinline__ double fw_u6(const PgState& st, const FwCx& x, unsigned long long upd, int r) {
    const int buf = (int)(upd % 3ull);
    if (r < FW_DR && x.f->d_ok[buf][r]) return x.f->d_u6[buf][r];
    return fw_cold_draw(fw_rng_view(st), x.c, 0, PG_D_RESAMPLE, upd, r, 0, 0.0, 0.0);
}
__device__ __forceinline__ double fw_u7(const PgState& st, const FwCx& x, unsigned long long upd) {
    const int buf = (int)(upd % 3ull);
    if (x.f->d_ok[buf][FW_DR]) return x.f->d_u7[buf];
    return fw_cold_draw(fw_rng_view(st), x.c, 0, PG_D_SELECT, upd, 0, 0, 0.0, 0.0);
}

// ---- per-lane job view ------------------------------------------------------------------------
struct FwJob {
    int slot, node, var;
    unsigned seg_off, len;
    float lo, hi, thr;
    double split;
    bool live;
};

// Compact live jobs (order preserved) into the GLOBAL job arrays of parity `pb` (read by the grid pass) and,
// when `to_cache`, into the shared job cache + slot_job + the control block (the jobs of the update in progress).
// Returns the new J; *ngroups_out = number of index segments the jobs span.
__device__ __forceinline__ int fw_write_jobs(const PgState& st, const FwCx& x, const FwJob& jb, int lane, bool write_all,
                                             int pb, bool to_global, bool to_cache, int* ngroups_out) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const size_t jo = (size_t)pb * st.S;
    const unsigned live = __ballot_sync(PG_FULL, jb.live);
    const int J = __popc(live);
    const int pos = __popc(live & ((1u << lane) - 1u));
    if (to_cache && lane < st.S) f->hsjob[lane] = -1;
    __syncwarp();
    if (jb.live) {
        if (to_global) {
            st.job_slot[jo + pos] = jb.slot; st.job_node[jo + pos] = jb.node; st.job_var[jo + pos] = jb.var;
            st.job_seg_off[jo + pos] = jb.seg_off; st.job_len[jo + pos] = jb.len;
            st.job_order[jo + pos] = pos;
            if (write_all) {
                st.job_thr32[jo + pos] = jb.thr; st.job_split[jo + pos] = jb.split;
                st.job_lo[jo + pos] = jb.lo; st.job_hi[jo + pos] = jb.hi;
            }
        }
        if (to_cache) {
            f->jslot[pos] = jb.slot; f->jnode[pos] = jb.node; f->jvar[pos] = jb.var;
            f->jseg[pos] = jb.seg_off; f->jlen[pos] = jb.len;
            if (write_all) { f->jthr[pos] = jb.thr; f->jsplit[pos] = jb.split; }
            f->hsjob[jb.slot] = pos;
        }
    }
    const int first = __ffs(live) - 1;
    const unsigned off0 = __shfl_sync(PG_FULL, jb.seg_off, first < 0 ? 0 : first);
    const unsigned len0 = __shfl_sync(PG_FULL, jb.len, first < 0 ? 0 : first);
    const bool same = !jb.live || (jb.seg_off == off0 && jb.len == len0);
    const bool one = __all_sync(PG_FULL, same);
    int G = J > 0 ? 1 : 0;
    if (one) {
        if (lane == 0 && to_global) { int* gf = st.grp_first + (size_t)pb * (st.S + 1); gf[0] = 0; gf[J > 0 ? 1 : 0] = J; }
    } else {
        __threadfence_block();
        __syncwarp();
        if (lane == 0)   // rare (several distinct survivors): generic grouping on the global arrays
            G = pg_group_jobs_arrays(st.job_seg_off + jo, st.job_len + jo, st.job_order + jo, st.grp_first + (size_t)pb * (st.S + 1), J);
        G = __shfl_sync(PG_FULL, G, 0);
    }
    if (to_cache && lane == 0) { c->n_jobs = J; c->n_groups = G; }
    __syncwarp();
    *ngroups_out = G;
    return J;
}

// Group the first JT jobs of the cached job list (f->jseg / f->jlen) by index segment, groups in first-seen order, jobs of
// a group in job order; writes job_order / grp_first of buffer parity pb (same result as pg_group_jobs_arrays, computed
// by the warp from shared memory instead of one lane walking global arrays).  Returns the number of groups.
__device__ __forceinline__ int fw_group_jobs(const PgState& st, const FwCx& x, int JT, int lane, int pb, unsigned long long* stats_bytes) {
    PgFastSm* f = x.f;
    const bool in = lane < JT;
    const unsigned seg = in ? f->jseg[lane] : 0xFFFFFFF0u - (unsigned)lane, len = in ? f->jlen[lane] : 0u;
    const unsigned m = __match_any_sync(PG_FULL, seg) & __match_any_sync(PG_FULL, len) & __ballot_sync(PG_FULL, in);
    const int leader = in ? __ffs(m) - 1 : 31;
    const unsigned leaders = __ballot_sync(PG_FULL, in && leader == lane);
    const int G = __popc(leaders);
    const int g = __popc(leaders & ((1u << leader) - 1u));              // my group: rank of my leader among the leaders
    int offset = 0;                                                     // jobs in the groups before mine
    {
        unsigned rest = leaders;
#pragma unroll 1
        while (rest) {
            const int l = __ffs(rest) - 1;
            rest &= rest - 1u;
            const int cnt = __popc(__shfl_sync(PG_FULL, m, l));
            if (l < leader) offset += cnt;
        }
    }
    const size_t jo = (size_t)pb * st.S;
    int* gf = st.grp_first + (size_t)pb * (st.S + 1);
    if (in) {
        st.job_order[jo + offset + __popc(m & ((1u << lane) - 1u))] = lane;
        if (leader == lane) gf[g] = offset;
    }
    if (lane == 0) gf[G] = JT;
    // bytes a statistics pass moves for these groups: (idx 4 + A 8 + y 4) L + 4 L per job of the group
    unsigned long long bm = (in && leader == lane) ? (16ull + 4ull * (unsigned long long)__popc(m)) * (unsigned long long)len : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bm += __shfl_xor_sync(PG_FULL, bm, o);
    *stats_bytes = bm;
    return G;
}

__device__ __forceinline__ int fw_unique(int var, bool live, int lane) {
    const unsigned lv = __ballot_sync(PG_FULL, live);
    const unsigned m = __match_any_sync(PG_FULL, live ? var : -1 - lane);
    const bool leader = live && ((__ffs(m & lv) - 1) == lane);
    return __popc(__ballot_sync(PG_FULL, leader));
}

// splitting.rs:39-56 for one job lane of a non-root round (u3 precomputed when available)
__device__ __forceinline__ void fw_decide(const PgState& st, const FwCx& x, FwJob& jb) {
    PgCtrl* c = x.c;
    if (!jb.live) return;
    const int r = c->round;
    double* tr = pg_trace_slot(st, c, r, jb.slot);
    if (tr) { tr[9] = (double)jb.lo; tr[10] = (double)jb.hi; }
    if (jb.lo >= jb.hi) { if (tr) tr[0] = 3; jb.live = false; return; }
    if (2 * jb.node + 2 >= st.maxn) { c->error = PG_ERR_DEPTH; jb.live = false; return; }
    const int buf = (int)(c->upd % 3ull);
    bool done = false;
    if (r < FW_DR && x.f->d_ok[buf][r]) {
        const double lo = (double)jb.lo, hi = (double)jb.hi;
        const double v = __dadd_rn(__dmul_rn(x.f->d_u3[buf][r][jb.slot], hi - lo), lo);
        if (v < hi) { jb.split = v; done = true; }
    }
    if (!done) jb.split = fw_cold_draw(fw_rng_view(st), c, 2, PG_D_SPLIT, c->upd, r, jb.slot, (double)jb.lo, (double)jb.hi);
    jb.thr = pg_thr32(jb.split);
    if (tr) tr[3] = jb.split;
}

// ---- sub-steps (warp 0; every lane calls) -------------------------------------------------------

__device__ __forceinline__ int fw_begin_step(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    if (lane == 0) {
        const double frac = c->tune ? st.batch_tune : st.batch_post;
        long long b = (long long)floor(frac * (double)st.m + 0.5);
        if (b < 1) b = 1;
        if (b > st.m) b = st.m;
        c->batch = (int)b; c->k = 0; c->t_add = -1; c->t_add_internal = 0;
        c->info_loglik = 0.0; c->info_acc = 0; c->info_n = 0;
        if (st.family == PG_GAUSSIAN) {
            c->inv_s2 = 1.0 / (st.lik_sigma * st.lik_sigma);
            c->Cconst = -(double)st.n_global * (log(st.lik_sigma) + 0.5 * log(6.283185307179586476925286766559));
        } else { c->inv_s2 = 1.0; c->Cconst = 0.0; }
    }
    __syncwarp();
    return SN_BEGIN_UPDATE;
}

// Round-0 proposals of tree update `upd` (precomputed by the helper warps) as a job lane view.
__device__ __forceinline__ FwJob fw_root_job(const PgState& st, const FwCx& x, unsigned long long upd, int lane, bool* passed, int* status_out) {
    PgFastSm* f = x.f;
    const int buf = (int)(upd % 3ull);
    FwJob jb;
    jb.live = false; jb.slot = lane; jb.node = 0; jb.var = 0; jb.seg_off = PG_SEG_IDENTITY; jb.len = (unsigned)st.n;
    jb.lo = jb.hi = jb.thr = 0.f; jb.split = 0.0;
    *passed = false;
    *status_out = 0;
    if (lane < st.S) {
        const int status = f->d0_status[buf][lane];
        jb.var = f->d_var[buf][0][lane];
        jb.lo = f->d0_lo[buf][lane]; jb.hi = f->d0_hi[buf][lane]; jb.thr = f->d0_thr[buf][lane]; jb.split = f->d0_split[buf][lane];
        *passed = status != 1;
        jb.live = status == 5;
        *status_out = status;
    }
    return jb;
}

// Option prune_dead_proposals.  Stale-weight rule (Q1): in round 0 a slot that proposes nothing keeps the
// log-weight 0.0 while a slot that grew gets its particle's log-likelihood.  For the Gaussian families that log-likelihood
// is at most Cconst = -n (log sigma + 0.5 log 2 pi) whatever the data (the leaf terms are bounded by half the residual sum
// of squares, by Cauchy-Schwarz), so when Cconst < -800 every grown particle's normalised weight is exp(w - 0) = 0.0
// EXACTLY and only the resampling pointer overrunning the end of the weights can still select it.  This checks, with the
// reference's arithmetic (smc.rs:257-268, resampling.rs:24-44), that no grown slot is anybody's ancestor: then the update
// ends as a stump whatever the proposals' statistics are, and they need not be computed.
__device__ __forceinline__ bool fw_round0_all_dead(const PgState& st, const FwCx& x, unsigned long long upd, int lane, bool live) {
    PgCtrl* c = x.c;
    const int S = st.S;
    if (!st.prune || st.family != PG_GAUSSIAN || !(c->Cconst < -800.0) || st.tr_slot || st.tr_round || st.tr_upd) return false;
    const unsigned lm = __ballot_sync(PG_FULL, live);
    if (lm == 0u || __popc(lm) >= S) return false;          // nothing to prune / every slot grows: weights matter
    const bool act = lane < S;
    const double e = act ? (live ? 0.0 : 1.0) : 0.0;          // exp(w - max): max = 0.0 (a non-growing slot), exp(-1e6) = 0.0
    const double sum = fw_sum_seq(e, S);
    const double wn = e / sum;
    const double u6 = fw_u6(st, x, upd, 0);
    const double target = ((double)lane + u6) / (double)S;
    int cidx = 0;
    double cum = 0.0;
#pragma unroll 1
    for (int k = 0; k < S; ++k) {
        if (cum < target) cidx = k + 1;
        cum += __shfl_sync(PG_FULL, wn, k);
    }
    const int anc = act ? (cidx == 0 ? 0 : cidx - 1) : 0;
    const bool anc_live = (lm >> anc) & 1u;
    return !__any_sync(PG_FULL, act && anc_live);
}

// The same argument in rounds >= 1 (option prune_dead_proposals): a slot that does not grow keeps its previous NORMALISED
// weight (in [0, 1]) as log-weight, so the maximum is >= 0 and every growing particle (log-likelihood <= Cconst < -800)
// gets exp(w - max) = 0.0 exactly.  True when, with those exact zeros, no growing slot is anybody's ancestor.
__device__ __forceinline__ bool fw_round_all_dead(const PgState& st, const FwCx& x, int lane, bool live) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int S = st.S;
    if (!st.prune || st.family != PG_GAUSSIAN || !(c->Cconst < -800.0) || st.tr_slot || st.tr_round || st.tr_upd) return false;
    if (!(0.5 * c->T_rr * c->inv_s2 < 1.0e13)) return false;
    const unsigned lm = __ballot_sync(PG_FULL, live);
    if (lm == 0u || __popc(lm) >= S) return false;
    const bool act = lane < S;
    const double w = act ? (live ? -INFINITY : f->hw[lane]) : -INFINITY;
    const double mx = fw_max(w);
    const double e = act ? exp(w - mx) : 0.0;
    const double sum = fw_sum_seq(e, S);
    const double wn = e / sum;
    const double u6 = fw_u6(st, x, c->upd, c->round);
    const double target = ((double)lane + u6) / (double)S;
    int cidx = 0;
    double cum = 0.0;
#pragma unroll 1
    for (int k = 0; k < S; ++k) {
        if (cum < target) cidx = k + 1;
        cum += __shfl_sync(PG_FULL, wn, k);
    }
    const int anc = act ? (cidx == 0 ? 0 : cidx - 1) : 0;
    return !__any_sync(PG_FULL, act && ((lm >> anc) & 1u));
}

// smc.rs:42-53 + round 0 of smc.rs:61-86: fresh particles of the update that becomes current.  The precomputed
// root proposals become the job list: written to the global job arrays unless a speculative root pass already did.
__device__ __forceinline__ void fw_start_update(const PgState& st, const FwCx& x, int lane, bool jobs_published) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int S = st.S;
    if (lane == 0) {
        const int t = (c->next_tree_idx + c->k) % st.m;
        c->t_cur = t; c->t_sub = t; c->round = 0; c->acc_update = 0; c->pool_top = 0;
    }
    __syncwarp();
    const unsigned long long upd = c->upd;
    const int buf = (int)(upd % 3ull);
    fw_wait_draws(st, f, upd);
    bool passed;
    int status;
    const FwJob jb = fw_root_job(st, x, upd, lane, &passed, &status);
    f->mm_valid[0][lane] = 0; f->mm_valid[1][lane] = 0; f->sc_valid[0][lane] = 0; f->sc_valid[1][lane] = 0;
    if (lane == 0) f->es_mask = 0;
    if (lane == 0) { f->n_mq = 0; f->pf_n = 0; f->from_cache = 0; f->pred_valid = 0; f->pred_ok = 0; f->pred_sel = 0; f->pred_tried = 0; f->sl_valid = 0; }
    if (lane < S) {
        const int s = lane;
        f->hw[s] = 0.0;       // smc.rs:53: once per update, never reset (Q1)
        f->hmut[s] = 0;
        if (!f->d_ok[buf][0] || status == 6) c->error = PG_ERR_TAPE;
        if (status == 7) c->error = PG_ERR_DEPTH;
        double* tr = pg_trace_slot(st, c, 0, s);
        if (tr) {
#pragma unroll 1
            for (int q = 0; q < PG_TS_FIELDS; ++q) tr[q] = pg_nan();
            tr[0] = status == 5 ? 1.0 : (double)status; tr[1] = 0.0; tr[2] = passed ? (double)jb.var : -1.0;
            if (passed) { tr[9] = (double)jb.lo; tr[10] = (double)jb.hi; }
            if (status == 5) tr[3] = jb.split;
        }
    }
    const int J0 = __popc(__ballot_sync(PG_FULL, passed));
    int G;
    const bool lazy = fw_round0_all_dead(st, x, upd, lane, jb.live);   // same decision the speculative issue took (fw_prestage)
    const int J = fw_write_jobs(st, x, jb, lane, true, (int)(upd & 1ull), !jobs_published && !lazy, true, &G);
    const int uniq = fw_unique(jb.var, jb.live, lane);
    if (lane == 0) {
        const unsigned long long n = (unsigned long long)st.n;
        const int cols = (lazy ? 0 : uniq) + c->t_add_internal + st.f_internal[c->t_sub];
        c->bytes_model += n * (8 + 8 + 4) + 4ull * n * (unsigned long long)cols;
        c->bytes_contract += 20ull * n + (lazy ? 0ull : 4ull * n * (unsigned long long)J0 + 16ull * n * (unsigned long long)J);
        c->phase = PH_ROOT; c->n_phases += 1;
        f->lazy_J = lazy ? J : 0;
        if (lazy) c->n_jobs = 0;                             // the root pass only commits / opens the trees and sums the residuals
    }
    __syncwarp();
}

__device__ __forceinline__ int fw_after_stats(const PgState& st, const FwCx& x, int lane, int phase);
__device__ __forceinline__ void fw_issue_part(const PgState& st, const FwCx& x, int lane);
__device__ __forceinline__ int fw_round_stats(const PgState& st, const FwCx& x, int lane, FwJob jb);
// (defined here, ahead of its first use: a call the front end cannot inline passes `st` by reference and puts the whole
// kernel parameter block on the stack)
__device__ __forceinline__ void fw_issue(const PgState& st, const FwCx& x, int lane, int phase, int t_add, int t_sub,
                                         int n_jobs, int n_groups, int n_pjobs, int flags) {
    PgFastSm* f = x.f;
    const unsigned g = f->issued + 1u;                       // this command's generation; its buffer: the one generation g - 2 used
    if (lane < 16) {
        int v = reinterpret_cast<const int*>(&f->lc)[lane];
        if (lane == 0) v = phase; else if (lane == 4) v = t_add; else if (lane == 5) v = t_sub; else if (lane == 7) v = n_jobs;
        else if (lane == 8) v = n_groups; else if (lane == 9) v = n_pjobs; else if (lane == 15) v = flags;
        st.ctrl->cmd[g & 1u][lane] = v;
    }
#ifdef PG_DIAG
    if (lane == 0) { st.ctrl->dbg[1] = 0ull; st.ctrl->dbg[2] = 0ull; st.ctrl->dbg[3] = 0ull; st.ctrl->dbg[0] = pg_globaltimer(); }
#endif
    __syncwarp();                                            // the lanes' stores are ordered before lane 0's release store
    if (lane == 0) {
        f->issued = g;
        f->out_phase[g & 1u] = phase; f->out_J[g & 1u] = n_jobs; f->out_pb[g & 1u] = (flags >> 2) & 1;
        pg_st_release(&st.ctrl->bar_gen, g);
    }
    __syncwarp();
}


// rounds >= 1: pop the next expandable node of every slot, depth-prior test, split variable (smc.rs:61-86 front half)
__device__ __forceinline__ int fw_draw_round(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int cur = c->cur, r = c->round, S = st.S;
    const unsigned long long upd = c->upd;
    const int buf = (int)(upd % 3ull);
    const bool pre = r < FW_DR && f->d_ok[buf][r];
    FwJob jb;
    jb.live = false; jb.slot = lane; jb.node = -1; jb.var = 0; jb.seg_off = 0; jb.len = 0; jb.lo = jb.hi = jb.thr = 0.f; jb.split = 0.0;
    if (lane < S) {
        const int s = lane;
        f->hmut[s] = 0;
        double* tr = pg_trace_slot(st, c, r, s);
        if (tr) {
#pragma unroll 1
            for (int q = 0; q < PG_TS_FIELDS; ++q) tr[q] = pg_nan();
            tr[0] = 0;
        }
        const int ss = pg_six(st, cur, s);
        const int qh = f->hqh[ss], qt = f->hqt[ss];
        // slots that still hold the untouched copy of the single grown particle of round 0: node / count / segment of the
        // leaf in front of the queue come from shared memory instead of two dependent reads of the slot tables
        const bool shallow = f->sl_valid && f->hlin[s] == f->sl_lin && f->hsize[ss] == 3 && qt == 2 && qh < 2;
        if (qh < qt) {
            const int node = shallow ? qh + 1 : st.pt_queue[pg_tix(st, cur, s, qh)];
            f->hqh[ss] = qh + 1;
            if (tr) { tr[0] = 1; tr[1] = (double)node; tr[2] = -1; }
            const double u1 = pre ? f->d_u1[buf][r][s] : fw_cold_draw(fw_rng_view(st), c, 0, PG_D_GROW, upd, r, s, 0.0, 0.0);
            if (!(x.thr[pg_depth_of(node)] > u1)) {
                const size_t iN = pg_tix(st, cur, s, node);
                const unsigned cnt = shallow ? f->sl_cnt[node - 1] : st.pt_cnt[iN];
                if (cnt == 0u) { if (tr) tr[0] = 2; }
                else {
                    int var;
                    if (pre) var = f->d_var[buf][r][s];
                    else {
                        const double u2 = fw_cold_draw(fw_rng_view(st), c, 0, PG_D_VAR, upd, r, s, 0.0, 0.0);
                        if (x.prior) var = fw_feature_from_probs(st, x.prior, u2);
                        else { var = (int)(u2 * (double)st.p); if (var > st.p - 1) var = st.p - 1; }
                    }
                    if (tr) tr[2] = (double)var;
                    jb.live = true; jb.node = node; jb.var = var;
                    jb.len = shallow ? f->sl_len[node - 1] : st.pt_len[iN];
                    jb.seg_off = shallow ? f->sl_seg[node - 1] : st.pt_seg_off[iN];
                }
            }
        }
    }
    const int pb = (int)(upd & 1ull);
    int G;
    if (lane == 0) f->from_cache = 0;
    __syncwarp();
    if (!__any_sync(PG_FULL, jb.live)) {
        (void)fw_write_jobs(st, x, jb, lane, false, pb, false, true, &G);
        return SN_FINISH_ROUND;
    }
    // ---- look-ahead caches: min/max computed by the partition pass, statistics prefetched by the previous stats pass
    const int ri = r - 1;
    const bool mm_hit = (r == 1 || r == 2) && (!jb.live || (f->mm_valid[ri][lane] && f->mm_seg[ri][lane] == jb.seg_off && f->mm_var[ri][lane] == jb.var));
    if (!__all_sync(PG_FULL, mm_hit)) {
        // the regular way: a min/max pass over the nodes' index segments
        const int J = fw_write_jobs(st, x, jb, lane, false, pb, true, true, &G);
        if (lane == 0) {
            const int* jord = st.job_order + (size_t)pb * st.S;
            const int* gf = st.grp_first + (size_t)pb * (st.S + 1);
            unsigned long long bm = 0, bc = 0;
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                const unsigned long long L = f->jlen[jord[gf[g]]];
                bm += 4ull * L + 4ull * L * (unsigned long long)(gf[g + 1] - gf[g]);
            }
#pragma unroll 1
            for (int j = 0; j < J; ++j) bc += 8ull * f->jlen[j];
            c->bytes_model += bm; c->bytes_contract += bc;
            c->phase = PH_MINMAX; c->n_phases += 1;
        }
        __syncwarp();
        return SN_NONE;
    }
    if (jb.live) { jb.lo = f->mm_lo[ri][lane]; jb.hi = f->mm_hi[ri][lane]; }
    fw_decide(st, x, jb);                                  // splitting.rs:39-56 (may drop the job: min >= max)
    if (!__any_sync(PG_FULL, jb.live)) {
        (void)fw_write_jobs(st, x, jb, lane, true, pb, false, true, &G);
        return SN_FINISH_ROUND;
    }
    if (fw_round_all_dead(st, x, lane, jb.live)) {
        // pruned: the proposals are accepted grows (they count) of particles that cannot be resampled; their statistics,
        // leaf values and weights are never looked at (the same normalisation runs below with exact zeros)
        if (jb.live) { f->hw[lane] = -INFINITY; f->hmut[lane] = 1; }
        __syncwarp();
        return SN_FINISH_ROUND;
    }
    if (f->es_pending) {
        // the early statistics pass (queued behind the partition pass) is computing exactly these proposals' statistics:
        // park the round here until it completes (pg_control_fast resumes it through fw_round_stats)
        f->wj_live[lane] = jb.live ? 1 : 0; f->wj_node[lane] = jb.node; f->wj_var[lane] = jb.var; f->wj_seg[lane] = jb.seg_off;
        f->wj_len[lane] = jb.len; f->wj_lo[lane] = jb.lo; f->wj_hi[lane] = jb.hi; f->wj_thr[lane] = jb.thr; f->wj_split[lane] = jb.split;
        if (lane == 0) f->es_resume = SN_DRAW_ROUND;
        __syncwarp();
        return SN_WAIT_EARLY;
    }
    return fw_round_stats(st, x, lane, jb);
}

// Second half of a round >= 1 (split values decided): statistics from the cache, or a statistics pass.
__device__ __forceinline__ int fw_round_stats(const PgState& st, const FwCx& x, int lane, FwJob jb) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int cur = c->cur, r = c->round, S = st.S;
    const unsigned long long upd = c->upd;
    const int buf = (int)(upd % 3ull);
    const int pb = (int)(upd & 1ull);
    const int ri = r - 1;
    int G;
    const bool cacheable = r == 1 || r == 2;
    const int rc = cacheable ? ri : 0;
    const bool sc_hit = !jb.live || (cacheable && f->sc_valid[rc][lane] && f->sc_seg[rc][lane] == jb.seg_off && f->sc_var[rc][lane] == jb.var && f->sc_thr[rc][lane] == jb.thr);
    if (__all_sync(PG_FULL, sc_hit)) {
        // the statistics were prefetched too: no pass at all for this round
        const int J = fw_write_jobs(st, x, jb, lane, true, pb, false, true, &G);
        if (lane == 0) f->from_cache = 1;       // per-block counts of these jobs sit at the prefetch jobs' indices (sc_job)
        if (lane < J) {
            const int s = f->jslot[lane];
            f->rCnt[lane] = f->sc_cnt[rc][s]; f->rCntL[lane] = f->sc_cntL[rc][s]; f->rSo[lane] = f->sc_So[rc][s]; f->rSr[lane] = f->sc_Sr[rc][s];
        }
        __syncwarp();
        if (lane == 0) { unsigned long long bc = 0; for (int j = 0; j < J; ++j) bc += 28ull * f->jlen[j]; c->bytes_contract += bc; }
        return fw_after_stats(st, x, lane, PH_STATS);
    }
    if (lane == 0 && cacheable && ((f->es_mask >> rc) & 1)) st.ctrl->n_es_miss += 1ull;   // an early statistics pass covered this round, and yet: a regular pass
    // statistics pass for this round's jobs ...
    const int J = fw_write_jobs(st, x, jb, lane, true, pb, true, true, &G);
    // ... carrying, for round 1, the round-2 proposals of every slot (node 2: the other child) as prefetch jobs
    int Jp = 0;
    unsigned long long grouped_bytes = 0ull;
    if (r == 1 && G == 1) {
        const int buf2 = buf;
        bool want = lane < S && f->mm_valid[1][lane] && f->d_ok[buf2][2] && f->mm_lo[1][lane] < f->mm_hi[1][lane];
        double split2 = 0.0;
        float thr2 = 0.f;
        if (want) {
            const double lo = (double)f->mm_lo[1][lane], hi = (double)f->mm_hi[1][lane];
            split2 = __dadd_rn(__dmul_rn(f->d_u3[buf2][2][lane], hi - lo), lo);
            want = split2 < hi;
            thr2 = pg_thr32(split2);
        }
        const unsigned wm = __ballot_sync(PG_FULL, want);
        Jp = __popc(wm);
        if (J + Jp > S) { Jp = 0; }
        else if (want) {
            const int q = __popc(wm & ((1u << lane) - 1u));
            const size_t jo = (size_t)pb * st.S + (size_t)(J + q);
            const unsigned len2 = (unsigned)st.n - 0u;
            (void)len2;
            st.job_var[jo] = f->mm_var[1][lane]; st.job_thr32[jo] = thr2; st.job_seg_off[jo] = f->mm_seg[1][lane];
            const unsigned len2v = st.pt_len[pg_tix(st, cur, lane, 2)];
            st.job_len[jo] = len2v;
            st.job_slot[jo] = lane; st.job_node[jo] = 2; st.job_split[jo] = split2;
            f->jseg[J + q] = f->mm_seg[1][lane]; f->jlen[J + q] = len2v;
            f->pf_slot[q] = lane;
            f->sc_seg[1][lane] = f->mm_seg[1][lane]; f->sc_var[1][lane] = f->mm_var[1][lane]; f->sc_thr[1][lane] = thr2;
        }
        __threadfence_block();
        __syncwarp();
        if (Jp > 0) G = fw_group_jobs(st, x, J + Jp, lane, pb, &grouped_bytes);
    }
    if (lane == 0) {
        f->pf_n = Jp;
        const size_t jo = (size_t)pb * st.S;
        const int* jord = st.job_order + jo;
        const int* gf = st.grp_first + (size_t)pb * (st.S + 1);
        unsigned long long bm = grouped_bytes, bc = 0;
        if (Jp == 0) {
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                const unsigned long long L = f->jlen[jord[gf[g]]];
                bm += (4ull + 8ull + 4ull) * L + 4ull * L * (unsigned long long)(gf[g + 1] - gf[g]);
            }
        }
#pragma unroll 1
        for (int j = 0; j < J; ++j) bc += 28ull * f->jlen[j];
        c->bytes_model += bm; c->bytes_contract += bc;
        c->n_groups = G;
        c->phase = PH_STATS; c->n_phases += 1;
    }
    __syncwarp();
    return SN_NONE;
}

__device__ __forceinline__ int fw_after_minmax(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int J0 = c->n_jobs;
    FwJob jb;
    jb.live = lane < J0;
    jb.slot = jb.live ? f->jslot[lane] : 0; jb.node = jb.live ? f->jnode[lane] : 0; jb.var = jb.live ? f->jvar[lane] : 0;
    jb.seg_off = jb.live ? f->jseg[lane] : 0u; jb.len = jb.live ? f->jlen[lane] : 0u;
    jb.thr = 0.f; jb.split = 0.0;
    jb.lo = jb.live ? pg_key2f(f->rMin[lane]) : 0.f; jb.hi = jb.live ? pg_key2f(f->rMax[lane]) : 0.f;
    fw_decide(st, x, jb);
    const int pb = (int)(c->upd & 1ull);
    int G;
    const int J = fw_write_jobs(st, x, jb, lane, true, pb, true, true, &G);
    if (J == 0) return SN_FINISH_ROUND;
    if (lane == 0) {
        const int* jord = st.job_order + (size_t)pb * st.S;
        const int* gf = st.grp_first + (size_t)pb * (st.S + 1);
        unsigned long long bm = 0, bc = 0;
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const unsigned long long L = f->jlen[jord[gf[g]]];
            bm += (4ull + 8ull + 4ull) * L + 4ull * L * (unsigned long long)(gf[g + 1] - gf[g]);
        }
#pragma unroll 1
        for (int j = 0; j < J; ++j) bc += 20ull * f->jlen[j];
        c->bytes_model += bm; c->bytes_contract += bc;
        c->phase = PH_STATS; c->n_phases += 1;
    }
    __syncwarp();
    return SN_NONE;
}

// particle.rs:91-148 bookkeeping half + smc.rs:76-81 for rounds >= 1 (mirrors pg_apply_mutation on the shared scalars)
// `cur`/`s`: the table (half, slot) the mutation is written to — the proposing slot itself, or (deferred mutations) the
// fresh copy a surviving descendant received; `record` = also set the slot's weight / trace (immediate application).
__device__ __forceinline__ void fw_apply(const PgState& st, const FwCx& x, int cur, int s, bool record, int j, double vL, double vR, double gL, double gR,
                                         unsigned cL, unsigned cR, double SoL, double SoR, double SrL, double SrR, unsigned lenL) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int node = f->jnode[j];
    const int L = 2 * node + 1, R = 2 * node + 2;
    const size_t iN = pg_tix(st, cur, s, node), iL = pg_tix(st, cur, s, L), iR = pg_tix(st, cur, s, R);
    const int ss = pg_six(st, cur, s);
    const int old_size = f->hsize[ss];
    const double g_node = st.pt_g[iN];
#pragma unroll 1
    for (int q = old_size; q < R + 1; ++q) {          // tree.rs:105-110 filler nodes
        const size_t iq = pg_tix(st, cur, s, q);
        st.pt_split_var[iq] = -1; st.pt_split_val[iq] = pg_nan(); st.pt_thr32[iq] = 0.0f; st.pt_leaf_val[iq] = 0.0;
        st.pt_seg_off[iq] = PG_SEG_NONE; st.pt_cnt[iq] = 0u; st.pt_len[iq] = 0u; st.pt_So[iq] = 0.0; st.pt_Sr[iq] = 0.0; st.pt_g[iq] = 0.0;
    }
    st.pt_split_var[iN] = f->jvar[j];
    st.pt_split_val[iN] = f->jsplit[j];
    st.pt_thr32[iN] = f->jthr[j];
    st.pt_leaf_val[iN] = pg_nan();
    st.pt_split_var[iL] = -1; st.pt_split_val[iL] = pg_nan(); st.pt_thr32[iL] = 0.0f; st.pt_leaf_val[iL] = vL;
    st.pt_split_var[iR] = -1; st.pt_split_val[iR] = pg_nan(); st.pt_thr32[iR] = 0.0f; st.pt_leaf_val[iR] = vR;
    st.pt_seg_off[iL] = PG_SEG_NONE; st.pt_seg_off[iR] = PG_SEG_NONE;
    st.pt_cnt[iL] = cL; st.pt_cnt[iR] = cR;
    st.pt_len[iL] = lenL; st.pt_len[iR] = st.pt_len[iN] - lenL;
    st.pt_So[iL] = SoL; st.pt_So[iR] = SoR; st.pt_Sr[iL] = SrL; st.pt_Sr[iR] = SrR;
    st.pt_g[iL] = gL; st.pt_g[iR] = gR;
    if (R + 1 > old_size) f->hsize[ss] = R + 1;
    const double G = (f->hG[ss] - g_node) + (gL + gR);
    f->hG[ss] = G;
    const int qt = f->hqt[ss];
    st.pt_queue[pg_tix(st, cur, s, qt)] = L;           // particle.rs:146-147
    st.pt_queue[pg_tix(st, cur, s, qt + 1)] = R;
    f->hqt[ss] = qt + 2;
    if (record) {
        f->hmut[s] = 1;
        const double w = c->C0 + G;                    // smc.rs:89-95
        f->hw[s] = w;
        double* tr = pg_trace_slot(st, c, c->round, s);
        if (tr) { tr[0] = 4; tr[4] = vL; tr[5] = vR; tr[6] = (double)cL; tr[7] = (double)cR; tr[8] = w; tr[11] = SoL; }
    }
}

// Leaf values of one job (smc.rs:219-239) and, for the Gaussian families, the leaf terms of the log-likelihood.
struct FwVals { double vL, vR, gL, gR, SoL, SoR, SrL, SrR; unsigned cL, cR; };
__device__ __forceinline__ FwVals fw_values(const PgState& st, const FwCx& x, int j, unsigned nc, double nSo, double nSr) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    FwVals v;
    const int s = f->jslot[j];
    const int r = c->round;
    const int buf = (int)(c->upd % 3ull);
    const bool pre = r < FW_DR && f->d_ok[buf][r];
    v.cL = f->rCnt[j]; v.cR = nc - v.cL;
    v.SoL = f->rSo[j]; v.SoR = nSo - v.SoL;
    v.SrL = f->rSr[j]; v.SrR = nSr - v.SrL;
    const double zL = pre ? f->d_zL[buf][r][s] : fw_cold_draw(fw_rng_view(st), c, 1, PG_D_ZLEFT, c->upd, r, s, 0.0, 0.0);   // drawn for both children, left first
    const double zR = pre ? f->d_zR[buf][r][s] : fw_cold_draw(fw_rng_view(st), c, 1, PG_D_ZRIGHT, c->upd, r, s, 0.0, 0.0);
    v.vL = v.cL == 0u ? zL * st.leaf_sigma : pg_mul_add(zL, st.leaf_sigma, v.SoL / ((double)v.cL * st.m_f));   // cL*m is exact
    v.vR = v.cR == 0u ? zR * st.leaf_sigma : pg_mul_add(zR, st.leaf_sigma, v.SoR / ((double)v.cR * st.m_f));
    v.gL = pg_gauss_g(v.vL, v.SrL, v.cL, c->inv_s2);
    v.gR = pg_gauss_g(v.vR, v.SrR, v.cR, c->inv_s2);
    return v;
}

__device__ __forceinline__ void fw_bern_bytes(const PgState& st, const FwCx& x, int J) {
    PgCtrl* c = x.c;
    const int pb = (int)(c->upd & 1ull);
    const int* jord = st.job_order + (size_t)pb * st.S;
    const int* gf = st.grp_first + (size_t)pb * (st.S + 1);
    unsigned long long bm = 0, bc = 0;
#pragma unroll 1
    for (int g = 0; g < c->n_groups; ++g) {
        const int j0 = jord[gf[g]];
        const unsigned long long L = x.f->jlen[j0];
        const unsigned long long fixed = x.f->jseg[j0] == PG_SEG_IDENTITY ? 12ull : 16ull;
        bm += fixed * L + 4ull * L * (unsigned long long)(gf[g + 1] - gf[g]);
    }
#pragma unroll 1
    for (int j = 0; j < J; ++j) bc += 12ull * x.f->jlen[j];
    c->bytes_model += bm; c->bytes_contract += bc;
    c->phase = PH_BERN; c->n_phases += 1;
}

// After PH_STATS, rounds >= 1 (and the Bernoulli completion after PH_BERN)
__device__ __forceinline__ int fw_after_stats(const PgState& st, const FwCx& x, int lane, int phase) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int J = c->n_jobs, cur = c->cur;
    const bool bern = st.family == PG_BERNOULLI_LOGIT;
    if (phase == PH_BERN) {
        if (lane < J) {
            const size_t iN = pg_tix(st, cur, f->jslot[lane], f->jnode[lane]);
            const unsigned cL = f->jr_cL[lane];
            fw_apply(st, x, cur, f->jslot[lane], true, lane, f->jr_vL[lane], f->jr_vR[lane], f->rllL[lane], f->rllR[lane], cL, st.pt_cnt[iN] - cL,
                     f->jr_SoL[lane], st.pt_So[iN] - f->jr_SoL[lane], f->jr_SrL[lane], st.pt_Sr[iN] - f->jr_SrL[lane], f->jr_cLl[lane]);
        }
        __syncwarp();
        return SN_FINISH_ROUND;
    }
    if (lane < J) {
        const size_t iN = pg_tix(st, cur, f->jslot[lane], f->jnode[lane]);
        // node statistics: from the shared-memory mirror when the slot still holds the untouched single lineage of round 0
        const int sj = f->jslot[lane], nj = f->jnode[lane];
        const bool shallow = f->sl_valid && f->hlin[sj] == f->sl_lin && f->hsize[pg_six(st, cur, sj)] == 3 && (nj == 1 || nj == 2);
        const unsigned n_cnt = shallow ? f->sl_cnt[nj - 1] : st.pt_cnt[iN];
        const double n_So = shallow ? f->sl_So[nj - 1] : st.pt_So[iN], n_Sr = shallow ? f->sl_Sr[nj - 1] : st.pt_Sr[iN];
        const FwVals v = fw_values(st, x, lane, n_cnt, n_So, n_Sr);
        f->jr_cL[lane] = v.cL;     // the partition reads the left count later
        f->jr_cLl[lane] = f->rCntL[lane];
        if (bern) {
            f->jr_vL[lane] = v.vL; f->jr_vR[lane] = v.vR; f->jr_SoL[lane] = v.SoL; f->jr_SrL[lane] = v.SrL;
            { const size_t jo = (size_t)(c->upd & 1ull) * st.S; st.job_vL[jo + lane] = v.vL; st.job_vR[jo + lane] = v.vR; }
        } else {
            // Deferred (particle.rs:91-148 is applied lazily): the mutated particle's weight is all resampling needs.  Its
            // table is written only if a copy of it survives (fw_finish_round) — under the stale-weight rule (Q1) it
            // almost never does, and then no slot table moves at all.
            const int s = f->jslot[lane];
            const int ss = pg_six(st, cur, s);
            f->jr_vL[lane] = v.vL; f->jr_vR[lane] = v.vR; f->jr_gL[lane] = v.gL; f->jr_gR[lane] = v.gR;
            f->jr_SoL[lane] = v.SoL; f->jr_SrL[lane] = v.SrL;
            const double G = (f->hG[ss] - (shallow ? f->sl_g[nj - 1] : st.pt_g[iN])) + (v.gL + v.gR);
            const double w = c->C0 + G;                // smc.rs:89-95
            f->hGp[s] = G;
            f->hw[s] = w;
            f->hmut[s] = 1;
            double* tr = pg_trace_slot(st, c, c->round, s);
            if (tr) { tr[0] = 4; tr[4] = v.vL; tr[5] = v.vR; tr[6] = (double)v.cL; tr[7] = (double)v.cR; tr[8] = w; tr[11] = v.SoL; }
        }
    }
    __syncwarp();
    if (bern && J > 0) {
        if (lane == 0) fw_bern_bytes(st, x, J);
        __syncwarp();
        return SN_NONE;
    }
    return SN_FINISH_ROUND;
}

// smc.rs:257-268 + resampling.rs:24-44 on the shared slot scalars; returns this lane's ancestor
__device__ __forceinline__ int fw_normalize_resample(const PgState& st, const FwCx& x, int lane, int* n_mut_out, int* mut_out) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int S = st.S;
    const bool act = lane < S;
    const int mut = act ? f->hmut[lane] : 0;
    const double w = act ? f->hw[lane] : -INFINITY;
    const int n_mut = __popc(__ballot_sync(PG_FULL, mut != 0));
    const double mx = fw_max(w);
    const double e = act ? exp(w - mx) : 0.0;
    const double sum = fw_sum_seq(e, S);
    const double wn = e / sum;
    if (act) f->hw[lane] = wn;          // kept for the next round: stale-weight rule (Q1)
    const double u6 = fw_u6(st, x, c->upd, c->round);
    const double target = ((double)lane + u6) / (double)S;
    int cidx = 0;
    double cum = 0.0;
#pragma unroll 8
    for (int k = 0; k < S; ++k) {          // cum == sum of the first k weights, in order, on every lane
        if (cum < target) cidx = k + 1;     // the reference's pointer passes k
        cum += __shfl_sync(PG_FULL, wn, k);
    }
    const int anc = act ? (cidx == 0 ? 0 : cidx - 1) : 0;
    if (act) f->hanc[lane] = anc;
    if (st.tr_round) {
        if ((long long)c->trace_n < st.tr_u_cap && c->round < st.tr_r_cap) {
            double* tr = st.tr_round + ((size_t)c->trace_n * st.tr_r_cap + c->round) * (size_t)(2 + 2 * S);
            if (lane == 0) { tr[0] = 1; tr[1] = (double)n_mut; }
            if (act) { tr[2 + lane] = wn; tr[2 + S + lane] = (double)anc; }
        } else if (lane == 0) c->trace_overflow = 1;
    }
    *n_mut_out = n_mut;
    *mut_out = mut;
    return anc;
}

// Partition jobs for the distinct surviving ancestors that grew this round; returns K and this lane's child offsets.
__device__ __forceinline__ int fw_plan_partitions(const PgState& st, const FwCx& x, int lane, int anc, bool need,
                                                  unsigned* off_out, unsigned* cntL_out, int* node_out) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const unsigned needm = __ballot_sync(PG_FULL, need);
    const unsigned same = __match_any_sync(PG_FULL, need ? anc : -1 - lane);
    const bool leader = need && ((__ffs(same & needm) - 1) == lane);
    const unsigned leadm = __ballot_sync(PG_FULL, leader);
    const int K = __popc(leadm);
    const int k = __popc(leadm & ((1u << lane) - 1u));
    int j = 0, node = 0;
    unsigned len = 0u, cntL = 0u;
    if (need) { j = f->hsjob[anc]; node = f->jnode[j]; len = f->jlen[j]; cntL = f->jr_cLl[j]; }   // this rank's rows: segment layout is local
    unsigned long long off = 0;
    {
        const unsigned long long mylen = leader ? (unsigned long long)len : 0ull;
        unsigned long long run = c->pool_top;
#pragma unroll 1
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
    if (lane < st.S) st.pj_of_anc[lane] = -1;
    __syncwarp();
    if (leader) {
        st.pj_of_anc[anc] = k;
        const int cj = f->from_cache ? f->sc_job[(c->round == 2) ? 1 : 0][f->jslot[j]] : j;   // where the pass left this job's per-block counts
        st.pj_anc[k] = anc; st.pj_job[k] = cj;
        f->pq_job[k] = cj; f->pq_len[k] = len; f->pq_ident[k] = f->jseg[j] == PG_SEG_IDENTITY ? 1 : 0;
        st.pj_var[k] = f->jvar[j]; st.pj_thr32[k] = f->jthr[j];
        st.pj_src_off[k] = f->jseg[j]; st.pj_len[k] = len; st.pj_cntL[k] = cntL; st.pj_dst_off[k] = (unsigned)off;
    }
    const int lead_lane = need ? (__ffs(same & needm) - 1) : 0;
    *off_out = (unsigned)__shfl_sync(PG_FULL, off, lead_lane);
    *cntL_out = cntL;
    *node_out = node;
    return K;
}

// smc.rs:96-102 for rounds >= 1: clone survivors' tables (only when a table actually changes owner), apply the deferred
// mutations of survivors, plan the lazy partitions
__device__ __forceinline__ int fw_finish_round(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int S = st.S;
    const bool act = lane < S;
    const bool bern = st.family == PG_BERNOULLI_LOGIT;    // Bernoulli mutations are applied immediately (fw_after_stats)
    int n_mut, mut;
    const int anc = fw_normalize_resample(st, x, lane, &n_mut, &mut);
    const int mut_a = __shfl_sync(PG_FULL, mut, anc);
    const int my_lin = act ? f->hlin[lane] : 0;
    const int lin_a = __shfl_sync(PG_FULL, my_lin, anc);
    const bool moves = act && (mut_a != 0 || lin_a != my_lin || (bern && mut != 0));
    if (!__any_sync(PG_FULL, moves)) {
        // Every slot keeps a table identical to its ancestor's: same lineage, nobody's pending mutation survived, and the
        // per-slot scalars (size, queue head / tail, log-lik sum) of one lineage are equal because every slot pops one
        // node per round.  Nothing to copy.
        const int ss = pg_six(st, c->cur, lane);
        const int any_queue = __any_sync(PG_FULL, act && f->hqh[ss] < f->hqt[ss]);
        if (lane == 0) { c->acc_update += n_mut; c->n_pjobs = 0; c->any_queue = any_queue; f->n_mq = 0; }
        __syncwarp();
        if (c->error != PG_OK) return SN_NONE;
        if (any_queue) { if (lane == 0) c->round += 1; __syncwarp(); return SN_DRAW_ROUND; }
        return SN_FINALIZE;
    }
    const int src = c->cur, dst = 1 - c->cur;
    const int fs = pg_six(st, src, anc), ts = pg_six(st, dst, lane);
    const int sz = act ? f->hsize[fs] : 0;
    const int qh = act ? f->hqh[fs] : 0, qt = act ? f->hqt[fs] : 0;
    const double Gs = act ? f->hG[fs] : 0.0;
    const int mxs = fw_imax(sz);
#pragma unroll 1
    for (int node = 0; node < mxs; ++node) {
        if (node < sz) {
            const size_t a = pg_tix(st, src, anc, node), t = pg_tix(st, dst, lane, node);
            st.pt_split_var[t] = st.pt_split_var[a]; st.pt_thr32[t] = st.pt_thr32[a];
            st.pt_split_val[t] = st.pt_split_val[a]; st.pt_leaf_val[t] = st.pt_leaf_val[a];
            st.pt_seg_off[t] = st.pt_seg_off[a]; st.pt_cnt[t] = st.pt_cnt[a]; st.pt_len[t] = st.pt_len[a];
            st.pt_So[t] = st.pt_So[a]; st.pt_Sr[t] = st.pt_Sr[a]; st.pt_g[t] = st.pt_g[a];
            st.pt_queue[t] = st.pt_queue[a];
        }
    }
    __syncwarp();
    if (act) { f->hsize[ts] = sz; f->hqh[ts] = qh; f->hqt[ts] = qt; f->hG[ts] = Gs; f->hlin[lane] = mut_a ? 64 + 32 * (c->round & 0xffff) + anc : lin_a; }
    __syncwarp();
    const bool need = act && mut_a != 0;
    if (need && !bern) {                                   // the ancestor's deferred mutation, on this slot's fresh copy
        const int j = f->hsjob[anc];
        const size_t iN = pg_tix(st, dst, lane, f->jnode[j]);
        const unsigned cL = f->jr_cL[j];
        fw_apply(st, x, dst, lane, false, j, f->jr_vL[j], f->jr_vR[j], f->jr_gL[j], f->jr_gR[j], cL, st.pt_cnt[iN] - cL,
                 f->jr_SoL[j], st.pt_So[iN] - f->jr_SoL[j], f->jr_SrL[j], st.pt_Sr[iN] - f->jr_SrL[j], f->jr_cLl[j]);
    }
    __syncwarp();
    unsigned off, cntL;
    int node;
    const int K = fw_plan_partitions(st, x, lane, anc, need, &off, &cntL, &node);
    if (need) {
        st.pt_seg_off[pg_tix(st, dst, lane, 2 * node + 1)] = off;
        st.pt_seg_off[pg_tix(st, dst, lane, 2 * node + 2)] = off + cntL;
    }
    const int any_queue = __any_sync(PG_FULL, act && f->hqh[ts] < f->hqt[ts]);
    if (lane == 0) { c->acc_update += n_mut; c->cur = dst; c->n_pjobs = K; c->any_queue = any_queue; f->n_mq = 0; }
    __syncwarp();
    if (c->error != PG_OK) return SN_NONE;
    if (K > 0) { fw_issue_part(st, x, lane); return SN_PART_ISSUED; }
    if (any_queue) { if (lane == 0) c->round += 1; __syncwarp(); return SN_DRAW_ROUND; }
    return SN_FINALIZE;
}

// end of a tree update: diagnostics, counters, next update or the step's final pass (kernel.rs:104-117)
__device__ __forceinline__ int fw_end_update(const PgState& st, const FwCx& x, int lane, int depth, int internal, double Wn0) {
    PgCtrl* c = x.c;
    int next = SN_NONE;
    __syncwarp();   // every lane has finished reading this update's counters (trace_n, upd, round) before lane 0 advances them
    if (lane == 0) {
        const int t = c->t_cur;
        c->info_loglik += Wn0;
        c->info_acc += c->acc_update;
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
        }
    }
    __syncwarp();
    return __shfl_sync(PG_FULL, next, 0);
}

// smc.rs:105-122 final selection among the reference stump and the S slots (lane i = particle i)
__device__ __forceinline__ int fw_select(const PgState& st, const FwCx& x, int lane, double Wl, double* Wn0_out) {
    PgCtrl* c = x.c;
    const int P = st.P;
    const bool act = lane < P;
    const double mx = fw_max(act ? Wl : -INFINITY);
    const double e = act ? exp(Wl - mx) : 0.0;
    const double sum = fw_sum_seq(e, P);
    const double Wn = e / sum;
    const double total = fw_sum_seq(act ? Wn : 0.0, P);
    const double chosen = fw_u7(st, x, c->upd) * total;
    int sel = 0;
    {
        double cum = 0.0;
        bool open = true;
#pragma unroll 8
        for (int i = 0; i + 1 < P; ++i) {
            cum += __shfl_sync(PG_FULL, Wn, i);
            if (open && cum <= chosen) sel = i + 1; else open = false;
        }
    }
    if (lane == 0 && (!(total > 0.0) || !(total < INFINITY))) c->error = PG_ERR_WEIGHTS;
    if (st.tr_upd) {
        if ((long long)c->trace_n < st.tr_u_cap) {
            double* tr = st.tr_upd + (size_t)c->trace_n * (size_t)(4 + 2 * P);
            if (lane == 0) { tr[0] = (double)(c->round + 1); tr[1] = (double)sel; tr[2] = (double)c->t_cur; tr[3] = (double)c->acc_update; }
            if (act) { tr[4 + lane] = Wl; tr[4 + P + lane] = Wn; }
        } else if (lane == 0) c->trace_overflow = 1;
    }
    *Wn0_out = __shfl_sync(PG_FULL, Wn, 0);
    return sel;
}

__device__ __forceinline__ void fw_write_stump(const PgState& st, int t) {
    const size_t fb = (size_t)t * st.maxn;
    st.f_split_var[fb] = -1; st.f_split_val[fb] = pg_nan(); st.f_thr32[fb] = 0.0f; st.f_leaf_val[fb] = st.init_leaf;
    st.f_size[t] = 1;
}

// general finalize (rounds >= 1 happened): the selected particle's tree comes from the slot tables
__device__ __forceinline__ int fw_finalize(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int cur = c->cur, t = c->t_cur;
    const double Wl = lane < st.P ? c->C0 + (lane == 0 ? c->g_stump : f->hG[pg_six(st, cur, lane - 1)]) : -INFINITY;
    double Wn0;
    const int sel = fw_select(st, x, lane, Wl, &Wn0);
    const size_t fb = (size_t)t * st.maxn;
    int depth = 0, internal = 0;
    if (lane == 0)     // early commit: did the update end with the tree predicted after round 0 (same lineage, never grown further)?
    {
        f->pred_ok = f->pred_valid && ((sel == 0 && f->pred_sel == 0) ||
                                       (sel > 0 && f->pred_sel > 0 && f->hlin[sel - 1] == f->pred_lin && f->hsize[pg_six(st, cur, sel - 1)] == 3));
        if (f->pred_valid && !f->pred_ok) c->n_pred_miss += 1ull;
    }
    __syncwarp();
    if (f->pred_ok) {              // the forest already holds this very tree (fw_predict wrote it): nothing to copy
        if (sel != 0) { depth = 1; internal = 1; }
    }
    else if (sel == 0) { if (lane == 0) fw_write_stump(st, t); }
    else {
        const int s = sel - 1;
        const int sz = f->hsize[pg_six(st, cur, s)];
#pragma unroll 1
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
    return fw_end_update(st, x, lane, depth, internal, Wn0);
}

// ---- round 0, table-free ---------------------------------------------------------------------------
// After PH_ROOT: totals, leaf values and weights of the root proposals (kept in shared memory).
__device__ __forceinline__ int fw_r0_values(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int J = c->n_jobs;
    const bool bern = st.family == PG_BERNOULLI_LOGIT;
    const double T_o = f->tot[0], T_r = f->tot[1], T_rr = f->tot[2], T_ll0 = f->tot[3];
    double C0, g_stump;
    if (bern) { C0 = 0.0; g_stump = T_ll0; }
    else if (st.family == PG_GAUSSIAN) {
        C0 = c->Cconst - 0.5 * T_rr * c->inv_s2;
        g_stump = pg_gauss_g(st.init_leaf, T_r, (unsigned)st.n_global, c->inv_s2);
    } else { C0 = -0.5 * T_rr; g_stump = pg_gauss_g(st.init_leaf, T_r, (unsigned)st.n_global, 1.0); }
    if (lane == 0) { c->T_o = T_o; c->T_r = T_r; c->T_rr = T_rr; c->T_ll0 = T_ll0; c->C0 = C0; c->g_stump = g_stump; }
    __syncwarp();
    if (lane < J) {
        const FwVals v = fw_values(st, x, lane, (unsigned)st.n_global, T_o, T_r);
        f->jr_vL[lane] = v.vL; f->jr_vR[lane] = v.vR; f->jr_gL[lane] = v.gL; f->jr_gR[lane] = v.gR;
        f->jr_SoL[lane] = v.SoL; f->jr_SrL[lane] = v.SrL; f->jr_cL[lane] = v.cL; f->jr_cLl[lane] = f->rCntL[lane];
        if (bern) { { const size_t jo = (size_t)(c->upd & 1ull) * st.S; st.job_vL[jo + lane] = v.vL; st.job_vR[jo + lane] = v.vR; } }
    }
    __syncwarp();
    if (bern && J > 0) {
        if (lane == 0) fw_bern_bytes(st, x, J);
        __syncwarp();
        return SN_NONE;
    }
    return SN_FINISH_ROUND;
}

// Weights of the root proposals, stale-weight normalisation, resampling; then either the update ends as a
// stump (no survivor grew) or the survivors' tables are materialised and the general path takes over.
__device__ __forceinline__ int fw_r0_finish(const PgState& st, const FwCx& x, int lane, bool after_bern) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const int S = st.S, J = c->n_jobs, cur = c->cur;
    const bool act = lane < S;
    if (lane < J) {
        if (after_bern) { f->jr_gL[lane] = f->rllL[lane]; f->jr_gR[lane] = f->rllR[lane]; }
        const int s = f->jslot[lane];
        const double G = (c->g_stump - c->g_stump) + (f->jr_gL[lane] + f->jr_gR[lane]);   // as pg_apply_mutation on a root
        const double w = c->C0 + G;
        f->hw[s] = w;
        f->hmut[s] = 1;
        double* tr = pg_trace_slot(st, c, 0, s);
        if (tr) {
            tr[0] = 4; tr[4] = f->jr_vL[lane]; tr[5] = f->jr_vR[lane]; tr[6] = (double)f->jr_cL[lane];
            tr[7] = (double)((unsigned)st.n_global - f->jr_cL[lane]); tr[8] = w; tr[11] = f->jr_SoL[lane];
        }
    }
    __syncwarp();
    int n_mut = 0, mut = 0, anc = lane;
    const bool lazy = f->lazy_J > 0;                         // pruned round 0: fw_round0_all_dead already ran the resampling
    if (!lazy) anc = fw_normalize_resample(st, x, lane, &n_mut, &mut);
    const int mut_a = __shfl_sync(PG_FULL, mut, anc);
    const bool need = act && mut_a != 0;
    if (lane == 0) {
        c->acc_update += n_mut + f->lazy_J;                  // pruned proposals were accepted grows all the same (smc.rs:79-81)
        // the bound w <= Cconst holds up to the rounding of terms of size T_rr / (2 sigma^2): keep 50 units of margin honest
        if (lazy && !(0.5 * c->T_rr * c->inv_s2 < 1.0e13)) c->error = PG_ERR_PRUNE;
    }
    __syncwarp();
    if (!__any_sync(PG_FULL, need)) {
        // every survivor is an untouched stump with an empty queue: the loop of smc.rs:61 ends here.  All P
        // particles are stumps with the same log-likelihood (smc.rs:105-111).
        // exp(0) = 1 for everyone, the sequential sum of P ones is P, so every normalised weight is 1/P; the sequential
        // partial sums of P copies of 1/P are launch constants (f->stump_cum) — same values the generic loops produce.
        const int P = st.P;
        const double Wn = 1.0 / (double)P;
        const double chosen = fw_u7(st, x, c->upd) * f->stump_total;
        const int sel = __popc(__ballot_sync(PG_FULL, lane + 1 < P && f->stump_cum[lane] <= chosen));
        if (st.tr_upd) {
            if ((long long)c->trace_n < st.tr_u_cap) {
                double* tr = st.tr_upd + (size_t)c->trace_n * (size_t)(4 + 2 * P);
                if (lane == 0) { tr[0] = (double)(c->round + 1); tr[1] = (double)sel; tr[2] = (double)c->t_cur; tr[3] = (double)c->acc_update; }
                if (lane < P) { tr[4 + lane] = c->C0 + c->g_stump; tr[4 + P + lane] = Wn; }
            } else if (lane == 0) c->trace_overflow = 1;
        }
        if (lane == 0) fw_write_stump(st, c->t_cur);
        return fw_end_update(st, x, lane, 0, 0, Wn);
    }
    // ---- survivors that grew.  The partition pass needs only the partition jobs and the look-ahead requests, so those are
    // prepared first and the pass is issued BEFORE the slot tables of all slots (copies of their ancestors) are written
    // into half `cur`: the table writes run while the pass streams.
    unsigned off, cntL;
    int node;
    const int K = fw_plan_partitions(st, x, lane, anc, need, &off, &cntL, &node);
    if (lane == 0) { c->n_pjobs = K; c->any_queue = 1; }
    // ---- look-ahead: when every slot is a copy of ONE grown particle, the proposals of rounds 1 and 2
    // are known now — slot s will try node 1 with (u1, var)(round 1, s) and node 2 with (u1, var)(round 2, s), whatever
    // the resampling does in between.  Let the partition pass compute the min/max those proposals need.
    const int buf = (int)(c->upd % 3ull);
    const bool single = K == 1 && st.family != PG_BERNOULLI_LOGIT && __all_sync(PG_FULL, !act || need) && st.maxn > 6;
    int es_n = 0;
    {
        int nq = 0;
        if (single && !st.tr_upd && c->k + 1 < c->batch && lane == 0) { f->pred_lin = anc; f->pred_sel = -1 - f->hsjob[anc]; }   // fw_predict, after PART is issued
        if (single) {
            const int j = f->hsjob[anc];
            const unsigned cL = f->jr_cL[act ? j : 0], cR = (unsigned)st.n_global - cL, cLl = f->jr_cLl[act ? j : 0];
#pragma unroll 1
            for (int r = 1; r <= 2; ++r) {
                const bool want = act && f->d_ok[buf][r] && !(x.thr[1] > f->d_u1[buf][r][lane]) && (r == 1 ? cL : cR) > 0u;
                const unsigned m = __ballot_sync(PG_FULL, want);
                const int q = nq + __popc(m & ((1u << lane) - 1u));
                if (want && q < 32) {
                    st.mq_pj[q] = 0; st.mq_side[q] = r - 1; st.mq_var[q] = f->d_var[buf][r][lane];
                    st.mq_u3[q] = f->d_u3[buf][r][lane];
                    f->rq_r[q] = r; f->rq_s[q] = lane; f->rq_seg[q] = r == 1 ? off : off + cLl;
                }
                nq += __popc(m);
            }
            // Early statistics pass: every proposal of rounds 1 and 2 is one of these requests, so the statistics pass can be
            // queued right behind the partition pass — it reduces the look-ahead min / max itself, derives the split values
            // from the recorded draws exactly as fw_decide will, and leaves its results for both rounds — instead of waiting
            // for this block to digest the partition pass first (that section was ~20 us of idle pass blocks per update).
            es_n = (st.early_stats && st.world == 1 && !st.prune && nq > 0 && nq <= 32 && nq <= S && f->d_ok[buf][1] && f->d_ok[buf][2]) ? nq : 0;
            if (nq > 32) nq = 32;
        }
        if (lane == 0) { f->n_mq = nq; f->es_n = es_n; }
    }
    __syncwarp();
    if (c->error != PG_OK) return SN_NONE;
    fw_issue_part(st, x, lane);          // K >= 1 here; after PH_PART the general rounds continue
    if (es_n > 0) {
        const int pb = (int)(c->upd & 1ull);
        if (lane == 0) {
            const unsigned cLl = f->jr_cLl[f->hsjob[anc]];
            const unsigned long long Ll = cLl, Lr = (unsigned long long)st.n - cLl;
            int n1 = 0;
#pragma unroll 1
            for (int q = 0; q < es_n; ++q) n1 += f->rq_r[q] == 1 ? 1 : 0;
            c->bytes_model += (n1 > 0 ? 16ull * Ll + 4ull * Ll * (unsigned long long)n1 : 0ull)
                            + (es_n - n1 > 0 ? 16ull * Lr + 4ull * Lr * (unsigned long long)(es_n - n1) : 0ull);
            c->n_phases += 1;
        }
        __syncwarp();
        fw_issue(st, x, lane, PH_STATS, c->t_add, -1, es_n, 0, 0, c->a_cur | (c->a_cur << 1) | (pb << 2) | 8);
        if (lane == 0) { f->early_now = 1; st.ctrl->n_es += 1ull; }
        __syncwarp();
    }
    // ---- the slot tables, while the partition pass streams
    if (single) {
        const int j = f->hsjob[anc];
        const unsigned cL = f->jr_cL[act ? j : 0], cR = (unsigned)st.n_global - cL, cLl = f->jr_cLl[act ? j : 0];
        if (lane == 0) {   // the two leaves every slot will propose on in rounds 1 and 2: kept at hand (fw_draw_round)
            f->sl_lin = anc; f->sl_cnt[0] = cL; f->sl_cnt[1] = cR; f->sl_len[0] = cLl; f->sl_len[1] = (unsigned)st.n - cLl;
            f->sl_seg[0] = off; f->sl_seg[1] = off + cLl;
            f->sl_So[0] = f->jr_SoL[j]; f->sl_So[1] = c->T_o - f->jr_SoL[j]; f->sl_Sr[0] = f->jr_SrL[j]; f->sl_Sr[1] = c->T_r - f->jr_SrL[j];
            f->sl_g[0] = f->jr_gL[j]; f->sl_g[1] = f->jr_gR[j]; f->sl_valid = 1;
        }
    }
    if (act) {
        const int ss = pg_six(st, cur, lane);
        const size_t i0 = pg_tix(st, cur, lane, 0);
        if (need) {
            const int j = f->hsjob[anc];
            const size_t i1 = i0 + 1, i2 = i0 + 2;
            const unsigned cL = f->jr_cL[j], cR = (unsigned)st.n_global - cL, cLl = f->jr_cLl[j];
            st.pt_split_var[i0] = f->jvar[j]; st.pt_split_val[i0] = f->jsplit[j]; st.pt_thr32[i0] = f->jthr[j];
            st.pt_leaf_val[i0] = pg_nan(); st.pt_seg_off[i0] = PG_SEG_IDENTITY; st.pt_cnt[i0] = (unsigned)st.n_global; st.pt_len[i0] = (unsigned)st.n;
            st.pt_So[i0] = c->T_o; st.pt_Sr[i0] = c->T_r; st.pt_g[i0] = c->g_stump;
            st.pt_split_var[i1] = -1; st.pt_split_val[i1] = pg_nan(); st.pt_thr32[i1] = 0.0f; st.pt_leaf_val[i1] = f->jr_vL[j];
            st.pt_split_var[i2] = -1; st.pt_split_val[i2] = pg_nan(); st.pt_thr32[i2] = 0.0f; st.pt_leaf_val[i2] = f->jr_vR[j];
            st.pt_seg_off[i1] = off; st.pt_seg_off[i2] = off + cLl;
            st.pt_cnt[i1] = cL; st.pt_cnt[i2] = cR;
            st.pt_len[i1] = cLl; st.pt_len[i2] = (unsigned)st.n - cLl;
            st.pt_So[i1] = f->jr_SoL[j]; st.pt_So[i2] = c->T_o - f->jr_SoL[j];
            st.pt_Sr[i1] = f->jr_SrL[j]; st.pt_Sr[i2] = c->T_r - f->jr_SrL[j];
            st.pt_g[i1] = f->jr_gL[j]; st.pt_g[i2] = f->jr_gR[j];
            st.pt_queue[i0] = 1; st.pt_queue[i1] = 2;      // particle.rs:146-147: left then right
            f->hsize[ss] = 3; f->hqh[ss] = 0; f->hqt[ss] = 2;
            f->hG[ss] = (c->g_stump - c->g_stump) + (f->jr_gL[j] + f->jr_gR[j]);
            f->hlin[lane] = anc;                           // copies of one grown particle share a lineage
        } else {
            f->hlin[lane] = 32;                            // the untouched stump
            st.pt_split_var[i0] = -1; st.pt_split_val[i0] = pg_nan(); st.pt_thr32[i0] = 0.0f; st.pt_leaf_val[i0] = st.init_leaf;
            st.pt_seg_off[i0] = PG_SEG_IDENTITY; st.pt_cnt[i0] = (unsigned)st.n_global; st.pt_len[i0] = (unsigned)st.n;
            st.pt_So[i0] = c->T_o; st.pt_Sr[i0] = c->T_r; st.pt_g[i0] = c->g_stump;
            f->hsize[ss] = 1; f->hqh[ss] = 1; f->hqt[ss] = 1;
            f->hG[ss] = c->g_stump;
        }
    }
    __syncwarp();
    return SN_PART_ISSUED;
}

// ---- commands to the pass blocks ---------------------------------------------------------------------
// The pass blocks read the first 16 ints of the global PgCtrl (phase, t_add, t_sub, n_jobs, n_groups, n_pjobs,
// flags, ...) after observing bar_gen; issuing = publish those words, then release-store the new generation.
// passes of one generation parity completed when generation g has: 147 x (generations <= g of g's parity)
__device__ __forceinline__ unsigned fw_cnt_target(unsigned g) { return ((g + 1u) >> 1) * (gridDim.x - 1u); }

// Issue the pass the control block asks for (c->phase), on the buffers of the update in progress.
__device__ __forceinline__ void fw_issue_ctrl(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    const int phase = c->phase;
    const int pb = (int)(c->upd & 1ull);
    int src = c->a_cur, dst = c->a_cur;
    if (phase == PH_ROOT) dst = 1 - src;       // ping-pong: a discarded speculative pass never touches the live array
    else if (phase == PH_FINAL) dst = 0;       // `predictions` always end in buffer 0
    __syncwarp();
    if (lane == 0 && (phase == PH_ROOT || phase == PH_FINAL)) c->a_cur = dst;
    const int nj = phase == PH_STATS ? c->n_jobs + x.f->pf_n : c->n_jobs;
    fw_issue(st, x, lane, phase, c->t_add, phase == PH_ROOT ? c->t_sub : -1, nj, c->n_groups, c->n_pjobs,
             src | (dst << 1) | (pb << 2));
    if (phase == PH_PART && lane == 0) x.f->out_J[x.f->issued & 1u] = x.f->n_mq;   // what the next section reduces: the look-ahead min/max
    __syncwarp();
}

// wait until every pass issued so far has completed on every pass block
__device__ __forceinline__ void fw_await(const PgState& st, const FwCx& x, int lane) {
    if (lane == 0) {
        const unsigned g = x.f->issued;
        if (g >= 2u) pg_spin_until(&st.ctrl->cnt[(g - 1u) & 1u][0], fw_cnt_target(g - 1u), st.ctrl);
        if (g >= 1u) pg_spin_until(&st.ctrl->cnt[g & 1u][0], fw_cnt_target(g), st.ctrl);
    }
    __syncwarp();
}

// Issue the partition pass of the round (bytes accounting, then the command).  A root pass issued ahead on a guess that
// this very round refuted is drained first: the partition pass reuses the pass blocks.
__device__ __forceinline__ void fw_issue_part(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    __syncwarp();
    if (f->spec_now) {
        fw_await(st, x, lane);
        if (lane == 0) { c->n_drain += 1ull; f->spec_now = 0; }
        __syncwarp();
    }
    if (lane == 0) {
        const int K = c->n_pjobs;
        unsigned long long bm = 0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const unsigned long long L = f->pq_len[k];
            bm += (f->pq_ident[k] ? 0ull : 4ull * L) + 4ull * L + 4ull * L;
        }
        c->bytes_model += bm + 4ull * (unsigned long long)f->n_mq * (K > 0 ? (unsigned long long)f->pq_len[0] : 0ull);
        c->n_groups = f->n_mq;          // look-ahead min/max requests riding on the partition pass
        c->phase = PH_PART; c->n_phases += 1;
    }
    __syncwarp();
    fw_issue_ctrl(st, x, lane);
}

// Speculation: the root pass of update U+1 only needs (i) U+1's root proposals, which are
// data-independent and already precomputed, and (ii) the tree update U commits.  When some slot of update U has no
// proposal, U ends as a stump unless its grown particles out-weigh the stale weights (Q1) — so issue U+1's root pass
// NOW, assuming "U commits a stump", and validate while it streams.  It reads the live A buffer and writes the other
// one, with its own parity of job / partial buffers: a wrong guess costs nothing but the bandwidth of an idle GPU.
__device__ __forceinline__ bool fw_try_speculate(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const unsigned long long U1 = c->upd + 1ull;
    const int buf = (int)(U1 % 3ull);
    bool ok = st.family != PG_BERNOULLI_LOGIT && c->k + 1 < c->batch && st.m >= 2 && c->n_jobs < st.S && c->error == PG_OK
              && f->d_upd[buf] == U1 && f->d_ok[buf][0];
#pragma unroll 1
    for (int q = 0; q <= FW_DR && ok; ++q) ok = f->d_flag[buf][q] == U1 + 1ull;
    // The helper warps write these flags while this warp reads them: lanes can see different values.  The decision must be
    // the warp's, not the lane's — a lane that left here while the others went on into the warp-wide votes below hung the
    // control block once in ~1500 sweeps.
    ok = __all_sync(PG_FULL, ok);
    if (!ok) return false;
    __threadfence_block();
    bool passed;
    int status;
    const FwJob jb = fw_root_job(st, x, U1, lane, &passed, &status);
    if (__any_sync(PG_FULL, status == 6 || status == 7)) return false;   // let the regular path raise the error
    int G;
    const int pb1 = (int)(U1 & 1ull);
    int J1 = 0;
    G = 0;
    if (!fw_round0_all_dead(st, x, U1, lane, jb.live)) J1 = fw_write_jobs(st, x, jb, lane, true, pb1, true, false, &G);
    const int t_sub1 = (c->next_tree_idx + c->k + 1) % st.m;
    const int src = c->a_cur, dst = 1 - c->a_cur;
    fw_issue(st, x, lane, PH_ROOT, -2 /* a stump */, t_sub1, J1, G, 0, src | (dst << 1) | (pb1 << 2));
    if (lane == 0) c->n_spec += 1;
    return true;
}

// Stage the root jobs of the NEXT update in the global job arrays of its parity while nothing waits for us, so that
// the speculative command itself is only the 16 published words (fw_issue_prestaged).
__device__ __forceinline__ void fw_prestage(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const unsigned long long U1 = c->upd + 1ull;
    const int buf = (int)(U1 % 3ull);
    bool ok = st.family != PG_BERNOULLI_LOGIT && c->k + 1 < c->batch && st.m >= 2 && c->error == PG_OK
              && f->d_upd[buf] == U1 && f->d_ok[buf][0];
#pragma unroll 1
    for (int q = 0; q <= FW_DR && ok; ++q) ok = f->d_flag[buf][q] == U1 + 1ull;
    ok = __all_sync(PG_FULL, ok);          // warp-uniform: the flags are being written by the helper warps (see fw_try_speculate)
    if (!ok) return;
    __threadfence_block();
    bool passed;
    int status;
    const FwJob jb = fw_root_job(st, x, U1, lane, &passed, &status);
    if (__any_sync(PG_FULL, status == 6 || status == 7)) return;
    int G;
    int J1 = 0;
    G = 0;
    if (!fw_round0_all_dead(st, x, U1, lane, jb.live)) J1 = fw_write_jobs(st, x, jb, lane, true, (int)(U1 & 1ull), true, false, &G);
    if (lane == 0) { f->pre_upd = U1; f->pre_J = J1; f->pre_G = G; st.ctrl->pre_J[(int)(U1 & 1ull)] = J1; }
    __syncwarp();
}

__device__ __forceinline__ bool fw_issue_prestaged(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const unsigned long long U1 = c->upd + 1ull;
    if (!(f->pre_upd == U1 && c->n_jobs < st.S && c->k + 1 < c->batch && c->error == PG_OK)) return false;
    const int t_sub1 = (c->next_tree_idx + c->k + 1) % st.m;
    const int src = c->a_cur, dst = 1 - c->a_cur;
    fw_issue(st, x, lane, PH_ROOT, -2 /* a stump */, t_sub1, f->pre_J, f->pre_G, 0, src | (dst << 1) | ((int)(U1 & 1ull) << 2));
    if (lane == 0) c->n_spec += 1;
    return true;
}

// Early commit, prediction step (runs right after the partition pass of an all-grew update has been issued, while it streams):
__device__ __forceinline__ void fw_predict(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    if (f->pred_valid || f->pred_sel >= 0) return;          // nothing pending (pred_sel = -1 - job marks a pending prediction)
    const int j0 = -1 - f->pred_sel;
    __syncwarp();
    // Early commit: every slot now holds a copy of ONE grown particle.  Under the stale-weight
    // rule later rounds cannot change that (their mutated particles die), so the final selection is between the
    // stump and this tree with weights that are known now: run the selection (smc.rs:105-122) on those weights,
    // write the predicted tree to the forest, and let the section after the update's last data pass issue the next
    // update's root pass before it digests the results.  fw_finalize checks the prediction; a miss is drained.
    const int P = st.P;
    const double Gl = (c->g_stump - c->g_stump) + (f->jr_gL[j0] + f->jr_gR[j0]);
    const bool pact = lane < P;
    const double Wl = pact ? c->C0 + (lane == 0 ? c->g_stump : Gl) : -INFINITY;
    const double mx = fw_max(Wl);
    const double e = pact ? exp(Wl - mx) : 0.0;
    const double sum = fw_sum_seq(e, P);
    const double Wn = e / sum;
    const double total = fw_sum_seq(pact ? Wn : 0.0, P);
    const double chosen = fw_u7(st, x, c->upd) * total;
    int sel = 0;
    {
        double cum = 0.0;
        bool open = true;
#pragma unroll 8
        for (int i = 0; i + 1 < P; ++i) {
            cum += __shfl_sync(PG_FULL, Wn, i);
            if (open && cum <= chosen) sel = i + 1; else open = false;
        }
    }
    if (lane == 0) {
        const int t = c->t_cur, j = j0;
        const size_t fb = (size_t)t * st.maxn;
        if (sel == 0) fw_write_stump(st, t);
        else {
            st.f_split_var[fb] = f->jvar[j]; st.f_split_val[fb] = f->jsplit[j]; st.f_thr32[fb] = f->jthr[j]; st.f_leaf_val[fb] = pg_nan();
            st.f_split_var[fb + 1] = -1; st.f_split_val[fb + 1] = pg_nan(); st.f_thr32[fb + 1] = 0.0f; st.f_leaf_val[fb + 1] = f->jr_vL[j];
            st.f_split_var[fb + 2] = -1; st.f_split_val[fb + 2] = pg_nan(); st.f_thr32[fb + 2] = 0.0f; st.f_leaf_val[fb + 2] = f->jr_vR[j];
            st.f_size[t] = 3;
        }
        f->pred_sel = sel; f->pred_valid = 1;
        c->n_pred += 1ull;
    }
    __syncwarp();
}

// Early commit: the root pass of the next update, adding the tree predicted for this one (already in the forest).
__device__ __forceinline__ bool fw_issue_predicted(const PgState& st, const FwCx& x, int lane) {
    PgCtrl* c = x.c;
    PgFastSm* f = x.f;
    const unsigned long long U1 = c->upd + 1ull;
    if (!(f->pred_valid && !f->pred_tried && f->pre_upd == U1 && c->k + 1 < c->batch && c->error == PG_OK && st.family != PG_BERNOULLI_LOGIT)) return false;
    const int t_sub1 = (c->next_tree_idx + c->k + 1) % st.m;
    const int src = c->a_cur, dst = 1 - c->a_cur;
    fw_issue(st, x, lane, PH_ROOT, f->pred_sel == 0 ? -2 : c->t_cur, t_sub1, f->pre_J, f->pre_G, 0, src | (dst << 1) | ((int)(U1 & 1ull) << 2));
    if (lane == 0) { c->n_spec += 1; f->pred_tried = 1; }     // once per update: a drained guess is not repeated
    return true;
}

// ---- the serial block's whole-kernel control loop ------------------------------------------------------
__device__ __forceinline__ void pg_control_fast(const PgState& st, const PgSmemView& sm) {
    __shared__ PgCtrl fw_scratch[FW_DR + 1];   // helpers must not raise errors for draws that may never be consumed
    PgFastSm* f = sm.fast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FwCx x;
    x.f = f;
    x.c = &f->lc;
    const bool staged = sm.base->cstaged != 0;
    x.thr = staged ? sm.base->cthr : st.thr_depth;
    x.prior = st.prior ? (staged ? sm.cprior : st.prior) : nullptr;
    x.cmin = staged ? sm.cmin : st.col_min;
    x.cmax = staged ? sm.cmax : st.col_max;
    PgCtrl* c = x.c;

    fw_load_state(st, f);
    if (threadIdx.x == 0) {   // final selection among P identical stumps: sequential sums of P copies of 1/P (smc.rs:105-114)
        const double wn = 1.0 / (double)st.P;
        double cum = 0.0;
        for (int i = 0; i < st.P && i < 32; ++i) { cum += wn; f->stump_cum[i] = cum; }
        f->stump_total = cum;
    }
    __syncthreads();
    if (threadIdx.x == 0) c->a_cur = 0;        // `predictions` live in buffer 0 between steps
    bool first = true;
#define FW_BEAT(slot, code) do { if (st.heartbeat && threadIdx.x == 0) { volatile unsigned long long* hb_ = reinterpret_cast<volatile unsigned long long*>(st.ctrl->hostdiag); if (hb_) hb_[8 + (slot)] = ((unsigned long long)f->issued << 32) | ((unsigned long long)(c->round & 0xff) << 24) | ((unsigned long long)(c->k & 0xfff) << 8) | (unsigned long long)(code); } } while (0)
#define FW_TICK(slot) do { if (lane == 0) { const long long now_ = clock64(); c->prof2[slot] += (unsigned long long)(now_ - tq); c->prof2[8 + slot] += 1ull; tq = now_; } } while (0)

#pragma unroll 1
    for (long long guard = 0; guard < (1ll << 40); ++guard) {
        // ---- (A) the outstanding pass (if any) completes
        if (!first && threadIdx.x == 0)
            f->cur_gen = f->issued - ((f->spec_now || f->early_now) ? 1u : 0u);   // the OLDEST outstanding pass: a pass issued ahead queues behind it
        FW_BEAT(0, 1);
        if (st.heartbeat && (threadIdx.x & 31) == 0) { volatile unsigned long long* hw_ = reinterpret_cast<volatile unsigned long long*>(st.ctrl->hostdiag); if (hw_) hw_[16 + (threadIdx.x >> 5)] = ((unsigned long long)f->issued << 32) | 1ull; }
        __syncthreads();
        const unsigned gcur = f->cur_gen;
        const int phase = first ? PH_BEGIN : f->out_phase[gcur & 1u];
        const int J = first ? 0 : f->out_J[gcur & 1u];
        const int pb = first ? 0 : f->out_pb[gcur & 1u];
        // Warp 0 waits for the completion COUNT (everything below that is not the reduce needs the whole pass done: the issue
        // logic reuses the command buffer of the pass before it, the control code reads what the pass wrote); warps 1.. go
        // straight into fw_reduce and fetch each block's partials as that block completes.
        if (!first && threadIdx.x == 0) {
            const long long tw = clock64();
            const unsigned g = gcur;
            pg_spin_until(&st.ctrl->cnt[g & 1u][0], fw_cnt_target(g), st.ctrl);
            const int ph = f->out_phase[g & 1u];
            c->prof3[ph == PH_ROOT ? 4 : ph == PH_STATS ? 5 : ph == PH_PART ? 6 : 7] += (unsigned long long)(clock64() - tw);
#ifdef PG_DIAG
            const int cls = ph == PH_ROOT ? 0 : ph == PH_STATS ? 1 : ph == PH_PART ? 2 : -1;
            if (cls >= 0) {
                const unsigned long long now = pg_globaltimer();
                volatile unsigned long long* d = st.ctrl->dbg;
                const unsigned long long t_is = d[0];
                d[4 + 4 * cls] += d[1] - t_is; d[5 + 4 * cls] += d[2] - t_is; d[6 + 4 * cls] += now - t_is; d[7 + 4 * cls] += d[3];
            }
#endif
        }
        if (warp == 0) __syncwarp();
        first = false;
        const long long t_entry = clock64();
        const int gerr = threadIdx.x == 0 ? __ldcg(&st.ctrl->error) : PG_OK;   // a pass may have flagged an error
        if (warp == 0) {
            // A root pass for the NEXT update is normally in flight already, issued ahead at the end of the previous section
            // (it queues behind the pass that just completed, so the pass blocks never idle between the two).  If it could
            // not be issued then (draws / staged jobs not ready), issue it now before anything else: after a root pass the
            // next update's root pass on a stump; after the last data pass of an update whose outcome was predicted after
            // round 0 (the statistics pass; with pruning the partition pass) the next root pass on the predicted tree.
            if (phase == PH_PART && f->early_now) {          // the early statistics pass queued behind this partition pass is now the pass in progress
                __syncwarp();
                if (lane == 0) { f->early_now = 0; f->es_pending = 1; f->es_resume = SN_NONE; }
                __syncwarp();
            }
            bool sp = f->spec_now != 0;
            if (!sp) sp = (phase == PH_ROOT && fw_issue_prestaged(st, x, lane)) ||
                          ((phase == PH_STATS || (phase == PH_PART && (st.prune || f->es_pending))) && fw_issue_predicted(st, x, lane));
            __syncwarp();
            if (lane == 0) { f->spec_now = sp ? 1 : 0; f->stat_spec = sp ? 1 : 0; }
        }
        FW_BEAT(0, 2);
        fw_reduce(st, f, phase, J, pb, gcur);   // contains a __syncthreads before its combine step
        FW_BEAT(0, 3);
        if (st.world > 1 && warp == 0) fw_exchange(st, f, phase, J, lane);
        if (threadIdx.x == 0 && gerr != PG_OK) c->error = gerr;
        if (st.heartbeat && (threadIdx.x & 31) == 0) { volatile unsigned long long* hw_ = reinterpret_cast<volatile unsigned long long*>(st.ctrl->hostdiag); if (hw_) hw_[16 + (threadIdx.x >> 5)] = ((unsigned long long)f->issued << 32) | 3ull; }
        __syncthreads();
        const unsigned long long upd0 = c->upd;   // the update in progress at entry
        bool need[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) need[q] = f->d_upd[(int)((upd0 + q) % 3ull)] != upd0 + q;
        if (st.heartbeat && (threadIdx.x & 31) == 0) { volatile unsigned long long* hw_ = reinterpret_cast<volatile unsigned long long*>(st.ctrl->hostdiag); if (hw_) hw_[16 + (threadIdx.x >> 5)] = ((unsigned long long)f->issued << 32) | 4ull; }
        __syncthreads();                           // ... all read by every warp before anything is advanced
        long long tq = clock64();
        if (threadIdx.x == 0) { c->prof2[0] += (unsigned long long)(tq - t_entry); c->prof2[8] += 1ull; }

        // ---- (B) helper warps: draws of the update in progress (missing only right after launch) and of the next two
        if (warp >= 1 && warp <= FW_DR + 1) {
#pragma unroll 1
            for (int q = 0; q < 3; ++q)
                if (need[q]) fw_prefetch_draws(st, x, upd0 + q, warp, lane, &fw_scratch[warp - 1]);
        }

        // ---- (C) warp 0: the control logic
        if (warp == 0) {
            int sn = SN_NONE;
            bool spec = false;
            switch (phase) {
                case PH_BEGIN: sn = fw_begin_step(st, x, lane); break;
                case PH_ROOT:
                    spec = f->spec_now != 0;
                    if (!spec) { spec = fw_try_speculate(st, x, lane); if (spec && lane == 0) f->spec_now = 1; }
                    FW_TICK(6);
                    sn = fw_r0_values(st, x, lane);
                    if (sn == SN_FINISH_ROUND) sn = fw_r0_finish(st, x, lane, false);
                    FW_TICK(1);
                    break;
                case PH_STATS:
                    if (f->es_pending) {
                        // the early statistics pass: job q = look-ahead request q of the partition pass; its results are the
                        // statistics of rounds 1 and 2 (keyed by the threshold the pass derived: fw_round_stats compares)
                        const int nq = f->es_n;
                        if (lane < nq) {
                            const int ri = f->rq_r[lane] - 1, s = f->rq_s[lane];
                            f->sc_cnt[ri][s] = f->rCnt[lane]; f->sc_cntL[ri][s] = f->rCntL[lane]; f->sc_So[ri][s] = f->rSo[lane]; f->sc_Sr[ri][s] = f->rSr[lane];
                            f->sc_job[ri][s] = lane; f->sc_seg[ri][s] = f->rq_seg[lane]; f->sc_var[ri][s] = st.mq_var[lane];
                            const float th = __ldcg(&st.es_thr[lane]);
                            f->sc_thr[ri][s] = th; f->sc_valid[ri][s] = th == th ? 1 : 0;
                        }
                        __syncwarp();
                        const int resume = f->es_resume;
                        const unsigned m1 = __ballot_sync(PG_FULL, lane < nq && f->rq_r[lane] == 1), m2 = __ballot_sync(PG_FULL, lane < nq && f->rq_r[lane] == 2);
                        if (lane == 0) { f->es_pending = 0; f->es_n = 0; f->es_mask = (m1 ? 1 : 0) | (m2 ? 2 : 0); }
                        __syncwarp();
                        spec = f->spec_now != 0;
                        if (resume == SN_DRAW_ROUND) {
                            FwJob jb;
                            jb.live = f->wj_live[lane] != 0; jb.slot = lane; jb.node = f->wj_node[lane]; jb.var = f->wj_var[lane];
                            jb.seg_off = f->wj_seg[lane]; jb.len = f->wj_len[lane]; jb.lo = f->wj_lo[lane]; jb.hi = f->wj_hi[lane];
                            jb.thr = f->wj_thr[lane]; jb.split = f->wj_split[lane];
                            sn = fw_round_stats(st, x, lane, jb);
                        } else sn = resume == SN_FINALIZE ? SN_FINALIZE : SN_WAIT_EARLY;   // SN_NONE cannot happen: the round logic always parks first
                        FW_TICK(1);
                        break;
                    }
                    if (f->pf_n > 0) {   // results of the prefetch jobs (round-2 proposals) go to the statistics cache
                        const int J0 = c->n_jobs;
                        if (lane >= J0 && lane < J0 + f->pf_n) {
                            const int s = f->pf_slot[lane - J0];
                            f->sc_cnt[1][s] = f->rCnt[lane]; f->sc_cntL[1][s] = f->rCntL[lane]; f->sc_So[1][s] = f->rSo[lane]; f->sc_Sr[1][s] = f->rSr[lane]; f->sc_job[1][s] = lane; f->sc_valid[1][s] = 1;
                        }
                        __syncwarp();
                        if (lane == 0) f->pf_n = 0;
                        __syncwarp();
                    }
                    spec = f->spec_now != 0;
                    sn = fw_after_stats(st, x, lane, phase);
                    FW_TICK(1);
                    break;
                case PH_BERN:
                    if (c->round == 0) sn = fw_r0_finish(st, x, lane, true); else sn = fw_after_stats(st, x, lane, phase);
                    FW_TICK(1);
                    break;
                case PH_MINMAX: sn = fw_after_minmax(st, x, lane); FW_TICK(2); break;
                case PH_PART: {
                    const int nq = f->n_mq;
                    __syncwarp();
                    if (lane < nq) {   // look-ahead min/max computed by the partition pass
                        const int ri = f->rq_r[lane] - 1, s = f->rq_s[lane];
                        f->mm_seg[ri][s] = f->rq_seg[lane]; f->mm_var[ri][s] = st.mq_var[lane];
                        f->mm_lo[ri][s] = pg_key2f(f->rMin[lane]); f->mm_hi[ri][s] = pg_key2f(f->rMax[lane]);
                        f->mm_valid[ri][s] = 1;
                    }
                    if (lane == 0) { c->round += 1; f->n_mq = 0; }
                    __syncwarp();
                    spec = f->spec_now != 0;
                    sn = SN_DRAW_ROUND;
                    break;
                }
                case PH_FINAL:
                    if (lane == 0) { c->next_tree_idx = (c->next_tree_idx + c->batch) % st.m; c->t_add = -1; c->phase = PH_DONE; }
                    __syncwarp();
                    break;
                default: break;
            }
            bool spec_commit = false;
            // warp-only sub-steps chain without block barriers until a grid pass or a block-wide step is needed
#pragma unroll 1
            for (int it = 0; it < (1 << 20); ++it) {
                FW_BEAT(1, 0x10 + sn);
                if (lane == 0 && st.heartbeat) { volatile unsigned long long* hb_ = reinterpret_cast<volatile unsigned long long*>(st.ctrl->hostdiag); if (hb_) hb_[10] = (unsigned long long)it; }
                if (c->error != PG_OK) { sn = SN_NONE; break; }
                if (sn == SN_NONE || sn == SN_PART_ISSUED || sn == SN_WAIT_EARLY) break;
                if (sn == SN_FINALIZE && f->es_pending) {      // nothing asked for the early statistics after all: let the pass finish first
                    __syncwarp();
                    if (lane == 0) f->es_resume = SN_FINALIZE;
                    __syncwarp();
                    sn = SN_WAIT_EARLY;
                    break;
                }
                if (sn == SN_BEGIN_UPDATE) {
                    __syncwarp();
                    if (spec && f->spec_now && (phase == PH_ROOT || f->pred_ok)) {   // the guess held: the root pass in flight is the next update's
                        __syncwarp();
                        if (lane == 0) { c->a_cur = 1 - c->a_cur; f->spec_now = 0; }   // adopted: no longer "ahead"
                        __syncwarp();
                        fw_start_update(st, x, lane, true);
                        spec_commit = true;
                    } else fw_start_update(st, x, lane, false);
                    sn = SN_NONE;
                    FW_TICK(7);
                }
                else if (sn == SN_DRAW_ROUND) { sn = fw_draw_round(st, x, lane); FW_TICK(7); }
                else if (sn == SN_FINISH_ROUND) { sn = fw_finish_round(st, x, lane); FW_TICK(3); }
                else if (sn == SN_FINALIZE) { sn = fw_finalize(st, x, lane); FW_TICK(5); }
                else { if (lane == 0) c->error = PG_ERR_INTERNAL; sn = SN_NONE; }
            }
            if (c->error != PG_OK) { if (lane == 0) c->phase = PH_DONE; sn = SN_NONE; __syncwarp(); }
            if (sn == SN_PART_ISSUED) {
                FW_TICK(4);
                fw_predict(st, x, lane);            // off the critical path: the partition pass is streaming
            } else if (sn == SN_WAIT_EARLY) {
                // nothing to issue: the early statistics pass is in flight (and, normally, the predicted root pass behind it)
                __syncwarp();
            } else {
                __syncwarp();
                if (f->es_pending || f->early_now) {    // odd paths (an error, a pass nobody planned for): the early statistics are dropped
                    fw_await(st, x, lane);
                    if (lane == 0) { f->es_pending = 0; f->early_now = 0; f->es_n = 0; }
                    __syncwarp();
                }
                if (f->spec_now && !spec_commit) {      // drain a wrong guess (odd paths, and before PH_DONE on an error)
                    fw_await(st, x, lane);
                    if (lane == 0) { c->n_drain += 1ull; f->spec_now = 0; }
                    __syncwarp();
                }
                if (!spec_commit) fw_issue_ctrl(st, x, lane);                              // next pass, or PH_DONE
                // ---- issue AHEAD: the pass blocks find the next root pass queued behind the pass they are running
                if (c->phase == PH_ROOT && c->round == 0) {
                    // an update just started (its root pass was issued, or adopted): stage the jobs of the update after it and,
                    // if this one is expected to end as a stump (some slot proposes nothing: Q1), issue that update's root pass
                    fw_prestage(st, x, lane);
                    if (st.issue_ahead && !f->spec_now && fw_issue_prestaged(st, x, lane)) { if (lane == 0) f->spec_now = 1; }
                } else if (st.issue_ahead && c->phase == PH_STATS && !f->spec_now) {
                    // early commit: the statistics pass just issued is normally the update's last data pass and the outcome was
                    // predicted after round 0: the next update's root pass on the predicted tree goes right behind it
                    if (fw_issue_predicted(st, x, lane)) { if (lane == 0) f->spec_now = 1; }
                }
                __syncwarp();
            }
            if (lane == 0) { f->sn = sn; f->quit = c->phase == PH_DONE ? 1 : 0; }
            FW_BEAT(0, 8);
        }
        if (warp == 1 && lane == 0 && st.heartbeat) { volatile unsigned long long* hb_ = reinterpret_cast<volatile unsigned long long*>(st.ctrl->hostdiag); if (hb_) hb_[11] = ((unsigned long long)f->issued << 32) | 9ull; }
        if (st.heartbeat && (threadIdx.x & 31) == 0) { volatile unsigned long long* hw_ = reinterpret_cast<volatile unsigned long long*>(st.ctrl->hostdiag); if (hw_) hw_[16 + (threadIdx.x >> 5)] = ((unsigned long long)f->issued << 32) | 5ull; }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long d = (unsigned long long)(clock64() - t_entry);
            c->prof_ser[phase & 7] += d;
            if (phase == PH_ROOT) { const int q = f->stat_spec ? 0 : 2; c->prof3[q] += d; c->prof3[q + 1] += 1ull; }
        }
        if (f->quit) break;
    }
    __syncthreads();
    fw_store_state(st, f);
}
