// pg_sweep.cu — kernels of the PGBART sweep for sm_100a and their launchers.
//
//   pg_sweep_persistent*  one launch per PySampler.step(): all tree updates of the step's batch (kernel.rs:81-115)
//                         run inside it, no host round trip.  P <= 32: block 0 is the control block
//                         (pg_serial_fast.cuh) and issues commands that the other 147 blocks execute as grid passes
//                         (pg_passes.cuh); one kernel per pass variant (PGV_*).  P > 32 (_generic): every block runs
//                         the passes and block 0 runs the team-based serial section of pg_serial.h between them.
//   pg_phase_kernel /     the same passes / serial section as separate launches with a host read-back between them
//   pg_serial_kernel      ("stepwise" driver: slow, used to cross-check the persistent kernels).
//   pg_predict_kernel     out-of-sample forest prediction (tree.rs:169-192).
#include <cooperative_groups.h>
#include <cstdio>
#include "pg_passes.cuh"
#include "pg_serial_fast.cuh"
#include "pg_sweep.h"

namespace cg = cooperative_groups;

extern __shared__ __align__(128) unsigned char pg_smem_raw[];   // dynamic shared memory of every kernel here (pg_smem_view)

// Grid barrier with a serial section.  Block 0 is ALWAYS the serial block: its SM keeps the (large,
// straight-line) control code in its instruction caches, which is what the section's latency is made of.
// Other blocks arrive (release fence + counter) and wait for the generation flag; block 0 waits for the
// counter, runs the serial section, publishes, and flips the flag.  `gen` counts barriers since launch
// (bar_count / bar_gen are zeroed by pg_prepare_step).  phase / J describe the pass that just ended.
template <bool FAST>
__device__ void pg_barrier_serial(const PgState& st, const PgSmemView& smv, unsigned& gen, int phase, int J) {
    PgCtrl* c = st.ctrl;
    __syncthreads();
    if (blockIdx.x != 0) {
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&c->bar_count, 1u);
            pg_spin_until(&c->bar_gen, gen + 1u, c);
        }
        __syncthreads();
    } else {
        if (threadIdx.x == 0) {
            __threadfence();
            pg_spin_until(&c->bar_count, (gen + 1u) * (gridDim.x - 1u), c);
        }
        __syncthreads();
        long long dur;
        {                                                    // generic team version of the control logic
            PgTeam tm{(int)threadIdx.x, (int)blockDim.x};
            const long long ts = clock64();
            pg_serial(st, tm);
            dur = clock64() - ts;
        }
        (void)smv; (void)J;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            pg_st_release(&c->bar_gen, gen + 1u);
            c->prof[16 + (phase & 7)] += (unsigned long long)dur;   // after the release: off the critical path
        }
    }
    gen += 1u;
}

template <int VAR>
__device__ __forceinline__ void pg_sweep_body(const PgState& st) {
    constexpr bool FAST = VAR != PGV_GENERIC;
    if (__ldcg(&st.ctrl->poisoned)) return;      // an earlier step failed on the device: every block leaves at once
    const PgSmemView sm = pg_smem_view(pg_smem_raw, st.S);
    // sampler constants the serial section reads: staged once per launch
    for (int i = threadIdx.x; i < st.max_depth + 2 && i < 16; i += blockDim.x) sm.base->cthr[i] = st.thr_depth[i];
    if (threadIdx.x == 0) { sm.base->cstaged = (st.p <= FW_MAXP && st.max_depth + 2 <= 16) ? 1 : 0; sm.base->pass_bid = (int)blockIdx.x - 1; }
    if (st.p <= FW_MAXP && FAST && blockIdx.x == 0)     // control block only: in the pass blocks these bytes are the staging ring
        for (int i = threadIdx.x; i < st.p; i += blockDim.x) {
            sm.cprior[i] = st.prior ? st.prior[i] : 0.0;
            sm.cmin[i] = st.col_min[i];
            sm.cmax[i] = st.col_max[i];
        }
    __syncthreads();
    if (FAST) {
        // ---- P <= 32: block 0 runs the whole control loop (pg_control_fast) and issues commands; the other blocks
        // execute them.  A command = the published words of PgCtrl + a new generation in bar_gen; completion = every
        // pass block adds 1 to bar_count.  The control block may issue the next root pass before it has finished
        // digesting the previous one (speculation), so passes and control overlap.
        if (blockIdx.x == 0) { pg_control_fast(st, sm); return; }
        PgCtrl* c = st.ctrl;
        unsigned gen = 0;
        long long t_idle = clock64();
        for (;;) {
            const unsigned g = gen + 1u;                         // the command this block executes next
            if (threadIdx.x < 32) {
                if (threadIdx.x == 0) {
                    pg_spin_until(&c->bar_gen, g, c);
#ifdef PG_DIAG
                    atomicMax(&c->dbg[1], pg_globaltimer());
#endif
                }
                __syncwarp();
                pg_fetch_cmd(c->cmd[g & 1u], sm.base);
            }
            __syncthreads();
            const int phase = sm.base->cmd.phase;
            if (phase == PH_DONE) break;
            const long long t0 = clock64();
            const int flags = sm.base->cmd.flags;
            if (threadIdx.x == 0) sm.base->pass_gen = g;
            pg_run_pass<VAR>(st, sm, phase);
            __syncthreads();
            const long long t1 = clock64();
            if (threadIdx.x == 0) {
#ifdef PG_DIAG
                atomicMax(&c->dbg[3], (unsigned long long)(t1 - t0));
                st.diag_blocks[(size_t)(blockIdx.x - 1) * 8 + (phase & 7)] += (unsigned long long)(t1 - t0);
                atomicMax(&c->dbg[2], pg_globaltimer());
#endif
                pg_fence_acq_rel();                                   // one release fence for both signals
                pg_st_relaxed(&st.blk_done[blockIdx.x - 1], g);       // per block: the control block's reduce polls these
                pg_red_relaxed_add(&c->cnt[g & 1u][0], 1u);           // count: pass-side barriers, the control block's issue logic
            }
            if (phase == PH_ROOT && st.prefetch_next) pg_prefetch_next_root(st, flags);
            if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {   // stopwatch kept by the last pass block
                c->prof[phase & 7] += (unsigned long long)(t1 - t0);
                c->prof[8 + (phase & 7)] += (unsigned long long)(t0 - t_idle);   // time spent waiting for this command
                c->prof[24 + (phase & 7)] += 1ull;
            }
            t_idle = t1;
            gen = g;
        }
        return;
    }
    unsigned gen = 0;
    pg_barrier_serial<FAST>(st, sm, gen, PH_BEGIN, 0);   // serial section of PH_BEGIN prepares the first pass
    for (;;) {
        pg_fetch_cmd(reinterpret_cast<const int*>(st.ctrl), sm.base);   // the generic control code publishes in PgCtrl itself
        __syncthreads();
        const int phase = sm.base->cmd.phase;
        const int J = sm.base->cmd.n_jobs;
        if (phase == PH_DONE) break;
        const long long t0 = clock64();
        if (blockIdx.x != 0) pg_run_pass<PGV_GENERIC>(st, sm, phase);   // block 0 = serial block: streams nothing
        __syncthreads();
        const long long t1 = clock64();
        pg_barrier_serial<FAST>(st, sm, gen, phase, J);
        if (blockIdx.x != 0 && threadIdx.x == 0) st.diag_blocks[(size_t)(blockIdx.x - 1) * 8 + (phase & 7)] += (unsigned long long)(t1 - t0);
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {   // stopwatch kept by a block that never runs the serial section
            PgCtrl* c = st.ctrl;
            c->prof[phase & 7] += (unsigned long long)(t1 - t0);
            c->prof[8 + (phase & 7)] += (unsigned long long)(clock64() - t1);
            c->prof[24 + (phase & 7)] += 1ull;
        }
    }
}

// P <= 32: latency-organised control block, one kernel per pass variant; P > 32: generic team serial section.
extern "C" __global__ void __launch_bounds__(PG_THREADS, 1) pg_sweep_persistent_ring(const PgState st) { pg_sweep_body<PGV_RING>(st); }
extern "C" __global__ void __launch_bounds__(PG_THREADS, 1) pg_sweep_persistent_direct(const PgState st) { pg_sweep_body<PGV_DIRECT>(st); }
extern "C" __global__ void __launch_bounds__(PG_THREADS, 1) pg_sweep_persistent_bern(const PgState st) { pg_sweep_body<PGV_BERN>(st); }
extern "C" __global__ void __launch_bounds__(PG_THREADS, 1) pg_sweep_persistent_generic(const PgState st) { pg_sweep_body<PGV_GENERIC>(st); }

extern "C" __global__ void __launch_bounds__(PG_THREADS, 1) pg_phase_kernel(const PgState st) {
    const PgSmemView sm = pg_smem_view(pg_smem_raw, st.S);
    if (threadIdx.x == 0) sm.base->pass_bid = (int)blockIdx.x;
    pg_fetch_cmd(reinterpret_cast<const int*>(st.ctrl), sm.base);
    __syncthreads();
    pg_run_pass<PGV_GENERIC>(st, sm, sm.base->cmd.phase);
}

extern "C" __global__ void __launch_bounds__(PG_THREADS, 1) pg_serial_kernel(const PgState st) {
    if (__ldcg(&st.ctrl->poisoned)) return;
    PgTeam tm{(int)threadIdx.x, (int)blockDim.x};
    pg_serial(st, tm);
}

extern "C" __global__ void pg_prepare_step(const PgState st, int tune) {
    PgCtrl* c = st.ctrl;
    c->poisoned = c->error != PG_OK ? 1 : 0;
    if (c->poisoned) { c->phase = PH_DONE; return; }
    if (tune >= 0) c->tune = tune;
    c->phase = PH_BEGIN;
    c->snext = SN_NONE;
    c->bar_count = 0u;
    c->bar_gen = 0u;
    c->cnt[0][0] = 0u;
    c->cnt[1][0] = 0u;
    for (int b = 0; b < st.grid + 32; ++b) st.blk_done[b] = 0u;
}

extern "C" __global__ void pg_fill_f64(double* p, long long n, double v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// Forest of m stumps (kernel.rs:43-45).
extern "C" __global__ void pg_init_forest(const PgState st) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= st.m) return;
    const size_t fb = (size_t)t * st.maxn;
    st.f_split_var[fb] = -1;
    st.f_split_val[fb] = pg_nan();
    st.f_thr32[fb] = 0.0f;
    st.f_leaf_val[fb] = st.init_leaf;
    st.f_size[t] = 1;
}

// leaf_indices of a forest tree by traversal (tree.rs:17 `leaf_indices`, materialised on demand).
extern "C" __global__ void pg_leaf_indices_kernel(const PgState st, int t, unsigned* out) {
    const size_t fb = (size_t)t * st.maxn;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < st.n; i += (long long)gridDim.x * blockDim.x) {
        int node = 0, sv;
        while ((sv = st.f_split_var[fb + node]) >= 0) {
            const float x = st.X[(size_t)sv * st.ld + i];
            node = 2 * node + 1 + (x < st.f_thr32[fb + node] ? 0 : 1);
        }
        out[i] = (unsigned)node;
    }
}

// Out-of-sample prediction of the whole forest (tree.rs:169-192 per tree, summed over trees in tree order):
// x is column-major fp32 [p][ldx], compared with thr32 = the smallest float >= split_val exactly as the training passes
// do; for fp32-representable x that IS the reference's `sample[sv] < split_val` in f64 (pg_thr32).
extern "C" __global__ void pg_predict_kernel(const PgState st, const float* x, long long n_new, long long ldx, double* out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_new; i += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int t = 0; t < st.m; ++t) {
            const size_t fb = (size_t)t * st.maxn;
            const int size = st.f_size[t];
            int node = 0, sv;
            while (node < size && (sv = st.f_split_var[fb + node]) >= 0)
                node = 2 * node + 1 + (x[(size_t)sv * ldx + i] < st.f_thr32[fb + node] ? 0 : 1);
            if (node < size) acc += st.f_leaf_val[fb + node];
        }
        out[i] = acc;
    }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------

int pg_query_grid(int device, int S, int smem_kb, int* grid_out, size_t* smem_out, int* stage_out, char* err, size_t errlen) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return 1; }
    // dynamic shared memory = the pass arrays + the union area (pg_smem_view): control state in block 0, staging ring of
    // the root pass in the pass blocks.  smem_kb = total to ask for (P <= 32; 0 = no ring, just what the control state needs).
    const void* kernels[] = {(const void*)pg_sweep_persistent_ring, (const void*)pg_sweep_persistent_direct, (const void*)pg_sweep_persistent_bern,
                             (const void*)pg_sweep_persistent_generic, (const void*)pg_phase_kernel};
    size_t static_smem = 0;
    for (const void* k : kernels) {
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, k);
        if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncGetAttributes: %s", cudaGetErrorString(e)); return 1; }
        if (fa.sharedSizeBytes > static_smem) static_smem = fa.sharedSizeBytes;
    }
    const size_t pass = pg_smem_pass_bytes(S), ctrl = pg_smem_ctrl_bytes();
    const size_t optin = prop.sharedMemPerBlockOptin;
    if (pass + ctrl + static_smem > optin) { snprintf(err, errlen, "shared memory: %zu B needed, %zu B available", pass + ctrl + static_smem, optin); return 1; }
    size_t want = (size_t)smem_kb * 1024;
    if (want + static_smem > optin) want = optin - static_smem;
    size_t stage = want > pass + 1024 ? ((want - pass) / 1024) * 1024 : 0;
    if (stage < ctrl) stage = 0;                         // no room for a useful ring: the union area only holds the control state
    const size_t smem = pass + (stage > ctrl ? stage : ctrl);
    for (const void* k : kernels) {
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { snprintf(err, errlen, "shared memory request %zu B: %s", smem, cudaGetErrorString(e)); return 1; }
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pg_sweep_persistent_ring, PG_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) { snprintf(err, errlen, "occupancy query failed: %s", cudaGetErrorString(e)); return 1; }
    if (per_sm > 1) per_sm = 1;   // one 512-thread block per SM: 148 partials per pass for the serial reduce
    if (!prop.cooperativeLaunch) { snprintf(err, errlen, "device lacks cooperative launch"); return 1; }
    *grid_out = per_sm * prop.multiProcessorCount - 1;   // pass blocks; one more SM hosts the serial block
    if (*grid_out < 1) { snprintf(err, errlen, "device has too few SMs"); return 1; }
    *smem_out = smem;
    *stage_out = (int)stage;
    return 0;
}

cudaError_t pg_launch_prepare(const PgState& st, int tune, cudaStream_t s) {
    pg_prepare_step<<<1, 1, 0, s>>>(st, tune);
    return cudaGetLastError();
}

cudaError_t pg_launch_persistent(const PgState& st, size_t smem, cudaStream_t s) {
    void* args[] = {(void*)&st};
    void* fn = (void*)pg_sweep_persistent_generic;
    if (!st.generic) {
        if (st.family == PG_BERNOULLI_LOGIT) fn = (void*)pg_sweep_persistent_bern;
        else if (pg_use_ring(st)) fn = (void*)pg_sweep_persistent_ring;
        else fn = (void*)pg_sweep_persistent_direct;
    }
    return cudaLaunchCooperativeKernel(fn, dim3(st.grid + 1), dim3(PG_THREADS), args, smem, s);   // + the serial block
}

cudaError_t pg_launch_phase(const PgState& st, size_t smem, cudaStream_t s) {
    pg_phase_kernel<<<st.grid, PG_THREADS, smem, s>>>(st);
    return cudaGetLastError();
}

cudaError_t pg_launch_serial(const PgState& st, cudaStream_t s) {
    pg_serial_kernel<<<1, PG_THREADS, 0, s>>>(st);
    return cudaGetLastError();
}

cudaError_t pg_launch_fill(double* p, long long n, double v, cudaStream_t s) {
    pg_fill_f64<<<296, 256, 0, s>>>(p, n, v);
    return cudaGetLastError();
}

cudaError_t pg_launch_init_forest(const PgState& st, cudaStream_t s) {
    pg_init_forest<<<(st.m + 127) / 128, 128, 0, s>>>(st);
    return cudaGetLastError();
}

cudaError_t pg_launch_predict(const PgState& st, const float* x, long long n_new, long long ldx, double* out, cudaStream_t s) {
    pg_predict_kernel<<<592, 256, 0, s>>>(st, x, n_new, ldx, out);
    return cudaGetLastError();
}

cudaError_t pg_launch_leaf_indices(const PgState& st, int t, unsigned* out, cudaStream_t s) {
    pg_leaf_indices_kernel<<<296, 256, 0, s>>>(st, t, out);
    return cudaGetLastError();
}
