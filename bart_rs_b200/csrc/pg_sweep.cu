// pg_sweep.cu — kernels of the PGBART sweep for sm_100a and their launchers.
//
//   pg_sweep_persistent  one cooperative launch per PySampler.step(): all tree updates of the
//                        step's batch (kernel.rs:81-115) run inside it.  Grid passes alternate
//                        with a serial section executed by the last block to reach the grid
//                        barrier (pg_serial.h); no host round trip inside a step.
//   pg_phase_kernel /    the same passes / serial section as separate launches with a host
//   pg_serial_kernel     read-back between them ("stepwise" driver: slow, used to cross-check
//                        the persistent kernel and under compute-sanitizer).
#include <cooperative_groups.h>
#include <cstdio>
#include "pg_passes.cuh"
#include "pg_sweep.h"

namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned pg_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pg_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid barrier whose last arriver runs the serial section before releasing the others.
// `gen` counts barriers passed by this block since kernel start (bar_count/bar_gen are zeroed by
// pg_prepare_step before every launch).  A watchdog turns a lost block into an error, not a hang.
__device__ void pg_barrier_serial(const PgState& st, PgSmem* sm, unsigned& gen) {
    PgCtrl* c = st.ctrl;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned ticket = atomicAdd(&c->bar_count, 1u);
        sm->is_last = (ticket == (gen + 1u) * gridDim.x - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (sm->is_last) {
        __threadfence();
        PgTeam tm{(int)threadIdx.x, (int)blockDim.x};
        pg_serial(st, tm);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) pg_st_release(&c->bar_gen, gen + 1u);
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            unsigned ns = 32;
            while (pg_ld_acquire(&c->bar_gen) < gen + 1u) {
                __nanosleep(ns);
                if (ns < 256) ns <<= 1;
                if (clock64() - t0 > 40000000000ll) {   // ~20 s at 2 GHz
                    c->error = PG_ERR_WATCHDOG;
                    __threadfence();
                    __trap();
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
    gen += 1u;
}

extern "C" __global__ void __launch_bounds__(PG_THREADS, 2) pg_sweep_persistent(const PgState st) {
    extern __shared__ __align__(16) unsigned char pg_smem_raw[];
    const PgSmemView sm = pg_smem_view(pg_smem_raw, st.S);
    unsigned gen = 0;
    pg_barrier_serial(st, sm.base, gen);   // serial section of PH_BEGIN prepares the first pass
    for (;;) {
        const int phase = __ldcg(&st.ctrl->phase);
        if (phase == PH_DONE) break;
        pg_run_pass(st, sm, phase);
        pg_barrier_serial(st, sm.base, gen);
    }
}

extern "C" __global__ void __launch_bounds__(PG_THREADS, 2) pg_phase_kernel(const PgState st) {
    extern __shared__ __align__(16) unsigned char pg_smem_raw[];
    const PgSmemView sm = pg_smem_view(pg_smem_raw, st.S);
    pg_run_pass(st, sm, __ldcg(&st.ctrl->phase));
}

extern "C" __global__ void __launch_bounds__(PG_THREADS, 1) pg_serial_kernel(const PgState st) {
    PgTeam tm{(int)threadIdx.x, (int)blockDim.x};
    pg_serial(st, tm);
}

extern "C" __global__ void pg_prepare_step(const PgState st, int tune) {
    PgCtrl* c = st.ctrl;
    if (tune >= 0) c->tune = tune;
    c->phase = PH_BEGIN;
    c->snext = SN_NONE;
    c->bar_count = 0u;
    c->bar_gen = 0u;
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

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------

int pg_query_grid(int device, int S, int* grid_out, size_t* smem_out, char* err, size_t errlen) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return 1; }
    const size_t smem = pg_smem_bytes(S);
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(pg_sweep_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pg_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { snprintf(err, errlen, "shared memory request %zu B: %s", smem, cudaGetErrorString(e)); return 1; }
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pg_sweep_persistent, PG_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) { snprintf(err, errlen, "occupancy query failed: %s", cudaGetErrorString(e)); return 1; }
    if (per_sm > 2) per_sm = 2;
    if (!prop.cooperativeLaunch) { snprintf(err, errlen, "device lacks cooperative launch"); return 1; }
    *grid_out = per_sm * prop.multiProcessorCount;
    *smem_out = smem;
    return 0;
}

cudaError_t pg_launch_prepare(const PgState& st, int tune, cudaStream_t s) {
    pg_prepare_step<<<1, 1, 0, s>>>(st, tune);
    return cudaGetLastError();
}

cudaError_t pg_launch_persistent(const PgState& st, size_t smem, cudaStream_t s) {
    void* args[] = {(void*)&st};
    return cudaLaunchCooperativeKernel((void*)pg_sweep_persistent, dim3(st.grid), dim3(PG_THREADS), args, smem, s);
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

cudaError_t pg_launch_leaf_indices(const PgState& st, int t, unsigned* out, cudaStream_t s) {
    pg_leaf_indices_kernel<<<296, 256, 0, s>>>(st, t, out);
    return cudaGetLastError();
}
