// microbench_stream.cu — how fast can 147 blocks x 512 threads (one block per SM, persistent-kernel shape) stream U
// columns of n floats plus an fp64 array, by (a) a cp.async ring in shared memory, (b) plain float4 register loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mb tools/microbench_stream.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define T 512
#define THREADS 512
#define P 160          // columns: 640 MB at n = 1M, far beyond L2; every rep touches different ones

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(unsigned dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NST>
__global__ void __launch_bounds__(THREADS, 1) k_ring(const float* X, const double* A, long long n, long long ld, int U, int reps, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const long long per = ((n / 4 + gridDim.x - 1) / gridDim.x) * 4;
    const long long B0 = blockIdx.x * per, B1 = min(n, B0 + per);
    const int n_tiles = (int)((B1 - B0 + T - 1) / T);
    const unsigned stage_bytes = T * (8 + 4 * U);
    const int n_chunks = 256 + 128 * U;
    float acc = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
        auto issue = [&](int t) {
            if (t < n_tiles) {
                const long long r0 = B0 + (long long)t * T;
                const int rows = (int)min((long long)T, B1 - r0);
                const unsigned sb = smem_u32(smem) + (t % NST) * stage_bytes;
                for (int ch = tid; ch < n_chunks; ch += THREADS) {
                    if (ch < 256) { if (ch * 2 < rows) cp16(sb + ch * 16, A + r0 + ch * 2); }
                    else { const int u = (ch - 256) >> 7, q = (ch - 256) & 127; if (q * 4 < rows) cp16(sb + T * 8 + u * (T * 4) + q * 16, X + (size_t)((u * 7 + rep * 17) % P) * ld + r0 + q * 4); }
                }
            }
            cp_commit();
        };
        for (int t = 0; t < NST - 1; ++t) issue(t);
        int stg = 0;
        for (int t = 0; t < n_tiles; ++t) {
            cp_wait<NST - 2>();
            __syncthreads();
            issue(t + NST - 1);
            const unsigned char* base = smem + stg * stage_bytes;
            stg = stg + 1 == NST ? 0 : stg + 1;
            const float* cols = (const float*)(base + T * 8);
            for (int u = tid >> 7; u < U; u += 4) acc += cols[u * T + (tid & 127) * 4];
        }
        cp_wait<0>();
        __syncthreads();
    }
    if (acc == 12345.678f) out[0] = acc;
}

__global__ void __launch_bounds__(THREADS, 1) k_regs(const float* X, const double* A, long long n, long long ld, int U, int reps, float* out) {
    const int tid = threadIdx.x;
    const long long per = ((n / 4 + gridDim.x - 1) / gridDim.x) * 4;
    const long long B0 = blockIdx.x * per, B1 = min(n, B0 + per);
    float acc = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
        for (long long r = B0 + tid * 4; r < B1; r += THREADS * 4) {
            const double2 a = *(const double2*)(A + r);
            acc += (float)a.x;
            for (int u0 = 0; u0 < U; u0 += 8) {
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = u0 + q < U ? __ldcs((const float4*)(X + (size_t)(((u0 + q) * 7 + rep * 17) % P) * ld + r)) : make_float4(0, 0, 0, 0);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc += v[q].x + v[q].w;
            }
        }
    }
    if (acc == 12345.678f) out[0] = acc;
}

int main(int argc, char** argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 1000000;
    const long long ld = (n + 31) / 32 * 32;
    const int p = P, reps = 20;
    float *X, *out; double* A;
    cudaMalloc(&X, ld * p * 4); cudaMalloc(&A, ld * 8); cudaMalloc(&out, 4);
    cudaMemset(X, 0, ld * p * 4); cudaMemset(A, 0, ld * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 147;
    for (int U : {0, 1, 4, 8, 15, 19}) {
        const size_t stage = T * (8 + 4 * U);
        auto run = [&](auto kern, int nst, const char* name) {
            const size_t smem = stage * nst;
            if (smem > 200 * 1024) return;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, THREADS, smem>>>(X, A, n, ld, U, 2, out);
            cudaEventRecord(e0);
            kern<<<grid, THREADS, smem>>>(X, A, n, ld, U, reps, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double us = ms * 1000.0 / reps, gb = (double)n * (8 + 4.0 * U) / 1e9;
            printf("U=%2d %-10s nst=%d smem=%6zu  %7.2f us/pass  %7.1f GB/s  err=%s\n", U, name, nst, smem, us, gb / (us * 1e-6), cudaGetErrorString(cudaGetLastError()));
        };
        run(k_ring<3>, 3, "ring");
        run(k_ring<4>, 4, "ring");
        run(k_ring<6>, 6, "ring");
        run(k_ring<8>, 8, "ring");
        cudaEventRecord(e0);
        k_regs<<<grid, THREADS>>>(X, A, n, ld, U, reps, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1000.0 / reps, gb = (double)n * (8 + 4.0 * U) / 1e9;
        printf("U=%2d %-10s                     %7.2f us/pass  %7.1f GB/s\n", U, "regs", us, gb / (us * 1e-6));
    }
    return 0;
}
