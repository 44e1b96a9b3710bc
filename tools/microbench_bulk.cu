// microbench_bulk.cu — what the bulk-copy engine of one SM sustains: 147 blocks (one per SM, the persistent-kernel shape)
// each stream their share of U columns + A + y of n rows into shared memory with cp.async.bulk, nobody consuming the data.
// Varied: bytes per copy, how many warps issue concurrently (lane 0 of W warps, each with its own ring of D slots).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mbb tools/microbench_bulk.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define THREADS 512
#define P 160

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk(unsigned dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// W issuing warps, each lane 0 owns D slots of `cb` bytes; the block's byte range of every array is cut into copies of cb bytes
__global__ void __launch_bounds__(THREADS, 1) k_bulk(const float* X, long long n, long long ld, int U, int cb, int W, int D, int reps, long long* cyc_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bars[16 * 16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long per = ((n / 4 + gridDim.x - 1) / gridDim.x) * 4;
    const long long B0 = blockIdx.x * per, B1 = min(n, B0 + per);
    const long long bytes_per_arr = (B1 - B0) * 4;
    const int copies_per_arr = (int)((bytes_per_arr + cb - 1) / cb);
    const int n_copies = copies_per_arr * (U + 3);          // A counts as two float arrays, y as one
    if (tid < 256) mbar_init(&bars[tid], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    long long t_issue = 0;
    if (warp < W && lane == 0) {
        unsigned long long* mb = &bars[warp * 16];
        const unsigned base = smem_u32(smem) + (unsigned)warp * D * cb;
        int k = 0;                                            // copies issued by this warp so far
        for (int rep = 0; rep < reps; ++rep) {
            for (int c = warp; c < n_copies; c += W, ++k) {
                const int slot = k % D;
                if (k >= D) mbar_wait(&mb[slot], (unsigned)((k / D - 1) & 1));
                const int arr = c / copies_per_arr, piece = c % copies_per_arr;
                const long long off = (long long)piece * cb;
                const unsigned bytes = (unsigned)min((long long)cb, bytes_per_arr - off);
                const char* src = (const char*)(X + (size_t)((arr * 7 + rep * 17) % P) * ld + B0) + off;
                const long long t0 = clock64();
                mbar_expect(&mb[slot], bytes);
                bulk(base + slot * cb, src, bytes, &mb[slot]);
                t_issue += clock64() - t0;
            }
        }
        for (int q = 0; q < D && q < k; ++q) { const int kk = k - 1 - q; mbar_wait(&mb[kk % D], (unsigned)((kk / D) & 1)); }
        if (blockIdx.x == 0 && warp == 0) cyc_out[0] = t_issue / (k > 0 ? k : 1);
    }
    __syncthreads();
}

int main(int argc, char** argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 1000000;
    const long long ld = (n + 31) / 32 * 32;
    const int reps = 20, grid = 147, U = 15;
    float* X; long long* cyc;
    cudaMalloc(&X, ld * P * 4); cudaMalloc(&cyc, 8);
    cudaMemset(X, 0, ld * P * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const double gb = (double)n * 4.0 * (U + 3) / 1e9;
    for (int cb : {1024, 2048, 4096, 8192, 16384, 28672})
        for (int W : {1, 2, 4, 8, 16})
            for (int D : {2, 4, 8}) {
                if ((size_t)W * D * cb > 192 * 1024) continue;
                k_bulk<<<grid, THREADS, 200 * 1024>>>(X, n, ld, U, cb, W, D, 2, cyc);
                cudaEventRecord(e0);
                k_bulk<<<grid, THREADS, 200 * 1024>>>(X, n, ld, U, cb, W, D, reps, cyc);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
                const double us = ms * 1000.0 / reps;
                printf("copy=%5d B  issuers=%2d  depth=%d  in-flight/SM=%6d B  %7.2f us/pass  %7.1f GB/s  issue %lld cyc/copy  %s\n", cb, W, D, W * D * cb, us,
                       gb / (us * 1e-6), h, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
