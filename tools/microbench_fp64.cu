// FP64 pipe of one SM (B200): dependent-chain latency and throughput of DFMA, and the cost of the Bernoulli row-job
// (csrc/pg_passes.cuh: pg_pass_bern) in isolation.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _mbf tools/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(512, 1) k_dfma(double* out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ double2 lds_d2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double log_ge1(double u, unsigned tab) {
    const int hi = __double2hiint(u);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(u));
    const double2 t = lds_d2(tab + (unsigned)(((hi >> 14) & 63) * 128));
    const double e = (double)((hi >> 20) - 1023);
    const double r = fma(m, t.x, -1.0);
    double p = 1.0 / 7.0;
    p = fma(p, r, -1.0 / 6.0); p = fma(p, r, 0.2); p = fma(p, r, -0.25); p = fma(p, r, 1.0 / 3.0); p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    return fma(p, r, fma(e, 0.6931471805599453, t.y));
}

// R rows per lane held in registers, `jobs` jobs: the arithmetic of the Bernoulli pass without its memory traffic
template <int R>
__global__ void __launch_bounds__(512, 1) k_bern(double* out, int jobs, int reps, const double* tbl) {
    __shared__ __align__(16) double tab[64 * 8 * 2];
    __shared__ double jv[4][256];
    __shared__ float jt[256];
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) { tab[2 * i] = tbl[2 * (i / 8)]; tab[2 * i + 1] = tbl[2 * (i / 8) + 1]; }
    for (int j = threadIdx.x; j < 256; j += blockDim.x) { jv[0][j] = 0.01 * j; jv[1][j] = -0.01 * j; jv[2][j] = exp(0.01 * j); jv[3][j] = exp(-0.01 * j); jt[j] = 0.3f + 0.001f * j; }
    __syncthreads();
    const unsigned taddr = (unsigned)__cvta_generic_to_shared(tab) + (threadIdx.x & 7) * 16;
    double o[R], eo[R], yd[R];
    float x[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { o[k] = 0.001 * threadIdx.x - 0.2 + k; eo[k] = exp(o[k]); yd[k] = (threadIdx.x + k) & 1; x[k] = 0.002f * ((threadIdx.x * 7 + k * 13) % 500); }
    double acc = 0.0;
    for (int rep = 0; rep < reps; ++rep)
        for (int j = 0; j < jobs; ++j) {
            const float thr = jt[j];
            const double vL = jv[0][j], vR = jv[1][j], evL = jv[2][j], evR = jv[3][j];
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const bool left = x[k] < thr;
                const double mu = o[k] + (left ? vL : vR);
                const double f = yd[k] * mu - log_ge1(fma(eo[k], left ? evL : evR, 1.0), taddr);
                a = fma(left ? 1.0 : 0.0, f, a);
                b = fma(left ? 0.0 : 1.0, f, b);
            }
            acc += a - b;
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const double ghz = prop.clockRate / 1e6;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 512);
    double htbl[128];
    for (int k = 0; k < 64; ++k) { const double c = 1.0 + (k + 0.5) / 64; const double inv = floor(4096.0 / c + 0.5) / 4096.0; htbl[2 * k] = inv; htbl[2 * k + 1] = -log(inv); }
    double* tbl; cudaMalloc(&tbl, sizeof htbl); cudaMemcpy(tbl, htbl, sizeof htbl, cudaMemcpyHostToDevice);
    printf("%s, %d SMs, %.3f GHz (nominal)\n", prop.name, sms, ghz);
    const int iters = 20000;
#define RUN_DFMA(ILP) { float ms = time_ms([&] { k_dfma<ILP><<<sms, 512>>>(out, iters, 1.0000001, 1e-9); }); \
        const double per_clk = (double)iters * ILP * 512 / (ms * 1e-3 * ghz * 1e9); \
        printf("DFMA ILP %d, 16 warps/SM: %.1f lanes/clk/SM, %.1f TFLOP/s chip; cycles per dependent step per warp %.1f\n", ILP, per_clk, per_clk * 2 * sms * ghz / 1e3, ms * 1e-3 * ghz * 1e9 / iters); }
    RUN_DFMA(1) RUN_DFMA(2) RUN_DFMA(4) RUN_DFMA(8)
    const int jobs = 255, reps = 40;
#define RUN_BERN(R) { float ms = time_ms([&] { k_bern<R><<<sms, 512>>>(out, jobs, reps, tbl); }); \
        const double rowjobs = (double)jobs * reps * R * 512; \
        printf("Bernoulli row-job, R=%d rows/lane: %.2f clk per warp-row-job per SM (%.2f ns per 1e3 row-jobs per SM), C5 pass of 2.55e8 row-jobs on %d SMs: %.0f us\n", R, ms * 1e-3 * ghz * 1e9 / (rowjobs / 32), ms * 1e6 / rowjobs * 1e3, sms - 1, 2.55e8 / (sms - 1) * (ms * 1e3 / rowjobs)); }
    RUN_BERN(1) RUN_BERN(2) RUN_BERN(4) RUN_BERN(8)
    return 0;
}
