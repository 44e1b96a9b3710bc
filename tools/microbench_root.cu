// microbench_root.cu — the root pass (19 jobs: per row a compare and two masked f64 FMAs per job, plus A r/w and y) in
// isolation, persistent-kernel shape (147 blocks x 512 threads), to compare loop structures without the monolith:
//   direct   : registers, 4 jobs x 256 rows per phase (the shipped pass's structure)
//   pipeline : one producer warp issues cp.async.bulk copies (one per array per tile) into a ring of NST stages guarded by
//              full / empty mbarriers; 15 consumer warps = 3 row parts x 5 job groups with per-lane accumulators
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/_mbr tools/microbench_root.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define THREADS 512
#define WARPS 16
#define P 50
#define J 19
#define FULLM 0xFFFFFFFFu

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    const unsigned a = smem_u32(b);
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 lds_f4(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_d2(unsigned a) { double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void acc_left(double& a, double& b, unsigned& c, float x, float thr, double o, double r) {
    const bool left = x < thr; const double m = left ? 1.0 : 0.0; a = fma(m, o, a); b = fma(m, r, b); c += left ? 1u : 0u;
}
__device__ __forceinline__ double warp_sum(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o); return v; }

struct Args { const float* X; const float* y; const double* A; double* A2; long long n, ld; int var[J]; float thr[J]; double* out; int rot; int nocompute; };

// ---- the shipped structure: 4 jobs x 256 rows per phase, cross-lane reduction per phase into shared accumulators
__global__ void __launch_bounds__(THREADS, 1) k_direct(const Args a) {
    __shared__ double accA[WARPS][J], accB[WARPS][J];
    __shared__ unsigned accC[WARPS][J];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < WARPS * J; i += THREADS) { (&accA[0][0])[i] = 0; (&accB[0][0])[i] = 0; (&accC[0][0])[i] = 0; }
    __syncthreads();
    const long long per = ((a.n / 4 + gridDim.x - 1) / gridDim.x) * 4;
    const long long B0 = blockIdx.x * per, B1 = min(a.n, B0 + per);
    const long long wper = (((B1 - B0) / 4 + WARPS - 1) / WARPS) * 4;
    const long long beg = B0 + warp * wper, end = min(B1, beg + wper);
    double tot = 0;
    for (long long base = beg; base < end; base += 256) {
        const long long r0 = base + 4 * lane, r1 = r0 + 128;
        const bool g0 = r0 < end, g1 = r1 < end;
        double o[8], r[8];
        {
            double2 v; float4 yy;
            if (g0) { v = *(const double2*)(a.A + r0); o[0] = v.x; o[1] = v.y; v = *(const double2*)(a.A + r0 + 2); o[2] = v.x; o[3] = v.y; yy = __ldcs((const float4*)(a.y + r0)); r[0] = yy.x; r[1] = yy.y; r[2] = yy.z; r[3] = yy.w; }
            else { for (int q = 0; q < 4; ++q) { o[q] = 0; r[q] = 0; } }
            if (g1) { v = *(const double2*)(a.A + r1); o[4] = v.x; o[5] = v.y; v = *(const double2*)(a.A + r1 + 2); o[6] = v.x; o[7] = v.y; yy = __ldcs((const float4*)(a.y + r1)); r[4] = yy.x; r[5] = yy.y; r[6] = yy.z; r[7] = yy.w; }
            else { for (int q = 4; q < 8; ++q) { o[q] = 0; r[q] = 0; } }
        }
        for (int q = 0; q < 8; ++q) { o[q] += 0.125; r[q] = r[q] - o[q]; tot += r[q] * r[q]; }
        if (g0) { *(double2*)(a.A2 + r0) = make_double2(o[0], o[1]); *(double2*)(a.A2 + r0 + 2) = make_double2(o[2], o[3]); }
        if (g1) { *(double2*)(a.A2 + r1) = make_double2(o[4], o[5]); *(double2*)(a.A2 + r1 + 2) = make_double2(o[6], o[7]); }
        for (int jb = 0; jb < J; jb += 4) {
            float4 xa[4], xb[4];
            const float inf = INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool live = jb + q < J;
                const float* col = a.X + (size_t)((a.var[live ? jb + q : 0] + a.rot) % P) * a.ld;
                xa[q] = (live && g0) ? __ldcs((const float4*)(col + r0)) : make_float4(inf, inf, inf, inf);
                xb[q] = (live && g1) ? __ldcs((const float4*)(col + r1)) : make_float4(inf, inf, inf, inf);
            }
            double so[4], sr[4]; unsigned cn[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float thr = jb + q < J ? a.thr[jb + q] : -inf;
                const float xs[8] = {xa[q].x, xa[q].y, xa[q].z, xa[q].w, xb[q].x, xb[q].y, xb[q].z, xb[q].w};
                double sa = 0, sb = 0; unsigned c = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc_left(sa, sb, c, xs[e], thr, o[e], r[e]);
                so[q] = sa; sr[q] = sb; cn[q] = c;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {      // (plain butterflies here; the shipped pass uses a 4-job transposed reduction)
                const double ta = warp_sum(so[q]), tb = warp_sum(sr[q]);
                const unsigned tc = __reduce_add_sync(FULLM, cn[q]);
                if (lane == 0 && jb + q < J) { accA[warp][jb + q] += ta; accB[warp][jb + q] += tb; accC[warp][jb + q] += tc; }
            }
        }
    }
    __syncthreads();
    if (tid < J) { double s = 0; for (int w = 0; w < WARPS; ++w) s += accA[w][tid] + accB[w][tid] + accC[w][tid]; a.out[blockIdx.x * 32 + tid] = s + tot; }
}

// ---- producer / consumer pipeline: NP producer warps (a bulk copy costs its issuing warp ~150 cycles, lanes of one warp
// are served one after the other), 12 consumer warps = 3 row parts x 4 job groups of up to 5 jobs
template <int NST, int TR, int NP, int RP>      // stages; rows per tile; producer warps; row parts (consumer warps = RP x 12/RP job groups)
__global__ void __launch_bounds__(THREADS, 1) k_pipe(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long full[NST], empty[NST];
    __shared__ double racc[4][J][2];
    __shared__ unsigned rcnt[4][J];
    constexpr int NC = 12;
    constexpr int JG = NC / RP, JPG = (J + JG - 1) / JG;   // job groups, jobs per group
    constexpr int RPL = TR / (32 * RP);          // rows per lane per tile (4 or 8)
    constexpr unsigned STAGE = TR * (12u + 4u * J);
    constexpr int NARR = 2 + J;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < NST) { mbar_init(&full[tid], NP); mbar_init(&empty[tid], NC); }
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const long long per = ((a.n / 4 + gridDim.x - 1) / gridDim.x) * 4;
    const long long B0 = blockIdx.x * per, B1 = min(a.n, B0 + per);
    const int n_tiles = (int)((B1 - B0 + TR - 1) / TR);
    const unsigned s0 = smem_u32(smem);
    if (warp >= NC) {
        if (warp >= NC + NP) return;
        const int pw = warp - NC;
        const int arr = pw + NP * lane;          // this lane's array
        for (int t = 0; t < n_tiles; ++t) {
            const int st = t % NST;
            if (t >= NST) mbar_wait(&empty[st], (unsigned)((t / NST - 1) & 1));
            const long long r0 = B0 + (long long)t * TR;
            const unsigned rows = (unsigned)min((long long)TR, B1 - r0);
            unsigned mine = arr < NARR ? rows * (arr == 0 ? 8u : 4u) : 0u;
            unsigned bytes = mine;
            for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(FULLM, bytes, o);
            if (lane == 0) { asm volatile("fence.proxy.async;" ::: "memory"); mbar_expect_tx(&full[st], bytes); }
            __syncwarp();
            const unsigned sb = s0 + st * STAGE;
            if (arr == 0) bulk_g2s(sb, a.A + r0, rows * 8u, &full[st]);
            else if (arr == 1) bulk_g2s(sb + TR * 8u, a.y + r0, rows * 4u, &full[st]);
            else if (arr < NARR) bulk_g2s(sb + TR * 12u + (arr - 2) * (TR * 4u), a.X + (size_t)((a.var[arr - 2] + a.rot) % P) * a.ld + r0, rows * 4u, &full[st]);
        }
        return;
    }
    const int rp = warp / JG, jg = warp % JG;    // row part, job group (jobs JPG jg .. )
    const int nj = max(0, min(JPG, J - JPG * jg));
    double accA[JPG], accB[JPG];
    unsigned accC[JPG];
    float thr[JPG];
    for (int q = 0; q < JPG; ++q) { accA[q] = 0; accB[q] = 0; accC[q] = 0; thr[q] = q < nj ? a.thr[JPG * jg + q] : -INFINITY; }
    double tot = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const int st = t % NST;
        mbar_wait(&full[st], (unsigned)((t / NST) & 1));
        const unsigned sb = s0 + st * STAGE;
        const long long r0 = B0 + (long long)t * TR;
        const int rows = (int)min((long long)TR, B1 - r0);
#pragma unroll
        for (int h = 0; h < RPL / 4; ++h) {
            const int lr = rp * (TR / RP) + h * 128 + lane * 4;
            if (lr < rows && !a.nocompute) {
                const double2 a01 = lds_d2(sb + lr * 8u), a23 = lds_d2(sb + lr * 8u + 16u);
                const float4 yy = lds_f4(sb + TR * 8u + lr * 4u);
                double o[4] = {a01.x + 0.125, a01.y + 0.125, a23.x + 0.125, a23.y + 0.125};
                double r[4] = {yy.x - o[0], yy.y - o[1], yy.z - o[2], yy.w - o[3]};
                if (jg == 0) {
                    *(double2*)(a.A2 + r0 + lr) = make_double2(o[0], o[1]); *(double2*)(a.A2 + r0 + lr + 2) = make_double2(o[2], o[3]);
                    tot += r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
                }
#pragma unroll
                for (int q = 0; q < JPG; ++q) {
                    if (q < nj) {
                        const float4 xv = lds_f4(sb + TR * 12u + (JPG * jg + q) * (TR * 4u) + lr * 4u);
                        acc_left(accA[q], accB[q], accC[q], xv.x, thr[q], o[0], r[0]);
                        acc_left(accA[q], accB[q], accC[q], xv.y, thr[q], o[1], r[1]);
                        acc_left(accA[q], accB[q], accC[q], xv.z, thr[q], o[2], r[2]);
                        acc_left(accA[q], accB[q], accC[q], xv.w, thr[q], o[3], r[3]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
    for (int q = 0; q < JPG; ++q) {
        const double sa = warp_sum(accA[q]), sb2 = warp_sum(accB[q]);
        const unsigned sc = __reduce_add_sync(FULLM, accC[q]);
        if (lane == 0 && q < nj) { racc[rp][JPG * jg + q][0] = sa; racc[rp][JPG * jg + q][1] = sb2; rcnt[rp][JPG * jg + q] = sc; }
    }
    tot = warp_sum(tot);
    asm volatile("bar.sync 1, 384;" ::: "memory");      // the 12 consumer warps
    if (tid < J) { double s = 0; for (int q = 0; q < RP; ++q) s += racc[q][tid][0] + racc[q][tid][1] + rcnt[q][tid]; a.out[blockIdx.x * 32 + tid] = s + (tid == 0 ? tot : 0.0); }
}

int main(int argc, char** argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 1000000;
    const long long ld = (n + 31) / 32 * 32;
    float *X, *y; double *A, *A2, *out;
    cudaMalloc(&X, ld * P * 4); cudaMalloc(&y, ld * 4); cudaMalloc(&A, ld * 8); cudaMalloc(&A2, ld * 8); cudaMalloc(&out, 148 * 32 * 8);
    {
        float* h = (float*)malloc(ld * P * 4);
        unsigned s = 12345u;
        for (long long i = 0; i < ld * P; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) * (1.0f / 16777216.0f); }
        cudaMemcpy(X, h, ld * P * 4, cudaMemcpyHostToDevice); cudaMemcpy(y, h, ld * 4, cudaMemcpyHostToDevice); free(h);
    }
    cudaMemset(A, 0, ld * 8);
    Args a; a.X = X; a.y = y; a.A = A; a.A2 = A2; a.n = n; a.ld = ld; a.out = out; a.rot = 0;
    for (int j = 0; j < J; ++j) { a.var[j] = (j * 7 + 3) % P; a.thr[j] = 0.3f + 0.02f * j; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 147, reps = 40;
    const int rotate = argc > 2 ? atoi(argv[2]) : 1;
    a.nocompute = argc > 3 ? atoi(argv[3]) : 0;
    printf("n=%lld rotate=%d nocompute=%d\n", n, rotate, a.nocompute);
    auto run = [&](const char* name, auto launch) {
        a.rot = 0; launch(); launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) { a.rot = rotate ? (r * 11) % P : 0; launch(); }    // different columns every pass: X (200 MB) does not stay in L2
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double h[32]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        printf("%-28s %7.2f us per pass (incl. ~3 us launch)  check %.6e  %s\n", name, ms * 1000.0 / reps, h[1], cudaGetErrorString(cudaGetLastError()));
    };
    run("direct (registers)", [&] { k_direct<<<grid, THREADS>>>(a); });
#define PIPE(NST, TR, NP, RP) { const size_t sm = (size_t)NST * TR * (12 + 4 * J); cudaFuncSetAttribute(k_pipe<NST, TR, NP, RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        char nm[64]; snprintf(nm, sizeof nm, "pipe NST=%d tile=%d NP=%d RP=%d (%zu KB)", NST, TR, NP, RP, sm / 1024); run(nm, [&] { k_pipe<NST, TR, NP, RP><<<grid, THREADS, sm>>>(a); }); }
    PIPE(2, 768, 2, 3) PIPE(3, 512, 2, 2) PIPE(3, 512, 4, 2) PIPE(4, 512, 4, 2) PIPE(3, 512, 4, 4) PIPE(2, 1024, 4, 4) PIPE(2, 1024, 4, 2) PIPE(3, 768, 4, 3)
    return 0;
}
