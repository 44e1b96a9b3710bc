// pg_capi.cu — host side of the C ABI declared in include/pgbart.h.
//
// Owns all device memory of one sampler (the reference's PySampler owns copies of X, y, the
// forest and predictions: lib.rs:106-111,122-123) and one CUDA stream per handle
// (`#[pyclass(unsendable)]`, lib.rs:106: one sampler per host thread).
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>
#include "../../include/pgbart.h"
#include "pg_sweep.h"
#include "pg_types.h"

namespace {

thread_local std::string g_create_error;   // per host thread: concurrent pgbart_create calls do not race

struct Sampler {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tm0 = nullptr, tm1 = nullptr;
    PgState st{};
    std::vector<void*> allocs;
    std::string err;
    bool dead = false;
    bool stepwise = false;
    size_t smem = 0;
    double launches = 0;
    bool timing_pending = false;
    unsigned long long pool_factor = 2;
    // owned optional buffers
    double *tape_slot = nullptr, *tape_round = nullptr, *tape_upd = nullptr;
    double *tr_slot = nullptr, *tr_round = nullptr, *tr_upd = nullptr;
    long long tr_u_cap = 0;
    int tr_r_cap = 0;
    PgCtrl host_ctrl{};
    bool allow_rounding = false;     // created with PGBART_ALLOW_FP32_ROUNDING: x / y (and predict inputs) are rounded to fp32 knowingly
    double* pin_pred = nullptr;      // pinned staging buffer of pgbart_step's device->host copy (lazily allocated, n doubles)
    std::vector<cudaEvent_t> chunk_ev;   // one event per staged chunk of that copy
    std::vector<void*> ipc_open;     // peer mailboxes opened through CUDA IPC (observation sharding)
    long long steps_done = 0;
    unsigned long long* hostdiag = nullptr;   // host-mapped: filled by a block that traps on a watchdog
    bool any_onehot = false;         // some split_rules entry is OneHotSplit: the step runs on the generic control code
};

const char* pg_error_text(int code) {
    switch (code) {
        case PG_ERR_DEPTH: return "Tree mutation would exceed maximum capacity (the reference panics here, tree.rs:98-102)";
        case PG_ERR_POOL: return "segment pool exhausted (raise option pool_factor)";
        case PG_ERR_TAPE: return "RNG tape exhausted or a taped split draw is not below the node maximum";
        case PG_ERR_WATCHDOG: return "grid barrier watchdog fired";
        case PG_ERR_WEIGHTS: return "final particle weights are not finite (the reference panics in WeightedIndex::new)";
        case PG_ERR_INTERNAL: return "internal consistency check failed";
        case PG_ERR_PRUNE: return "prune_dead_proposals: residual sum of squares too large for the weight bound (disable the option)";
        default: return "unknown device error";
    }
}

bool ck(Sampler* s, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    s->err = std::string(what) + ": " + cudaGetErrorString(e);
    if (e != cudaErrorInvalidValue && e != cudaErrorMemoryAllocation) s->dead = true;
    return false;
}

template <typename T>
bool dalloc(Sampler* s, T** out, size_t count, bool zero = true) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if (!ck(s, cudaMalloc(&p, bytes), "cudaMalloc")) return false;
    s->allocs.push_back(p);
    if (zero && !ck(s, cudaMemsetAsync(p, 0, bytes, s->stream), "cudaMemset")) return false;
    *out = (T*)p;
    return true;
}

// ndarray's mean(): sum()/n with the 8-way unrolled fold of numeric_util (kernel.rs:40, smc.rs:40).
double unrolled_mean(const double* xs, size_t n) {
    if (n == 0) return 0.0;
    double p[8] = {0, 0, 0, 0, 0, 0, 0, 0}, acc = 0.0;
    size_t len = n;
    while (len >= 8) {
        for (int j = 0; j < 8; ++j) p[j] += xs[j];
        xs += 8; len -= 8;
    }
    acc += p[0] + p[4];
    acc += p[1] + p[5];
    acc += p[2] + p[6];
    acc += p[3] + p[7];
    for (size_t i = 0; i < len && i < 7; ++i) acc += xs[i];
    return acc / (double)n;
}

// The device stores X and y as fp32 (north_star: "fp32 data, fp64 accumulation"); the reference computes on the f64 values.
// A value that is not exactly representable in fp32 would be silently quantised: distinct values can collapse (a column
// like 1e9 + k becomes constant), ties and child counts change, and decisions diverge from the reference.  So f64 input is
// checked and refused unless the caller passes PGBART_ALLOW_FP32_ROUNDING.  Returns the index of the first offending
// element, or -1.  NaN / inf pass (they round to themselves).
template <typename T>
long long first_not_fp32(const T* v, size_t count) {
    for (size_t i = 0; i < count; ++i) {
        const double d = (double)v[i];
        if ((double)(float)d != d && d == d) return (long long)i;
    }
    return -1;
}

// X (n x p, f64 or f32, C or Fortran order, host) -> column-major fp32 [p][ld] in HBM, with per-column min / max
// (splitting.rs:43-49 at the root is a per-column constant).  Row blocks are converted by a few host threads into pinned
// staging tiles [p][R] and copied with one 2-D async copy each, so a 10M x 100 matrix (BASELINE C4) uploads at memory
// speed instead of p strided walks over the whole matrix.  Padding rows [n, ld) hold +inf: they compare as "not left"
// for every threshold, so block-unit passes need no row masks on x.
bool upload_columns(Sampler* s, const void* x, int x_dtype, int x_order, uint64_t n, uint64_t p, uint64_t ld,
                    float* dX, float* cmin, float* cmax) {
    const uint64_t R = 8192;                                   // rows per tile (a multiple of 32)
    const uint64_t n_tiles = (ld + R - 1) / R;
    unsigned T = std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 8) T = 8;
    if (T > n_tiles) T = (unsigned)n_tiles;
    std::vector<std::vector<float>> tmin(T, std::vector<float>(p, std::numeric_limits<float>::infinity()));
    std::vector<std::vector<float>> tmax(T, std::vector<float>(p, -std::numeric_limits<float>::infinity()));
    std::vector<cudaError_t> terr(T, cudaSuccess);
    auto work = [&](unsigned t) {
        cudaSetDevice(s->device);
        cudaStream_t st_ = nullptr;
        float* stage[2] = {nullptr, nullptr};
        cudaEvent_t done[2] = {nullptr, nullptr};
        cudaError_t e = cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
            e = cudaHostAlloc((void**)&stage[b], (size_t)p * R * 4, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
        }
        int buf = 0;
        for (uint64_t tile = t; tile < n_tiles && e == cudaSuccess; tile += T, buf ^= 1) {
            const uint64_t i0 = tile * R, i1 = std::min(i0 + R, ld), rows = i1 - i0;
            e = cudaEventSynchronize(done[buf]);               // the copy that last used this staging tile has finished
            if (e != cudaSuccess) break;
            float* sg = stage[buf];
            const uint64_t iv = std::min(i1, n);               // rows below iv are real
            auto conv = [&](auto* src) {
                if (x_order == PGBART_ORDER_F) {
                    for (uint64_t j = 0; j < p; ++j) {
                        float mn = tmin[t][j], mx = tmax[t][j];
                        float* d = sg + j * R;
                        const auto* c = src + (size_t)j * n;
                        for (uint64_t i = i0; i < iv; ++i) { const float v = (float)c[i]; d[i - i0] = v; mn = std::fmin(mn, v); mx = std::fmax(mx, v); }
                        tmin[t][j] = mn; tmax[t][j] = mx;
                    }
                } else {
                    for (uint64_t i = i0; i < iv; ++i) {
                        const auto* r = src + (size_t)i * p;
                        for (uint64_t j = 0; j < p; ++j) sg[j * R + (i - i0)] = (float)r[j];
                    }
                    for (uint64_t j = 0; j < p; ++j) {
                        float mn = tmin[t][j], mx = tmax[t][j];
                        const float* d = sg + j * R;
                        for (uint64_t i = 0; i < iv - std::min(iv, i0); ++i) { mn = std::fmin(mn, d[i]); mx = std::fmax(mx, d[i]); }
                        tmin[t][j] = mn; tmax[t][j] = mx;
                    }
                }
            };
            if (iv > i0) { if (x_dtype == PGBART_X_F32) conv((const float*)x); else conv((const double*)x); }
            for (uint64_t j = 0; j < p; ++j)
                for (uint64_t i = std::max(iv, i0); i < i1; ++i) sg[j * R + (i - i0)] = std::numeric_limits<float>::infinity();
            e = cudaMemcpy2DAsync(dX + i0, (size_t)ld * 4, sg, (size_t)R * 4, (size_t)rows * 4, (size_t)p, cudaMemcpyHostToDevice, st_);
            if (e == cudaSuccess) e = cudaEventRecord(done[buf], st_);
        }
        if (st_) { const cudaError_t e2 = cudaStreamSynchronize(st_); if (e == cudaSuccess) e = e2; }
        for (int b = 0; b < 2; ++b) { if (done[b]) cudaEventDestroy(done[b]); if (stage[b]) cudaFreeHost(stage[b]); }
        if (st_) cudaStreamDestroy(st_);
        terr[t] = e;
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& q : th) q.join();
    for (unsigned t = 0; t < T; ++t) if (!ck(s, terr[t], "upload X")) return false;
    for (uint64_t j = 0; j < p; ++j) {
        float mn = std::numeric_limits<float>::infinity(), mx = -mn;
        for (unsigned t = 0; t < T; ++t) { mn = std::fmin(mn, tmin[t][j]); mx = std::fmax(mx, tmax[t][j]); }
        cmin[j] = mn; cmax[j] = mx;
    }
    return true;
}

void free_all(Sampler* s) {
    for (void* p : s->ipc_open) cudaIpcCloseMemHandle(p);
    s->ipc_open.clear();
    for (void* p : s->allocs) cudaFree(p);
    s->allocs.clear();
    for (double* p : {s->tape_slot, s->tape_round, s->tape_upd, s->tr_slot, s->tr_round, s->tr_upd})
        if (p) cudaFree(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->tm0) cudaEventDestroy(s->tm0);
    if (s->tm1) cudaEventDestroy(s->tm1);
    if (s->pin_pred) cudaFreeHost(s->pin_pred);
    if (s->hostdiag) cudaFreeHost(s->hostdiag);
    for (cudaEvent_t e : s->chunk_ev) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
}

bool fetch_ctrl(Sampler* s) {
    if (!ck(s, cudaMemcpyAsync(&s->host_ctrl, s->st.ctrl, sizeof(PgCtrl), cudaMemcpyDeviceToHost, s->stream), "read ctrl")) return false;
    return ck(s, cudaStreamSynchronize(s->stream), "synchronize");
}

int finish_step(Sampler* s) {
    if (!fetch_ctrl(s)) {
        if (s->hostdiag && s->hostdiag[0]) {      // a block recorded what it was waiting for before it trapped
            char b[512];
            snprintf(b, sizeof b, " [watchdog record: code %llx, offset/update %llu, wanted %llu, saw %llu, aux %llu %llu %llu %llx; control heartbeat %llx %llx %llx %llx]",
                     s->hostdiag[0], s->hostdiag[1], s->hostdiag[2], s->hostdiag[3], s->hostdiag[4], s->hostdiag[5], s->hostdiag[6], s->hostdiag[7],
                     s->hostdiag[8], s->hostdiag[9], s->hostdiag[10], s->hostdiag[11]);
            s->err += b;
            std::string w = " [control block warps, (commands issued << 32 | barrier) :";
            for (int q = 16; q < 32; ++q) { snprintf(b, sizeof b, " %llx", s->hostdiag[q]); w += b; }
            s->err += w + "]";
        }
        return 2;
    }
    if (s->host_ctrl.error != PG_OK) {
        s->err = pg_error_text(s->host_ctrl.error);
        // the sampler state is no longer meaningful after a device-side failure
        s->dead = true;
        return 3;
    }
    return 0;
}

int enqueue_step(Sampler* s, int tune) {
    if (s->dead) { if (s->err.empty()) s->err = "sampler is unusable after an earlier failure"; return 2; }
    if (!ck(s, cudaSetDevice(s->device), "cudaSetDevice")) return 2;
    if (!ck(s, pg_launch_prepare(s->st, tune, s->stream), "launch prepare")) return 2;
    s->launches += 1;
    if (!ck(s, cudaEventRecord(s->ev0, s->stream), "event")) return 2;
    if (!s->stepwise) {
        if (!ck(s, pg_launch_persistent(s->st, s->smem, s->stream), "cooperative launch")) return 2;
        s->launches += 1;
    } else {
        // serial(BEGIN) -> [pass -> serial]* until DONE, reading the phase back between launches
        if (!ck(s, pg_launch_serial(s->st, s->stream), "launch serial")) return 2;
        s->launches += 1;
        for (long long guard = 0; guard < (1ll << 40); ++guard) {
            int phase = PH_DONE;
            if (!ck(s, cudaMemcpyAsync(&phase, &s->st.ctrl->phase, sizeof(int), cudaMemcpyDeviceToHost, s->stream), "read phase")) return 2;
            if (!ck(s, cudaStreamSynchronize(s->stream), "synchronize")) return 2;
            if (phase == PH_DONE) break;
            if (!ck(s, pg_launch_phase(s->st, s->smem, s->stream), "launch pass")) return 2;
            if (!ck(s, pg_launch_serial(s->st, s->stream), "launch serial")) return 2;
            s->launches += 2;
        }
    }
    if (!ck(s, cudaEventRecord(s->ev1, s->stream), "event")) return 2;
    s->timing_pending = true;
    return 0;
}

// Predictions (f64[n], HBM) -> caller memory.  Page-locked destinations (pgbart_host_alloc, cudaHostRegister) take one
// asynchronous copy at link speed.  Pageable destinations go through a pinned staging buffer in 2 MB chunks: chunk k is
// copied out of the staging buffer by the host while chunk k + 1 is still on the wire (a plain pageable cudaMemcpy of
// 8 MB cost 1.3 ms per step at BASELINE C3, more than a tenth of the sweep itself).
// The pinned staging buffer and the per-chunk events of the pageable path.  Created on first use — or ahead of time
// (option "prewarm_host"): an observation-sharded sampler must not allocate inside a step, because its peers' kernels
// are already waiting for this rank's by then.
constexpr size_t PRED_CHUNK = (size_t)1 << 18;
bool prepare_staging(Sampler* s) {
    const size_t n = (size_t)s->st.n;
    if (!s->pin_pred && !ck(s, cudaHostAlloc((void**)&s->pin_pred, std::max<size_t>(n, 1) * 8, cudaHostAllocDefault), "cudaHostAlloc")) return false;
    const size_t nch = (n + PRED_CHUNK - 1) / PRED_CHUNK;
    while (s->chunk_ev.size() < nch) {
        cudaEvent_t e;
        if (!ck(s, cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event")) return false;
        s->chunk_ev.push_back(e);
    }
    return true;
}

bool copy_predictions(Sampler* s, double* out) {
    const size_t n = (size_t)s->st.n;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();      // (older drivers report an unregistered pointer as an error: clear it)
    if (pinned) return ck(s, cudaMemcpyAsync(out, s->st.A, n * 8, cudaMemcpyDeviceToHost, s->stream), "copy predictions");
    const size_t CH = PRED_CHUNK;
    const size_t nch = (n + CH - 1) / CH;
    if (!prepare_staging(s)) return false;
    for (size_t c = 0; c < nch; ++c) {
        const size_t off = c * CH, cnt = std::min(CH, n - off);
        if (!ck(s, cudaMemcpyAsync(s->pin_pred + off, s->st.A + off, cnt * 8, cudaMemcpyDeviceToHost, s->stream), "copy predictions") ||
            !ck(s, cudaEventRecord(s->chunk_ev[c], s->stream), "event")) return false;
    }
    for (size_t c = 0; c < nch; ++c) {
        const size_t off = c * CH, cnt = std::min(CH, n - off);
        if (!ck(s, cudaEventSynchronize(s->chunk_ev[c]), "copy predictions")) return false;
        std::memcpy(out + off, s->pin_pred + off, cnt * 8);
    }
    return true;
}

}  // namespace

struct pgbart_sampler { Sampler s; };

extern "C" {

const char* pgbart_last_error(pgbart_handle h) { return h ? h->s.err.c_str() : g_create_error.c_str(); }

int pgbart_create(const void* x, int x_dtype, int x_order, uint64_t n, uint64_t p, const double* y,
                  const pgbart_settings* cfg, int device, pgbart_handle* out) {
    g_create_error.clear();
    if (!out) { g_create_error = "out handle is null"; return 1; }
    *out = nullptr;
    auto fail = [&](const std::string& m) { g_create_error = m; return 1; };
    if (!x || !y || !cfg) return fail("null input");
    const bool allow_rounding = (x_dtype & PGBART_ALLOW_FP32_ROUNDING) != 0;
    x_dtype &= 0xff;
    if (x_dtype != PGBART_X_F64 && x_dtype != PGBART_X_F32) return fail("x_dtype must be PGBART_X_F64 or PGBART_X_F32");
    if (x_order != PGBART_ORDER_C && x_order != PGBART_ORDER_F) return fail("x_order must be PGBART_ORDER_C or PGBART_ORDER_F");
    if (n < 1 || p < 1) return fail("x must have at least one row and one column");
    if (n >= 0xFFFFFFF0ull) return fail("n must fit in 32 bits (sample indices are u32 as in the reference)");
    if (cfg->n_trees < 1) return fail("n_trees must be >= 1");
    if (cfg->n_particles < 2) return fail("n_particles must be >= 2 (one reference stump + at least one growing particle)");
    if (cfg->n_particles > 4096) return fail("n_particles > 4096 is not supported");
    if (cfg->max_depth > 12) return fail("max_depth > 12 is not supported (node tables are heap-indexed)");
    if (cfg->n_split_prior != 0 && cfg->n_split_prior != p) return fail("split_prior must be empty or have one weight per column");
    for (uint64_t i = 0; i < cfg->n_split_rules; ++i) {   // lib.rs:129-132
        const char* r = cfg->split_rules ? cfg->split_rules[i] : nullptr;
        if (!r) return fail("null split rule");
        if (std::strcmp(r, "ContinuousSplit") == 0) continue;
        if (std::strcmp(r, "OneHotSplit") == 0) continue;
        return fail(std::string("Unknown split rule: '") + r + "'. Supported: 'ContinuousSplit', 'OneHotSplit'.");
    }
    // OneHotSplit columns (lib.rs:134-144 pads the rule list with ContinuousSplit): integer codes spanning < 64 values
    std::vector<unsigned char> onehot((size_t)p, 0);
    std::vector<int> col_imin((size_t)p, 0);
    std::vector<unsigned long long> col_bits((size_t)p, 0ull);
    bool any_onehot = false;
    for (uint64_t j = 0; j < cfg->n_split_rules && j < p; ++j) {
        if (std::strcmp(cfg->split_rules[j], "OneHotSplit") != 0) continue;
        onehot[j] = 1; any_onehot = true;
        auto at = [&](uint64_t i) -> double {
            const size_t k = x_order == PGBART_ORDER_F ? (size_t)j * n + i : (size_t)i * p + j;
            return x_dtype == PGBART_X_F64 ? ((const double*)x)[k] : (double)((const float*)x)[k];
        };
        int lo = INT32_MAX, hi = INT32_MIN;
        for (uint64_t i = 0; i < n; ++i) {
            const double v = at(i);
            if (!(std::fabs(v) < 16777216.0)) return fail("OneHotSplit column " + std::to_string(j) + ": values must be finite and below 2^24 in magnitude");
            const int c = (int)v;      // Rust `as i32`: truncation toward zero (splitting.rs:131)
            lo = c < lo ? c : lo; hi = c > hi ? c : hi;
        }
        if ((long long)hi - (long long)lo >= 64)
            return fail("OneHotSplit column " + std::to_string(j) + ": integer codes span " + std::to_string((long long)hi - lo + 1) +
                        " values; at most 64 consecutive codes per column are supported");
        col_imin[j] = lo;
        for (uint64_t i = 0; i < n; ++i) col_bits[j] |= 1ull << ((int)at(i) - lo);
    }
    if (!allow_rounding) {
        char msg[512];
        long long bad = x_dtype == PGBART_X_F64 ? first_not_fp32((const double*)x, (size_t)n * p) : -1;
        if (bad >= 0) {
            const long long i = x_order == PGBART_ORDER_F ? bad % (long long)n : bad / (long long)p, j = x_order == PGBART_ORDER_F ? bad / (long long)n : bad % (long long)p;
            snprintf(msg, sizeof msg, "x[%lld, %lld] = %.17g is not exactly representable in fp32: the device stores X as fp32 and would "
                     "quantise it (split values, ties and partitions could then differ from the reference). Round the data yourself or pass "
                     "PGBART_ALLOW_FP32_ROUNDING to accept round-to-nearest fp32 storage of x and y.", i, j, ((const double*)x)[bad]);
            return fail(msg);
        }
        bad = first_not_fp32(y, (size_t)n);
        if (bad >= 0) {
            snprintf(msg, sizeof msg, "y[%lld] = %.17g is not exactly representable in fp32: the device stores y as fp32. Round the data "
                     "yourself or pass PGBART_ALLOW_FP32_ROUNDING to accept round-to-nearest fp32 storage of x and y.", bad, y[bad]);
            return fail(msg);
        }
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {   // runtime failure, not a bad argument: status 2
        g_create_error = std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e);
        return 2;
    }
    if (device < 0 || device >= ndev) return fail("invalid CUDA device ordinal");

    auto* h = new pgbart_sampler();
    Sampler* s = &h->s;
    auto bail = [&]() { g_create_error = s->err; free_all(s); delete h; return 2; };
    s->device = device;
    s->allow_rounding = allow_rounding;
    if (!ck(s, cudaSetDevice(device), "cudaSetDevice")) return bail();
    if (!ck(s, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "stream")) return bail();
    if (!ck(s, cudaEventCreate(&s->ev0), "event") || !ck(s, cudaEventCreate(&s->ev1), "event") ||
        !ck(s, cudaEventCreate(&s->tm0), "event") || !ck(s, cudaEventCreate(&s->tm1), "event")) return bail();

    PgState& st = s->st;
    st.n = (long long)n;
    st.ld = (long long)((n + 31) / 32 * 32);
    st.p = (int)p; st.m = (int)cfg->n_trees; st.P = (int)cfg->n_particles; st.S = st.P - 1;
    st.max_depth = (int)cfg->max_depth;
    st.maxn = (1 << (st.max_depth + 1)) - 1;
    st.leaf_sigma = cfg->sigma;
    st.m_f = (double)st.m;
    st.batch_tune = cfg->batch_tune; st.batch_post = cfg->batch_post;
    st.family = PG_GAUSSIAN; st.lik_sigma = 1.0;
    st.y_mean = unrolled_mean(y, n);
    st.init_leaf = st.y_mean / (double)st.m;

    char ebuf[256];
    int grid = 0;
    int stage_bytes = 0;
    // P <= 32: 164 KB of shared memory per block — ~145 KB of staging ring for the root pass in the pass blocks, and the L1
    // the control block's SM keeps is still 92 KB (a 225 KB carve-out made every control section slower)
    int smem_kb = st.P <= 32 ? 163 : 0;   // + 1 KB the system reserves per block = the 164 KB carve-out
    if (const char* e = std::getenv("PGBART_SMEM_KB")) smem_kb = std::atoi(e);
    if (pg_query_grid(device, st.S, smem_kb, &grid, &s->smem, &stage_bytes, ebuf, sizeof ebuf)) { s->err = ebuf; return bail(); }
    st.grid = grid;
    st.stage_bytes = stage_bytes;
    st.prefetch_next = 0;
    st.prune = 0;
    if (const char* e = std::getenv("PGBART_PRUNE")) st.prune = std::atoi(e) != 0;
    st.root_ring = 2;
    st.root_tile_groups = 0;
    st.root_stages = 0;
    st.root_dbg = 0;
    st.issue_ahead = 1;
    st.early_stats = 1;
    st.heartbeat = 0;
    if (const char* e = std::getenv("PGBART_HEARTBEAT")) st.heartbeat = std::atoi(e) != 0;
    if (const char* e = std::getenv("PGBART_EARLY_STATS")) st.early_stats = std::atoi(e) != 0;
    st.pg_correct = 0;
    st.generic = (st.P > 32 || any_onehot) ? 1 : 0;
    if (const char* e = std::getenv("PGBART_ISSUE_AHEAD")) st.issue_ahead = std::atoi(e) != 0;
    if (const char* e = std::getenv("PGBART_ROOT_PASS")) st.root_ring = std::strcmp(e, "direct") == 0 ? 0 : std::strcmp(e, "ring") == 0 ? 1 : 2;
    if (const char* e = std::getenv("PGBART_ROOT_TILE_GROUPS")) st.root_tile_groups = std::atoi(e);
    if (const char* e = std::getenv("PGBART_ROOT_STAGES")) st.root_stages = std::atoi(e);
    if (const char* e = std::getenv("PGBART_PREFETCH_NEXT")) st.prefetch_next = std::atoi(e) != 0;
    st.n_gwarps = grid * PG_WARPS;

    // ---- X -> column-major fp32 with per-column min/max; y -> fp32
    {
        std::vector<float> colbuf((size_t)st.ld);
        std::vector<float> cmin(p), cmax(p);
        float* dX = nullptr;
        if (!dalloc(s, &dX, (size_t)st.ld * p, false)) return bail();
        if (!upload_columns(s, x, x_dtype, x_order, n, p, (uint64_t)st.ld, dX, cmin.data(), cmax.data())) return bail();
        st.X = dX;
        float *dmin = nullptr, *dmax = nullptr, *dy = nullptr;
        if (!dalloc(s, &dmin, p, false) || !dalloc(s, &dmax, p, false) || !dalloc(s, &dy, (size_t)st.ld, false)) return bail();
        if (!ck(s, cudaMemcpy(dmin, cmin.data(), p * 4, cudaMemcpyHostToDevice), "upload") ||
            !ck(s, cudaMemcpy(dmax, cmax.data(), p * 4, cudaMemcpyHostToDevice), "upload")) return bail();
        for (uint64_t i = 0; i < n; ++i) colbuf[i] = (float)y[i];
        for (uint64_t i = n; i < (uint64_t)st.ld; ++i) colbuf[i] = 0.0f;
        if (!ck(s, cudaMemcpy(dy, colbuf.data(), (size_t)st.ld * 4, cudaMemcpyHostToDevice), "upload y")) return bail();
        st.col_min = dmin; st.col_max = dmax; st.y = dy;
    }
    // ---- prior and depth thresholds
    if (cfg->n_split_prior) {
        double* d = nullptr;
        if (!dalloc(s, &d, p, false)) return bail();
        if (!ck(s, cudaMemcpy(d, cfg->split_prior, p * 8, cudaMemcpyHostToDevice), "upload prior")) return bail();
        st.prior = d;
        double tot = 0.0;
        for (uint64_t j = 0; j < p; ++j) tot += cfg->split_prior[j];   // smc.rs:243 (sequential sum)
        st.prior_total = tot;
    }
    {
        std::vector<double> thr((size_t)st.max_depth + 2);
        for (size_t d = 0; d < thr.size(); ++d) thr[d] = 1.0 - (cfg->alpha * std::pow(1.0 + (double)d, -cfg->beta));  // smc.rs:143
        double* dt = nullptr;
        if (!dalloc(s, &dt, thr.size(), false)) return bail();
        if (!ck(s, cudaMemcpy(dt, thr.data(), thr.size() * 8, cudaMemcpyHostToDevice), "upload thr")) return bail();
        st.thr_depth = dt;
    }
    {
        double tbl[2 * PG_LOGT_K];
        pg_log_table_fill(tbl);
        double* dt = nullptr;
        if (!dalloc(s, &dt, 2 * PG_LOGT_K, false)) return bail();
        if (!ck(s, cudaMemcpy(dt, tbl, sizeof tbl, cudaMemcpyHostToDevice), "upload log table")) return bail();
        st.log_tbl = dt;
    }
    s->any_onehot = any_onehot;
    if (any_onehot) {
        unsigned char* d_oh = nullptr; int* d_im = nullptr; unsigned long long* d_cb = nullptr;
        if (!dalloc(s, &d_oh, (size_t)p, false) || !dalloc(s, &d_im, (size_t)p, false) || !dalloc(s, &d_cb, (size_t)p, false) ||
            !dalloc(s, &st.job_bits, (size_t)st.S)) return bail();
        if (!ck(s, cudaMemcpy(d_oh, onehot.data(), (size_t)p, cudaMemcpyHostToDevice), "upload split rules") ||
            !ck(s, cudaMemcpy(d_im, col_imin.data(), (size_t)p * 4, cudaMemcpyHostToDevice), "upload code minima") ||
            !ck(s, cudaMemcpy(d_cb, col_bits.data(), (size_t)p * 8, cudaMemcpyHostToDevice), "upload code masks")) return bail();
        st.onehot = d_oh; st.col_imin = d_im; st.col_bits = d_cb;
    }
    // ---- state arrays
    const size_t S = (size_t)st.S, MN = (size_t)st.maxn, M = (size_t)st.m, G = (size_t)st.grid, W = (size_t)st.n_gwarps;
    bool ok = dalloc(s, &st.A, (size_t)st.ld) && dalloc(s, &st.A2, (size_t)st.ld)
        && dalloc(s, &st.f_split_var, M * MN) && dalloc(s, &st.f_thr32, M * MN) && dalloc(s, &st.f_split_val, M * MN)
        && dalloc(s, &st.f_leaf_val, M * MN) && dalloc(s, &st.f_size, M) && dalloc(s, &st.f_internal, M)
        && dalloc(s, &st.pt_split_var, 2 * S * MN) && dalloc(s, &st.pt_thr32, 2 * S * MN) && dalloc(s, &st.pt_split_val, 2 * S * MN)
        && dalloc(s, &st.pt_leaf_val, 2 * S * MN) && dalloc(s, &st.pt_seg_off, 2 * S * MN) && dalloc(s, &st.pt_cnt, 2 * S * MN) && dalloc(s, &st.pt_len, 2 * S * MN)
        && dalloc(s, &st.pt_So, 2 * S * MN) && dalloc(s, &st.pt_Sr, 2 * S * MN) && dalloc(s, &st.pt_g, 2 * S * MN)
        && dalloc(s, &st.pt_queue, 2 * S * MN) && dalloc(s, &st.pt_size, 2 * S) && dalloc(s, &st.pt_qhead, 2 * S)
        && dalloc(s, &st.pt_qtail, 2 * S) && dalloc(s, &st.pt_G, 2 * S)
        && dalloc(s, &st.w, S) && dalloc(s, &st.mutated, S) && dalloc(s, &st.anc, S) && dalloc(s, &st.slot_job, S)
        && dalloc(s, &st.pend_node, S) && dalloc(s, &st.pend_var, S) && dalloc(s, &st.Wfinal, 2 * (S + 1))
        && dalloc(s, &st.job_slot, 2 * S) && dalloc(s, &st.job_node, 2 * S) && dalloc(s, &st.job_var, 2 * S) && dalloc(s, &st.job_thr32, 2 * S)
        && dalloc(s, &st.job_split, 2 * S) && dalloc(s, &st.job_seg_off, 2 * S) && dalloc(s, &st.job_len, 2 * S) && dalloc(s, &st.job_lo, 2 * S)
        && dalloc(s, &st.job_hi, 2 * S) && dalloc(s, &st.job_vL, 2 * S) && dalloc(s, &st.job_vR, 2 * S) && dalloc(s, &st.job_order, 2 * S)
        && dalloc(s, &st.grp_first, 2 * (S + 1)) && dalloc(s, &st.red_So, S) && dalloc(s, &st.red_Sr, S) && dalloc(s, &st.red_llL, S)
        && dalloc(s, &st.red_llR, S) && dalloc(s, &st.red_cnt, S)
        && dalloc(s, &st.pj_anc, S) && dalloc(s, &st.pj_job, S) && dalloc(s, &st.pj_var, S) && dalloc(s, &st.pj_thr32, S)
        && dalloc(s, &st.pj_src_off, S) && dalloc(s, &st.pj_len, S) && dalloc(s, &st.pj_dst_off, S) && dalloc(s, &st.pj_cntL, S)
        && dalloc(s, &st.pj_warp_left, S * W) && dalloc(s, &st.pj_of_anc, S) && dalloc(s, &st.scan_tmp, PG_THREADS + 1)
        && dalloc(s, &st.part_sum, 2 * G * S * 2) && dalloc(s, &st.part_cnt, 2 * W * S) && dalloc(s, &st.part_ccnt, 2 * G * S)
        && dalloc(s, &st.part_tot, 2 * G * 4) && dalloc(s, &st.part_mm, G * 2 * S * 2) && dalloc(s, &st.mq_pj, 2 * S) && dalloc(s, &st.mq_side, 2 * S) && dalloc(s, &st.mq_var, 2 * S) && dalloc(s, &st.mq_u3, 2 * S) && dalloc(s, &st.es_thr, 2 * S) && dalloc(s, &st.part_ll, G * S * 2) && dalloc(s, &st.diag_blocks, G * 16) && dalloc(s, &st.part_old, G) && dalloc(s, &st.blk_done, G + 32)
        && dalloc(s, &st.info_depths, M) && dalloc(s, &st.ctrl, 1);
    if (!ok) return bail();
    {
        // every surviving lineage appends at most n entries per tree level; 2S*n covers the reference's
        // behaviour (one split per tree, Q1), (max_depth+2)*n covers one lineage growing to full depth
        unsigned long long cap = std::max<unsigned long long>(s->pool_factor * (unsigned long long)S, (unsigned long long)st.max_depth + 2) * n;
        // The bound that cannot be exceeded in any mode is S * (max_depth + 1) * n (every lineage distinct, each growing one
        // level per round).  It is 0.7 GB at BASELINE C3 and 9 GB at C5, so take it whenever it costs less than a quarter of
        // the free device memory: PG_ERR_POOL (which ends the chain; the reference has no such failure) is then unreachable.
        {
            const unsigned long long worst = (unsigned long long)S * (unsigned long long)(st.max_depth + 1) * n;
            size_t free_b = 0, total_b = 0;
            if (worst > cap && worst <= 0xFFFFFFF0ull && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess
                && worst * 4ull <= (unsigned long long)free_b / 4ull)
                cap = worst;
        }
        if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;   // u32 offsets
        if (cap < n) cap = n;
        st.pool_cap = cap;
        if (!dalloc(s, &st.pool, (size_t)cap, false)) return bail();
    }
    st.rank = 0; st.world = 1; st.n_global = st.n;
    for (int r = 0; r < PG_MAX_RANKS; ++r) st.mail[r] = nullptr;
    st.rng_mode = PG_RNG_PHILOX;
    st.philox_key0 = (unsigned)cfg->seed;
    st.philox_key1 = (unsigned)(cfg->seed >> 32);

    PgCtrl c0{};
    c0.phase = PH_DONE; c0.tune = 1; c0.t_add = -1; c0.t_sub = -1;
    if (cudaHostAlloc((void**)&s->hostdiag, 256, cudaHostAllocMapped) == cudaSuccess) {
        std::memset(s->hostdiag, 0, 256);
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, s->hostdiag, 0) == cudaSuccess) c0.hostdiag = (unsigned long long)dp;
    } else { s->hostdiag = nullptr; cudaGetLastError(); }
    if (!ck(s, cudaMemcpyAsync(st.ctrl, &c0, sizeof c0, cudaMemcpyHostToDevice, s->stream), "init ctrl")) return bail();
    if (!ck(s, pg_launch_fill(st.A, st.n, st.y_mean, s->stream), "fill predictions")) return bail();   // kernel.rs:48
    if (!ck(s, pg_launch_init_forest(st, s->stream), "init forest")) return bail();
    s->launches += 2;
    if (!ck(s, cudaStreamSynchronize(s->stream), "synchronize")) return bail();
    *out = h;
    return 0;
}

int pgbart_destroy(pgbart_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->s.device);
    cudaStreamSynchronize(h->s.stream);
    free_all(&h->s);
    delete h;
    return 0;
}

int pgbart_set_likelihood(pgbart_handle h, int family, const double* params, int nparams) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (family != PGBART_GAUSSIAN && family != PGBART_BERNOULLI_LOGIT && family != PGBART_GAUSSIAN_KERNEL) {
        s->err = "unknown likelihood family (device path supports Gaussian and Bernoulli-logit; no CPU fallback)";
        return 1;
    }
    if (family == PGBART_GAUSSIAN) {
        if (nparams < 1 || !params || !(params[0] > 0.0)) { s->err = "Gaussian likelihood needs sigma > 0"; return 1; }
        s->st.lik_sigma = params[0];
    }
    s->st.family = family;
    return 0;
}

int pgbart_step_device(pgbart_handle h, int tune) { if (!h) return 1; h->s.steps_done += 1; return enqueue_step(&h->s, tune); }

int pgbart_sync(pgbart_handle h) { return h ? finish_step(&h->s) : 1; }

int pgbart_step(pgbart_handle h, int tune, double* out_pred) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    s->steps_done += 1;
    int rc = enqueue_step(s, tune);
    if (rc) return rc;
    if (out_pred && !copy_predictions(s, out_pred)) return 2;
    return finish_step(s);
}

const double* pgbart_predictions_device(pgbart_handle h) { return h ? h->s.st.A : nullptr; }

int pgbart_get_predictions(pgbart_handle h, double* out_pred) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!out_pred) { s->err = "null output"; return 1; }
    if (!copy_predictions(s, out_pred)) return 2;
    return ck(s, cudaStreamSynchronize(s->stream), "synchronize") ? 0 : 2;
}

int pgbart_get_info(pgbart_handle h, double* ll, int64_t* acc, uint8_t* depths, int* n_depths) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!fetch_ctrl(s)) return 2;
    *ll = s->host_ctrl.info_loglik;
    *acc = s->host_ctrl.info_acc;
    const int cap = *n_depths;
    *n_depths = s->host_ctrl.info_n;
    const int k = std::min(cap, s->host_ctrl.info_n);
    if (k > 0 && !ck(s, cudaMemcpy(depths, s->st.info_depths, (size_t)k, cudaMemcpyDeviceToHost), "copy depths")) return 2;
    return 0;
}

int pgbart_tree_size(pgbart_handle h, uint64_t t, int* size) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (t >= (uint64_t)s->st.m) { s->err = "tree index out of range"; return 1; }
    if (!ck(s, cudaStreamSynchronize(s->stream), "synchronize")) return 2;
    return ck(s, cudaMemcpy(size, s->st.f_size + t, 4, cudaMemcpyDeviceToHost), "copy") ? 0 : 2;
}

int pgbart_get_tree(pgbart_handle h, uint64_t t, int32_t* split_var, double* split_val, double* leaf_val) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    int size = 0;
    int rc = pgbart_tree_size(h, t, &size);
    if (rc) return rc;
    const size_t fb = (size_t)t * s->st.maxn;
    if (!ck(s, cudaMemcpy(split_var, s->st.f_split_var + fb, (size_t)size * 4, cudaMemcpyDeviceToHost), "copy") ||
        !ck(s, cudaMemcpy(split_val, s->st.f_split_val + fb, (size_t)size * 8, cudaMemcpyDeviceToHost), "copy") ||
        !ck(s, cudaMemcpy(leaf_val, s->st.f_leaf_val + fb, (size_t)size * 8, cudaMemcpyDeviceToHost), "copy")) return 2;
    return 0;
}

int pgbart_get_leaf_indices(pgbart_handle h, uint64_t t, uint32_t* out) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (t >= (uint64_t)s->st.m) { s->err = "tree index out of range"; return 1; }
    unsigned* d = nullptr;
    if (!ck(s, cudaMalloc(&d, (size_t)s->st.n * 4), "cudaMalloc")) return 2;
    bool ok = ck(s, pg_launch_leaf_indices(s->st, (int)t, d, s->stream), "launch")
           && ck(s, cudaMemcpyAsync(out, d, (size_t)s->st.n * 4, cudaMemcpyDeviceToHost, s->stream), "copy")
           && ck(s, cudaStreamSynchronize(s->stream), "synchronize");
    s->launches += 1;
    cudaFree(d);
    return ok ? 0 : 2;
}

int pgbart_get_state(pgbart_handle h, uint64_t* next_tree_idx, int* tune, uint64_t* n_updates) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!fetch_ctrl(s)) return 2;
    if (next_tree_idx) *next_tree_idx = (uint64_t)s->host_ctrl.next_tree_idx;
    if (tune) *tune = s->host_ctrl.tune;
    if (n_updates) *n_updates = s->host_ctrl.n_updates;
    return 0;
}

int pgbart_set_rng_philox(pgbart_handle h, uint64_t seed) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    s->st.rng_mode = PG_RNG_PHILOX;
    s->st.philox_key0 = (unsigned)seed;
    s->st.philox_key1 = (unsigned)(seed >> 32);
    return 0;
}

int pgbart_set_rng_tape(pgbart_handle h, const double* slot, const double* round, const double* upd,
                        int64_t n_upd, int r_cap) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (n_upd < 1 || r_cap < 1 || !slot || !round || !upd) { s->err = "bad tape"; return 1; }
    if (!fetch_ctrl(s)) return 2;
    for (double** p : {&s->tape_slot, &s->tape_round, &s->tape_upd}) if (*p) { cudaFree(*p); *p = nullptr; }
    const size_t ns = (size_t)n_upd * r_cap * s->st.S * 5, nr = (size_t)n_upd * r_cap, nu = (size_t)n_upd;
    if (!ck(s, cudaMalloc(&s->tape_slot, ns * 8), "cudaMalloc") || !ck(s, cudaMalloc(&s->tape_round, nr * 8), "cudaMalloc") ||
        !ck(s, cudaMalloc(&s->tape_upd, nu * 8), "cudaMalloc")) return 2;
    if (!ck(s, cudaMemcpy(s->tape_slot, slot, ns * 8, cudaMemcpyHostToDevice), "upload tape") ||
        !ck(s, cudaMemcpy(s->tape_round, round, nr * 8, cudaMemcpyHostToDevice), "upload tape") ||
        !ck(s, cudaMemcpy(s->tape_upd, upd, nu * 8, cudaMemcpyHostToDevice), "upload tape")) return 2;
    s->st.rng_mode = PG_RNG_TAPE;
    s->st.tape_slot = s->tape_slot; s->st.tape_round = s->tape_round; s->st.tape_upd = s->tape_upd;
    s->st.tape_n_upd = n_upd; s->st.tape_r_cap = r_cap;
    s->st.tape_base_upd = s->host_ctrl.upd;
    return 0;
}

int pgbart_trace_enable(pgbart_handle h, int64_t u_cap, int r_cap) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    for (double** p : {&s->tr_slot, &s->tr_round, &s->tr_upd}) if (*p) { cudaFree(*p); *p = nullptr; }
    s->st.tr_slot = s->st.tr_round = s->st.tr_upd = nullptr;
    if (u_cap <= 0) return 0;
    if (r_cap < 1) { s->err = "trace_enable: r_cap must be >= 1 when u_cap > 0"; return 1; }
    const size_t S = s->st.S, P = s->st.P;
    const size_t ns = (size_t)u_cap * r_cap * S * PG_TS_FIELDS, nr = (size_t)u_cap * r_cap * (2 + 2 * S), nu = (size_t)u_cap * (4 + 2 * P);
    if (!ck(s, cudaMalloc(&s->tr_slot, ns * 8), "cudaMalloc") || !ck(s, cudaMalloc(&s->tr_round, nr * 8), "cudaMalloc") ||
        !ck(s, cudaMalloc(&s->tr_upd, nu * 8), "cudaMalloc")) return 2;
    // NaN-fill the slot records (all-ones bytes are a NaN) and zero the rest, like the oracle's buffers
    if (!ck(s, cudaMemset(s->tr_slot, 0xFF, ns * 8), "memset") || !ck(s, cudaMemset(s->tr_round, 0, nr * 8), "memset") ||
        !ck(s, cudaMemset(s->tr_upd, 0, nu * 8), "memset")) return 2;
    s->st.tr_slot = s->tr_slot; s->st.tr_round = s->tr_round; s->st.tr_upd = s->tr_upd;
    s->st.tr_u_cap = u_cap; s->st.tr_r_cap = r_cap;
    s->tr_u_cap = u_cap; s->tr_r_cap = r_cap;
    // restart the trace counter
    if (!fetch_ctrl(s)) return 2;
    s->host_ctrl.trace_n = 0; s->host_ctrl.trace_overflow = 0;
    if (!ck(s, cudaMemcpy(&s->st.ctrl->trace_n, &s->host_ctrl.trace_n, 8, cudaMemcpyHostToDevice), "reset") ||
        !ck(s, cudaMemcpy(&s->st.ctrl->trace_overflow, &s->host_ctrl.trace_overflow, 4, cudaMemcpyHostToDevice), "reset")) return 2;
    return 0;
}

int pgbart_trace_read(pgbart_handle h, double* slot, double* round, double* upd, int64_t* n_upd, int* overflow) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!s->tr_slot) { s->err = "trace not enabled"; return 1; }
    if (!fetch_ctrl(s)) return 2;
    const size_t S = s->st.S, P = s->st.P;
    const size_t ns = (size_t)s->tr_u_cap * s->tr_r_cap * S * PG_TS_FIELDS, nr = (size_t)s->tr_u_cap * s->tr_r_cap * (2 + 2 * S),
                 nu = (size_t)s->tr_u_cap * (4 + 2 * P);
    if (!ck(s, cudaMemcpy(slot, s->tr_slot, ns * 8, cudaMemcpyDeviceToHost), "copy trace") ||
        !ck(s, cudaMemcpy(round, s->tr_round, nr * 8, cudaMemcpyDeviceToHost), "copy trace") ||
        !ck(s, cudaMemcpy(upd, s->tr_upd, nu * 8, cudaMemcpyDeviceToHost), "copy trace")) return 2;
    *n_upd = (int64_t)s->host_ctrl.trace_n;
    *overflow = s->host_ctrl.trace_overflow;
    return 0;
}

int pgbart_set_option(pgbart_handle h, const char* key, const char* value) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!key || !value) { s->err = "null option"; return 1; }
    if (std::strcmp(key, "driver") == 0) {
        if (std::strcmp(value, "persistent") == 0) { s->stepwise = false; return 0; }
        if (std::strcmp(value, "stepwise") == 0 || std::strcmp(value, "stepwise_fast") == 0) { s->stepwise = true; return 0; }
        s->err = "driver must be 'persistent' or 'stepwise'";
        return 1;
    }
    if (std::strcmp(key, "prewarm_host") == 0) { return prepare_staging(s) ? 0 : 2; }
    if (std::strcmp(key, "weight_mode") == 0) {
        // "reference" = the reference's stale-weight rule and fresh-stump reference particle (Q1, Q2; default);
        // "pg_correct" = log-weights carried with the particles and the tree being replaced as the conditional reference particle
        if (s->steps_done != 0) { s->err = "weight_mode must be set before the first step"; return 1; }
        if (s->st.world > 1) { s->err = "weight_mode pg_correct is not available with observation sharding"; return 1; }
        if (std::strcmp(value, "reference") == 0) { s->st.pg_correct = 0; s->st.generic = (s->st.P > 32 || s->any_onehot) ? 1 : 0; return 0; }
        if (std::strcmp(value, "pg_correct") == 0) {
            if (s->st.S > 512) { s->err = "weight_mode pg_correct supports at most 513 particles"; return 1; }
            // In this mode distinct lineages do survive and grow level by level: every surviving lineage that grew appends the
            // rows of its node, so a tree update can append up to S * (max_depth + 1) * n entries to the segment pool (the
            // reference mode's default, 2 S n, covers one split per tree).  Size the pool for that now.
            unsigned long long cap = (unsigned long long)s->st.S * (unsigned long long)(s->st.max_depth + 1) * (unsigned long long)s->st.n;
            if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
            if (cap > s->st.pool_cap) {
                unsigned* np = nullptr;
                if (!ck(s, cudaStreamSynchronize(s->stream), "synchronize")) return 2;
                if (!ck(s, cudaMalloc(&np, (size_t)cap * 4), "cudaMalloc pool (weight_mode pg_correct)")) return 2;
                auto it = std::find(s->allocs.begin(), s->allocs.end(), (void*)s->st.pool);
                if (it != s->allocs.end()) { cudaFree(*it); *it = np; } else s->allocs.push_back(np);
                s->st.pool = np; s->st.pool_cap = cap;
            }
            s->st.pg_correct = 1; s->st.generic = 1; return 0;
        }
        s->err = "weight_mode must be 'reference' or 'pg_correct'";
        return 1;
    }
    if (std::strcmp(key, "root_pass") == 0) {
        if (std::strcmp(value, "ring") == 0 || std::strcmp(value, "staged") == 0) {
            if (s->st.stage_bytes <= 0) { s->err = "the ring root pass needs a staging area (PGBART_SMEM_KB, n_particles <= 32)"; return 1; }
            s->st.root_ring = 1; return 0;
        }
        if (std::strcmp(value, "direct") == 0) { s->st.root_ring = 0; return 0; }
        if (std::strcmp(value, "auto") == 0) { s->st.root_ring = 2; return 0; }
        s->err = "root_pass must be 'ring', 'direct' or 'auto'";
        return 1;
    }
    if (std::strcmp(key, "root_tile_groups") == 0) { s->st.root_tile_groups = std::atoi(value); return 0; }
    if (std::strcmp(key, "root_stages") == 0) { s->st.root_stages = std::atoi(value); return 0; }
    if (std::strcmp(key, "root_dbg") == 0) { s->st.root_dbg = std::atoi(value); return 0; }
    if (std::strcmp(key, "heartbeat") == 0) { s->st.heartbeat = std::atoi(value) != 0; return 0; }
    if (std::strcmp(key, "early_stats") == 0) { s->st.early_stats = std::atoi(value) != 0; return 0; }
    if (std::strcmp(key, "issue_ahead") == 0) { s->st.issue_ahead = std::atoi(value) != 0; return 0; }
    if (std::strcmp(key, "prune_dead_proposals") == 0) { s->st.prune = std::atoi(value) != 0; return 0; }
    if (std::strcmp(key, "prefetch_next") == 0) { s->st.prefetch_next = std::atoi(value) != 0; return 0; }
    if (std::strcmp(key, "pool_factor") == 0) {
        const long long f = std::atoll(value);
        if (f < 1) { s->err = "pool_factor must be >= 1"; return 1; }
        unsigned long long cap = (unsigned long long)f * (unsigned long long)s->st.S * (unsigned long long)s->st.n;
        if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
        unsigned* np = nullptr;
        if (!ck(s, cudaStreamSynchronize(s->stream), "synchronize")) return 2;
        if (!ck(s, cudaMalloc(&np, (size_t)cap * 4), "cudaMalloc pool")) return 2;
        auto it = std::find(s->allocs.begin(), s->allocs.end(), (void*)s->st.pool);
        if (it != s->allocs.end()) { cudaFree(*it); *it = np; } else s->allocs.push_back(np);
        s->st.pool = np; s->st.pool_cap = cap; s->pool_factor = (unsigned long long)f;
        return 0;
    }
    s->err = std::string("unknown option '") + key + "'";
    return 1;
}

int pgbart_get_counters(pgbart_handle h, double* out8) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!fetch_ctrl(s)) return 2;
    float ms = 0.0f;
    if (s->timing_pending) cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    out8[0] = (double)s->host_ctrl.bytes_model;
    out8[1] = (double)s->host_ctrl.bytes_contract;
    out8[2] = (double)s->host_ctrl.n_phases;
    out8[3] = (double)s->host_ctrl.n_updates;
    out8[4] = s->launches;
    out8[5] = (double)ms * 1000.0;
    out8[6] = (double)s->st.grid;
    out8[7] = (double)s->smem;
    return 0;
}

int pgbart_get_spec_counters(pgbart_handle h, uint64_t* out4) {
    if (!h || !out4) return 1;
    Sampler* s = &h->s;
    if (!fetch_ctrl(s)) return 2;
    out4[0] = s->host_ctrl.n_spec; out4[1] = s->host_ctrl.n_pred; out4[2] = s->host_ctrl.n_pred_miss; out4[3] = s->host_ctrl.n_drain;
    return 0;
}

int pgbart_get_early_stats_counters(pgbart_handle h, uint64_t* out2) {
    if (!h || !out2) return 1;
    Sampler* s = &h->s;
    if (!fetch_ctrl(s)) return 2;
    out2[0] = s->host_ctrl.n_es; out2[1] = s->host_ctrl.n_es_miss;
    return 0;
}

int pgbart_get_profile(pgbart_handle h, double* out32) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!fetch_ctrl(s)) return 2;
    for (int i = 0; i < 32; ++i) out32[i] = (double)s->host_ctrl.prof[i];
    if (!s->st.generic && !s->stepwise) for (int i = 0; i < 8; ++i) out32[16 + i] = (double)s->host_ctrl.prof_ser[i];
    for (int i = 0; i < 16; ++i) out32[32 + i] = (double)s->host_ctrl.prof2[i];
    out32[48] = (double)s->host_ctrl.n_spec;
    for (int i = 0; i < 16; ++i) out32[49 + i] = (double)s->host_ctrl.prof3[i];
    for (int i = 0; i < 12; ++i) out32[57 + i] = (double)s->host_ctrl.dbg[4 + i];   // prof3[8..15] unused: diagnostics ride there
    return 0;
}

int pgbart_debug_state(pgbart_handle h, double* out8) {
    if (!h || !out8) return 1;
    Sampler* s = &h->s;
    // no error check on purpose: this is what one reads AFTER a step failed (the copy itself fails if the context is lost)
    if (cudaMemcpy(&s->host_ctrl, s->st.ctrl, sizeof(PgCtrl), cudaMemcpyDeviceToHost) != cudaSuccess) { s->err = "device state is not readable (context lost)"; return 2; }
    const PgCtrl& c = s->host_ctrl;
    out8[0] = c.phase; out8[1] = c.k; out8[2] = c.round; out8[3] = (double)c.upd; out8[4] = (double)c.xseq;
    out8[5] = (double)c.n_phases; out8[6] = c.error; out8[7] = c.t_cur;
    return 0;
}

int pgbart_debug_block_cycles(pgbart_handle h, double* out, int cap) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    const int G = s->st.grid;
    std::vector<unsigned long long> tmp((size_t)G * 16);
    if (!ck(s, cudaStreamSynchronize(s->stream), "sync")) return 2;
    if (!ck(s, cudaMemcpy(tmp.data(), s->st.diag_blocks, tmp.size() * 8, cudaMemcpyDeviceToHost), "copy diag")) return 2;
    for (int i = 0; i < G * 16 && i < cap; ++i) out[i] = (double)tmp[i];
    return 0;
}

// ---- observation sharding (include/pgbart.h) ---------------------------------------------------------
int pgbart_shard_config(pgbart_handle h, int rank, int world, uint64_t n_global, double y_mean_global,
                        const float* col_min_global, const float* col_max_global) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    PgState& st = s->st;
    if (world < 1 || world > PG_MAX_RANKS || rank < 0 || rank >= world) { s->err = "shard_config: need 0 <= rank < world <= 8"; return 1; }
    if (st.generic) { s->err = "observation sharding is implemented for n_particles <= 32 in the reference weight mode with ContinuousSplit columns"; return 1; }
    if (s->stepwise) { s->err = "observation sharding needs the persistent driver"; return 1; }
    if (s->steps_done != 0) { s->err = "shard_config must precede the first step"; return 1; }
    cudaSetDevice(s->device);
    st.rank = rank; st.world = world; st.n_global = (long long)n_global;
    st.y_mean = y_mean_global;
    st.init_leaf = st.y_mean / (double)st.m;
    if (!ck(s, cudaMemcpyAsync((void*)st.col_min, col_min_global, (size_t)st.p * 4, cudaMemcpyHostToDevice, s->stream), "upload col_min") ||
        !ck(s, cudaMemcpyAsync((void*)st.col_max, col_max_global, (size_t)st.p * 4, cudaMemcpyHostToDevice, s->stream), "upload col_max")) return 2;
    if (!ck(s, pg_launch_fill(st.A, st.n, st.y_mean, s->stream), "fill predictions")) return 2;
    if (!ck(s, pg_launch_init_forest(st, s->stream), "init forest")) return 2;
    if (!st.mail[rank]) {
        double* mb = nullptr;
        if (!dalloc(s, &mb, PG_MAIL_BYTES / 8)) return 2;
        st.mail[rank] = mb;
    }
    return ck(s, cudaStreamSynchronize(s->stream), "synchronize") ? 0 : 2;
}

int pgbart_shard_mailbox(pgbart_handle h, unsigned char* handle64) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!s->st.mail[s->st.rank]) { s->err = "shard_mailbox: call shard_config first"; return 1; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd;
    cudaSetDevice(s->device);
    if (!ck(s, cudaIpcGetMemHandle(&hd, s->st.mail[s->st.rank]), "cudaIpcGetMemHandle")) return 2;
    std::memcpy(handle64, &hd, 64);
    return 0;
}

int pgbart_shard_connect(pgbart_handle h, const unsigned char* handles) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    PgState& st = s->st;
    cudaSetDevice(s->device);
    for (int r = 0; r < st.world; ++r) {
        if (r == st.rank) continue;
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        if (!ck(s, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return 2;
        st.mail[r] = (double*)p;
        s->ipc_open.push_back(p);
    }
    return 0;
}

// Out-of-sample prediction (tree.rs:169-192 summed over the forest) and variable inclusion (state.rs:10, declared by
// the reference and never filled): the "next" rows of SURVEY.md 8(f).
int pgbart_predict(pgbart_handle h, const void* x, int x_dtype, int x_order, uint64_t n_new, double* out) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    const PgState& st = s->st;
    if (!x || !out) { s->err = "predict: null argument"; return 1; }
    const bool allow = s->allow_rounding || (x_dtype & PGBART_ALLOW_FP32_ROUNDING) != 0;
    x_dtype &= 0xff;
    if (x_dtype != PGBART_X_F32 && x_dtype != PGBART_X_F64) { s->err = "predict: x_dtype must be f32 or f64"; return 1; }
    if (n_new == 0) return 0;
    const uint64_t p = (uint64_t)st.p;
    if (x_dtype == PGBART_X_F64 && !allow) {
        const long long bad = first_not_fp32((const double*)x, (size_t)n_new * p);
        if (bad >= 0) {
            char msg[256];
            snprintf(msg, sizeof msg, "predict: x element %lld = %.17g is not exactly representable in fp32 (the forest was trained on fp32 "
                     "data); round it or pass PGBART_ALLOW_FP32_ROUNDING", bad, ((const double*)x)[bad]);
            s->err = msg;
            return 1;
        }
    }
    cudaSetDevice(s->device);
    std::vector<float> col((size_t)n_new * p);
    for (uint64_t j = 0; j < p; ++j)
        for (uint64_t i = 0; i < n_new; ++i) {
            const size_t src = x_order == PGBART_ORDER_F ? (size_t)j * n_new + i : (size_t)i * p + j;
            col[(size_t)j * n_new + i] = x_dtype == PGBART_X_F32 ? ((const float*)x)[src] : (float)((const double*)x)[src];
        }
    float* dx = nullptr;
    double* dout = nullptr;
    if (!ck(s, cudaMalloc(&dx, col.size() * 4), "cudaMalloc") || !ck(s, cudaMalloc(&dout, (size_t)n_new * 8), "cudaMalloc")) { if (dx) cudaFree(dx); return 2; }
    bool ok = ck(s, cudaMemcpyAsync(dx, col.data(), col.size() * 4, cudaMemcpyHostToDevice, s->stream), "upload x")
           && ck(s, pg_launch_predict(st, dx, (long long)n_new, (long long)n_new, dout, s->stream), "launch predict")
           && ck(s, cudaMemcpyAsync(out, dout, (size_t)n_new * 8, cudaMemcpyDeviceToHost, s->stream), "copy predictions")
           && ck(s, cudaStreamSynchronize(s->stream), "synchronize");
    s->launches += 1;
    cudaFree(dx); cudaFree(dout);
    return ok ? 0 : 2;
}

int pgbart_variable_inclusion(pgbart_handle h, uint32_t* out_p) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    const PgState& st = s->st;
    cudaSetDevice(s->device);
    std::vector<int> sv((size_t)st.m * st.maxn), size((size_t)st.m);
    if (!ck(s, cudaStreamSynchronize(s->stream), "synchronize") ||
        !ck(s, cudaMemcpy(sv.data(), st.f_split_var, sv.size() * 4, cudaMemcpyDeviceToHost), "copy forest") ||
        !ck(s, cudaMemcpy(size.data(), st.f_size, size.size() * 4, cudaMemcpyDeviceToHost), "copy forest")) return 2;
    for (int j = 0; j < st.p; ++j) out_p[j] = 0u;
    for (int t = 0; t < st.m; ++t)
        for (int q = 0; q < size[(size_t)t]; ++q) {
            const int v = sv[(size_t)t * st.maxn + q];
            if (v >= 0 && v < st.p) out_p[v] += 1u;
        }
    return 0;
}

// Page-locked host memory for callers that want pgbart_step's device->host copy to run at link speed straight into
// their own buffer (the Python mirror hands such buffers out as the arrays PySampler.step returns).
int pgbart_host_alloc(uint64_t bytes, void** out) {
    if (!out) return 1;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, std::max<uint64_t>(bytes, 1), cudaHostAllocPortable);
    if (e != cudaSuccess) { g_create_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); cudaGetLastError(); return 2; }
    return 0;
}

int pgbart_host_free(void* p) {
    if (!p) return 0;
    return cudaFreeHost(p) == cudaSuccess ? 0 : 2;
}

int pgbart_timer_start(pgbart_handle h) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    return ck(s, cudaEventRecord(s->tm0, s->stream), "event record") ? 0 : 2;
}

int pgbart_timer_stop(pgbart_handle h, double* elapsed_ms) {
    if (!h) return 1;   // a null handle is a bad argument, not a crash
    Sampler* s = &h->s;
    if (!ck(s, cudaEventRecord(s->tm1, s->stream), "event record")) return 2;
    if (!ck(s, cudaEventSynchronize(s->tm1), "event synchronize")) return 2;
    float ms = 0.0f;
    if (!ck(s, cudaEventElapsedTime(&ms, s->tm0, s->tm1), "event elapsed")) return 2;
    *elapsed_ms = (double)ms;
    return 0;
}

}  // extern "C"
