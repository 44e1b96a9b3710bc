// pgbart_oracle.cpp — CPU restatement of the bart-rs Particle-Gibbs BART step.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under bart_rs_b200/ may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it (as the checker / the timed CPU
// baseline).  The product path is the CUDA library in bart_rs_b200/csrc.
//
// What it restates (paths relative to /root/reference):
//   src/kernel.rs:38-126   BartKernel::init / step  (batch loop over trees)
//   src/smc.rs:26-268      smc_step, propose_mutation, propose_leaf_values,
//                          sample_feature_from_probs, normalize_weights_inplace
//   src/particle.rs:13-148 Particle, LeafSamplesFlat, apply_mutation
//   src/tree.rs:9-198      TreeArrays, split_node, get_depth, predict_training
//   src/splitting.rs:39-56 ContinuousSplit::sample_split_value
//   src/resampling.rs:24-44 SystematicResampling::resample_into
//   src/weight.rs:5-30     WeightFn (C callback slot kept; built-in Gaussian /
//                          Bernoulli-logit / bench mock added)
// It deliberately keeps the reference's cost structure (f64 data, full-n
// predict + log-likelihood per mutated particle, per-round particle clones of
// the 4n-byte index array, per-tree allocation, unstable two-pointer
// partition, the stale-weight rule of smc.rs:53,89-96) because it is also the
// timed single-core CPU baseline.
//
// PARITY STATUS: "parity unpinned" at the RNG boundary.  The reference's
// randomness comes from rand 0.9.1 / rand_distr 0.5.1 / rand_core 0.9.3
// (Cargo.lock), whose sources are not on this machine, and no reference test
// seeds an RNG and checks an output.  Deterministic mechanics ARE pinned: the
// 12 Rust unit tests (src/tree.rs:205-275, src/particle.rs:157-237) are ported
// as known-answer tests in tests/test_oracle_kat.py through the orc_kat_*
// entry points below.  Randomness is exchanged with the device through a
// typed RNG tape or a coordinate-addressed Philox4x32-10
// stream; both are restated here independently of the product code.
//
// Restatement notes (recollection of third-party crates, unverifiable here):
//   * SmallRng = xoshiro256++ seeded through SplitMix64; random::<f64>() =
//     (next_u64 >> 11) * 2^-53; random_range(lo..hi) for f64 =
//     ((next_u64 >> 12 | exponent 0) - 1.0) * (hi-lo) + lo, redrawn while the
//     result is >= hi (no FMA: Rust never contracts).
//   * StandardNormal in rand_distr is a 256-layer ziggurat whose tables are
//     not reproducible here; this file draws normals by Box-Muller from two
//     uniforms.  The typed tape stores the normal itself, so the device is
//     unaffected.
//   * random_range(0..p) for usize (only reachable from Rust callers, Q3) is
//     restated as floor(u * p).
//   * ndarray's mean() = sum()/n with an 8-way unrolled fold (numeric_util).
//
// Build: see oracle/Makefile (g++ -O3 -march=x86-64-v3 -ffp-contract=off).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------
// Random sources
// ---------------------------------------------------------------------------

enum DrawKind : int {
    D_GROW = 0,    // u1: depth-prior test           smc.rs:145
    D_VAR = 1,     // u2: split variable             smc.rs:154-158
    D_SPLIT = 2,   // u3: split value in [lo,hi)     splitting.rs:55
    D_ZLEFT = 3,   // zL: left leaf noise            smc.rs:221
    D_ZRIGHT = 4,  // zR: right leaf noise           smc.rs:230
    D_RESAMPLE = 5,  // u6: systematic offset        resampling.rs:27
    D_SELECT = 6     // u7: final categorical draw   smc.rs:113-114
};

struct DrawCtx {
    uint64_t upd = 0;  // tree-update counter since sampler creation
    int round = 0;     // SMC round inside this tree update (0-based)
    int slot = 0;      // growing-particle slot (0..P-2)
};

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

struct Xoshiro256pp {
    uint64_t s[4];
    explicit Xoshiro256pp(uint64_t seed = 0) { seed_from_u64(seed); }
    void seed_from_u64(uint64_t state) {
        for (int i = 0; i < 4; ++i) {
            state += 0x9e3779b97f4a7c15ULL;
            uint64_t z = state;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            s[i] = z ^ (z >> 31);
        }
    }
    uint64_t next() {
        const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl64(s[3], 45);
        return result;
    }
};

// Philox4x32-10 (Salmon et al., SC'11), restated from the published algorithm.
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static const double TWO_PI = 6.283185307179586476925286766559;

struct TapeView {
    double* slot = nullptr;   // [U][R][S][5]
    double* round = nullptr;  // [U][R]
    double* upd = nullptr;    // [U]
    int64_t n_upd = 0;
    int r_cap = 0;
    int n_slots = 0;
    uint64_t base_upd = 0;
    bool overflow = false;
    double* at(DrawKind k, const DrawCtx& c) {
        if (c.upd < base_upd) { overflow = true; return nullptr; }
        const uint64_t u = c.upd - base_upd;
        if ((int64_t)u >= n_upd || c.round >= r_cap) { overflow = true; return nullptr; }
        if (k <= D_ZRIGHT) return slot + (((u * r_cap + c.round) * n_slots + c.slot) * 5 + (int)k);
        if (k == D_RESAMPLE) return round + (u * r_cap + c.round);
        return upd + u;
    }
};

// One object, four modes: sequential xoshiro (optionally recording a typed
// tape), coordinate-addressed Philox, replay of a supplied tape, or a FLAT
// sequential log: the words a reference run consumed, in consumption order —
// raw generator outputs (u64, stored in the bit pattern of a double) for every
// uniform / range draw, and the VALUE of every standard normal (the ziggurat of
// rand_distr is not restated here; the reference-side dumper of INTEGRATION.md
// section 1b logs normals at their call sites and mutes the raw words they
// consume).  FLAT converts raw words exactly like SEQ and still records the typed
// tape, so a log taken from the real Rust sampler drives oracle AND device.
struct Draws {
    enum Mode { SEQ = 0, PHILOX = 1, REPLAY = 2, FLAT = 3 } mode = SEQ;
    Xoshiro256pp seq;
    const double* flat = nullptr;   // FLAT: the log being consumed
    size_t flat_n = 0, flat_pos = 0;
    double* flat_out = nullptr;     // SEQ: optional flat log being written (tests the FLAT path without a Rust toolchain)
    size_t flat_cap = 0, flat_w = 0;

    uint64_t raw_word() {           // next raw generator output: the sequential generator, or the log
        if (mode == FLAT) {
            if (flat_pos >= flat_n) { error = true; return 0; }
            uint64_t b; std::memcpy(&b, &flat[flat_pos++], 8);
            return b;
        }
        const uint64_t b = seq.next();
        if (flat_out) { if (flat_w < flat_cap) std::memcpy(&flat_out[flat_w], &b, 8); flat_w++; }
        return b;
    }
    uint32_t key[2] = {0, 0};
    TapeView tape;  // SEQ: written if non-null; REPLAY: read
    bool error = false;

    void set_seed(uint64_t seed) {
        seq.seed_from_u64(seed);
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
    }

    static double bits53(uint64_t b) { return (double)(b >> 11) * (1.0 / 9007199254740992.0); }

    void philox_pair(DrawKind k, const DrawCtx& c, int attempt, double& ua, double& ub) const {
        uint32_t ctr[4] = {(uint32_t)c.upd, (uint32_t)(c.upd >> 32), (uint32_t)c.round,
                           ((uint32_t)c.slot << 16) | ((uint32_t)k << 8) | (uint32_t)(attempt & 0xff)};
        uint32_t o[4];
        philox4x32_10(ctr, key, o);
        ua = bits53(((uint64_t)o[0] << 32) | o[1]);
        ub = bits53(((uint64_t)o[2] << 32) | o[3]);
    }

    static double box_muller(double ua, double ub) {
        return std::sqrt(-2.0 * std::log(1.0 - ua)) * std::cos(TWO_PI * ub);
    }

    void record(DrawKind k, const DrawCtx& c, double v) {
        if (tape.slot == nullptr) return;
        double* p = tape.at(k, c);
        if (p) *p = v;
    }

    double replay(DrawKind k, const DrawCtx& c) {
        double* p = tape.at(k, c);
        if (!p) { error = true; return 0.0; }
        return *p;
    }

    // rand: random::<f64>()
    double uniform(DrawKind k, const DrawCtx& c) {
        double v;
        if (mode == REPLAY) return replay(k, c);
        if (mode == PHILOX) { double ub; philox_pair(k, c, 0, v, ub); return v; }
        v = bits53(raw_word());
        record(k, c, v);
        return v;
    }

    // rand: random_range(lo..hi) for f64.  Returns the split; the typed tape
    // stores the accepted fraction u3.
    double range(const DrawCtx& c, double lo, double hi) {
        const double scale = hi - lo;
        if (mode == REPLAY) {
            const double u = replay(D_SPLIT, c);
            const double r = u * scale + lo;
            if (!(r < hi)) error = true;
            return r;
        }
        for (int attempt = 0; attempt < 64; ++attempt) {
            double u;
            if (mode == PHILOX) { double ub; philox_pair(D_SPLIT, c, attempt, u, ub); }
            else {
                // value in [1,2) from the top 52 bits, minus 1
                const uint64_t bits = (raw_word() >> 12) | 0x3ff0000000000000ULL;
                double v12; std::memcpy(&v12, &bits, 8);
                u = v12 - 1.0;
            }
            const double r = u * scale + lo;
            if (r < hi) { if (mode == SEQ || mode == FLAT) record(D_SPLIT, c, u); return r; }
        }
        if (mode == SEQ || mode == FLAT) record(D_SPLIT, c, 0.0);
        return lo;
    }

    // rand_distr: Normal(0,1).sample()
    double normal(DrawKind k, const DrawCtx& c) {
        if (mode == REPLAY) return replay(k, c);
        double ua, ub;
        if (mode == FLAT) {             // the normal itself was logged
            if (flat_pos >= flat_n) { error = true; return 0.0; }
            const double z = flat[flat_pos++];
            record(k, c, z);
            return z;
        }
        if (mode == PHILOX) philox_pair(k, c, 0, ua, ub);
        else { ua = bits53(seq.next()); ub = bits53(seq.next()); }
        const double z = box_muller(ua, ub);
        if (mode == SEQ) {
            record(k, c, z);
            if (flat_out) { if (flat_w < flat_cap) flat_out[flat_w] = z; flat_w++; }
        }
        return z;
    }
};

// ---------------------------------------------------------------------------
// tree.rs
// ---------------------------------------------------------------------------

static const uint32_t LEAF_SENTINEL = 0xFFFFFFFFu;  // tree.rs:25

static inline size_t max_nodes_for_depth(uint8_t depth) {  // tree.rs:196-198
    return ((size_t)1 << ((size_t)depth + 1)) - 1;
}

struct TreeArrays {  // tree.rs:9-23
    std::vector<uint32_t> split_var;
    std::vector<double> split_val;
    std::vector<double> leaf_val;
    std::vector<uint32_t> leaf_indices;
    size_t size = 0;
    uint8_t max_depth = 0;

    TreeArrays() = default;
    TreeArrays(double init_leaf_value, size_t n_samples, uint8_t md) {  // tree.rs:29-49
        const size_t max_nodes = max_nodes_for_depth(md);
        split_var.reserve(std::min<size_t>(max_nodes, 1u << 20));
        split_val.reserve(std::min<size_t>(max_nodes, 1u << 20));
        leaf_val.reserve(std::min<size_t>(max_nodes, 1u << 20));
        split_var.push_back(LEAF_SENTINEL);
        split_val.push_back(std::numeric_limits<double>::quiet_NaN());
        leaf_val.push_back(init_leaf_value);
        leaf_indices.assign(n_samples, 0u);
        size = 1;
        max_depth = md;
    }

    static size_t get_depth(size_t node_idx) {  // tree.rs:54-57
        return 63 - (size_t)__builtin_clzll((unsigned long long)(node_idx + 1));
    }

    bool is_leaf(size_t node_idx) const {  // tree.rs:60-62
        return node_idx < split_var.size() && split_var[node_idx] == LEAF_SENTINEL;
    }

    // returns false where the reference panics (tree.rs:98-102)
    bool split_node(size_t leaf_idx, uint32_t var, double val, double left_val, double right_val) {
        const size_t left_child = 2 * leaf_idx + 1;
        const size_t right_child = 2 * leaf_idx + 2;
        if (!(right_child < max_nodes_for_depth(max_depth))) return false;
        const size_t required = right_child + 1;
        while (split_var.size() < required) {  // tree.rs:105-110
            split_var.push_back(LEAF_SENTINEL);
            split_val.push_back(std::numeric_limits<double>::quiet_NaN());
            leaf_val.push_back(0.0);
        }
        split_var[leaf_idx] = var;
        split_val[leaf_idx] = val;
        leaf_val[leaf_idx] = std::numeric_limits<double>::quiet_NaN();
        split_var[left_child] = LEAF_SENTINEL;
        split_val[left_child] = std::numeric_limits<double>::quiet_NaN();
        leaf_val[left_child] = left_val;
        split_var[right_child] = LEAF_SENTINEL;
        split_val[right_child] = std::numeric_limits<double>::quiet_NaN();
        leaf_val[right_child] = right_val;
        size = std::max(size, right_child + 1);
        return true;
    }

    void predict_training_into(double* out) const {  // tree.rs:157-166
        const double* lv = leaf_val.data();
        const uint32_t* li = leaf_indices.data();
        const size_t n = leaf_indices.size();
        for (size_t i = 0; i < n; ++i) out[i] = lv[li[i]];
    }
};

// ---------------------------------------------------------------------------
// particle.rs
// ---------------------------------------------------------------------------

struct LeafSamplesFlat {  // particle.rs:13-46
    std::vector<uint32_t> data, node_start, node_len;
    LeafSamplesFlat() = default;
    LeafSamplesFlat(size_t n_samples, uint8_t max_depth) {
        const size_t max_nodes = max_nodes_for_depth(max_depth);
        data.resize(n_samples);
        for (size_t i = 0; i < n_samples; ++i) data[i] = (uint32_t)i;
        node_start.assign(max_nodes, 0u);
        node_len.assign(max_nodes, 0u);
        node_len[0] = (uint32_t)n_samples;
    }
};

struct TreeProposal {  // update.rs:9-16
    size_t node_idx;
    uint32_t split_var;
    double split_val, left_value, right_value;
};

struct Particle {  // particle.rs:48-86
    std::shared_ptr<TreeArrays> tree;
    std::deque<uint32_t> expandable_nodes;
    LeafSamplesFlat sample_map;

    static Particle fresh(double init_leaf, size_t n, uint8_t md, bool reference) {
        Particle p;
        p.tree = std::make_shared<TreeArrays>(init_leaf, n, md);
        if (!reference) p.expandable_nodes.push_back(0);
        p.sample_map = LeafSamplesFlat(n, md);
        return p;
    }

    // particle.rs:91-148.  x is column-major (F-order) f64, n rows.
    bool apply_mutation(const TreeProposal& pr, const double* x, size_t n) {
        const size_t node_idx = pr.node_idx;
        const size_t left_child = 2 * node_idx + 1, right_child = 2 * node_idx + 2;
        const uint32_t lc = (uint32_t)left_child, rc = (uint32_t)right_child;
        if (tree.use_count() > 1) tree = std::make_shared<TreeArrays>(*tree);  // Arc::make_mut
        if (!tree->split_node(node_idx, pr.split_var, pr.split_val, pr.left_value, pr.right_value))
            return false;
        const size_t start = sample_map.node_start[node_idx];
        const size_t len = sample_map.node_len[node_idx];
        const double* col = x + (size_t)pr.split_var * n;
        uint32_t* slice = sample_map.data.data() + start;
        uint32_t* leaf_indices = tree->leaf_indices.data();
        size_t l = 0, r = len;
        while (l < r) {  // unstable two-pointer partition, particle.rs:121-137
            const uint32_t s = slice[l];
            if (col[s] < pr.split_val) {
                leaf_indices[s] = lc;
                ++l;
            } else {
                leaf_indices[s] = rc;
                --r;
                std::swap(slice[l], slice[r]);
            }
        }
        sample_map.node_start[left_child] = (uint32_t)start;
        sample_map.node_len[left_child] = (uint32_t)l;
        sample_map.node_start[right_child] = (uint32_t)(start + l);
        sample_map.node_len[right_child] = (uint32_t)(len - l);
        sample_map.node_len[node_idx] = 0;
        expandable_nodes.push_back(lc);
        expandable_nodes.push_back(rc);
        return true;
    }
};

// ---------------------------------------------------------------------------
// weight.rs — log-likelihood of mu = predictions
// ---------------------------------------------------------------------------

enum Family : int { FAM_GAUSSIAN = 0, FAM_BERNOULLI_LOGIT = 1, FAM_GAUSSIAN_KERNEL = 2, FAM_CALLBACK = 3 };

typedef double (*LogpFunc)(const double*, size_t);  // lib.rs:33

struct WeightFn {
    int family = FAM_GAUSSIAN;
    double sigma = 1.0;
    const double* y = nullptr;
    size_t n = 0;
    LogpFunc cb = nullptr;

    double log_weight(const double* pred) const {
        switch (family) {
            case FAM_GAUSSIAN: {
                // Normal logp: -0.5((y-mu)/sigma)^2 - log(sigma) - 0.5 log(2 pi), summed.
                const double c = -std::log(sigma) - 0.5 * std::log(TWO_PI);
                double acc = 0.0;
                for (size_t i = 0; i < n; ++i) {
                    const double z = (y[i] - pred[i]) / sigma;
                    acc += -0.5 * z * z + c;
                }
                return acc;
            }
            case FAM_BERNOULLI_LOGIT: {
                // Bernoulli(logit_p = mu): y*mu - softplus(mu)
                double acc = 0.0;
                for (size_t i = 0; i < n; ++i) {
                    const double mu = pred[i];
                    const double sp = std::max(mu, 0.0) + std::log1p(std::exp(-std::fabs(mu)));
                    acc += y[i] * mu - sp;
                }
                return acc;
            }
            case FAM_GAUSSIAN_KERNEL: {  // benches/particle_gibbs.rs:22-32
                double acc = 0.0;
                for (size_t i = 0; i < n; ++i) {
                    const double r = pred[i] - y[i];
                    acc += -0.5 * r * r;
                }
                return acc;
            }
            default:
                return cb ? cb(pred, n) : 0.0;
        }
    }
};

// ---------------------------------------------------------------------------
// config.rs / state.rs
// ---------------------------------------------------------------------------

struct BartConfig {  // config.rs:4-18
    size_t n_trees = 50, n_particles = 10;
    uint8_t max_depth = 6;
    double alpha = 0.95, beta = 2.0, sigma = 1.0;
    std::vector<double> splitting_probs;  // empty = None
    double batch_tune = 0.1, batch_post = 0.1;
    // SURVEY.md 8(f4) modes (NOT the reference's behaviour; each has its device counterpart and its own parity tests):
    //   pg_correct: per-slot log-weights are kept as log-likelihoods, normalised into a copy, and permuted with the
    //               ancestors (instead of the stale-weight rule Q1, smc.rs:53,89-102); particle 0 is the tree being
    //               replaced, with nothing to expand (conditional SMC; instead of the fresh stump of smc.rs:42-50, Q2).
    bool pg_correct = false;
    // split_rules[j] == OneHotSplit (splitting.rs:72-104 through SplitRules::sample_split_value :127-137)
    std::vector<uint8_t> onehot;
};

struct BartInfo {  // state.rs:16-20
    double log_likelihood = 0.0;
    size_t acceptance_count = 0;
    std::vector<uint8_t> tree_depths;
};

// ---------------------------------------------------------------------------
// Trace (shared layout with the device library; see include/pgbart.h)
// ---------------------------------------------------------------------------

static const int TS_FIELDS = 12;
struct Trace {
    double* slot = nullptr;   // [U][R][S][12]
    double* round = nullptr;  // [U][R][2+2S]
    double* upd = nullptr;    // [U][4+2P]
    int64_t u_cap = 0;
    int r_cap = 0;
    int64_t n_upd = 0;
    bool overflow = false;
};

// ndarray numeric_util::unrolled_fold restated (8 partial sums)
static double unrolled_sum(const double* xs, size_t len) {
    double acc = 0.0, p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (len >= 8) {
        for (int j = 0; j < 8; ++j) p[j] += xs[j];
        xs += 8; len -= 8;
    }
    acc += p[0] + p[4];
    acc += p[1] + p[5];
    acc += p[2] + p[6];
    acc += p[3] + p[7];
    for (size_t i = 0; i < len && i < 7; ++i) acc += xs[i];
    return acc;
}

// smc.rs:257-268
static void normalize_weights_inplace(double* w, size_t len) {
    double mx = -std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < len; ++i) mx = std::max(mx, w[i]);
    for (size_t i = 0; i < len; ++i) w[i] = std::exp(w[i] - mx);
    double sum = 0.0;
    for (size_t i = 0; i < len; ++i) sum += w[i];
    for (size_t i = 0; i < len; ++i) w[i] /= sum;
}

// resampling.rs:24-44
static void systematic_resample(double u, const double* weights, size_t n, size_t* out) {
    size_t current_idx = 0;
    double current_cum = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double target = ((double)i + u) / (double)n;
        while (current_cum < target && current_idx < n) {
            current_cum += weights[current_idx];
            current_idx += 1;
        }
        out[i] = current_idx == 0 ? 0 : current_idx - 1;
    }
}

// smc.rs:242-254
static size_t sample_feature_from_probs(double u, const double* probs, size_t p) {
    double total = 0.0;
    for (size_t i = 0; i < p; ++i) total += probs[i];
    double target = u * total;
    for (size_t i = 0; i < p; ++i) {
        target -= probs[i];
        if (target <= 0.0) return i;
    }
    return p - 1;
}

// rand WeightedIndex: cumulative weights without the last; uniform in [0,total);
// index = number of cumulative weights <= chosen.
static size_t weighted_index(double u, const double* w, size_t len) {
    double total = 0.0;
    std::vector<double> cum;
    cum.reserve(len);
    for (size_t i = 0; i < len; ++i) {
        total += w[i];
        if (i + 1 < len) cum.push_back(total);
    }
    const double chosen = u * total;
    size_t idx = 0;
    while (idx < cum.size() && cum[idx] <= chosen) ++idx;
    return idx;
}

struct Sampler {
    // data.rs: OwnedData (F-order x)
    std::vector<double> x, y;
    size_t n = 0, p = 0;
    BartConfig cfg;
    WeightFn wf;
    Draws draws;
    Trace trace;
    // state.rs: BartState
    std::vector<TreeArrays> forest;
    std::vector<double> predictions;
    size_t next_tree_idx = 0;
    bool tune = true;
    uint64_t upd_counter = 0;
    BartInfo last_info;
    std::string error;
    std::vector<double> thr;  // depth-prior thresholds 1 - alpha (1+d)^-beta

    void init() {  // kernel.rs:38-58
        const double y_mean = n ? unrolled_sum(y.data(), n) / (double)n : 0.0;
        const double init_leaf = y_mean / (double)cfg.n_trees;
        forest.clear();
        forest.reserve(cfg.n_trees);
        for (size_t t = 0; t < cfg.n_trees; ++t) forest.emplace_back(init_leaf, n, cfg.max_depth);
        predictions.assign(n, y_mean);
        next_tree_idx = 0;
        tune = true;
        thr.resize((size_t)cfg.max_depth + 2);
        for (size_t d = 0; d < thr.size(); ++d)
            thr[d] = 1.0 - (cfg.alpha * std::pow(1.0 + (double)d, -cfg.beta));  // smc.rs:143
    }

    double* ts(const DrawCtx& c) {
        if (!trace.slot) return nullptr;
        const int64_t u = trace.n_upd;
        if (u >= trace.u_cap || c.round >= trace.r_cap) { trace.overflow = true; return nullptr; }
        const size_t S = cfg.n_particles - 1;
        return trace.slot + (((size_t)u * trace.r_cap + c.round) * S + c.slot) * TS_FIELDS;
    }

    // smc.rs:133-187 + 189-240.  Returns true on Accept.
    bool propose_mutation(const Particle& particle, size_t node_idx, const double* others,
                          const DrawCtx& c, TreeProposal& out) {
        double* tr = ts(c);
        if (tr) {
            for (int f = 0; f < TS_FIELDS; ++f) tr[f] = std::numeric_limits<double>::quiet_NaN();
            tr[0] = 1; tr[1] = (double)node_idx; tr[2] = -1;
        }
        const size_t depth = TreeArrays::get_depth(node_idx);
        const double prob_not_expanding = depth < thr.size()
            ? thr[depth] : 1.0 - (cfg.alpha * std::pow(1.0 + (double)depth, -cfg.beta));
        if (prob_not_expanding > draws.uniform(D_GROW, c)) return false;

        const uint32_t* node_samples = particle.sample_map.data.data() + particle.sample_map.node_start[node_idx];
        const size_t len = particle.sample_map.node_len[node_idx];
        if (len == 0) { if (tr) tr[0] = 2; return false; }

        size_t split_var;
        {
            const double u = draws.uniform(D_VAR, c);
            if (!cfg.splitting_probs.empty())
                split_var = sample_feature_from_probs(u, cfg.splitting_probs.data(), cfg.splitting_probs.size());
            else
                split_var = std::min<size_t>((size_t)(u * (double)p), p - 1);
        }
        if (tr) tr[2] = (double)split_var;
        const double* col = x.data() + split_var * n;

        if (split_var < cfg.onehot.size() && cfg.onehot[split_var]) {
            // splitting.rs:127-137 casts the node's values to i32, :79-90 collects the distinct ones and picks one
            // uniformly.  The reference iterates a HashSet (process-random order, Q7); here the distinct codes are taken
            // in ascending order, which has the same distribution.  random_range(0..len) for usize is restated as
            // floor of the f64 range draw on [0, len), like the uniform variable draw above.  The partition and the
            // leaf statistics below use `v < split_val` for every rule (particle.rs:128, smc.rs:208: Q7).
            std::vector<int32_t> codes(len);
            for (size_t i = 0; i < len; ++i) {
                const double v = col[node_samples[i]];
                codes[i] = std::isnan(v) ? 0 : (v >= 2147483647.0 ? INT32_MAX : (v <= -2147483648.0 ? INT32_MIN : (int32_t)v));  // Rust `as i32`
            }
            std::sort(codes.begin(), codes.end());
            codes.erase(std::unique(codes.begin(), codes.end()), codes.end());
            if (tr) { tr[9] = (double)codes.front(); tr[10] = (double)codes.back(); }
            if (codes.size() <= 1) { if (tr) tr[0] = 3; return false; }
            const size_t k = std::min<size_t>((size_t)draws.range(c, 0.0, (double)codes.size()), codes.size() - 1);
            return finish_proposal(particle, node_idx, others, c, split_var, (double)codes[k], out, tr);
        }

        // splitting.rs:39-56
        double mn = col[node_samples[0]], mx = mn;
        for (size_t i = 1; i < len; ++i) {
            const double v = col[node_samples[i]];
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        if (tr) { tr[9] = mn; tr[10] = mx; }
        if (mn >= mx) { if (tr) tr[0] = 3; return false; }
        const double split_val = draws.range(c, mn, mx);
        return finish_proposal(particle, node_idx, others, c, split_var, split_val, out, tr);
    }

    // smc.rs:189-240 propose_leaf_values + the Accept arm of propose_mutation
    bool finish_proposal(const Particle& particle, size_t node_idx, const double* others, const DrawCtx& c,
                         size_t split_var, double split_val, TreeProposal& out, double* tr) {
        const uint32_t* node_samples = particle.sample_map.data.data() + particle.sample_map.node_start[node_idx];
        const size_t len = particle.sample_map.node_len[node_idx];
        const double* col = x.data() + split_var * n;
        // smc.rs:197-217
        double l_sum = 0.0, r_sum = 0.0;
        size_t l_n = 0, r_n = 0;
        for (size_t i = 0; i < len; ++i) {
            const uint32_t idx = node_samples[i];
            const double v = col[idx];
            const double pp = others[idx];
            if (v < split_val) { l_sum += pp; l_n += 1; }
            else { r_sum += pp; r_n += 1; }
        }
        // smc.rs:219-239 (noise drawn for both children, left first)
        const double noise_l = draws.normal(D_ZLEFT, c) * cfg.sigma;
        const double left_value = l_n == 0 ? noise_l : l_sum / (double)l_n / (double)cfg.n_trees + noise_l;
        const double noise_r = draws.normal(D_ZRIGHT, c) * cfg.sigma;
        const double right_value = r_n == 0 ? noise_r : r_sum / (double)r_n / (double)cfg.n_trees + noise_r;

        out.node_idx = node_idx;
        out.split_var = (uint32_t)split_var;
        out.split_val = split_val;
        out.left_value = left_value;
        out.right_value = right_value;
        if (tr) {
            tr[0] = 4; tr[3] = split_val; tr[4] = left_value; tr[5] = right_value;
            tr[6] = (double)l_n; tr[7] = (double)r_n; tr[11] = l_sum;
        }
        return true;
    }

    // smc.rs:26-130
    bool smc_step(const std::vector<double>& others, TreeArrays& out_tree, double& info_loglik,
                  size_t& info_acc, size_t tree_idx, const TreeArrays* old_tree = nullptr) {
        const size_t P = cfg.n_particles;
        const size_t S = P - 1;
        const double init_leaf = (n ? unrolled_sum(y.data(), n) / (double)n : 0.0) / (double)cfg.n_trees;  // smc.rs:40

        std::vector<Particle> particles;
        particles.reserve(P);
        for (size_t i = 0; i < P; ++i) particles.push_back(Particle::fresh(init_leaf, n, cfg.max_depth, i == 0));
        if (cfg.pg_correct && old_tree) particles[0].tree = std::make_shared<TreeArrays>(*old_tree);   // conditional SMC (not the reference: Q2)

        std::vector<double> inner_weights(S, 0.0);  // smc.rs:53 (never reset: Q1)
        size_t acceptance_count = 0;
        std::vector<double> predictions_buf(n, 0.0);
        std::vector<double> slot_logw;              // pg_correct: log-likelihood of each slot's particle, permuted with the ancestors
        if (cfg.pg_correct) {
            for (size_t r = 0; r < n; ++r) predictions_buf[r] = init_leaf + others[r];
            slot_logw.assign(S, wf.log_weight(predictions_buf.data()));
        }
        std::vector<size_t> ancestors(S);
        std::vector<Particle> scratch;
        std::vector<char> mutated(S, 0);

        DrawCtx c;
        c.upd = upd_counter;
        int round = 0;
        auto any_expandable = [&]() {
            for (size_t i = 1; i < P; ++i) if (!particles[i].expandable_nodes.empty()) return true;
            return false;
        };
        while (any_expandable()) {
            c.round = round;
            std::fill(mutated.begin(), mutated.end(), 0);
            for (size_t i = 0; i < S; ++i) {
                Particle& particle = particles[1 + i];
                c.slot = (int)i;
                if (particle.expandable_nodes.empty()) {
                    double* tr = ts(c);
                    if (tr) { for (int f = 0; f < TS_FIELDS; ++f) tr[f] = std::numeric_limits<double>::quiet_NaN(); tr[0] = 0; }
                    continue;
                }
                const size_t node_idx = particle.expandable_nodes.front();
                TreeProposal pr;
                if (propose_mutation(particle, node_idx, others.data(), c, pr)) {
                    particle.expandable_nodes.pop_front();
                    if (!particle.apply_mutation(pr, x.data(), n)) {
                        error = "Tree mutation would exceed maximum capacity (reference panics, tree.rs:98-102)";
                        return false;
                    }
                    acceptance_count += 1;
                    mutated[i] = 1;
                } else {
                    particle.expandable_nodes.pop_front();
                }
            }
            size_t n_mut = 0;
            for (size_t i = 0; i < S; ++i) {
                if (!mutated[i]) continue;
                ++n_mut;
                const Particle& particle = particles[1 + i];
                particle.tree->predict_training_into(predictions_buf.data());
                for (size_t r = 0; r < n; ++r) predictions_buf[r] += others[r];
                inner_weights[i] = wf.log_weight(predictions_buf.data());
                if (cfg.pg_correct) slot_logw[i] = inner_weights[i];
                c.slot = (int)i;
                double* tr = ts(c);
                if (tr) tr[8] = inner_weights[i];
            }
            if (cfg.pg_correct) inner_weights = slot_logw;    // every slot enters with its particle's log-likelihood
            normalize_weights_inplace(inner_weights.data(), S);

            c.slot = 0;
            const double u6 = draws.uniform(D_RESAMPLE, c);
            systematic_resample(u6, inner_weights.data(), S, ancestors.data());
            if (trace.round && trace.n_upd < trace.u_cap && round < trace.r_cap) {
                double* tr = trace.round + ((size_t)trace.n_upd * trace.r_cap + round) * (2 + 2 * S);
                tr[0] = 1; tr[1] = (double)n_mut;
                for (size_t i = 0; i < S; ++i) { tr[2 + i] = inner_weights[i]; tr[2 + S + i] = (double)ancestors[i]; }
            } else if (trace.round) trace.overflow = true;

            if (cfg.pg_correct) {
                std::vector<double> lw(S);
                for (size_t i = 0; i < S; ++i) lw[i] = slot_logw[ancestors[i]];
                slot_logw.swap(lw);
            }
            scratch.clear();
            for (size_t i = 0; i < S; ++i) scratch.push_back(particles[1 + ancestors[i]]);  // clone (copies CSR data)
            particles.resize(1);
            for (auto& q : scratch) particles.push_back(std::move(q));
            scratch.clear();
            ++round;
        }

        std::vector<double> weights(P, 0.0), logw(P, 0.0);
        for (size_t i = 0; i < P; ++i) {  // smc.rs:105-110
            particles[i].tree->predict_training_into(predictions_buf.data());
            for (size_t r = 0; r < n; ++r) predictions_buf[r] += others[r];
            weights[i] = wf.log_weight(predictions_buf.data());
            logw[i] = weights[i];
        }
        normalize_weights_inplace(weights.data(), P);
        c.round = 0; c.slot = 0;
        const double u7 = draws.uniform(D_SELECT, c);
        const size_t selected = weighted_index(u7, weights.data(), P);

        out_tree = *particles[selected].tree;
        info_loglik = weights[0];
        info_acc = acceptance_count;

        if (trace.upd) {
            if (trace.n_upd < trace.u_cap) {
                double* tr = trace.upd + (size_t)trace.n_upd * (4 + 2 * P);
                tr[0] = (double)round; tr[1] = (double)selected; tr[2] = (double)tree_idx; tr[3] = (double)acceptance_count;
                for (size_t i = 0; i < P; ++i) { tr[4 + i] = logw[i]; tr[4 + P + i] = weights[i]; }
                trace.n_upd += 1;
            } else trace.overflow = true;
        }
        upd_counter += 1;
        return !draws.error;
    }

    // kernel.rs:60-126
    bool step() {
        const size_t m = cfg.n_trees;
        const double batch_frac = tune ? cfg.batch_tune : cfg.batch_post;
        size_t batch_size = (size_t)std::llround(batch_frac * (double)m);
        batch_size = std::min(std::max<size_t>(batch_size, 1), m);

        last_info = BartInfo();
        std::vector<double> residuals_buf(n, 0.0), tree_pred_buf(n, 0.0);
        for (size_t k = 0; k < batch_size; ++k) {
            const size_t tree_idx = (next_tree_idx + k) % m;
            forest[tree_idx].predict_training_into(tree_pred_buf.data());
            for (size_t i = 0; i < n; ++i) residuals_buf[i] = predictions[i];
            for (size_t i = 0; i < n; ++i) residuals_buf[i] -= tree_pred_buf[i];

            TreeArrays new_tree;
            double ll = 0.0; size_t acc = 0;
            if (!smc_step(residuals_buf, new_tree, ll, acc, tree_idx, &forest[tree_idx])) {
                if (error.empty()) error = "RNG tape exhausted or inconsistent";
                return false;
            }
            new_tree.predict_training_into(tree_pred_buf.data());
            for (size_t i = 0; i < n; ++i) predictions[i] = residuals_buf[i];
            for (size_t i = 0; i < n; ++i) predictions[i] += tree_pred_buf[i];

            last_info.log_likelihood += ll;
            last_info.acceptance_count += acc;
            uint8_t depth = 0;
            for (size_t i = 0; i < new_tree.size; ++i)
                if (new_tree.is_leaf(i)) depth = std::max<uint8_t>(depth, (uint8_t)TreeArrays::get_depth(i));
            last_info.tree_depths.push_back(depth);
            forest[tree_idx] = std::move(new_tree);
        }
        next_tree_idx = (next_tree_idx + batch_size) % m;
        return true;
    }
};

}  // namespace orc

// ---------------------------------------------------------------------------
// C ABI (ctypes from tests / bench)
// ---------------------------------------------------------------------------

extern "C" {

struct orc_settings {
    uint64_t n_trees, n_particles;
    uint32_t max_depth;
    double alpha, beta, sigma;
    const double* split_prior;
    uint64_t n_split_prior;
    double batch_tune, batch_post;
    uint64_t seed;
};

void* orc_create(const double* x_f, uint64_t n, uint64_t p, const double* y, const orc_settings* st) {
    auto* s = new orc::Sampler();
    s->n = n; s->p = p;
    s->x.assign(x_f, x_f + n * p);
    s->y.assign(y, y + n);
    s->cfg.n_trees = st->n_trees;
    s->cfg.n_particles = st->n_particles;
    s->cfg.max_depth = (uint8_t)st->max_depth;
    s->cfg.alpha = st->alpha; s->cfg.beta = st->beta; s->cfg.sigma = st->sigma;
    if (st->n_split_prior) s->cfg.splitting_probs.assign(st->split_prior, st->split_prior + st->n_split_prior);
    s->cfg.batch_tune = st->batch_tune; s->cfg.batch_post = st->batch_post;
    s->draws.set_seed(st->seed);
    s->wf.y = s->y.data();
    s->wf.n = n;
    s->init();
    return s;
}

void orc_destroy(void* h) { delete (orc::Sampler*)h; }

// SURVEY.md 8(f4) modes, set before the first step: pg_correct (0/1), onehot[p] flags (may be NULL)
int orc_set_modes(void* h, int pg_correct, const uint8_t* onehot) {
    auto* s = (orc::Sampler*)h;
    s->cfg.pg_correct = pg_correct != 0;
    if (onehot) s->cfg.onehot.assign(onehot, onehot + s->p); else s->cfg.onehot.clear();
    return 0;
}

int orc_set_likelihood(void* h, int family, const double* params, int nparams) {
    auto* s = (orc::Sampler*)h;
    s->wf.family = family;
    if (family == orc::FAM_GAUSSIAN && nparams >= 1) s->wf.sigma = params[0];
    return 0;
}

int orc_set_callback(void* h, orc::LogpFunc f) {
    auto* s = (orc::Sampler*)h;
    s->wf.family = orc::FAM_CALLBACK;
    s->wf.cb = f;
    return 0;
}

// mode: 0 sequential xoshiro (records into the tape if one is attached),
//       1 Philox by coordinates, 2 replay the attached tape.
int orc_set_rng(void* h, int mode, uint64_t seed) {
    auto* s = (orc::Sampler*)h;
    s->draws.mode = (orc::Draws::Mode)mode;
    s->draws.set_seed(seed);
    return 0;
}

int orc_attach_tape(void* h, double* slot, double* round, double* upd, int64_t n_upd, int r_cap) {
    auto* s = (orc::Sampler*)h;
    s->draws.tape.slot = slot; s->draws.tape.round = round; s->draws.tape.upd = upd;
    s->draws.tape.n_upd = n_upd; s->draws.tape.r_cap = r_cap;
    s->draws.tape.n_slots = (int)s->cfg.n_particles - 1;
    s->draws.tape.base_upd = s->upd_counter;
    s->draws.tape.overflow = false;
    return 0;
}

// FLAT mode input (mode 3 of orc_set_rng) / SEQ mode output: see struct Draws.
int orc_attach_flat(void* h, const double* log, uint64_t n) {
    auto* s = (orc::Sampler*)h;
    s->draws.flat = log; s->draws.flat_n = (size_t)n; s->draws.flat_pos = 0;
    return 0;
}
int orc_record_flat(void* h, double* out, uint64_t cap) {
    auto* s = (orc::Sampler*)h;
    s->draws.flat_out = out; s->draws.flat_cap = (size_t)cap; s->draws.flat_w = 0;
    return 0;
}
uint64_t orc_flat_count(void* h) {
    auto* s = (orc::Sampler*)h;
    return s->draws.mode == orc::Draws::FLAT ? (uint64_t)s->draws.flat_pos : (uint64_t)s->draws.flat_w;
}

int orc_attach_trace(void* h, double* slot, double* round, double* upd, int64_t u_cap, int r_cap) {
    auto* s = (orc::Sampler*)h;
    s->trace.slot = slot; s->trace.round = round; s->trace.upd = upd;
    s->trace.u_cap = u_cap; s->trace.r_cap = r_cap; s->trace.n_upd = 0; s->trace.overflow = false;
    return 0;
}

int64_t orc_trace_count(void* h) { return ((orc::Sampler*)h)->trace.n_upd; }
int orc_overflowed(void* h) {
    auto* s = (orc::Sampler*)h;
    return (s->trace.overflow ? 1 : 0) | (s->draws.tape.overflow ? 2 : 0);
}

int orc_step(void* h, int tune, double* out_pred) {
    auto* s = (orc::Sampler*)h;
    if (tune >= 0) s->tune = tune != 0;
    if (!s->step()) return 1;
    if (out_pred) std::memcpy(out_pred, s->predictions.data(), s->n * sizeof(double));
    return 0;
}

const char* orc_last_error(void* h) { return ((orc::Sampler*)h)->error.c_str(); }

int orc_get_info(void* h, double* loglik, int64_t* acc, uint8_t* depths, int* n_depths) {
    auto* s = (orc::Sampler*)h;
    *loglik = s->last_info.log_likelihood;
    *acc = (int64_t)s->last_info.acceptance_count;
    const int cap = *n_depths;
    *n_depths = (int)s->last_info.tree_depths.size();
    for (int i = 0; i < std::min(cap, *n_depths); ++i) depths[i] = s->last_info.tree_depths[i];
    return 0;
}

int orc_tree_size(void* h, uint64_t t) { return (int)((orc::Sampler*)h)->forest[t].size; }

// split_var: -1 for leaves.  Arrays must hold orc_tree_size entries.
int orc_get_tree(void* h, uint64_t t, int32_t* split_var, double* split_val, double* leaf_val) {
    const auto& tr = ((orc::Sampler*)h)->forest[t];
    for (size_t i = 0; i < tr.size; ++i) {
        split_var[i] = tr.split_var[i] == orc::LEAF_SENTINEL ? -1 : (int32_t)tr.split_var[i];
        split_val[i] = tr.split_val[i];
        leaf_val[i] = tr.leaf_val[i];
    }
    return 0;
}

int orc_get_leaf_indices(void* h, uint64_t t, uint32_t* out) {
    const auto& tr = ((orc::Sampler*)h)->forest[t];
    std::memcpy(out, tr.leaf_indices.data(), tr.leaf_indices.size() * sizeof(uint32_t));
    return 0;
}

uint64_t orc_next_tree_idx(void* h) { return ((orc::Sampler*)h)->next_tree_idx; }
uint64_t orc_upd_counter(void* h) { return ((orc::Sampler*)h)->upd_counter; }

// ---- known-answer-test entry points (mirror the Rust unit tests) ----------

uint64_t orc_kat_max_nodes_for_depth(uint32_t d) { return orc::max_nodes_for_depth((uint8_t)d); }
uint64_t orc_kat_get_depth(uint64_t node) { return orc::TreeArrays::get_depth(node); }

// Build a tree, apply `n_splits` split_node calls, return arrays.  Returns -1
// where the reference panics.
int orc_kat_tree(double init_leaf, uint64_t n_samples, uint32_t max_depth, int n_splits,
                 const uint64_t* leaf_idx, const uint32_t* var, const double* val, const double* lv,
                 const double* rv, int32_t* split_var_out, double* leaf_val_out, uint32_t* leaf_indices_out,
                 double* predict_out, int cap) {
    orc::TreeArrays t(init_leaf, n_samples, (uint8_t)max_depth);
    for (int i = 0; i < n_splits; ++i)
        if (!t.split_node(leaf_idx[i], var[i], val[i], lv[i], rv[i])) return -1;
    for (size_t i = 0; i < t.size && (int)i < cap; ++i) {
        split_var_out[i] = t.split_var[i] == orc::LEAF_SENTINEL ? -1 : (int32_t)t.split_var[i];
        leaf_val_out[i] = t.leaf_val[i];
    }
    if (leaf_indices_out) std::memcpy(leaf_indices_out, t.leaf_indices.data(), n_samples * 4);
    if (predict_out) t.predict_training_into(predict_out);
    return (int)t.size;
}

// Particle::new + apply_mutation (particle.rs:170-201).  x is F-order.
// Outputs: CSR data, node_start/len for the children, leaf_indices, and the
// use_count of the tree after the mutation; `share_first` clones the particle
// before mutating (COW test, particle.rs:204-228) and reports whether the
// original stayed a stump.
int orc_kat_apply_mutation(const double* x_f, uint64_t n, uint64_t p, uint32_t max_depth, uint64_t node,
                           uint32_t var, double split, double lv, double rv, int share_first,
                           uint32_t* data_out, uint32_t* start_len_out /*4*/, uint32_t* leaf_indices_out,
                           int* use_count_out, int* original_is_stump_out) {
    orc::Particle part = orc::Particle::fresh(0.0, n, (uint8_t)max_depth, false);
    orc::Particle other;
    if (share_first) other = part;
    orc::TreeProposal pr{(size_t)node, var, split, lv, rv};
    (void)p;
    if (!part.apply_mutation(pr, x_f, n)) return -1;
    std::memcpy(data_out, part.sample_map.data.data(), n * 4);
    start_len_out[0] = part.sample_map.node_start[2 * node + 1];
    start_len_out[1] = part.sample_map.node_len[2 * node + 1];
    start_len_out[2] = part.sample_map.node_start[2 * node + 2];
    start_len_out[3] = part.sample_map.node_len[2 * node + 2];
    std::memcpy(leaf_indices_out, part.tree->leaf_indices.data(), n * 4);
    *use_count_out = (int)part.tree.use_count();
    *original_is_stump_out = share_first ? (other.tree->is_leaf(0) ? 1 : 0) : -1;
    return 0;
}

void orc_kat_normalize(double* w, uint64_t len) { orc::normalize_weights_inplace(w, len); }
void orc_kat_systematic(double u, const double* w, uint64_t n, uint64_t* out) {
    std::vector<size_t> o(n);
    orc::systematic_resample(u, w, n, o.data());
    for (uint64_t i = 0; i < n; ++i) out[i] = o[i];
}
uint64_t orc_kat_sample_feature(double u, const double* probs, uint64_t p) {
    return orc::sample_feature_from_probs(u, probs, p);
}
uint64_t orc_kat_weighted_index(double u, const double* w, uint64_t len) { return orc::weighted_index(u, w, len); }
void orc_kat_xoshiro(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3, int count, uint64_t* out) {
    orc::Xoshiro256pp g(0);
    g.s[0] = s0; g.s[1] = s1; g.s[2] = s2; g.s[3] = s3;
    for (int i = 0; i < count; ++i) out[i] = g.next();
}
void orc_kat_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { orc::philox4x32_10(ctr, key, out); }
double orc_kat_unrolled_mean(const double* xs, uint64_t len) { return orc::unrolled_sum(xs, len) / (double)len; }
double orc_kat_loglik(int family, double sigma, const double* y, const double* pred, uint64_t n) {
    orc::WeightFn w; w.family = family; w.sigma = sigma; w.y = y; w.n = n;
    return w.log_weight(pred);
}

}  // extern "C"
