// pg_rng.h — random draws of the sweep, addressed by (tree update, round, slot, kind).
//
// Two sources (north_star "replay mode" / "free-running device Philox"):
//   PG_RNG_TAPE    typed tape recorded from the reference-order draws of the CPU
//                  side: tape_slot[U][R][S][5] = u1,u2,u3,zL,zR;
//                  tape_round[U][R] = u6; tape_upd[U] = u7.
//   PG_RNG_PHILOX  Philox4x32-10 keyed by the sampler seed, counter =
//                  (upd lo, upd hi, round, slot<<16 | kind<<8 | attempt).
// Draw sites in the reference: smc.rs:145 (u1), :154-158 (u2), splitting.rs:55
// (u3), smc.rs:219-231 (zL,zR), resampling.rs:27 (u6), smc.rs:113-114 (u7).
#pragma once
#include "pg_types.h"
#include <math.h>

PG_HD void pg_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                            uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)M0 * c0;
        const unsigned long long p1 = (unsigned long long)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

PG_HD double pg_bits53(uint32_t hi, uint32_t lo) {
    const unsigned long long b = ((unsigned long long)hi << 32) | lo;
    return (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

PG_HD void pg_philox_pair(const PgState& st, int kind, unsigned long long upd, int round, int slot, int attempt,
                          double* ua, double* ub) {
    uint32_t o[4];
    pg_philox4x32_10((uint32_t)upd, (uint32_t)(upd >> 32), (uint32_t)round,
                     ((uint32_t)slot << 16) | ((uint32_t)kind << 8) | (uint32_t)(attempt & 0xff),
                     st.philox_key0, st.philox_key1, o);
    *ua = pg_bits53(o[0], o[1]);
    *ub = pg_bits53(o[2], o[3]);
}

// Address of a tape entry, or null (and the error flag raised) when out of range.
PG_HD const double* pg_tape_at(const PgState& st, PgCtrl* c, int kind, unsigned long long upd, int round, int slot) {
    if (upd < st.tape_base_upd) { c->error = PG_ERR_TAPE; return 0; }
    const unsigned long long u = upd - st.tape_base_upd;
    if ((long long)u >= st.tape_n_upd || round >= st.tape_r_cap) { c->error = PG_ERR_TAPE; return 0; }
    if (kind <= PG_D_ZRIGHT)
        return st.tape_slot + (((u * (unsigned long long)st.tape_r_cap + (unsigned long long)round) * (unsigned long long)st.S
                                + (unsigned long long)slot) * 5ull + (unsigned long long)kind);
    if (kind == PG_D_RESAMPLE) return st.tape_round + (u * (unsigned long long)st.tape_r_cap + (unsigned long long)round);
    return st.tape_upd + u;
}

// `c` = the control block to flag errors in (the global one, or the serial section's working copy)
PG_HD double pg_uniform(const PgState& st, PgCtrl* c, int kind, unsigned long long upd, int round, int slot) {
    if (st.rng_mode == PG_RNG_TAPE) {
        const double* p = pg_tape_at(st, c, kind, upd, round, slot);
        return p ? *p : 0.0;
    }
    double ua, ub;
    pg_philox_pair(st, kind, upd, round, slot, 0, &ua, &ub);
    return ua;
}

PG_HD double pg_normal(const PgState& st, PgCtrl* c, int kind, unsigned long long upd, int round, int slot) {
    if (st.rng_mode == PG_RNG_TAPE) {
        const double* p = pg_tape_at(st, c, kind, upd, round, slot);
        return p ? *p : 0.0;
    }
    double ua, ub;
    pg_philox_pair(st, kind, upd, round, slot, 0, &ua, &ub);
    return sqrt(-2.0 * log(1.0 - ua)) * cos(6.283185307179586476925286766559 * ub);
}

// Split value: lo + u3 (hi - lo) with separately rounded multiply and add (the
// reference's `value0_1 * scale + low` is never contracted), redrawn while the
// result is not below hi (rand's UniformFloat::sample_single).
PG_HD double pg_split_draw(const PgState& st, PgCtrl* c, unsigned long long upd, int round, int slot, double lo, double hi) {
    const double scale = hi - lo;
    if (st.rng_mode == PG_RNG_TAPE) {
        const double* p = pg_tape_at(st, c, PG_D_SPLIT, upd, round, slot);
        const double u = p ? *p : 0.0;
#if defined(__CUDA_ARCH__)
        const double r = __dadd_rn(__dmul_rn(u, scale), lo);
#else
        volatile double prod = u * scale;
        const double r = prod + lo;
#endif
        if (!(r < hi)) c->error = PG_ERR_TAPE;
        return r;
    }
    for (int attempt = 0; attempt < 64; ++attempt) {
        double u, ub;
        pg_philox_pair(st, PG_D_SPLIT, upd, round, slot, attempt, &u, &ub);
#if defined(__CUDA_ARCH__)
        const double r = __dadd_rn(__dmul_rn(u, scale), lo);
#else
        volatile double prod = u * scale;
        const double r = prod + lo;
#endif
        if (r < hi) return r;
    }
    return lo;
}

// convenience forms on the global control block
PG_HD double pg_uniform(const PgState& st, int kind, unsigned long long upd, int round, int slot) { return pg_uniform(st, st.ctrl, kind, upd, round, slot); }
PG_HD double pg_normal(const PgState& st, int kind, unsigned long long upd, int round, int slot) { return pg_normal(st, st.ctrl, kind, upd, round, slot); }
PG_HD double pg_split_draw(const PgState& st, unsigned long long upd, int round, int slot, double lo, double hi) { return pg_split_draw(st, st.ctrl, upd, round, slot, lo, hi); }
