"""Independent pure-Python restatement of the reference's step semantics.

Second implementation used only to cross-check the C++ oracle on tiny cases: it reads the typed
RNG tape the oracle recorded and must reproduce the oracle's trace.  Test infrastructure only.
Follows /root/reference: src/kernel.rs:60-126, src/smc.rs:26-268, src/particle.rs:91-148,
src/resampling.rs:24-44, src/splitting.rs:39-56.
"""
from __future__ import annotations

import math

import numpy as np

TS_FIELDS = 12
NAN = float("nan")


def loglik(family, sigma, y, mu):
    if family == 0:
        c = -math.log(sigma) - 0.5 * math.log(2 * math.pi)
        acc = 0.0
        for yi, mi in zip(y, mu):
            z = (yi - mi) / sigma
            acc += -0.5 * z * z + c
        return acc
    if family == 1:
        acc = 0.0
        for yi, mi in zip(y, mu):
            acc += yi * mi - (max(mi, 0.0) + math.log1p(math.exp(-abs(mi))))
        return acc
    acc = 0.0
    for yi, mi in zip(y, mu):
        r = mi - yi
        acc += -0.5 * r * r
    return acc


class Tree:
    def __init__(self, leaf, n):
        self.split_var = {0: -1}
        self.split_val = {0: NAN}
        self.leaf_val = {0: leaf}
        self.leaf_idx = [0] * n

    def copy(self):
        t = Tree(0.0, 0)
        t.split_var = dict(self.split_var)
        t.split_val = dict(self.split_val)
        t.leaf_val = dict(self.leaf_val)
        t.leaf_idx = list(self.leaf_idx)
        return t

    def predict(self):
        return [self.leaf_val[k] for k in self.leaf_idx]


class PyRestatement:
    def __init__(self, x, y, *, n_trees, n_particles, max_depth, alpha, beta, sigma, split_prior, batch_tune,
                 batch_post, family=0, lik_sigma=1.0):
        self.x = np.asarray(x, dtype=np.float64)
        self.y = [float(v) for v in y]
        self.n, self.p = self.x.shape
        self.m, self.P, self.max_depth = n_trees, n_particles, max_depth
        self.alpha, self.beta, self.sigma = alpha, beta, sigma
        self.prior = [float(v) for v in split_prior]
        self.batch = (batch_tune, batch_post)
        self.family, self.lik_sigma = family, lik_sigma
        ymean = float(np.mean(self.y)) if self.n else 0.0
        self.init_leaf = ymean / n_trees
        self.forest = [Tree(self.init_leaf, self.n) for _ in range(n_trees)]
        self.pred = [ymean] * self.n
        self.next = 0
        self.tune = True
        self.upd = 0

    def step(self, tape, trace, tune=None):
        if tune is not None:
            self.tune = tune
        frac = self.batch[0] if self.tune else self.batch[1]
        B = min(max(int(math.floor(frac * self.m + 0.5)), 1), self.m)  # f64::round = half away from zero
        for k in range(B):
            t = (self.next + k) % self.m
            old = self.forest[t].predict()
            others = [a - b for a, b in zip(self.pred, old)]
            new = self.smc(others, tape, trace, t)
            nv = new.predict()
            self.pred = [a + b for a, b in zip(others, nv)]
            self.forest[t] = new
        self.next = (self.next + B) % self.m
        return list(self.pred)

    def smc(self, others, tape, trace, t):
        S, P, n = self.P - 1, self.P, self.n
        u = self.upd
        trees = [Tree(self.init_leaf, n) for _ in range(S)]
        queues = [[0] for _ in range(S)]
        members = [{0: list(range(n))} for _ in range(S)]
        w = [0.0] * S
        acc = 0
        rnd = 0
        while any(queues):
            mutated = [False] * S
            for s in range(S):
                rec = [NAN] * TS_FIELDS
                if not queues[s]:
                    rec[0] = 0
                    trace.slot[u, rnd, s] = rec
                    continue
                node = queues[s].pop(0)
                rec[0], rec[1], rec[2] = 1, node, -1
                u1, u2, u3, zl, zr = tape.slot[u, rnd, s]
                depth = int(math.floor(math.log2(node + 1)))
                if 1.0 - self.alpha * (1.0 + depth) ** (-self.beta) > u1:
                    trace.slot[u, rnd, s] = rec
                    continue
                rows = members[s][node]
                if not rows:
                    rec[0] = 2
                    trace.slot[u, rnd, s] = rec
                    continue
                if self.prior:
                    tgt = u2 * sum(self.prior)
                    var = len(self.prior) - 1
                    for j, q in enumerate(self.prior):
                        tgt -= q
                        if tgt <= 0:
                            var = j
                            break
                else:
                    var = min(int(u2 * self.p), self.p - 1)
                rec[2] = var
                col = self.x[:, var]
                lo = min(col[i] for i in rows)
                hi = max(col[i] for i in rows)
                rec[9], rec[10] = lo, hi
                if lo >= hi:
                    rec[0] = 3
                    trace.slot[u, rnd, s] = rec
                    continue
                split = u3 * (hi - lo) + lo
                L = [i for i in rows if col[i] < split]
                R = [i for i in rows if not col[i] < split]
                sl = 0.0
                sr = 0.0
                for i in rows:  # reference order of accumulation: node order
                    if col[i] < split:
                        sl += others[i]
                    else:
                        sr += others[i]
                vl = zl * self.sigma if not L else sl / len(L) / self.m + zl * self.sigma
                vr = zr * self.sigma if not R else sr / len(R) / self.m + zr * self.sigma
                if 2 * node + 2 >= 2 ** (self.max_depth + 1) - 1:
                    raise RuntimeError("Tree mutation would exceed maximum capacity")
                tr = trees[s] = trees[s].copy()
                tr.split_var[node], tr.split_val[node], tr.leaf_val[node] = var, split, NAN
                for child, val, rr in ((2 * node + 1, vl, L), (2 * node + 2, vr, R)):
                    tr.split_var[child], tr.split_val[child], tr.leaf_val[child] = -1, NAN, val
                    for i in rr:
                        tr.leaf_idx[i] = child
                members[s] = dict(members[s])
                members[s][2 * node + 1], members[s][2 * node + 2] = L, R
                queues[s] += [2 * node + 1, 2 * node + 2]
                mutated[s] = True
                acc += 1
                rec[0], rec[3], rec[4], rec[5], rec[6], rec[7], rec[11] = 4, split, vl, vr, len(L), len(R), sl
                trace.slot[u, rnd, s] = rec
            for s in range(S):
                if mutated[s]:
                    mu = [a + b for a, b in zip(others, trees[s].predict())]
                    w[s] = loglik(self.family, self.lik_sigma, self.y, mu)
                    trace.slot[u, rnd, s, 8] = w[s]
            mx = max(w)
            w = [math.exp(v - mx) for v in w]
            tot = 0.0
            for v in w:
                tot += v
            w = [v / tot for v in w]
            u6 = tape.round[u, rnd]
            anc, cum, c = [], 0.0, 0
            for i in range(S):
                tgt = (i + u6) / S
                while cum < tgt and c < S:
                    cum += w[c]
                    c += 1
                anc.append(0 if c == 0 else c - 1)
            trace.round[u, rnd, 0] = 1
            trace.round[u, rnd, 1] = sum(mutated)
            trace.round[u, rnd, 2:2 + S] = w
            trace.round[u, rnd, 2 + S:2 + 2 * S] = anc
            trees = [trees[a] for a in anc]
            queues = [list(queues[a]) for a in anc]
            members = [members[a] for a in anc]
            rnd += 1
        stump = Tree(self.init_leaf, n)
        cands = [stump] + trees
        W = []
        for tr in cands:
            mu = [a + b for a, b in zip(others, tr.predict())]
            W.append(loglik(self.family, self.lik_sigma, self.y, mu))
        mx = max(W)
        Wn = [math.exp(v - mx) for v in W]
        tot = 0.0
        for v in Wn:
            tot += v
        Wn = [v / tot for v in Wn]
        chosen = tape.upd[u] * sum(Wn)
        cum, sel = 0.0, 0
        for i in range(P - 1):
            cum += Wn[i]
            if cum <= chosen:
                sel = i + 1
            else:
                break
        trace.upd[u, 0:4] = [rnd, sel, t, acc]
        trace.upd[u, 4:4 + P] = W
        trace.upd[u, 4 + P:4 + 2 * P] = Wn
        self.upd += 1
        return cands[sel]
