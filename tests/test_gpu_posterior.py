"""Free-running generators (BASELINE.json north_star): device chains driven by coordinate-addressed Philox and oracle chains
driven by the sequential xoshiro generator share no draws, so they can only agree in distribution.  Pooled split R-hat of
a few posterior summaries over all chains, and the difference of the posterior mean predictions against its Monte Carlo
standard error."""
import numpy as np
import pytest

from tests import datagen
from tests.test_gpu_parity import cfg_for, make_gpu


def split_rhat(chains):
    """chains: [n_chains][n_draws] -> Gelman-Rubin R-hat on half-chains."""
    x = np.asarray(chains, dtype=np.float64)
    h = x.shape[1] // 2
    x = np.concatenate([x[:, :h], x[:, h:2 * h]], axis=0)
    m, n = x.shape
    w = x.var(axis=1, ddof=1).mean()
    b = n * x.mean(axis=1).var(ddof=1)
    return float(np.sqrt(((n - 1) / n * w + b / n) / w))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_free_running_device_chains_match_oracle_posterior():
    from oracle import pyoracle as po
    n, p, m, P = 150, 3, 10, 10
    X, y = datagen.small(n, p, seed=21)
    burn, keep, n_chains = 300, 900, 3
    probes = [0, n // 3, (2 * n) // 3, n - 1]

    def summaries(pred):
        return [pred.mean()] + [pred[i] for i in probes]

    dev_draws, orc_draws = [], []
    for ch in range(n_chains):
        cfg = cfg_for(y, m, P, p, seed=100 + ch)
        dev = make_gpu()(X, y, cfg)
        dev.set_likelihood_code(0, [1.0])
        dev.set_rng_philox(100 + ch)
        rows = []
        for it in range(burn + keep):
            pred = dev.step(it < burn)
            if it >= burn:
                rows.append(summaries(pred))
        dev_draws.append(np.array(rows))
        dev.close()
        orc = po.OracleSampler(X.astype(np.float64), y, **cfg_for(y, m, P, p, seed=500 + ch))
        orc.set_likelihood(po.FAM_GAUSSIAN, [1.0])
        orc.set_rng(po.RNG_SEQ, 500 + ch)
        rows = []
        for it in range(burn + keep):
            pred = np.asarray(orc.step(it < burn))
            if it >= burn:
                rows.append(summaries(pred))
        orc_draws.append(np.array(rows))
        orc.close()
    dev_draws, orc_draws = np.array(dev_draws), np.array(orc_draws)          # [chain][draw][summary]
    for k in range(dev_draws.shape[2]):
        pooled = np.concatenate([dev_draws[:, :, k], orc_draws[:, :, k]], axis=0)
        r = split_rhat(pooled)
        print(f"summary {k}: pooled split R-hat {r:.4f}, device mean {dev_draws[:, :, k].mean():.4f}, oracle mean {orc_draws[:, :, k].mean():.4f}")
        assert r < 1.01, f"summary {k}: pooled split R-hat {r:.4f}"      # north_star: R-hat < 1.01
        # difference of means against the chain-to-chain Monte Carlo error
        md, mo = dev_draws[:, :, k].mean(axis=1), orc_draws[:, :, k].mean(axis=1)
        se = np.sqrt(md.var(ddof=1) / n_chains + mo.var(ddof=1) / n_chains) + 1e-12
        assert abs(md.mean() - mo.mean()) < 5.0 * se + 0.02 * abs(mo.mean()) + 1e-3, f"summary {k}: {md.mean()} vs {mo.mean()} (se {se})"
