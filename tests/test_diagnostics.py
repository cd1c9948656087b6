"""Ensemble diagnostics: the restated MultiChainStats formulas reproduce the paper's Table 2
(doc/ABMCMC.tex:849-850) from the sufficient statistics stored in the reference's own result archive, and the
.stat reader (when the archive is on this machine) parses the Boost layout."""
from pathlib import Path

import numpy as np
import pytest

from conftest import load_golden
from agentbasedmcmc_b200 import diagnostics as D

TABLE2_RHAT = [1.00727, 1.01167, 1.00451, 1.00153]
TABLE2_NEFF = [420, 417, 489, 1019]


def _from_fixture():
    g = load_golden("stat_PredPrey32-16.abmp.xz")
    m = D.MultiChainStats(kappa=float(g["kappa_exec_ms"][0]), exec_time_ms=int(g["kappa_exec_ms"][1]))
    n = g["n_samples"].size
    d = g["sum"].size // n
    for k in range(n):
        m.chains.append(D.ChainStats(int(g["n_samples"][k]), g["sum"][k * d:(k + 1) * d], g["sum_sq"][k * d:(k + 1) * d],
                                     g["vario"].reshape(n, -1, d)[k], int(g["vario_stride"][k]),
                                     g["mean_end_state"], g["n_accepted"].reshape(n, 2, 2)[k], g["n_proposals"].reshape(n, 2, 2)[k]))
    return m


def test_table2_reproduced_from_reference_archive():
    m = _from_fixture()
    assert m.n_chains() == 8 and m.n_samples_per_chain() == 5_000_000
    assert np.allclose(m.potential_scale_reduction(), TABLE2_RHAT, atol=5e-6)
    assert np.array_equal(np.rint(m.effective_samples()).astype(int), TABLE2_NEFF)
    # paper: 85 % acceptance, 70 % of the steps infeasible (doc/ABMCMC.tex:822)
    acc = sum(c.n_accepted.sum() for c in m.chains) / sum(c.n_proposals.sum() for c in m.chains)
    assert abs(acc - 0.85) < 0.005
    c = m.chains[0]
    infeasible = (c.n_proposals[0, 0] + c.n_accepted[1, 0] + c.n_proposals[0, 1] - c.n_accepted[0, 1]) / c.n_proposals.sum()
    assert abs(infeasible - 0.70) < 0.01


def test_stat_reader_on_reference_archive():
    path = Path("/root/reference/src/data/PredPrey32-16.stat")
    if not path.exists():
        pytest.skip("reference archive not on this machine")
    m = D.read_stat(path)
    f = _from_fixture()
    assert m.exec_time_ms == 1641426 and m.kappa == 10.0
    for a, b in zip(m.chains, f.chains):
        assert a.n_samples == b.n_samples and np.array_equal(a.sum, b.sum) and np.array_equal(a.vario, b.vario)


def test_variogram_and_moment_rhat_agree_with_series_definitions():
    rng = np.random.default_rng(0)
    # AR(1) sequences with known autocorrelation: n_eff ~ n (1-phi)/(1+phi)
    phi, n, m = 0.9, 40000, 6
    chains = []
    for _ in range(m):
        e = rng.normal(size=(n, 2))
        x = np.zeros((n, 2))
        for t in range(1, n):
            x[t] = phi * x[t - 1] + e[t]
        chains.append(D.ChainStats.from_series(x, n_lags=200, max_lag_proportion=0.01))
    ms = D.MultiChainStats(chains)
    assert np.all(np.abs(ms.potential_scale_reduction(reference_integer_division=False) - 1.0) < 0.01)
    neff = ms.effective_samples(reference_integer_division=False)
    expect = n * m * (1 - phi) / (1 + phi)
    assert np.all(np.abs(neff / expect - 1.0) < 0.25)
    r, W, B, vp = D.rhat_from_moments([c.n_samples for c in chains], np.stack([c.sum for c in chains]), np.stack([c.sum_sq for c in chains]))
    assert np.allclose(r, ms.potential_scale_reduction(reference_integer_division=False), rtol=1e-9)
    v = D.variogram(np.arange(10.0)[:, None], n_lags=3, max_lag_proportion=0.5)
    assert np.allclose(v[:, 0], [0.0, 4.0, 16.0])   # stride 2: lags 0, 2, 4
