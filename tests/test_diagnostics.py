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


def test_diagnostics_equal_the_references_own_code_on_a_fresh_series(tmp_path):
    """Row a17 pinned to the reference's code, not only to its published table: `abm_ref series` runs reference chains and stores
    their synopsis series; `abm_ref refstats` feeds that series to the reference's own ChainStats / MultiChainStats
    (diagnostics/ChainStats.h:27-40, MultiChainStats.h:81-181, built as FiguresForPaper::startStatsThread builds them: two
    half-sequences per chain, 200 lags, maxLagProportion 0.5) and prints W, B, var+, R-hat, n_eff; diagnostics.py must agree."""
    import json
    import lzma
    import subprocess
    from pathlib import Path
    import numpy as np
    import pytest
    from agentbasedmcmc_b200 import diagnostics as D
    from agentbasedmcmc_b200.abmp import read_abmp
    root = Path(__file__).resolve().parent.parent
    ref = root / "oracle" / "_ref" / "abm_ref"
    if not ref.exists():
        pytest.skip("oracle/_ref/abm_ref not built (needs /root/reference at build time)")
    prob, ser = tmp_path / "p.abmp", tmp_path / "s.abmp"
    prob.write_bytes(lzma.decompress((root / "tests" / "golden" / "pp8-4.abmp.xz").read_bytes()))
    subprocess.run([str(ref), "series", str(prob), "3", "6000", "500", str(ser)], check=True, capture_output=True, timeout=600)
    out = subprocess.run([str(ref), "refstats", str(ser), "1"], check=True, capture_output=True, text=True, timeout=600)
    r = json.loads(out.stdout.strip().splitlines()[-1])
    a = read_abmp(ser)
    n_seq, n, d = (int(v) for v in a["shape"])
    x = a["series"].reshape(n_seq, n, d).astype(np.float64)
    half = n // 2
    ms = D.MultiChainStats([D.ChainStats.from_series(x[c, h * half:(h + 1) * half], n_lags=200, max_lag_proportion=0.5)
                            for c in range(n_seq) for h in range(2)])
    assert r["n_sequences"] == ms.n_chains() and r["n_samples_per_sequence"] == ms.n_samples_per_chain()
    assert r["vario_stride"] == ms.chains[0].vario_stride
    assert np.allclose(ms.mean(), r["mean"], rtol=1e-12) and np.allclose(ms.W(), r["W"], rtol=1e-12)
    assert np.allclose(ms.B(), r["B"], rtol=1e-10)                       # the reference's integer division included
    assert np.allclose(ms.var_plus(), r["var_plus"], rtol=1e-12)
    assert np.allclose(ms.potential_scale_reduction(), r["rhat"], rtol=1e-12)
    assert np.allclose(ms.effective_samples(), r["n_eff"], rtol=1e-10)


def test_ensemble_effective_samples_on_autoregressive_chains():
    """The many-chain estimator (reference formula, variogram over all chains at a stride) on AR(1) chains whose autocorrelation
    time is known: tau = (1 + phi) / (1 - phi).  The reference's sum counts one whole stride at rho_0 = 1, so it over-estimates tau
    by up to one stride; R-hat of stationary sequences is sqrt(1 + 1 / n_eff per sequence), and pooling 8 chains per sequence
    brings it towards 1."""
    import numpy as np
    from agentbasedmcmc_b200 import diagnostics as D
    rng = np.random.default_rng(3)
    m, n, phi, thin = 1500, 20000, 0.99, 25
    x = np.zeros((m, n))
    e = rng.standard_normal((m, n)) * np.sqrt(1 - phi ** 2)
    x[:, 0] = rng.standard_normal(m)
    for i in range(1, n):
        x[:, i] = phi * x[:, i - 1] + e[:, i]
    r = D.ensemble_effective_samples(x[:, ::thin, None], thin, np.full(m, n), x.sum(1)[:, None], (x * x).sum(1)[:, None], group=8)
    tau = (1 + phi) / (1 - phi)
    assert tau <= r["tau"][0] <= tau + 2.2 * thin
    per_seq = n / r["tau"][0]
    assert abs(r["rhat"][0] - np.sqrt(1 + 1 / per_seq)) < 0.002
    assert r["rhat_grouped"][0] < 1 + 0.25 * (r["rhat"][0] - 1)
    assert abs(r["var_plus"][0] - 1.0) < 0.02
