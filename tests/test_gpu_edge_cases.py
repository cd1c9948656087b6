"""Edge cases of the C-ABI path on the GPU: ragged chain counts (groups / warps / CTAs only partly filled), a problem
without importance function, call-order and argument errors, unsupported problems rejected loudly."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _S():
    from agentbasedmcmc_b200 import sampler
    return sampler


@pytest.mark.parametrize("n_chains", [1, 3, 17, 33])
def test_ragged_chain_counts_match_oracle(n_chains):
    from oracle.oracle import OracleChain, OracleProblem
    S = _S()
    p = load_golden("pp8-4.abmp.xz")
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=5, first_chain_id=7)
    j, acc, inf = smp.step_traced(400)
    OP = OracleProblem(p)
    for c in sorted({0, n_chains // 2, n_chains - 1}):
        oc = OracleChain(OP)
        oc.seed_philox(5, 7 + c)
        assert oc.find_feasible() == int(smp.init_iterations[c])
        oj, oacc, oinf = oc.step(400, trace=True)
        assert np.array_equal(j[c], oj) and np.array_equal(acc[c], oacc) and np.array_equal(inf[c], oinf)
        assert np.array_equal(smp.state(c)[0], oc.X)


def test_without_importance_function():
    """grid_size == 0: log W == 0 (a plain FactoredConvexDistribution); decisions follow the ratio of sums alone."""
    from oracle.oracle import OracleChain, OracleProblem
    S = _S()
    p = load_golden("pp8-4.abmp.xz")
    P = S.Problem(p, with_importance=False)
    smp = S.SparseBasisSampler(P, n_chains=4, seed=11)
    j, acc, inf = smp.step_traced(1500)
    OP = OracleProblem(p, with_importance=False)
    for c in range(4):
        oc = OracleChain(OP)
        oc.seed_philox(11, c)
        oc.find_feasible()
        oj, oacc, oinf = oc.step(1500, trace=True)
        assert np.array_equal(j[c], oj) and np.array_equal(acc[c], oacc)
        assert smp.state(c)[4] == 0.0


def test_errors_are_reported_not_swallowed():
    S = _S()
    p = load_golden("pp8-4.abmp.xz")
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=2, seed=1, find_feasible=False)
    with pytest.raises(S.AbmError, match="feasible"):
        smp.step(10)                                    # stepping before findInitialFeasibleSolution
    with pytest.raises(S.AbmError):
        smp.state(5)                                    # chain index out of range
    with pytest.raises(S.AbmError, match="iteration limit"):
        smp.findInitialFeasibleSolution(max_iterations=3)
    smp.findInitialFeasibleSolution()                   # and the search can be resumed afterwards
    assert np.all(smp.infeasibility() == 0)
    # tableau entries beyond magnitude 4 are rejected, not mis-sampled (a proposal tabulates every move a coupled column
    # could add to a row, -maxAbs..maxAbs: the reference's PredPrey problems need 1, its 3x3 validation experiment 2)
    bad = dict(p)
    bad["nb_val"] = p["nb_val"].copy()
    bad["nb_val"][0] = 5
    with pytest.raises(S.AbmError, match=r"at most 4"):
        S.Problem(bad)
    # a factor table that does not cover its row's reachable range
    bad = dict(p)
    bad["ut_off"] = p["ut_off"].copy()
    bad["ut_off"][1:] -= 1
    bad["ut_off"] = np.maximum(bad["ut_off"], 0)
    with pytest.raises(S.AbmError):
        S.Problem(bad)


def test_sample_is_bounded_by_max_steps_and_counts_only_feasible_states():
    S = _S()
    p = load_golden("pp16-8.abmp.xz")
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=64, seed=2, prior_init=True)
    smp.sample(10 ** 9, max_steps=300)
    end, ss, sq, ns = smp.statistics()
    ctr = smp.counters()
    assert np.all(ctr[:, :4].sum(axis=1) == 300)        # exactly max_steps Metropolis steps per chain
    # feasible-ending steps (MCMCStatistics::nFeasibleSamples, MCMCStatistics.cpp:40-42) == samples taken
    feas = ctr[:, 3] + ctr[:, 4 + 1] + (ctr[:, 2] - ctr[:, 4 + 2])
    assert np.array_equal(ns, feas)
    assert end.sum() > 0 and np.all(sq >= 0)


@pytest.mark.parametrize("tree", ["blocked", "reference"])
def test_paper_scale_problem_matches_oracle(tree):
    """32x32 grid, 16 timesteps (215040 x 164219 basis, tree depth 18: three storage tiers plus the shared-memory top):
    device feasibility search (~4e5 iterations per chain) and Metropolis steps, bit-exact against the oracle."""
    from oracle.oracle import OracleChain, OracleProblem
    S = _S()
    p = load_golden("pp32-16.abmp.xz")
    P = S.Problem(p, sum_tree=tree)
    assert (P.n_rows, P.n_cols) == (215040, 164219)
    smp = S.SparseBasisSampler(P, n_chains=3, seed=77, prior_init=True, return_prior_sample=True)
    j, acc, inf = smp.step_traced(3000)
    OP = OracleProblem(p)
    for c in (0, 2):
        oc = OracleChain(OP, smp.prior_sample[c])
        oc.seed_philox(77, c)
        assert oc.find_feasible() == int(smp.init_iterations[c])
        oj, oacc, oinf = oc.step(3000, trace=True)
        assert np.array_equal(j[c], oj) and np.array_equal(acc[c], oacc) and np.array_equal(inf[c], oinf)
        X, delta, DE, tr, logimp, infeas = smp.state(c)
        assert np.array_equal(X, oc.X) and np.array_equal(delta, oc.delta) and infeas == oc.infeasibility
        assert np.allclose(DE, oc.DE, rtol=1e-9, atol=1e-12) and np.allclose(tr, oc.tree, rtol=1e-9, atol=1e-11 * tr[0])
        assert np.isclose(logimp, oc.log_importance, rtol=1e-9)


def test_feasibility_search_tail_on_the_wide_kernel(monkeypatch):
    """The tail of a large ensemble's feasibility search runs on the 32-lanes-per-chain kernel (abm_chains_find_feasible);
    forced on here for a small ensemble: iteration counts and first feasible states equal the oracle's, and the Metropolis steps
    that follow (16-lane kernel) continue from them."""
    from oracle.oracle import OracleChain, OracleProblem
    monkeypatch.setenv("ABM_INIT_TAIL", "1")
    S = _S()
    p = load_golden("pp16-8.abmp.xz")
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=5, seed=21, first_chain_id=2)
    monkeypatch.delenv("ABM_INIT_TAIL")
    j, acc, inf = smp.step_traced(600)
    OP = OracleProblem(p)
    for c in range(5):
        oc = OracleChain(OP)
        oc.seed_philox(21, 2 + c)
        assert oc.find_feasible() == int(smp.init_iterations[c])
        oj, oacc, oinf = oc.step(600, trace=True)
        assert np.array_equal(j[c], oj) and np.array_equal(acc[c], oacc) and np.array_equal(inf[c], oinf)
        assert np.array_equal(smp.state(c)[0], oc.X)
