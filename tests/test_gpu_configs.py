"""`-m gpu` parity on the remaining configurations of BASELINE.json: configs[3] (64x64 grid, 16 timesteps, Poisson start state:
tree depth 20, four tiers of the blocked tree, coupled-column ids beyond 16 bits, i.e. the records' wide slots) and configs[4]
(the window sweep's problems: 32x32 grid, 4 / 8 / 32 timesteps).  A few chains against the oracle each, both sum-tree storages.

A feasibility search from a prior draw or from the empty trajectory takes 1e7 iterations at 64x64x16 (minutes for the
oracle), so the Metropolis runs start from the problem's own real trajectory -- the act-Fermionic trajectory the reference
drew the observations from (`rt_idx` / `rt_val` of the fixture), feasible by construction -- and the feasibility search
itself is compared on a fixed budget of iterations from the empty trajectory (state compared while still infeasible)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
TREES = ["blocked", "reference"]


def _real_trajectory(p):
    init = np.zeros(int(p["pp_meta"][2]), np.int8)
    init[p["rt_idx"]] = np.minimum(p["rt_val"], 127)
    return init


@pytest.mark.parametrize("tree", TREES)
@pytest.mark.parametrize("prob,start,n_chains,n_steps", [("pp64-16-poisson.abmp.xz", "real", 3, 2500), ("pp32-32.abmp.xz", "real", 3, 2500),
                                                       ("pp32-8.abmp.xz", "empty", 2, 2500), ("pp32-4.abmp.xz", "empty", 3, 2500)])
def test_metropolis_steps_match_oracle(prob, start, n_chains, n_steps, tree):
    from agentbasedmcmc_b200 import sampler as S
    from oracle.oracle import OracleChain, OracleProblem
    p = load_golden(prob)
    init = _real_trajectory(p) if start == "real" else np.zeros(0, np.int8)
    seed, first = 977, 40
    P = S.Problem(p, sum_tree=tree)
    smp = S.SparseBasisSampler(P, None if start == "empty" else np.tile(init, (n_chains, 1)), n_chains=n_chains, seed=seed, first_chain_id=first)
    if start == "real":
        assert np.all(smp.init_iterations == 0)            # the real trajectory satisfies every constraint
    j, acc, pinf = smp.step_traced(n_steps)
    ctr = smp.counters()
    OP = OracleProblem(p)
    for c in range(n_chains):
        oc = OracleChain(OP, init)
        oc.seed_philox(seed, first + c)
        assert oc.find_feasible() == int(smp.init_iterations[c])
        oj, oacc, oinf = oc.step(n_steps, trace=True)
        assert np.array_equal(j[c], oj) and np.array_equal(acc[c], oacc) and np.array_equal(pinf[c], oinf)
        X, delta, DE, tr, logimp, inf = smp.state(c)
        assert np.array_equal(X, oc.X) and np.array_equal(delta, oc.delta) and inf == oc.infeasibility
        assert np.allclose(DE, oc.DE, rtol=1e-9, atol=1e-12) and np.isclose(logimp, oc.log_importance, rtol=1e-9, atol=1e-12)
        assert np.array_equal(ctr[c], oc.counters)


@pytest.mark.parametrize("prob,budget", [("pp64-16-poisson.abmp.xz", 150000), ("pp32-32.abmp.xz", 150000)])
def test_feasibility_search_on_a_budget_matches_oracle(prob, budget):
    """findInitialFeasibleSolution (SparseBasisSampler.h:254-263) from the empty trajectory, stopped after `budget` iterations:
    the chains are still infeasible (the C-ABI says so) and their states equal the oracle's after the same iterations."""
    from agentbasedmcmc_b200 import sampler as S
    from oracle.oracle import OracleChain, OracleProblem
    p = load_golden(prob)
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=2, seed=31, first_chain_id=5, find_feasible=False)
    with pytest.raises(S.AbmError, match="iteration limit"):
        smp.findInitialFeasibleSolution(max_iterations=budget)
    OP = OracleProblem(p)
    for c in range(2):
        oc = OracleChain(OP)
        oc.seed_philox(31, 5 + c)
        assert oc.find_feasible(budget) == budget and oc.infeasibility > 0
        X, delta, DE, tr, logimp, inf = smp.state(c)
        assert np.array_equal(X, oc.X) and np.array_equal(delta, oc.delta) and inf == oc.infeasibility
        assert np.allclose(DE, oc.DE, rtol=1e-9, atol=1e-12)
