"""Agents other than PredPrey on the GPU: TrajectoryImportance<AGENT> described by tables
(abm_problem_desc::agent_*, filled by include/abmcmc_b200/reference_flatten.hpp from the agent's own timestep / marginalTimestep /
neighbours / consequences).  Fixtures: the reference's validation experiments on BinomialAgent<3> and CatMouseAgent
(Experiments.cpp:21-47, 70-125; `abm_ref toy`), with the answers those experiments compare against.

The bit-level parity of these problems (reference golden traces, sampling statistics) runs with the PredPrey cases in
tests/test_gpu_parity.py; the sample-for-sample C++ drop-in of the experiments in tests/test_gpu_dropin.py.  Here: free-running
device streams against the oracle, and the posterior itself against the exhaustive solver."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

TOYS = ["binomial", "catmouse", "catmouse-assim"]


def _gpu():
    from agentbasedmcmc_b200 import sampler
    return sampler


@pytest.mark.parametrize("tree", ["blocked", "reference"])
@pytest.mark.parametrize("name", TOYS + ["pp8-4-astable"])
def test_table_agent_free_running_matches_oracle(name, tree):
    """Device Philox streams keyed by (seed, chain): every chain equals the oracle's chain on the same stream -- columns, accept
    decisions, infeasibilities, X, delta, log importance, end state."""
    from oracle.oracle import OracleChain, OracleProblem
    S = _gpu()
    p = load_golden(f"toy_{name}.abmp.xz" if name in TOYS else f"{name}.abmp.xz")
    n_chains, n_steps, seed, first = 12, 3000, 0x5EED1234, 50
    P = S.Problem(p, sum_tree=tree)
    assert P.grid_size == 0 and P.n_agent_states == int(p["ag_meta"][0])
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=seed, first_chain_id=first)
    OP = OracleProblem(p)
    j, acc, pinf = smp.step_traced(n_steps)
    ctr = smp.counters()
    for c in range(n_chains):
        oc = OracleChain(OP)
        oc.seed_philox(seed, first + c)
        assert oc.find_feasible() == int(smp.init_iterations[c])
        oj, oacc, oinf = oc.step(n_steps, trace=True)
        assert np.array_equal(j[c], oj) and np.array_equal(acc[c], oacc) and np.array_equal(pinf[c], oinf)
        X, delta, DE, tr, logimp, inf = smp.state(c)
        assert np.array_equal(X, oc.X) and np.array_equal(delta, oc.delta) and inf == oc.infeasibility
        assert np.allclose(DE, oc.DE, rtol=1e-9, atol=1e-12) and np.isclose(logimp, oc.log_importance, rtol=1e-9, atol=1e-12)
        assert np.array_equal(ctr[c], oc.counters)


@pytest.mark.parametrize("name", TOYS)
def test_table_agent_posterior_matches_exact_solver(name):
    """What the reference's experiments check (Experiments.cpp:29-32, 113-124): the sampler's answer against an independent
    solver.  2048 device chains x 2000 samples: the end-state means against ExactSolver's (closed form 0.525 / 0.525 / 0.05 for the
    binomial experiment, Experiments.cpp:32) and against the reference RejectionSampler's; and one draw per chain -- the state
    each chain stopped in -- against the exact posterior over whole trajectories."""
    S = _gpu()
    p = load_golden(f"toy_{name}.abmp.xz")
    n_chains, per_chain = 2048, 2000
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=2024)
    smp.step(2000)
    smp.sample(1)   # every chain at a feasible state before the statistics start
    smp.reset_statistics()
    smp.sample(per_chain)
    end, _, _, ns = smp.statistics()
    assert np.all(ns == per_chain)
    mean = end / float(ns.sum())
    assert np.abs(mean - p["exact_endstate_mean"]).max() < 2.5e-3
    assert np.abs(mean - p["rejection_endstate_mean"]).max() < 4e-3
    if name == "binomial":
        assert np.allclose(p["exact_endstate_mean"], [0.525, 0.525, 0.05], atol=1e-12)
    # whole trajectories: chains are independent, so their final states are 2048 independent draws
    n_events = P.n_events
    support = p["exact_traj"].reshape(-1, n_events)
    prob = p["exact_p"]
    traj = smp.trajectories()
    assert np.all(smp.infeasibility() == 0)
    index = {tuple(row.tolist()): k for k, row in enumerate(support)}
    counts = np.zeros(prob.size)
    for row in traj:
        assert tuple(row.tolist()) in index, "a chain stopped outside the exact support"
        counts[index[tuple(row.tolist())]] += 1
    freq = counts / n_chains
    sigma = np.sqrt(prob * (1 - prob) / n_chains)
    assert np.all(np.abs(freq - prob) < 5 * sigma + 2e-3)


def _random_table_agent(p, rng, n_acts):
    """The tableau of a PredPrey problem with a RANDOM agent description laid over its events: S' x n_acts acts per timestep
    (same event count), random classes, 0..4 influencers per state, neighbours() = the states a state influences plus random
    extras in random order, random consequences, random log-ratio rows.  Not a model of anything -- a differential input for
    the table path (oracle/sbs_oracle.c against the kernels)."""
    q = {k: v for k, v in p.items() if k not in ("pp_meta", "pp_actpmf", "pp_params")}
    G, T, n_events = (int(v) for v in p["pp_meta"][:3])
    S = n_events // (T * n_acts)
    assert S * T * n_acts == n_events
    n_cls = 5
    infl = np.full((S, 4), -1, np.int32)
    nb = [[] for _ in range(S)]
    for s in range(S):
        k = int(rng.integers(0, 5))
        cand = rng.choice(S, size=k, replace=False) if k else []
        infl[s, :k] = cand
        for c in cand:
            nb[int(c)].append(s)
    for s in range(S):
        extra = rng.choice(S, size=int(rng.integers(0, 3)), replace=False).tolist()
        lst = list(dict.fromkeys(nb[s] + extra))
        rng.shuffle(lst)
        nb[s] = lst
    cq = [rng.choice(S, size=int(rng.integers(0, 3)), replace=False).tolist() for _ in range(S * n_acts)]
    q["ag_meta"] = np.array([S, n_acts, T, n_cls], np.int32)
    q["ag_class"] = rng.integers(0, n_cls, S).astype(np.int32)
    q["ag_infl"] = infl.ravel()
    q["ag_nb_ptr"] = np.concatenate([[0], np.cumsum([len(v) for v in nb])]).astype(np.int32)
    q["ag_nb"] = np.array([x for v in nb for x in v], np.int32)
    q["ag_cq_ptr"] = np.concatenate([[0], np.cumsum([len(v) for v in cq])]).astype(np.int32)
    q["ag_cq"] = np.array([x for v in cq for x in v], np.int32)
    q["imp_lgamma"] = np.array([0.0, 0.0] + [float(np.log(np.arange(1, k + 1)).sum()) for k in range(2, 7)])
    q["imp_logratio"] = rng.normal(0.0, 0.4, n_cls * 2 * n_acts) * (rng.random(n_cls * 2 * n_acts) < 0.8)
    return q


@pytest.mark.parametrize("tree", ["blocked", "reference"])
@pytest.mark.parametrize("n_acts,seed", [(6, 1), (4, 2), (3, 3), (2, 4), (1, 5)])
def test_random_table_agents_match_oracle(n_acts, seed, tree):
    """Differential run of the table path on inputs no agent of the reference produces: 1..6 acts, 0..4 influencers, several
    classes, non-trivial log-ratios, asymmetric neighbour lists.  Device chains against the oracle's chains on the same Philox
    streams (decisions, states, log importance), then the sampling statistics (end-state sums over the feasible states)."""
    from oracle.oracle import OracleChain, OracleProblem
    S = _gpu()
    rng = np.random.default_rng(100 + seed)
    p = _random_table_agent(load_golden("pp8-4.abmp.xz"), rng, n_acts)
    n_chains, n_steps, key = 6, 2500, 0xA6E27 + seed
    P = S.Problem(p, sum_tree=tree)
    init = (rng.random((n_chains, P.n_events)) < 0.02).astype(np.int8)
    smp = S.SparseBasisSampler(P, init, seed=key)
    OP = OracleProblem(p)
    j, acc, pinf = smp.step_traced(n_steps)
    chains = []
    for c in range(n_chains):
        oc = OracleChain(OP, init[c])
        oc.seed_philox(key, c)
        assert oc.find_feasible() == int(smp.init_iterations[c])
        oj, oacc, oinf = oc.step(n_steps, trace=True)
        assert np.array_equal(j[c], oj) and np.array_equal(acc[c], oacc) and np.array_equal(pinf[c], oinf)
        X, delta, DE, tr, logimp, inf = smp.state(c)
        assert np.array_equal(X, oc.X) and np.array_equal(delta, oc.delta) and inf == oc.infeasibility
        assert np.isclose(logimp, oc.log_importance, rtol=1e-9, atol=1e-10)
        chains.append(oc)
    # statistics: 300 further feasible samples per chain
    smp.sample(300)
    end, _, _, ns = smp.statistics()
    ref_end = np.zeros(P.n_agent_states, np.int64)
    for c, oc in enumerate(chains):
        s1, s2 = np.zeros(8, np.int64), np.zeros(8, np.int64)
        assert oc.sample_stats(300, 10 ** 9, ref_end, s1, s2) == 300 == int(ns[c])
        assert np.array_equal(smp.state(c)[0], oc.X)
    assert np.array_equal(end, ref_end)
