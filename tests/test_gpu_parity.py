"""GPU parity: the CUDA sampler (through the C-ABI) against the oracle and the reference's golden traces.

Contract (BASELINE.json north_star): integer trajectory updates and accept decisions bit-exact under the same
uniform / proposal streams; doubles (currentDE, sum-tree, log importance) within 1e-9 relative -- they differ from
the host only through exp() (CUDA libdevice vs glibc, <= 1 ulp each)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _gpu():
    from agentbasedmcmc_b200 import sampler
    return sampler


def _close(a, b, rtol=RTOL, atol=1e-12):
    return np.allclose(a, b, rtol=rtol, atol=atol)


TREES = ["blocked", "reference"]   # ABM_TREE_BLOCKED (default, fast) and ABM_TREE_REFERENCE (node-for-node), abmcmc_b200.h


@pytest.mark.parametrize("tree", TREES)
@pytest.mark.parametrize("prob,trace,n_steps", [("pp8-4.abmp.xz", "trace_pp8-4_101.abmp.xz", 20000),
                                                ("pp8-4.abmp.xz", "trace_pp8-4_102.abmp.xz", 20000),
                                                ("pp16-8.abmp.xz", "trace_pp16-8_202.abmp.xz", 20000),
                                                ("pp8-4-poisson.abmp.xz", "trace_pp8-4-poisson_103.abmp.xz", 20000),
                                                # 3x3 grid, Poisson start, a tableau entry of magnitude 2 (Experiments.cpp:49-65)
                                                ("exp_pp3-2-single.abmp.xz", "trace_exp_pp3-2-single_301.abmp.xz", 20000),
                                                # the reference's experiments on its other agent types (Experiments.cpp:21-47,
                                                # 70-125), agents described by tables (abm_problem_desc::agent_*)
                                                ("toy_binomial.abmp.xz", "trace_toy_binomial_401.abmp.xz", 20000),
                                                ("toy_catmouse.abmp.xz", "trace_toy_catmouse_402.abmp.xz", 20000),
                                                ("toy_catmouse-assim.abmp.xz", "trace_toy_catmouse-assim_403.abmp.xz", 20000),
                                                # PredPrey re-described as a table agent: the same chains must come out
                                                ("pp8-4-astable.abmp.xz", "trace_pp8-4_101.abmp.xz", 20000),
                                                ("exp_pp3-2-single-astable.abmp.xz", "trace_exp_pp3-2-single_301.abmp.xz", 20000)])
def test_reference_golden_trace(prob, trace, n_steps, tree):
    """Same mt19937 uniform stream as the reference run -> same columns, decisions, infeasibilities and X, in both
    sum-tree storages.  The blocked tree holds exact leaves, the reference's get(u) is a difference of drifting sums
    (off by up to ~1e-9 absolute after 5e4 updates), hence the wider leaf tolerance there."""
    from oracle.oracle import mt_uniforms
    S = _gpu()
    p, t = load_golden(prob), load_golden(trace)
    n_init = int(t["init_iters"][0])
    u = mt_uniforms(int(t["seeds"][1]), n_init + 2 * n_steps)
    P = S.Problem(p, sum_tree=tree)
    leaf_atol = 1e-13 if tree == "reference" else 1e-11
    smp = S.SparseBasisSampler(P, t["init_state"], find_feasible=False)
    it = smp.findInitialFeasibleSolution(uniforms=u[:n_init + 8].copy() if n_init + 8 <= u.size else u)
    assert int(it[0]) == n_init
    X, delta, DE, tr, logimp, inf = smp.state(0)
    assert inf == 0
    assert np.array_equal(X, t["X0"])
    assert np.array_equal(delta, t["delta0"])
    assert _close(DE, t["DE0"])
    # get(u) is a difference of O(tree[0]) sums: compare with an absolute tolerance scaled by the root
    assert _close(S.leaf_probabilities(tr), t["leaf0"], atol=leaf_atol * tr[0])
    assert _close([tr[0], logimp], t["scal0"])
    j, acc, pinf = smp.step_traced(n_steps, uniforms=u[n_init:n_init + 2 * n_steps])
    assert np.array_equal(j[0], t["tr_j"][:n_steps])
    assert np.array_equal(acc[0], t["tr_acc"][:n_steps])
    assert np.array_equal(pinf[0], t["tr_inf"][:n_steps])
    X, delta, DE, tr, logimp, inf = smp.state(0)
    assert np.array_equal(X, t["XN"])
    assert np.array_equal(delta, t["deltaN"])
    assert _close(DE, t["DEN"])
    assert _close(S.leaf_probabilities(tr), t["leafN"], atol=leaf_atol * tr[0])
    assert _close([tr[0], logimp], t["scalN"])
    c = smp.counters()[0]
    assert np.array_equal(c, t["counters"])


@pytest.mark.parametrize("tree", TREES)
@pytest.mark.parametrize("prob,trace", [("pp8-4.abmp.xz", "trace_pp8-4_101.abmp.xz"), ("pp16-8.abmp.xz", "trace_pp16-8_202.abmp.xz"),
                                        ("pp8-4-poisson.abmp.xz", "trace_pp8-4-poisson_103.abmp.xz"),
                                        # table agents: end-state sums (no synopsis)
                                        ("toy_binomial.abmp.xz", "trace_toy_binomial_401.abmp.xz"),
                                        ("toy_catmouse.abmp.xz", "trace_toy_catmouse_402.abmp.xz"),
                                        ("toy_catmouse-assim.abmp.xz", "trace_toy_catmouse-assim_403.abmp.xz"),
                                        ("pp8-4-astable.abmp.xz", "trace_pp8-4_101.abmp.xz")])
def test_sample_statistics_match_reference(prob, trace, tree):
    """Row a16 pinned to the reference: a sampling launch fed the reference run's mt19937 stream must accumulate exactly the
    sums the reference's own statistics pipeline formed over the feasible states of that run -- ModelState(traj,T,T)
    (ModelState.h:27-36), synopsis() (FiguresForPaper.h:245-263), MeanAndVariance (diagnostics/MeanAndVariance.h:35-45) --
    which `abm_ref trace` recorded with the reference's code (smp_* arrays of the trace fixtures)."""
    from oracle.oracle import mt_uniforms
    S = _gpu()
    p, t = load_golden(prob), load_golden(trace)
    n_init, n = int(t["init_iters"][0]), t["tr_j"].size
    u = mt_uniforms(int(t["seeds"][1]), n_init + 2 * n)
    P = S.Problem(p, sum_tree=tree)
    smp = S.SparseBasisSampler(P, t["init_state"], find_feasible=False)
    smp.findInitialFeasibleSolution(uniforms=u[:n_init + 8].copy())
    # two launches: the incremental end state must carry over a launch boundary
    smp.sample_injected(n // 2, u[n_init:n_init + n])
    smp.sample_injected(n - n // 2, u[n_init + n:n_init + 2 * n])
    end, ss, sq, ns = smp.statistics()
    assert int(ns[0]) == int(t["smp_n"][0])
    assert np.array_equal(end, t["smp_endstate_sum"].astype(np.int64))
    if P.n_synopsis:   # (agents described by tables have end states only)
        assert np.array_equal(ss[0], t["smp_syn_sum"].astype(np.int64))
        assert np.array_equal(sq[0], t["smp_syn_sumsq"].astype(np.int64))
    assert np.array_equal(smp.state(0)[0], t["XN"])
    # the device-reduced moments carry the same numbers (abm_chains_get_moments)
    ex, re = smp.moments()
    d = ss.shape[1]
    assert ex[0] == 1 and ex[1] == ns[0] and np.array_equal(ex[2:10], smp.counters()[0])
    assert np.array_equal(ex[10:10 + d], ss[0]) and np.array_equal(ex[18:18 + d], sq[0]) and np.array_equal(ex[26:], end)
    mean = ss[0] * (1.0 / ns[0])
    var = (sq[0] - ss[0].astype(np.float64) * ss[0] * (1.0 / ns[0])) * (1.0 / (ns[0] - 1))
    assert np.allclose(re[0:d], mean, rtol=1e-14) and np.allclose(re[8:8 + d], mean * mean, rtol=1e-14)
    assert np.allclose(re[16:16 + d], var, rtol=1e-12)


@pytest.mark.parametrize("tree", TREES)
@pytest.mark.parametrize("prob,n_chains,n_steps", [("pp8-4.abmp.xz", 16, 4000), ("pp16-8.abmp.xz", 6, 3000),
                                                   ("pp16-8-poisson.abmp.xz", 4, 3000)])
def test_philox_free_running_matches_oracle(prob, n_chains, n_steps, tree):
    """Free-running mode: device Philox4x32-10 streams keyed by (seed, chain) == the oracle's; many chains per launch."""
    from oracle.oracle import OracleChain, OracleProblem
    S = _gpu()
    p = load_golden(prob)
    rng = np.random.default_rng(5)
    n_events = int(p["pp_meta"][2])
    init = (rng.random((n_chains, n_events)) < 0.02).astype(np.int8)
    seed, first = 0xABCDEF12345, 1000
    P = S.Problem(p, sum_tree=tree)
    smp = S.SparseBasisSampler(P, init, seed=seed, first_chain_id=first)
    OP = OracleProblem(p)
    j, acc, pinf = smp.step_traced(n_steps)
    ctr = smp.counters()
    for c in range(n_chains):
        oc = OracleChain(OP, init[c])
        oc.seed_philox(seed, first + c)
        assert oc.find_feasible() == int(smp.init_iterations[c])
        oj, oacc, oinf = oc.step(n_steps, trace=True)
        assert np.array_equal(j[c], oj) and np.array_equal(acc[c], oacc) and np.array_equal(pinf[c], oinf)
        X, delta, DE, tr, logimp, inf = smp.state(c)
        assert np.array_equal(X, oc.X) and np.array_equal(delta, oc.delta) and inf == oc.infeasibility
        assert _close(DE, oc.DE) and _close(tr, oc.tree, atol=1e-13 * tr[0] if tree == "reference" else 1e-11 * tr[0]) and _close(logimp, oc.log_importance)
        assert np.array_equal(ctr[c], oc.counters)


def test_injected_proposals_replay():
    """replay mode: column stream and accept uniforms injected."""
    from oracle.oracle import OracleChain, OracleProblem
    S = _gpu()
    p = load_golden("pp8-4.abmp.xz")
    rng = np.random.default_rng(9)
    n_steps, n_chains = 3000, 4
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=3)
    OP = OracleProblem(p)
    props = rng.integers(0, P.n_cols, size=(n_chains, n_steps), dtype=np.int32)
    u = rng.random((n_chains, 2 * n_steps))
    j, acc, pinf = smp.step_traced(n_steps, uniforms=u, proposals=props)
    assert np.array_equal(j, props)
    for c in range(n_chains):
        oc = OracleChain(OP)
        oc.seed_philox(3, c)
        oc.find_feasible()
        oc.set_injected(u[c], proposals=props[c])
        oj, oacc, oinf = oc.step(n_steps, trace=True)
        assert np.array_equal(acc[c], oacc) and np.array_equal(pinf[c], oinf)
        assert np.array_equal(smp.state(c)[0], oc.X)


def test_async_step_matches_traced_step():
    """abm_chains_step (no trace, async) advances the chains exactly like the traced entry point."""
    S = _gpu()
    p = load_golden("pp8-4.abmp.xz")
    P = S.Problem(p)
    a = S.SparseBasisSampler(P, n_chains=8, seed=77)
    b = S.SparseBasisSampler(P, n_chains=8, seed=77)
    a.step(1500)
    a.step(500)
    a.synchronize()
    b.step_traced(2000)
    for c in range(8):
        sa, sb = a.state(c), b.state(c)
        assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
        assert np.array_equal(sa[2], sb[2]) and np.array_equal(sa[3], sb[3]) and sa[4] == sb[4]
    assert np.array_equal(a.counters(), b.counters())
    assert a.stats.nSamples() == 8 * 2000


def test_sample_statistics_match_oracle():
    """abm_chains_sample: every feasible state is a sample (SparseBasisSampler.h:247); end-state occupancy sums and
    synopsis sums / sums of squares (FiguresForPaper.h:93-104,245-263) are exact integers == the oracle's."""
    from oracle.oracle import OracleChain, OracleProblem
    S = _gpu()
    p = load_golden("pp8-4.abmp.xz")
    n_chains, seed = 12, 991
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=seed)
    smp.record_series(4, 3, 400)
    smp.step(300)                      # burn-in: no statistics
    smp.sample(700)
    smp.sample(500, max_steps=450)     # second launch, bounded by steps
    end, ss, sq, ns = smp.statistics()
    series = smp.series()
    OP = OracleProblem(p)
    nsyn = P.n_synopsis
    assert nsyn == 2
    o_end = np.zeros(128, np.int64)
    for c in range(n_chains):
        oc = OracleChain(OP)
        oc.seed_philox(seed, c)
        oc.find_feasible()
        oc.step(300)
        s1, s2 = np.zeros(8, np.int64), np.zeros(8, np.int64)
        ser = np.zeros((400, 8), np.int32)
        n1 = oc.sample_stats(700, 10**9, o_end, s1, s2, 0, ser, 3)
        n2 = oc.sample_stats(500, 450, o_end, s1, s2, n1, ser, 3)
        assert ns[c] == n1 + n2
        assert np.array_equal(ss[c], s1[:nsyn]) and np.array_equal(sq[c], s2[:nsyn])
        assert np.array_equal(smp.state(c)[0], oc.X)
        if c < 4:
            assert np.array_equal(series[c], ser[:, :nsyn])
    assert np.array_equal(end, o_end)
    # the non-blocking read-back returns the same numbers (and the counters) while a later launch is already queued
    cnt = smp.counters()
    smp.enqueue_statistics()
    smp.step(50)
    end2, ss2, sq2, ns2, cnt2 = smp.wait_statistics()
    assert np.array_equal(end2, end) and np.array_equal(ss2, ss) and np.array_equal(sq2, sq) and np.array_equal(ns2, ns)
    assert np.array_equal(cnt2, cnt)
    with pytest.raises(S.AbmError):
        smp.wait_statistics()          # nothing pending any more
    smp.reset_statistics()
    end, ss, sq, ns = smp.statistics()
    assert end.sum() == 0 and ss.sum() == 0 and ns.sum() == 0


def test_posterior_marginals_match_reference_long_run():
    """Statistical parity (BASELINE.json north_star): end-state marginals of 2048 GPU chains against the reference's own
    long CPU run (64 chains x 60000 feasible samples after 60000 burn-in, tests/golden/marginals_pp8-4, made by
    oracle/ref/abm_ref.cpp `marginals`).  Tolerance: 5 standard errors of the reference estimate (from the spread of
    its 64 chains; two such reference runs differ from each other by up to 4 of these standard errors because a few
    states mix slowly) plus 0.004 absolute; the expected number of agents to 5 %."""
    from agentbasedmcmc_b200 import ensemble as E
    S = _gpu()
    p = load_golden("pp8-4.abmp.xz")
    ref = load_golden("marginals_pp8-4.abmp.xz")
    ref_mean = ref["endstate_sum"] / ref["n_samples"][0]
    per_chain = ref["perchain_endstate_mean"].reshape(64, -1)
    se = per_chain.std(axis=0, ddof=1) / np.sqrt(64.0)
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=2048, seed=20240)
    smp.step(80000)     # the first feasible state holds ~80 agents; the chain needs ~5e4 steps to relax to ~9
    smp.sample(1500)
    summ = E.sampler_summary(smp)
    assert summ.n_samples == 2048 * 1500
    diff = np.abs(summ.end_state_mean - ref_mean)
    assert np.all(diff <= 5.0 * se + 0.004), f"worst state off by {diff.max():.4f} (se {se[diff.argmax()]:.4f})"
    assert abs(summ.end_state_mean.sum() / ref_mean.sum() - 1.0) < 0.05
    # acceptance / infeasibility statistics of the chain itself (83.4 % / 1.7 % at 8x8x4)
    c = summ.counters
    acc = c[4:].sum() / c[:4].sum()
    assert abs(acc - 0.834) < 0.02
    assert np.all(np.isfinite(summ.rhat)) and np.all(summ.rhat > 0.99)   # 1500 correlated samples per chain: R-hat is well above 1


@pytest.mark.parametrize("tree", TREES)
def test_reference_validation_experiment(tree):
    """Experiments::PredPreySingleObservation (Experiments.cpp:49-65 through doSingleObservationExperiment, :129-168) on the
    GPU: 3x3 grid, 2 timesteps, Poisson start, one observation, kappa = 3, chains started from the empty trajectory as
    SparseBasisSampler(posterior, kappa) does (:199-206).  The experiment compares the sampler's end-state means with the
    RejectionSampler's (the reference's own run: RMS error 0.0036 with 500000 samples).  Here 2048 chains x 1000 samples
    after 2000 steps of burn-in against the reference's 2000000 rejection samples: RMS error below 0.003, every state
    within 0.008.  The problem has a tableau entry of magnitude 2 and a grid that is not a power of two."""
    from agentbasedmcmc_b200 import ensemble as E
    S = _gpu()
    p, m = load_golden("exp_pp3-2-single.abmp.xz"), load_golden("exp_pp3-2-single_means.abmp.xz")
    assert np.abs(p["nb_val"]).max() == 2
    rej = m["rejection_endstate_mean"]
    P = S.Problem(p, sum_tree=tree)
    smp = S.SparseBasisSampler(P, n_chains=2048, seed=77)
    smp.step(2000)
    smp.sample(1000)
    summ = E.sampler_summary(smp)
    assert summ.n_samples == 2048 * 1000
    err = summ.end_state_mean - rej
    assert np.sqrt((err ** 2).mean()) < 0.003, f"RMS error {np.sqrt((err ** 2).mean()):.5f}"
    assert np.abs(err).max() < 0.008
    assert smp.stats.nInfeasibleSamples() < 0.6 * smp.stats.nSamples()


@pytest.mark.parametrize("name", ["pp8-4", "pp8-4-poisson"])
def test_device_prior_initial_states_match_reference_prior(name):
    """abm_chains_init_from_prior draws `prior.nextSample(false)` (Forecast.h:102-150) per chain on the device, from the
    Bernoulli or the Poisson start state.  Checked statistically against 20000 draws of the reference
    (tests/golden/prior_*, `abm_ref init`): for every (timestep, species, act) the probability that an event is set,
    within 5 binomial standard errors + 2e-4; for the Poisson start state also the share of start states that hold two
    agents with different acts (impossible under the Bernoulli one)."""
    S = _gpu()
    p, ref = load_golden(f"{name}.abmp.xz"), load_golden(f"prior_{name}.abmp.xz")
    n_chains = 8192
    P = S.Problem(p)
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=13, prior_init=True, return_prior_sample=True)
    ev = smp.prior_sample.reshape(n_chains, 4, 2, 64, 6)
    assert set(np.unique(ev)) <= {0, 1}
    got = ev.mean(axis=(0, 3)).ravel()
    want = ref["mean_indicator"]
    n_ref, n_gpu = 20000 * 64, n_chains * 64
    se = np.sqrt(want * (1 - want) * (1.0 / n_ref + 1.0 / n_gpu))
    assert np.all(np.abs(got - want) <= 5 * se + 2e-4), (got - want) / np.maximum(se, 1e-12)
    multi = (ev[:, 0].sum(axis=-1) >= 2).mean()
    if name.endswith("poisson"):
        want_multi = float(ref["multi_act_t0"][0])
        assert abs(multi - want_multi) <= 5 * np.sqrt(want_multi * (1.0 / (20000 * 128) + 1.0 / (n_chains * 128))) + 1e-5
    else:
        assert multi == 0.0
    # chains start from different trajectories and all reach feasibility
    assert len({ev[c].tobytes() for c in range(64)}) > 60
    assert np.all(smp.infeasibility() == 0)
    # delta reflects the drawn trajectory (SparseBasisSampler.h:147) before the feasibility search moves it: re-init only
    smp2 = S.SparseBasisSampler(P, n_chains=4, seed=13, prior_init=True, return_prior_sample=True, find_feasible=False)
    nb_col = p["nb_col"]
    for c in range(4):
        delta = smp2.state(c)[1]
        assert np.array_equal(delta == -1, smp2.prior_sample[c][nb_col] > 0)


def test_blocked_tree_tracks_reference_tree_and_its_sums_stay_exact():
    """Short version of scripts/tree_soak.py (1e7 steps: profiles/r2_tree_soak.json): the same chains on the same Philox
    streams in both storages of basisDistribution draw the same columns and take the same decisions; the blocked tree's
    incrementally kept sums stay within rounding of the exact sums of their leaves, and an explicit rebuild makes the sums of
    32 leaves exact."""
    S = _gpu()
    p = load_golden("pp16-8.abmp.xz")
    smp = {t: S.SparseBasisSampler(S.Problem(p, sum_tree=t), n_chains=8, seed=99, first_chain_id=3) for t in TREES}
    for _ in range(4):
        tr = {t: smp[t].step_traced(50000) for t in TREES}
        assert np.array_equal(tr["blocked"][0], tr["reference"][0]) and np.array_equal(tr["blocked"][1], tr["reference"][1])
    for c in range(8):
        assert np.array_equal(smp["blocked"].state(c)[0], smp["reference"].state(c)[0])
    d = smp["blocked"].tree_drift(0)
    assert d[0] > 0 and np.all(d[1:] <= 1e-11 * d[0])
    smp["blocked"].rebuild_tree()
    d2 = smp["blocked"].tree_drift(0)
    assert d2[1] <= 32 * 2.3e-16 * 32      # sums of 32 leaves (each <= 1) recomputed: ordinary summation rounding only


def test_device_variogram_and_ensemble_ess_match_the_host_route():
    """abm_chains_series_variogram (exact integer sums over all chains, on the device) against numpy on the same recorded
    series, and the effective sample size built from it and from the device-reduced moments against
    diagnostics.ensemble_effective_samples on the per-chain tables (rows a17 / f3)."""
    from agentbasedmcmc_b200 import diagnostics as D
    from agentbasedmcmc_b200 import ensemble as E
    S = _gpu()
    p = load_golden("pp8-4.abmp.xz")
    P = S.Problem(p)
    m, n_s, thin = 96, 6000, 25
    smp = S.SparseBasisSampler(P, n_chains=m, seed=5150)
    smp.step(20000)
    smp.reset_statistics()
    k = n_s // thin
    smp.record_series(m, thin, k + 1)
    smp.sample(n_s, max_steps=40 * n_s)
    end, ss, sq, ns = smp.statistics()
    assert np.all(ns == n_s)
    x = smp.series()[:, :k].astype(np.int64)
    n_lags = k // 2 + 1
    vs, mc = smp.series_variogram(k, n_lags, n_s)
    assert mc == m
    for j in (0, 1, 7, n_lags - 1):
        d = x[:, : k - j] - x[:, j:]
        assert np.array_equal(vs[j], (d * d).sum(axis=(0, 1)))
    ex, re = smp.moments()
    summ = E.summarize_moments(ex, re, P.n_synopsis)
    V = vs / (mc * (k - np.arange(n_lags))[:, None])
    n_eff, _ = D.effective_samples_from_variogram(V, summ.var_plus, n_s, mc, thin)
    r = D.ensemble_effective_samples(x.astype(np.float64), thin, ns, ss, sq)
    assert np.allclose(n_eff, r["n_eff"], rtol=1e-9) and np.allclose(summ.rhat, r["rhat"], rtol=1e-9)
