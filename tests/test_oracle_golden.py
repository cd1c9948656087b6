"""Pins oracle/sbs_oracle.c to the reference: golden step traces emitted by the UNMODIFIED reference
SparseBasisSampler (oracle/ref/abm_ref.cpp `trace`, fixtures made by tests/golden/make_golden.py).
Everything is compared bitwise -- integers, accept decisions and doubles."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.oracle import OracleChain, OracleProblem, mt_uniforms

TRACES = [("pp8-4.abmp.xz", "trace_pp8-4_101.abmp.xz"), ("pp8-4.abmp.xz", "trace_pp8-4_102.abmp.xz"),
          ("pp16-8.abmp.xz", "trace_pp16-8_202.abmp.xz"),
          # Poisson start state (PoissonStartState.h:18-31: unbounded occupation rows, lgamma factors), SURVEY 8(f)4
          ("pp8-4-poisson.abmp.xz", "trace_pp8-4-poisson_103.abmp.xz"),
          # the problem of the reference's own validation experiment (Experiments::PredPreySingleObservation, Experiments.cpp:49-65):
          # 3x3 grid (not a power of two), Poisson start, a tableau entry of magnitude 2
          ("exp_pp3-2-single.abmp.xz", "trace_exp_pp3-2-single_301.abmp.xz"),
          # the reference's validation experiments on its other agent types (Experiments.cpp:21-47, 70-125: BinomialAgent<3> with a
          # noisy observation, CatMouseAgent with one observation, CatMouseAgent assimilation), agents described by tables
          ("toy_binomial.abmp.xz", "trace_toy_binomial_401.abmp.xz"), ("toy_catmouse.abmp.xz", "trace_toy_catmouse_402.abmp.xz"),
          ("toy_catmouse-assim.abmp.xz", "trace_toy_catmouse-assim_403.abmp.xz"),
          # own test agent with non-trivial importance factors (oracle/ref/test_agents.h RingAgent<5>, T = 3)
          ("toy_ring.abmp.xz", "trace_toy_ring_404.abmp.xz"),
          # PredPrey re-described as a table agent (tables found by probing PredPreyAgent::timestep): the same traces must come out
          ("pp8-4-astable.abmp.xz", "trace_pp8-4_101.abmp.xz"), ("exp_pp3-2-single-astable.abmp.xz", "trace_exp_pp3-2-single_301.abmp.xz")]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def hash_X(X):
    # sum_i X_i * splitmix64(i) mod 2^64 (abm_ref.cpp hashX)
    i = np.arange(X.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = i + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int((X.astype(np.int64).astype(np.uint64) * z).sum(dtype=np.uint64))


@pytest.mark.parametrize("prob,trace", TRACES)
def test_oracle_matches_reference_trace(prob, trace):
    check_trace(load_golden(prob), load_golden(trace))


@pytest.mark.parametrize("grid,T,seed,start", [(3, 2, 21, "poisson"), (4, 3, 22, "bernoulli"), (8, 3, 23, "poisson"), (8, 5, 24, "bernoulli"),
                                               (16, 3, 25, "bernoulli")])
def test_oracle_matches_reference_on_fresh_problems(grid, T, seed, start, tmp_path):
    """Differential run on problems that are not fixtures: `abm_ref gen` (the reference's PredPreyProblem constructor) and
    `abm_ref trace` (the unmodified sampler, step by step) with other seeds, grids, horizons and both start states; the
    oracle must reproduce the trace bit for bit, doubles included."""
    import subprocess
    from pathlib import Path
    from agentbasedmcmc_b200.abmp import read_abmp
    ref = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "abm_ref"
    if not ref.exists():
        pytest.skip("oracle/_ref/abm_ref not built (needs /root/reference at build time)")
    prob, trace = tmp_path / "p.abmp", tmp_path / "t.abmp"
    dens = "0.1" if grid <= 4 else "0.05"
    subprocess.run([str(ref), "gen", str(grid), str(T), str(seed), "10.0", str(prob), dens, dens, "0.1", "1.0", start], check=True,
                   capture_output=True, timeout=600)
    subprocess.run([str(ref), "trace", str(prob), str(seed + 100), str(seed + 200), "4000", "500", str(trace)], check=True,
                   capture_output=True, timeout=600)
    check_trace(read_abmp(prob), read_abmp(trace))


@pytest.mark.parametrize("kind,seed", [("binomial", 31), ("catmouse", 32), ("catmouse-assim", 33), ("catmouse-assim", 34),
                                       # own test agent (oracle/ref/test_agents.h): two classes, every log-ratio non-zero
                                       ("ring", 35), ("ring", 36)])
def test_oracle_matches_reference_on_fresh_agent_experiments(kind, seed, tmp_path):
    """The same differential run for the agents described by tables: `abm_ref toy` builds the reference's validation
    experiments on BinomialAgent / CatMouseAgent (Experiments.cpp:21-47, 70-125) under another seed (another real trajectory
    and observation set for the assimilation experiment), flattens the agent by probing its timestep(), and traces the
    unmodified sampler; the oracle must reproduce the trace bit for bit from the tables alone."""
    import subprocess
    from pathlib import Path
    from agentbasedmcmc_b200.abmp import read_abmp
    ref = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "abm_ref"
    if not ref.exists():
        pytest.skip("oracle/_ref/abm_ref not built (needs /root/reference at build time)")
    prob, trace = tmp_path / "p.abmp", tmp_path / "t.abmp"
    subprocess.run([str(ref), "toy", kind, str(seed), str(seed + 100), "6000", "500", "0", str(prob), str(trace)], check=True,
                   capture_output=True, timeout=600)
    p = read_abmp(prob)
    assert "ag_meta" in p and "pp_meta" not in p
    check_trace(p, read_abmp(trace))


def test_agent_outside_the_table_form_is_refused(tmp_path):
    """flattenAgentImportance verifies the form it tabulates by probing timestep(); an agent whose act pmf depends on HOW MANY
    agents sit on another state (oracle/ref/test_agents.h CrowdedRingAgent) must be refused, not approximated."""
    import subprocess
    from pathlib import Path
    ref = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "abm_ref"
    if not ref.exists():
        pytest.skip("oracle/_ref/abm_ref not built (needs /root/reference at build time)")
    out = subprocess.run([str(ref), "toy", "crowded-ring", "1", "2", "100", "50", "0", str(tmp_path / "p.abmp"), str(tmp_path / "t.abmp")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "depends on more than whether any influencer state is occupied" in out.stderr
    assert not (tmp_path / "p.abmp").exists()


def check_trace(p, t):
    P = OracleProblem(p)
    c = OracleChain(P, t["init_state"])
    c.seed_mt(int(t["seeds"][1]))
    # constructor + findInitialFeasibleSolution (SparseBasisSampler.h:107-185, 254-263)
    assert c.find_feasible() == int(t["init_iters"][0])
    assert np.array_equal(c.X, t["X0"])
    assert np.array_equal(c.delta, t["delta0"])
    assert np.array_equal(bits(c.DE), bits(t["DE0"]))
    assert np.array_equal(bits(c.leaf), bits(t["leaf0"]))
    assert bits([c.sum, c.log_importance]).tolist() == bits(t["scal0"]).tolist()
    assert c.infeasibility == 0
    # Metropolis steps, checkpoint by checkpoint (SparseBasisSampler.h:219-249)
    every = int(t["ck_every"][0])
    n = t["tr_j"].size
    ck_hash = t["ck_hash"].view(np.uint32).astype(np.uint64)
    ck_hash = ck_hash[0::2] | (ck_hash[1::2] << np.uint64(32))
    for k in range(n // every):
        j, acc, inf = c.step(every, trace=True)
        sl = slice(k * every, (k + 1) * every)
        assert np.array_equal(j, t["tr_j"][sl])
        assert np.array_equal(acc, t["tr_acc"][sl])
        assert np.array_equal(inf, t["tr_inf"][sl])
        assert c.hash_X == int(ck_hash[k]) == hash_X(c.X)
        assert bits([c.log_importance])[0] == bits(t["ck_logimp"])[k]
        assert bits([c.sum])[0] == bits(t["ck_sum"])[k]
        assert c.infeasibility == int(t["ck_inf"][k])
    assert np.array_equal(c.X, t["XN"])
    assert np.array_equal(c.delta, t["deltaN"])
    assert np.array_equal(bits(c.DE), bits(t["DEN"]))
    assert np.array_equal(bits(c.leaf), bits(t["leafN"]))
    assert np.array_equal(c.counters, t["counters"])
    check_statistics(p, t)


def reference_synopsis(end, G):
    """FiguresForPaper::synopsis (FiguresForPaper.h:245-263) restated for the checks below: agents of both species inside
    nested squares of side G/2, G/4, ... > 1 along the diagonal.  end: ModelState, id = x + G y + G^2 type (PredPreyAgent.h:59)."""
    grid = np.asarray(end).reshape(2, G, G).sum(axis=0)   # [y][x]
    out, origin, part = [], 0, G // 2
    while part > 1:
        out.append(int(grid[origin:origin + part, origin:origin + part].sum()))
        origin += part
        part //= 2
    return out


def check_statistics(p, t):
    """Row a16: the per-sample statistics pinned to what the REFERENCE computed.  The trace carries, from the reference's own
    ModelState(X,T,T) (ModelState.h:27-36) and synopsis() (FiguresForPaper.h:245-263): the end state and synopsis at every
    checkpoint, and the sums its statistics pipeline forms over every feasible state of the run (FiguresForPaper.h:93-104).
    The oracle must reproduce all of them exactly."""
    if "ck_endstate" not in t or not ("pp_meta" in p or "ag_meta" in p):
        return
    every, n = int(t["ck_every"][0]), t["tr_j"].size
    if "pp_meta" in p:
        G = int(p["pp_meta"][0])
        S = 2 * G * G
        n_syn = len(reference_synopsis(np.zeros(S, np.int64), G))
    else:   # an agent described by tables: end states only, no synopsis
        G, S, n_syn = 0, int(p["ag_meta"][0]), 0
    ck_end = t["ck_endstate"].reshape(-1, S)
    ck_syn = t["ck_synopsis"].reshape(ck_end.shape[0], n_syn) if G else None
    P = OracleProblem(p)
    c = OracleChain(P, t["init_state"])
    c.seed_mt(int(t["seeds"][1]))
    c.find_feasible()
    for k in range(n // every):
        c.step(every)
        assert np.array_equal(c.end_state, ck_end[k])
        if G:
            assert reference_synopsis(ck_end[k], G) == ck_syn[k].tolist()   # the restated synopsis against the reference's
    # the running sums, through the oracle's own statistics loop (orc_chain_sample_stats)
    c = OracleChain(P, t["init_state"])
    c.seed_mt(int(t["seeds"][1]))
    c.find_feasible()
    end, s1, s2 = np.zeros(S, np.int64), np.zeros(8, np.int64), np.zeros(8, np.int64)
    taken = c.sample_stats(10 ** 12, n, end, s1, s2)
    assert taken == int(t["smp_n"][0])
    assert np.array_equal(end, t["smp_endstate_sum"].astype(np.int64))
    if G:
        assert np.array_equal(s1[:n_syn], t["smp_syn_sum"].astype(np.int64))
        assert np.array_equal(s2[:n_syn], t["smp_syn_sumsq"].astype(np.int64))


def test_injected_uniform_stream_equals_mt_stream():
    """Feeding the oracle the mt19937/generate_canonical doubles as an injected stream reproduces the
    trace: this is the route by which the GPU kernel is fed the reference's uniforms."""
    p, t = load_golden("pp8-4.abmp.xz"), load_golden("trace_pp8-4_101.abmp.xz")
    P = OracleProblem(p)
    c = OracleChain(P, t["init_state"])
    n_init, n = int(t["init_iters"][0]), 5000
    u = mt_uniforms(int(t["seeds"][1]), n_init + 2 * n)
    c.set_injected(u)
    assert c.find_feasible() == n_init
    j, acc, inf = c.step(n, trace=True)
    assert np.array_equal(j, t["tr_j"][:n]) and np.array_equal(acc, t["tr_acc"][:n]) and np.array_equal(inf, t["tr_inf"][:n])


def test_problem_structure_matches_survey():
    """Sizes quoted in SURVEY.md section 8 for the seed-1234 instances."""
    for name, rows, cols, nnz in (("pp8-4.abmp.xz", 3456, 2658, 13738), ("pp16-8.abmp.xz", 27136, 20788, 112552)):
        p = load_golden(name)
        P = OracleProblem(p)
        assert (P.n_rows, P.n_cols, int(p["dims"][4])) == (rows, cols, nnz)
        assert np.abs(p["nb_val"]).max() == 1  # PredPrey basis entries are all +-1
        # sampler row i == event id i for i < nEvents: no equality among the first nEvents rows
        ne = int(p["pp_meta"][2])
        assert np.all(p["tab_L"][:ne] != p["tab_U"][:ne])


def test_end_state_occupancy_counts_agents():
    p, t = load_golden("pp8-4.abmp.xz"), load_golden("trace_pp8-4_101.abmp.xz")
    c = OracleChain(OracleProblem(p), t["init_state"])
    c.seed_mt(5)
    c.find_feasible()
    c.step(200)
    while c.infeasibility:
        c.step(1)
    X = c.X
    A, S, T = 6, 128, 4
    last = X[(T - 1) * S * A:T * S * A].reshape(S, A)
    # every non-DIE act yields one agent at T, GIVEBIRTH two (PredPreyAgent.h:98-115)
    assert c.end_state.sum() == last[:, 1:].sum() + last[:, 5].sum()


def test_oracle_reproduces_reference_validation_experiment():
    """Experiments::PredPreySingleObservation (Experiments.cpp:49-65, 129-168), the reference's own check of this path: the
    sampler's end-state means against the RejectionSampler's.  Fixture: the reference's two estimates (500000 sampler
    samples, RMS error 0.0036 between them; 2000000 rejection samples).  Here: 8 oracle chains from the empty trajectory
    (SparseBasisSampler.h:199-206), 2000 samples burn-in, 100000 samples each; RMS error to the rejection means below
    0.006 (measured 0.0021) and every state within 0.012."""
    p, m = load_golden("exp_pp3-2-single.abmp.xz"), load_golden("exp_pp3-2-single_means.abmp.xz")
    rej = m["rejection_endstate_mean"]
    assert np.sqrt(((m["mcmc_endstate_mean"] - rej) ** 2).mean()) < 0.006      # the reference's own run
    P = OracleProblem(p)
    end, s1, s2, total = np.zeros(18, np.int64), np.zeros(8, np.int64), np.zeros(8, np.int64), 0
    for c in range(8):
        oc = OracleChain(P)
        oc.seed_mt(100 + c)
        oc.find_feasible()
        oc.sample_stats(2000, 10 ** 9, np.zeros(18, np.int64), s1, s2)
        total += oc.sample_stats(100000, 10 ** 9, end, s1, s2)
    mean = end / total
    assert np.sqrt(((mean - rej) ** 2).mean()) < 0.006
    assert np.abs(mean - rej).max() < 0.012
