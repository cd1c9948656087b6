#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory from the UNMODIFIED reference.

Needs /root/reference (read-only) and a host compiler; run from the repository root:
    make -C oracle ref && python tests/golden/make_golden.py [--with-32] [--with-sweep] [--traces-only | --agents-only]
--with-32 adds the paper-scale problem (32x32, T = 16; ~4 minutes of TableauNormMinimiser), --with-sweep the window sweep's
problems (32x32, T = 4, 8, 32; T = 32 at density 0.02 because the reference cannot draw an act-Fermionic real trajectory of
32 steps at 0.05, Forecast.h:145-147; ~20 minutes).
Problems: FiguresForPaper::generateStandardProblemFile protocol (FiguresForPaper.h:25-37): seed 1234,
pPredator = pPrey = 0.05, pMakeObservation = 0.05, pObserveIfPresent = 1.0, kappa = 10.
Traces: oracle/ref/abm_ref.cpp `trace` (initial state = prior.nextSample(false) under initSeed; sampler
constructed and stepped under chainSeed with the reference's own mt19937 Random::gen).
"""
import lzma
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = HERE.parent.parent / "oracle" / "_ref" / "abm_ref"


def xz(src: Path, dst: Path):
    dst.write_bytes(lzma.compress(src.read_bytes(), preset=9 | lzma.PRESET_EXTREME))
    print(f"{dst.name}: {dst.stat().st_size} bytes")


TRACES = [("pp8-4", 11, 101, 20000, 500), ("pp8-4", 12, 102, 20000, 500), ("pp16-8", 12, 202, 20000, 500),
          ("pp8-4-poisson", 11, 103, 20000, 500), ("exp_pp3-2-single", 31, 301, 20000, 500)]


def traces_only():
    """--traces-only: re-run `abm_ref trace` on the COMMITTED problem fixtures (same seeds, so the step arrays are reproduced
    bit for bit; used when the trace format gains arrays, e.g. the reference's own end state / synopsis per checkpoint)."""
    tmp = Path(tempfile.mkdtemp())
    for name, s0, s1, n, every in TRACES:
        raw = tmp / f"{name}.abmp"
        raw.write_bytes(lzma.decompress((HERE / f"{name}.abmp.xz").read_bytes()))
        out = tmp / f"trace_{name}_{s1}.abmp"
        subprocess.run([str(REF), "trace", str(raw), str(s0), str(s1), str(n), str(every), str(out)], check=True)
        xz(out, HERE / f"trace_{name}_{s1}.abmp.xz")


# The reference's validation experiments on its other agent types (Experiments.cpp:21-47, 70-125) -- `abm_ref toy`: problem
# (agent described by tables), the rejection sampler's end-state means over 2e6 samples, the exact posterior over trajectories,
# and a golden step trace of 20000 steps -- and PredPrey problems re-described as table agents (`abm_ref astable`), which the
# existing PredPrey traces must keep matching.
TOYS = [("binomial", 1234, 401), ("catmouse", 1234, 402), ("catmouse-assim", 1234, 403),
        ("ring", 1234, 404)]   # own test agent with non-trivial importance factors (oracle/ref/test_agents.h)
AS_TABLE = ["pp8-4", "exp_pp3-2-single"]


def agents_only():
    tmp = Path(tempfile.mkdtemp())
    for kind, seed, chain_seed in TOYS:
        prob, trace = tmp / f"toy_{kind}.abmp", tmp / f"trace_toy_{kind}_{chain_seed}.abmp"
        subprocess.run([str(REF), "toy", kind, str(seed), str(chain_seed), "20000", "500", "2000000", str(prob), str(trace)], check=True)
        xz(prob, HERE / prob.with_suffix(".abmp.xz").name)
        xz(trace, HERE / trace.with_suffix(".abmp.xz").name)
    for name in AS_TABLE:
        raw = tmp / f"{name}.abmp"
        raw.write_bytes(lzma.decompress((HERE / f"{name}.abmp.xz").read_bytes()))
        out = tmp / f"{name}-astable.abmp"
        subprocess.run([str(REF), "astable", str(raw), str(out)], check=True)
        xz(out, HERE / f"{name}-astable.abmp.xz")


def main():
    if "--traces-only" in sys.argv:
        return traces_only()
    if "--agents-only" in sys.argv:
        return agents_only()
    with32 = "--with-32" in sys.argv
    tmp = Path(tempfile.mkdtemp())
    problems = [(8, 4), (16, 8)] + ([(32, 16)] if with32 else [])
    for g, t in problems:
        raw = tmp / f"pp{g}-{t}.abmp"
        subprocess.run([str(REF), "gen", str(g), str(t), "1234", "10.0", str(raw)], check=True, stdout=subprocess.DEVNULL)
        xz(raw, HERE / f"pp{g}-{t}.abmp.xz")
    if "--with-sweep" in sys.argv:
        for t, dens in ((4, 0.05), (8, 0.05), (32, 0.02)):
            raw = tmp / f"pp32-{t}.abmp"
            subprocess.run([str(REF), "gen", "32", str(t), "1234", "10.0", str(raw), str(dens), str(dens)], check=True, stdout=subprocess.DEVNULL)
            xz(raw, HERE / f"pp32-{t}.abmp.xz")
    # Poisson start state (PoissonStartState.h; PredPreyProblem.h:60-62 keeps it as the commented-out alternative)
    for g, t in ((8, 4), (16, 8)):
        raw = tmp / f"pp{g}-{t}-poisson.abmp"
        subprocess.run([str(REF), "gen", str(g), str(t), "1234", "10.0", str(raw), "0.05", "0.05", "0.05", "1.0", "poisson"], check=True,
                       stdout=subprocess.DEVNULL)
        xz(raw, HERE / f"pp{g}-{t}-poisson.abmp.xz")
    # posterior.constraints of the problems, as the reference assembles them: the input of TableauNormMinimiser, against whose
    # tableau (in the problem fixtures) the own basis builder is checked (tests/test_basis_builder.py)
    for name in ["pp8-4", "pp16-8", "pp16-8-poisson"] + (["pp32-16"] if with32 else []):
        craw = tmp / f"constraints_{name}.abmp"
        subprocess.run([str(REF), "constraints", str(tmp / f"{name}.abmp"), str(craw), "0"], check=True)
        xz(craw, HERE / f"constraints_{name}.abmp.xz")
    # (problem, initSeed, chainSeed, nSteps, checkpointEvery)
    for name, s0, s1, n, every in TRACES[:4]:
        raw = tmp / f"trace_{name}_{s1}.abmp"
        subprocess.run([str(REF), "trace", str(tmp / f"{name}.abmp"), str(s0), str(s1), str(n), str(every), str(raw)], check=True)
        xz(raw, HERE / f"trace_{name}_{s1}.abmp.xz")
    # prior samples (prior.nextSample(false), FiguresForPaper.h:56): per (t, type, act) mean of [event > 0] over 20000 draws
    import numpy as np
    sys.path.insert(0, str(HERE.parent.parent))
    from agentbasedmcmc_b200.abmp import write_abmp
    for name in ("pp8-4", "pp8-4-poisson"):
        raw = tmp / f"prior_{name}.i8"
        subprocess.run([str(REF), "init", str(tmp / f"{name}.abmp"), "20000", "4711", str(raw)], check=True)
        x = np.fromfile(raw, dtype=np.int8).reshape(20000, 4, 2, 64, 6)
        write_abmp(HERE / f"prior_{name}.abmp.xz", {"mean_indicator": (x > 0).mean(axis=(0, 3)).ravel(),
                                                    "mean_count": x.astype(np.float64).mean(axis=(0, 3)).ravel(),
                                                    "n": np.array([20000, 64], np.int32),
                                                    # start states holding two agents that chose different acts (Poisson only)
                                                    "multi_act_t0": np.array([((x[:, 0] > 0).sum(axis=-1) >= 2).mean()])})
    # The reference's own validation experiment on PredPrey (Experiments::PredPreySingleObservation, Experiments.cpp:49-65): 3x3 grid,
    # 2 timesteps, Poisson start, one observation; its tableau has an entry of magnitude 2 and the grid is not a power of two.
    # experiment3 = problem + end-state means of the reference sampler (500000 samples after 10000) and of the reference
    # RejectionSampler (2000000 samples; 100000 in the experiment itself); then a golden step trace on the same problem.
    if True:
        raw, marg = tmp / "exp_pp3-2-single.abmp", tmp / "exp_pp3-2-single_means.abmp"
        subprocess.run([str(REF), "experiment3", str(raw), str(marg), "500000", "10000", "2000000", "4321"], check=True)
        xz(raw, HERE / "exp_pp3-2-single.abmp.xz")
        xz(marg, HERE / "exp_pp3-2-single_means.abmp.xz")
        craw = tmp / "constraints_exp_pp3-2-single.abmp"
        subprocess.run([str(REF), "constraints", str(raw), str(craw), "0"], check=True)
        xz(craw, HERE / "constraints_exp_pp3-2-single.abmp.xz")
        traw = tmp / "trace_exp_pp3-2-single_301.abmp"
        subprocess.run([str(REF), "trace", str(raw), "31", "301", "20000", "500", str(traw)], check=True)
        xz(traw, HERE / "trace_exp_pp3-2-single_301.abmp.xz")
    # long reference chains for posterior marginals (64 chains x 60000 feasible samples after 60000 burn-in)
    raw = tmp / "marginals_pp8-4.abmp"
    subprocess.run([str(REF), "marginals", str(tmp / "pp8-4.abmp"), "64", "60000", "60000", str(raw)], check=True)
    xz(raw, HERE / "marginals_pp8-4.abmp.xz")
    agents_only()


if __name__ == "__main__":
    main()
