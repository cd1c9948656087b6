#!/usr/bin/env python
"""Data-assimilation window sweep (BASELINE.json configs[4]): PredPrey 32x32, T in {4, 8, 16, 32}.

For every window length: chain-steps/s, ESS/s and R-hat of the GPU sampler (8192 chains by default, each from its own
prior sample) next to the unmodified reference sampler on the host cores (oracle/_ref/abm_ref `series`, one chain per
thread), both analysed with the reference's formulas (diagnostics/MultiChainStats.h:81-181 restated in
agentbasedmcmc_b200/diagnostics.py) on the synopsis series of the end state (FiguresForPaper.h:245-263).  The mean
synopsis of the two ensembles is printed side by side as the posterior cross-check.

    python scripts/window_sweep.py [--T 4 8 16 32] [--chains 8192] [--out profiles/window_sweep.json]

Problems: tests/golden/pp32-{4,8,16,32}.abmp.xz (reference generator, seed 1234; T = 32 at density 0.02, see
tests/golden/make_golden.py).  This is a measurement script, not a test: it needs a GPU and, for the CPU arm, the compiled
reference (oracle/_ref/abm_ref travels to the GPU box with the repository snapshot).
"""
from __future__ import annotations

import argparse
import json
import lzma
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import diagnostics as D  # noqa: E402
from agentbasedmcmc_b200 import sampler as S  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402


def analyse(series: np.ndarray):
    """series [sequences][n][d] -> (n_eff per sequence [d], R-hat [d]) with the reference's estimators."""
    ms = D.MultiChainStats([D.ChainStats.from_series(x.astype(np.float64), n_lags=200, max_lag_proportion=0.5) for x in series])
    n_eff = ms.effective_samples(reference_integer_division=False) / len(series)
    return n_eff, ms.potential_scale_reduction(reference_integer_division=False)


def gpu_arm(abmp, n_chains, burnin, mh, reps, ess_samples, ess_chains):
    t0 = time.time()
    P = S.Problem(abmp)
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=777, prior_init=True)
    setup = time.time() - t0
    done = 0
    while done < burnin:
        smp.step(min(250000, burnin - done))
        done += 250000
    smp.synchronize()
    rates = []
    for _ in range(reps):
        t0 = time.perf_counter()
        smp.step(mh)
        smp.synchronize()
        rates.append(n_chains * mh / (time.perf_counter() - t0))
    smp.reset_statistics()
    k = min(ess_chains, n_chains)
    smp.record_series(k, 1, ess_samples)
    t0 = time.perf_counter()
    smp.sample(ess_samples, max_steps=4 * ess_samples)
    smp.synchronize()
    ess_s = time.perf_counter() - t0
    end, ss, sq, ns = smp.statistics()
    n_rec = int(min(np.percentile(ns[:k], 25), ess_samples))
    use = [c for c in range(k) if ns[c] >= n_rec]
    out = {"chain_steps_per_s": float(np.median(rates)), "setup_seconds": round(setup, 1), "chains": n_chains,
           "chain_state_gb": round(P.info.chain_state_bytes * n_chains / 1e9, 2), "n_cols": P.n_cols,
           "mean_synopsis": [round(float(x), 3) for x in ss.sum(axis=0) / max(ns.sum(), 1)],
           "mean_agents_at_T": float(end.sum() / max(ns.sum(), 1))}
    st = smp.stats
    out["acceptance"] = round(st.totalAccepted() / max(st.nSamples(), 1), 4)
    out["infeasible_fraction"] = round(st.nInfeasibleSamples() / max(st.nSamples(), 1), 4)
    if n_rec >= 400 and len(use) >= 4:
        n_eff, rhat = analyse(smp.series()[use, :n_rec])
        out.update({"ess_per_s": float(n_eff.min() * n_chains / ess_s), "n_eff_per_chain": [round(float(x), 1) for x in n_eff],
                    "rhat": [round(float(x), 4) for x in rhat], "samples_per_chain": n_rec, "sequences": len(use),
                    "ess_seconds": round(ess_s, 3)})
    else:
        out["ess_per_s"] = None
    smp.close()
    P.close()
    return out


def cpu_arm(fixture: Path, threads, n_samples, burnin):
    exe = ROOT / "oracle" / "_ref" / "abm_ref"
    if not exe.exists():
        return None
    raw = tempfile.NamedTemporaryFile(prefix="abmp_", suffix=".abmp", delete=False)
    raw.write(lzma.decompress(fixture.read_bytes()))
    raw.close()
    out = raw.name + ".series"
    try:
        res = subprocess.run([str(exe), "series", raw.name, str(threads), str(n_samples), str(burnin), out], capture_output=True, text=True,
                             timeout=3000)
        if res.returncode != 0:
            return {"error": res.stderr[-300:]}
        r = read_abmp(out)
    finally:
        os.unlink(raw.name)
        if os.path.exists(out):
            os.unlink(out)
    nt, n, d = (int(x) for x in r["shape"])
    series = r["series"].reshape(nt, n, d)
    sec, steps, _ = r["timing"]
    n_eff, rhat = analyse(series)
    return {"chain_steps_per_s": float(steps / sec), "threads": nt, "ess_per_s": float(n_eff.min() * nt / sec),
            "n_eff_per_chain": [round(float(x), 1) for x in n_eff], "rhat": [round(float(x), 4) for x in rhat], "samples_per_chain": n,
            "seconds": round(float(sec), 2), "mean_synopsis": [round(float(x), 3) for x in series.mean(axis=(0, 1))]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, nargs="+", default=[4, 8, 16, 32])
    ap.add_argument("--chains", type=int, default=8192)
    ap.add_argument("--burnin", type=int, default=2000000, help="GPU Metropolis steps per chain before measuring")
    ap.add_argument("--mh-steps", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ess-samples", type=int, default=20000)
    ap.add_argument("--ess-chains", type=int, default=64)
    ap.add_argument("--cpu-samples", type=int, default=200000)
    ap.add_argument("--cpu-burnin", type=int, default=400000, help="reference feasible samples per chain before recording")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rows = []
    for T in args.T:
        fixture = ROOT / "tests" / "golden" / f"pp32-{T}.abmp.xz"
        abmp = read_abmp(fixture)
        row = {"workload": f"PredPrey 32x32, T={T}", "density": float(abmp["pp_params"][0]),
               "gpu": gpu_arm(abmp, args.chains, args.burnin, args.mh_steps, args.reps, args.ess_samples, args.ess_chains)}
        if not args.no_cpu:
            row["reference_cpu"] = cpu_arm(fixture, os.cpu_count() or 1, args.cpu_samples, args.cpu_burnin)
        print(json.dumps(row), flush=True)
        rows.append(row)
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(r) for r in rows) + "\n")


if __name__ == "__main__":
    main()
