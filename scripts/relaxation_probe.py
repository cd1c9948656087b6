"""Relaxation of an ensemble from its first feasible states: per interval of Metropolis steps, the ensemble mean of the synopsis
statistics over the interval's feasible samples (FiguresForPaper.h:245-263), the spread of the chain means, the infeasible
fraction and the acceptance.  Two ensembles run side by side on their own streams: every chain from its own prior sample
(FiguresForPaper.h:56, what the paper's runs do) and every chain from the empty trajectory (SparseBasisSampler(posterior,
kappa), SparseBasisSampler.h:199-206, what Experiments.cpp does).  Sizes the burn-in of bench.py and of the paper-scale
marginal test against the long reference run (tests/golden/posterior_pp32-16.abmp.xz).

    python scripts/relaxation_probe.py pp32-16 2048 250000 40 gpurun_out/relax.json
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S
from agentbasedmcmc_b200.abmp import read_abmp

name, n_chains, interval, n_int, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
kinds = sys.argv[6].split(",") if len(sys.argv) > 6 else ["prior", "empty"]
P = S.Problem(read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz"))
streams = {k: torch.cuda.Stream() for k in kinds}
t0 = time.time()
smp = {}
for k in kinds:
    smp[k] = S.SparseBasisSampler(P, n_chains=n_chains, seed=11 if k == "prior" else 12, stream=streams[k].cuda_stream,
                                  prior_init=(k == "prior"))
setup = time.time() - t0
res = {k: {"init_iterations_mean": float(smp[k].init_iterations.mean()), "init_iterations_max": int(smp[k].init_iterations.max()),
           "intervals": []} for k in kinds}
prev = {k: smp[k].counters().sum(0) for k in kinds}
for i in range(n_int):
    t1 = time.time()
    for k in kinds:
        smp[k].reset_statistics()
        smp[k].sample(10 ** 12, max_steps=interval)
    for k in kinds:
        end, ss, sq, ns = smp[k].statistics()
        c = smp[k].counters().sum(0)
        d = c - prev[k]
        prev[k] = c
        prop, acc = d[:4], d[4:]
        infeasible = float((prop[0] + acc[2] + prop[1] - acc[1]) / prop.sum())
        ok = ns > 0
        cm = ss[ok] / ns[ok, None]
        row = {"steps": (i + 1) * interval, "mean_synopsis": (ss.sum(0) / max(ns.sum(), 1)).tolist(),
               "mean_of_chain_means": cm.mean(0).tolist(), "sd_of_chain_means": cm.std(0).tolist(),
               "agents_at_T": float(end.sum() / max(ns.sum(), 1)), "infeasible": infeasible,
               "accept": float(acc.sum() / prop.sum()), "chains_without_sample": float((~ok).mean()),
               "feasible_samples_per_chain": float(ns.mean())}
        res[k]["intervals"].append(row)
        print(k, json.dumps(row), f"({time.time() - t1:.1f}s)", flush=True)
json.dump({"workload": name, "chains": n_chains, "interval_steps": interval, "setup_seconds": setup, "ensembles": res}, open(out, "w"))
