"""Development aid: a small pass over every kernel and mode (both tree storages, Bernoulli / Poisson start, the 3x3 experiment)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402

for name, prior in (("pp8-4", True), ("exp_pp3-2-single", False), ("pp8-4-poisson", True)):
    for tree in ("blocked", "reference"):
        P = S.Problem(read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz"), sum_tree=tree)
        smp = S.SparseBasisSampler(P, n_chains=13, seed=5, prior_init=prior)
        smp.step(200)
        smp.sample(60)
        smp.sample(10 ** 12, max_steps=150)
        smp.enqueue_statistics()
        smp.step(20)
        end, ss, sq, ns, c = smp.wait_statistics()
        j, acc, inf = smp.step_traced(50)
        smp.synchronize()
        print(name, tree, "samples", int(ns.sum()), "accepted", int(acc.sum()), flush=True)
        smp.close()
print("mode sweep done")
