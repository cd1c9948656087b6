"""Development aid: relaxation of the chains from their first feasible state (per-interval infeasible fraction, acceptance
and number of agents in the end state)."""
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S
from agentbasedmcmc_b200.abmp import read_abmp
name, n_chains, interval, n_int = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
P = S.Problem(read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz"))
smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=3, prior_init=True)
prev = smp.counters().sum(0)
for k in range(n_int):
    t0 = time.time()
    smp.reset_statistics()
    smp.sample(10 ** 12, max_steps=interval)
    end, ss, sq, ns = smp.statistics()
    c = smp.counters().sum(0)
    d = c - prev
    prev = c
    prop, acc = d[:4], d[4:]
    infeasible = (prop[0] + acc[2] + prop[1] - acc[1]) / prop.sum()
    stuck = float((ns == 0).mean())
    print(f"steps {(k+1)*interval:9d}: infeasible {infeasible:.3f} accept {acc.sum()/prop.sum():.3f} agents@T {end.sum()/max(ns.sum(),1):8.2f} "
          f"chains without a feasible state {stuck:.3f}  ({time.time()-t0:.1f}s)", flush=True)
