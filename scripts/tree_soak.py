"""Soak of the blocked sum tree against the node-for-node reference tree: the same chains on the same
Philox streams in both storages, traced launch by launch; reports the first step at which a drawn column or an accept decision
differs (expected: none -- a difference needs a uniform within rounding distance of an interval boundary), and the drift of
the blocked tree's incrementally kept sums against the exact sums of their leaves.

    python scripts/tree_soak.py pp16-8 16 100000 100 gpurun_out/tree_soak.json      (16 chains, 100 launches of 1e5 steps = 1e7 steps)
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S
from agentbasedmcmc_b200.abmp import read_abmp

name, n_chains, per_launch, n_launch, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
p = read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz")
smp = {t: S.SparseBasisSampler(S.Problem(p, sum_tree=t), n_chains=n_chains, seed=424242, first_chain_id=9) for t in ("blocked", "reference")}
assert np.array_equal(smp["blocked"].init_iterations, smp["reference"].init_iterations)
first_diff, drift_max, rows = None, np.zeros(3), []
t0 = time.time()
for k in range(n_launch):
    tr = {t: smp[t].step_traced(per_launch) for t in smp}
    same_j = tr["blocked"][0] == tr["reference"][0]
    same_a = tr["blocked"][1] == tr["reference"][1]
    if first_diff is None and not (same_j.all() and same_a.all()):
        c, s = np.argwhere(~(same_j & same_a))[0]
        first_diff = {"chain": int(c), "step": int(k * per_launch + s)}
    d = np.array([smp["blocked"].tree_drift(c) for c in range(min(n_chains, 4))])
    rel = (d[:, 1:] / d[:, :1]).max(axis=0)
    drift_max = np.maximum(drift_max, rel)
    if k % 10 == 9 or k == n_launch - 1:
        rows.append({"steps": (k + 1) * per_launch, "first_difference": first_diff, "max_relative_drift_32_1024_32768": rel.tolist(),
                     "seconds": round(time.time() - t0, 1)})
        print(rows[-1], flush=True)
    if first_diff is not None:
        break
same_state = all(np.array_equal(smp["blocked"].state(c)[0], smp["reference"].state(c)[0]) for c in range(n_chains)) if first_diff is None else False
json.dump({"workload": name, "chains": n_chains, "steps_per_chain": (k + 1) * per_launch, "first_difference": first_diff,
           "final_trajectories_identical": same_state, "max_relative_drift_of_sums_32_1024_32768": drift_max.tolist(),
           "note": "drift relative to the root; sums of 32 leaves are rebuilt from the leaves every 256th launch, the upper levels before every launch",
           "log": rows}, open(out, "w"), indent=1)
