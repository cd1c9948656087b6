#!/usr/bin/env python
"""Extracts the sufficient statistics of the reference's shipped result archive src/data/PredPrey32-16.stat (the run
behind Table 2 of the paper, doc/ABMCMC.tex:836-853) into tests/golden/stat_PredPrey32-16.abmp.xz.
Run from the repository root in the build container (needs /root/reference)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200.abmp import write_abmp  # noqa: E402
from agentbasedmcmc_b200.diagnostics import read_stat  # noqa: E402

m = read_stat("/root/reference/src/data/PredPrey32-16.stat")
arr = {
    "n_samples": np.array([c.n_samples for c in m.chains], np.int32),
    "sum": np.concatenate([c.sum for c in m.chains]),
    "sum_sq": np.concatenate([c.sum_sq for c in m.chains]),
    "vario": np.concatenate([c.vario.ravel() for c in m.chains]),
    "vario_stride": np.array([c.vario_stride for c in m.chains], np.int32),
    "mean_end_state": m.mean_end_state(),
    "n_accepted": np.concatenate([c.n_accepted.ravel() for c in m.chains]).astype(np.int32),
    "n_proposals": np.concatenate([c.n_proposals.ravel() for c in m.chains]).astype(np.int32),
    "kappa_exec_ms": np.array([m.kappa, m.exec_time_ms], np.float64),
}
out = ROOT / "tests" / "golden" / "stat_PredPrey32-16.abmp.xz"
write_abmp(out, arr)
print(out, out.stat().st_size)
