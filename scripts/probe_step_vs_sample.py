"""Development aid: a Metropolis launch and a sampling launch of the same length, alternating, at the same state -- for ncu
captures that compare stepKernel<MODE_MCMC> with stepKernel<MODE_SAMPLE_FIXED> (scripts/ncu_phases.py on each)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402

name, n_chains, n_steps, burn = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
P = S.Problem(read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz"))
smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=1, prior_init=True)
smp.step(burn // 2)
smp.step(burn - burn // 2)
for _ in range(3):
    smp.step(n_steps)
    smp.synchronize()
    a = smp.last_launch_ms()
    smp.sample(10 ** 12, max_steps=n_steps)
    smp.synchronize()
    print(f"step {a:.3f} ms, sample {smp.last_launch_ms():.3f} ms", flush=True)
