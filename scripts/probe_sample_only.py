import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S
from agentbasedmcmc_b200.abmp import read_abmp
name, n_chains, n_steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
P = S.Problem(read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz"))
smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=1)
for _ in range(3):
    smp.sample(10 ** 12, max_steps=n_steps)
smp.synchronize()
print("done")
