"""Development aid: step vs sample-mode throughput and the cost of a statistics read-back."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S, ensemble as E  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402

name, n_chains, n_steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
P = S.Problem(read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz"))
smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=1)
smp.step(n_steps); smp.synchronize()
for mode in ("step", "sample", "step", "sample"):
    t0 = time.time()
    if mode == "step":
        smp.step(n_steps)
    else:
        smp.sample(10 ** 12, max_steps=n_steps)
    smp.synchronize()
    dt = time.time() - t0
    print(f"{mode}: {n_chains * n_steps / dt / 1e6:.1f} M chain-steps/s (host clock), step kernel {smp.last_launch_ms():.2f} ms on the device", flush=True)
t0 = time.time(); st = smp.statistics(); t1 = time.time(); c = smp.counters(); t2 = time.time()
packed = E.pack_local(*st, c); t3 = time.time()
print(f"statistics() {1e3*(t1-t0):.2f} ms, counters() {1e3*(t2-t1):.2f} ms, pack_local {1e3*(t3-t2):.2f} ms")
