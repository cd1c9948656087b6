"""Quick throughput probe (development aid): steps/s of abm_chains_step at a given config."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "pp16-8"
    n_chains = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    p = read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz")
    t0 = time.time()
    P = S.Problem(p)
    print(f"problem: rows={P.n_rows} cols={P.n_cols} maxCoupled={P.info.max_coupled} shared={P.info.shared_bytes/1e6:.1f}MB "
          f"chain={P.info.chain_state_bytes/1e6:.3f}MB build={time.time()-t0:.2f}s", flush=True)
    t0 = time.time()
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=1, prior_init="--empty-init" not in sys.argv)
    dt = time.time() - t0
    it = smp.init_iterations
    print(f"init: {dt:.2f}s, iterations mean={it.mean():.0f} max={it.max()} -> {it.sum()/dt/1e6:.2f} M it/s", flush=True)
    burn = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--burnin=")]
    if burn:   # bring the ensemble to its stationary state first (bench.py does the same before it times anything)
        t0 = time.time()
        for _ in range(burn[0] // 250000):
            smp.step(250000)
        smp.synchronize()
        print(f"burn-in: {burn[0]} steps in {time.time()-t0:.1f}s", flush=True)
    smp.step(n_steps)
    smp.synchronize()
    for r in range(reps):
        t0 = time.time()
        smp.step(n_steps)
        smp.synchronize()
        dt = time.time() - t0
        print(f"rep {r}: {n_chains*n_steps/dt/1e6:.2f} M chain-steps/s ({dt*1e3:.1f} ms)", flush=True)
    st = smp.stats
    print(f"accept={st.totalAccepted()/st.nSamples():.3f} infeasible={st.nInfeasibleSamples()/st.nSamples():.3f}")


if __name__ == "__main__":
    main()
