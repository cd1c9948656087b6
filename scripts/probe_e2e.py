"""Development aid: where the end-to-end step time goes (sample launches alone / with blocking read-back / with queued read-back)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S, ensemble as E  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402

name, n_chains, mh, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
P = S.Problem(read_abmp(ROOT / "tests" / "golden" / f"{name}.abmp.xz"))
smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=20221, prior_init=True)
for _ in range(3):
    smp.step(mh)
smp.sample(10 ** 12, max_steps=mh)
smp.enqueue_statistics(); smp.wait_statistics()
smp.synchronize()


def report(tag, t0):
    smp.synchronize()
    dt = time.perf_counter() - t0
    print(f"{tag}: {1e3 * dt / steps:.2f} ms per step, {n_chains * mh * steps / dt / 1e6:.1f} M chain-steps/s", flush=True)


t0 = time.perf_counter()
for k in range(steps):
    smp.step(mh)
report("step launches only", t0)
t0 = time.perf_counter()
dev = 0.0
for k in range(steps):
    smp.sample(10 ** 12, max_steps=mh)
report("sample launches only", t0)
print(f"  last sample step kernel on the device: {smp.last_launch_ms():.2f} ms")
t0 = time.perf_counter()
for k in range(steps):
    smp.sample(10 ** 12, max_steps=mh)
    end, ss, sq, ns = smp.statistics()
    c = smp.counters()
report("sample + blocking read-back", t0)
t0 = time.perf_counter()
smp.sample(10 ** 12, max_steps=mh)
smp.enqueue_statistics()
for k in range(steps):
    if k + 1 < steps:
        smp.sample(10 ** 12, max_steps=mh)
    end, ss, sq, ns, c = smp.wait_statistics()
    if k + 1 < steps:
        smp.enqueue_statistics()
report("sample + queued read-back", t0)
t0 = time.perf_counter()
smp.sample(10 ** 12, max_steps=mh)
smp.enqueue_statistics()
for k in range(steps):
    if k + 1 < steps:
        smp.sample(10 ** 12, max_steps=mh)
    end, ss, sq, ns, c = smp.wait_statistics()
    if k + 1 < steps:
        smp.enqueue_statistics()
    packed = E.pack_local(end, ss, sq, ns, c)
    summary = E.allreduce_summary(packed, end.size, ss.shape[1])
report("sample + queued read-back + host reduction", t0)
