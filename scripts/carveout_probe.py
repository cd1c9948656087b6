"""Development aid: shared-memory carve-out of the step kernel (ABM_SMEM_CARVEOUT, percent; unset = the driver's choice)."""
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run(problem, steps, reps, carveout):
    env = dict(os.environ)
    env.pop("ABM_SMEM_CARVEOUT", None)
    if carveout is not None:
        env["ABM_SMEM_CARVEOUT"] = str(carveout)
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "probe.py"), problem, "8192", str(steps), str(reps)], env=env,
                         capture_output=True, text=True, timeout=120).stdout
    vals = [float(x) for x in re.findall(r"rep \d+: ([0-9.]+) M", out)]
    v = vals[-1] if vals else float("nan")
    print(f"{problem} carveout={carveout}: {v:.1f} M chain-steps/s", flush=True)
    return v


best, best_v = None, run("pp16-8", 1000, 4, None)
for c in (70, 78, 100):
    v = run("pp16-8", 1000, 4, c)
    if v > best_v:
        best, best_v = c, v
print("best on pp16-8:", best, flush=True)
run("pp32-16", 2000, 3, None)
if best is not None:
    run("pp32-16", 2000, 3, best)
