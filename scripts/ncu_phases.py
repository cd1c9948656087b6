"""Per-phase instruction / stall-sample breakdown of stepKernel from an `ncu --page source --csv --print-source cuda,sass` dump.
Rows are attributed to abm_kernels.cuh only when their text equals that line of the file (the dump mixes in CUDA header lines,
e.g. the shuffle intrinsics, under their own line numbers); everything else is reported as "cuda headers".

    python scripts/ncu_phases.py dump.csv <chain-steps in the profiled launch> [path of the kernel source the library was built from]
"""
import csv
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src_path = Path(sys.argv[3]) if len(sys.argv) > 3 else ROOT / "agentbasedmcmc_b200/csrc/abm_kernels.cuh"
src = {i: ln.strip() for i, ln in enumerate(src_path.read_text().split("\n"), start=1)}
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2])
hdr = rows[2]
ci = {}
for i, h in enumerate(hdr):
    ci.setdefault(h, i)
iI, iS = ci["Instructions Executed"], ci["# Samples"]
ours, other = {}, [0.0, 0.0]
for r in rows[3:]:
    if len(r) < len(hdr) or not r[0]:
        continue
    try:
        ln, ins, smp = int(r[0]), float(r[iI]), float(r[iS])
    except ValueError:
        continue
    if src.get(ln, None) == r[1].strip() and r[1].strip():
        a = ours.setdefault(ln, [0.0, 0.0])
        a[0] += ins
        a[1] += smp
    else:
        other[0] += ins
        other[1] += smp


def find(pat):
    for i in sorted(src):
        if pat in src[i]:
            return i
    raise KeyError(pat)


marks = [("helpers (exp, philox, select, ...)", 1), ("kernel prologue / re-derived ids and pointers", find("stepKernel(DevProblem P, DevChains S, RunParams R)")),
         ("loop head, uniforms", find("for (int s = 0; s < nStepsLaunch")), ("column draw (descent)", find("---- column draw")),
         ("record fetch + header", find("---- Proposal::Proposal (:288")), ("rows", find("const int nW = warpMax(n);")),
         ("importance", find("auto importance = [&]")), ("coupled columns (chunks)", find("double csLane = 0.0")),
         ("ratio / accept", find("bool accept = act;")), ("commit", find("---- applyProposal (:322")),
         ("sample statistics, loop end", find("the do-while at SparseBasisSampler.h:247")), ("epilogue", find("if (act) ++itersDone"))]
totI = sum(v[0] for v in ours.values()) + other[0]
totS = sum(v[1] for v in ours.values()) + other[1]
print(f"total {totI / steps:.1f} warp-instructions per chain-step, {totS:.0f} stall samples")
for k, (name, a) in enumerate(marks):
    b = marks[k + 1][1] if k + 1 < len(marks) else 10 ** 9
    I = sum(v[0] for ln, v in ours.items() if a <= ln < b)
    S = sum(v[1] for ln, v in ours.items() if a <= ln < b)
    print(f"{name:48s} inst/step={I / steps:7.1f} ({100 * I / totI:4.1f}%)  samples={100 * S / totS:4.1f}%")
print(f"{'cuda headers (shuffles, votes, atomics, math)':48s} inst/step={other[0] / steps:7.1f} ({100 * other[0] / totI:4.1f}%)  samples={100 * other[1] / totS:4.1f}%")
print("--- top lines by stall samples")
for ln, v in sorted(ours.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{ln:5d} inst/step={v[0] / steps:6.1f} samp={100 * v[1] / totS:4.1f}% | {src[ln][:110]}")
