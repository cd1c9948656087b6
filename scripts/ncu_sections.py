"""Per-section instruction / stall-sample breakdown of stepKernel from an ncu source-page CSV (cuda,sass view)."""
import csv
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
# section marks come from the source the profiled library was built from: argv[4], else the working tree
srcPath = Path(sys.argv[4]) if len(sys.argv) > 4 else ROOT / "agentbasedmcmc_b200/csrc/abm_kernels.cuh"
src = dict(enumerate(srcPath.read_text().split("\n"), start=1))
rows = list(csv.reader(open(sys.argv[1])))


def find(pat):
    for i in sorted(src):
        if pat in src[i]:
            return i
    raise KeyError(pat)


marks = [("rng", find("for (long long s = 0")), ("descent", find("---- column draw")), ("record+rows", find("Proposal::Proposal (:288")),
         ("columns: prefetch", find("coupled columns, one per lane")), ("importance", find("auto importance = [&]")),
         ("columns: chunks", find("double csLane = 0.0")), ("ratio/accept", find("calcRatioOfSums (:276")),
         ("apply X/delta", find("applyProposal (:322")), ("commit", find("const int Ca = accept")),
         ("sample stats", find("the do-while at SparseBasisSampler.h:247")), ("end", find("if (act) { ++iter"))]
steps = float(sys.argv[2])
hdr = rows[2]
ci = {h: i for i, h in enumerate(hdr)}
iI, iS = ci["Instructions Executed"], ci["# Samples"]
lines = {}
for r in rows[3:]:
    if len(r) < len(hdr) or not r[0]:
        continue
    try:
        lines[int(r[0])] = (float(r[iI]) / steps, float(r[iS]), r[1])
    except ValueError:
        pass
totS = sum(v[1] for v in lines.values())
print(f"total warp-instructions per chain-step: {sum(v[0] for v in lines.values()):.1f}")
for (name, a), (_, b) in zip(marks[:-1], marks[1:]):
    i = sum(v[0] for k, v in lines.items() if a <= k < b)
    s = sum(v[1] for k, v in lines.items() if a <= k < b)
    print(f"{name:16s} lines {a}-{b}: inst/chain-step {i:7.1f}  stall samples {100 * s / totS:5.1f}%")
i = sum(v[0] for k, v in lines.items() if k < marks[0][1])
s = sum(v[1] for k, v in lines.items() if k < marks[0][1])
print(f"{'inlined helpers':16s} (energy, exp, node(), shuffles): inst/chain-step {i:7.1f}  stall samples {100 * s / totS:5.1f}%")
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
print(f"--- top {n} lines by stall samples")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:n]:
    print(f"{k:4d} inst={v[0]:6.1f} samp%={100 * v[1] / totS:5.1f} | {v[2][:105]}")
print(f"--- top {n} lines by instructions")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{k:4d} inst={v[0]:6.1f} samp%={100 * v[1] / totS:5.1f} | {v[2][:105]}")
