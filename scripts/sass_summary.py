"""Static view of the built step kernels (no GPU): ptxas resource lines and the SASS opcode mix of each
stepKernel<MODE,16,blocked,FIX> (FIX = compile-time shared-memory layout, what the paper's problems run), and the opcodes that
show the bulk-copy variant (libabm_bulk.so, -DABM_REC_BULK=1: UBLKCP = cp.async.bulk, SYNCS = mbarrier).

    python scripts/sass_summary.py > profiles/<tag>_ptxas_sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "agentbasedmcmc_b200" / "csrc"
LIB = ROOT / "agentbasedmcmc_b200" / "libabmcmc_b200.so"
MODES = {0: "MODE_MCMC", 1: "MODE_INIT", 2: "MODE_SAMPLE", 3: "MODE_SAMPLE_FIXED"}

out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false", "-Xptxas=-v",
                      f"-I{ROOT / 'include'}", f"-I{SRC}", "-c", str(SRC / "abm_capi.cu"), "-o", "/dev/null"], capture_output=True, text=True).stderr
print("nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xptxas=-v   (W = 16 lanes per chain, blocked sum tree)")
lines = out.splitlines()
for k, ln in enumerate(lines):
    m = re.search(r"stepKernelILi(\d)ELi16ELb1ELb(\d)", ln)
    if m and "Compiling" in ln:
        props = " ".join(x.strip().replace("ptxas info    : ", "") for x in lines[k + 1:k + 4] if "Function properties" not in x)
        print(f"  stepKernel<{MODES[int(m.group(1))]},16,blocked,FIX={m.group(2)}>: {props}")


def opcode_mix(lib, mode, fix=1):
    sym = f"_ZN3abm10stepKernelILi{mode}ELi16ELb1ELb{fix}EEEvNS_10DevProblemENS_9DevChainsENS_9RunParamsE"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, str(lib)], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for ln in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            ops[m.group(1)] += 1
    return ops


for mode, name in MODES.items():
    ops = opcode_mix(LIB, mode)
    n = sum(ops.values())
    print(f"\nstepKernel<{name},16,blocked,FIX=1>: {n} SASS instructions ({n * 16 / 1024:.1f} KB), callees included")
    print("  " + ", ".join(f"{o} {c}" for o, c in ops.most_common(24)))
    mem = {k: ops[k] for k in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "REDG", "RED", "ATOMG", "ATOMS", "SHFL", "VOTE", "DADD", "DMUL", "DFMA", "MUFU",
                               "CCTL", "UBLKCP", "SYNCS")}
    print("  memory / FP64 / warp ops: " + ", ".join(f"{k} {v}" for k, v in mem.items() if v))
bulk = ROOT / "agentbasedmcmc_b200" / "libabm_bulk.so"
if bulk.exists():
    ops = opcode_mix(bulk, 0)
    print(f"\nlibabm_bulk.so (-DABM_REC_BULK=1), stepKernel<MODE_MCMC,16,blocked,FIX=1>: {sum(ops.values())} SASS instructions; "
          + ", ".join(f"{k} {ops[k]}" for k in ("UBLKCP", "SYNCS", "LDG", "LDS", "STS") if ops[k]))
