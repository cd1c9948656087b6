"""Builds agentbasedmcmc_b200/libabmcmc_b200.so for sm_100a with nvcc (in-tree, so it travels to the GPU box)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
SRC = PKG / "csrc"
OUT = PKG / "libabmcmc_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # no FMA contraction: the reference's double arithmetic is reproduced operation by operation
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def build(force: bool = False, verbose: bool = False) -> Path:
    sources = [SRC / "abm_capi.cu"]
    deps = list(SRC.glob("*")) + [ROOT / "include" / "abmcmc_b200.h"]
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, f"-I{ROOT / 'include'}", f"-I{SRC}", *(str(s) for s in sources), "-o", str(OUT)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
