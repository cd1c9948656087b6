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


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: Path = OUT) -> Path:
    """``defines``/``out`` build an experimental variant next to the product library (kernel A/B runs select it with
    the ABM_LIB environment variable, see _capi.py)."""
    sources = [SRC / "abm_capi.cu", SRC / "abm_basis.cpp"]
    deps = list(SRC.glob("*")) + [ROOT / "include" / "abmcmc_b200.h"]
    if not force and out.exists() and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return out
    cmd = ["nvcc", *NVCC_FLAGS, *(f"-D{d}" for d in defines), f"-I{ROOT / 'include'}", f"-I{SRC}", *(str(s) for s in sources),
           "-o", str(out)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="-v" in sys.argv, defines=defs,
                out=PKG / outs[0] if outs else OUT))
