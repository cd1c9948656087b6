#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory from the UNMODIFIED reference.

Needs /root/reference (read-only) and a host compiler; run from the repository root:
    make -C oracle ref && python tests/golden/make_golden.py [--with-32]
Problems: FiguresForPaper::generateStandardProblemFile protocol (FiguresForPaper.h:25-37): seed 1234,
pPredator = pPrey = 0.05, pMakeObservation = 0.05, pObserveIfPresent = 1.0, kappa = 10.
Traces: oracle/ref/abm_ref.cpp `trace` (initial state = prior.nextSample(false) under initSeed; sampler
constructed and stepped under chainSeed with the reference's own mt19937 Random::gen).
"""
import lzma
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = HERE.parent.parent / "oracle" / "_ref" / "abm_ref"


def xz(src: Path, dst: Path):
    dst.write_bytes(lzma.compress(src.read_bytes(), preset=9 | lzma.PRESET_EXTREME))
    print(f"{dst.name}: {dst.stat().st_size} bytes")


def main():
    with32 = "--with-32" in sys.argv
    tmp = Path(tempfile.mkdtemp())
    problems = [(8, 4), (16, 8)] + ([(32, 16)] if with32 else [])
    for g, t in problems:
        raw = tmp / f"pp{g}-{t}.abmp"
        subprocess.run([str(REF), "gen", str(g), str(t), "1234", "10.0", str(raw)], check=True, stdout=subprocess.DEVNULL)
        xz(raw, HERE / f"pp{g}-{t}.abmp.xz")
    # (problem, initSeed, chainSeed, nSteps, checkpointEvery)
    traces = [("pp8-4", 11, 101, 20000, 500), ("pp8-4", 12, 102, 20000, 500), ("pp16-8", 12, 202, 20000, 500)]
    for name, s0, s1, n, every in traces:
        raw = tmp / f"trace_{name}_{s1}.abmp"
        subprocess.run([str(REF), "trace", str(tmp / f"{name}.abmp"), str(s0), str(s1), str(n), str(every), str(raw)], check=True)
        xz(raw, HERE / f"trace_{name}_{s1}.abmp.xz")
    # prior samples (prior.nextSample(false), FiguresForPaper.h:56): per (t, type, act) mean of [event > 0] over 20000 draws
    import numpy as np
    sys.path.insert(0, str(HERE.parent.parent))
    from agentbasedmcmc_b200.abmp import write_abmp
    raw = tmp / "prior8-4.i8"
    subprocess.run([str(REF), "init", str(tmp / "pp8-4.abmp"), "20000", "4711", str(raw)], check=True)
    x = np.fromfile(raw, dtype=np.int8).reshape(20000, 4, 2, 64, 6)
    write_abmp(HERE / "prior_pp8-4.abmp.xz", {"mean_indicator": (x > 0).mean(axis=(0, 3)).ravel(),
                                              "mean_count": x.astype(np.float64).mean(axis=(0, 3)).ravel(),
                                              "n": np.array([20000, 64], np.int32)})
    # long reference chains for posterior marginals (64 chains x 60000 feasible samples after 60000 burn-in)
    raw = tmp / "marginals_pp8-4.abmp"
    subprocess.run([str(REF), "marginals", str(tmp / "pp8-4.abmp"), "64", "60000", "60000", str(raw)], check=True)
    xz(raw, HERE / "marginals_pp8-4.abmp.xz")


if __name__ == "__main__":
    main()
