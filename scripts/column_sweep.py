#!/usr/bin/env python
"""Basis-column sweep (BASELINE.json configs[3]): PredPrey 64x64, T = 16, Poisson start state -- the HBM gather stress.

The cost of a Metropolis step is set by the proposed basis column: its rows (nnz), the columns coupled to it through
those rows, and the sum-tree nodes the commit touches.  This script proposes, through the replay entry point
(`abm_chains_step_traced` with injected proposals), only columns of one nnz class at a time and
reports chain-steps/s of the step kernel per class (device time, `abm_chains_last_launch_ms`), next to the free-running
throughput where the sampler draws columns from its own categorical.

    python scripts/column_sweep.py [--chains 4096] [--steps 400] [--out profiles/column_sweep.json]

Problem: tests/golden/pp64-16-poisson.abmp.xz (reference problem assembly, seed 1234, density 0.0125, basis from
abm_basis_minimise -- the reference's own TableauNormMinimiser needs hours at this size).  Measurement script; needs a GPU.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import sampler as S  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402


def coupled_counts(abmp, cols):
    """Number of distinct columns sharing a sampler row with each of `cols` (the column itself included)."""
    ptr, row = abmp["nb_ptr"].astype(np.int64), abmp["nb_row"]
    order = np.argsort(row, kind="stable")
    col_of = np.repeat(np.arange(ptr.size - 1), np.diff(ptr))[order]
    rptr = np.searchsorted(row[order], np.arange(row.max() + 2))
    out = []
    for j in cols:
        u = set()
        for i in row[ptr[j]:ptr[j + 1]]:
            u.update(col_of[rptr[i]:rptr[i + 1]].tolist())
        out.append(len(u))
    return np.array(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="pp64-16-poisson")
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--burnin", type=int, default=20000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    abmp = read_abmp(ROOT / "tests" / "golden" / f"{args.problem}.abmp.xz")
    t0 = time.time()
    P = S.Problem(abmp)
    smp = S.SparseBasisSampler(P, n_chains=args.chains, seed=99, prior_init=True)
    setup = time.time() - t0
    smp.step(args.burnin)
    smp.synchronize()
    rates = []
    for _ in range(3):
        smp.step(2000)
        rates.append(args.chains * 2000 / (smp.last_launch_ms() * 1e-3))
    rows = [{"workload": f"{args.problem}: {P.n_rows} rows x {P.n_cols} basis columns, {args.chains} chains, "
                         f"{P.info.chain_state_bytes * args.chains / 1e9:.1f} GB chain state",
             "setup_seconds": round(setup, 1), "free_running_chain_steps_per_s": float(np.median(rates)),
             "max_coupled": int(P.info.max_coupled), "tree_depth": int(P.info.tree_depth)}]
    print(json.dumps(rows[0]), flush=True)
    nnz = np.diff(abmp["nb_ptr"])
    rng = np.random.default_rng(3)
    for v in np.unique(nnz):
        cand = np.flatnonzero(nnz == v)
        if cand.size < 64:
            continue
        props = rng.choice(cand, size=(args.chains, args.steps)).astype(np.int32)
        smp.step_traced(args.steps, proposals=props)       # warm-up of this class
        props = rng.choice(cand, size=(args.chains, args.steps)).astype(np.int32)
        j, acc, _ = smp.step_traced(args.steps, proposals=props)
        ms = smp.last_launch_ms()
        assert np.array_equal(j, props)
        cc = coupled_counts(abmp, rng.choice(cand, size=min(200, cand.size), replace=False))
        row = {"column_nnz": int(v), "columns": int(cand.size), "share_of_basis": round(cand.size / nnz.size, 4),
               "mean_coupled_columns": round(float(cc.mean()), 1), "max_coupled_columns_sampled": int(cc.max()),
               "chain_steps_per_s": args.chains * args.steps / (ms * 1e-3), "acceptance": round(float(acc.mean()), 3)}
        print(json.dumps(row), flush=True)
        rows.append(row)
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(r) for r in rows) + "\n")


if __name__ == "__main__":
    main()
