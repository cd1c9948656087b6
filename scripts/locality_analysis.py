#!/usr/bin/env python
"""Host-side analysis (no GPU): how many distinct 32-byte sectors / 128-byte lines of per-chain state the coupled columns of
a proposal touch, under the tableau's own column order and under candidate re-orderings of the basis columns.

A Metropolis step reads (and, when accepted, rewrites) one 16-byte `colRec` per column coupled to the proposed one, and adds
to one block sum per 32 columns; an L2 miss costs a 128-byte line from DRAM whatever part of it is wanted
(profiles/r1c_dram_random_probe*.json), so the number of distinct lines per proposal is what an ordering could improve.
Columns are weighted by how often the sampler proposes them: a trace of the C oracle (oracle/sbs_oracle.c, the reference's
chain restated) started from the empty trajectory, or uniformly with --uniform.

    python scripts/locality_analysis.py pp16-8 [--steps 40000] [--uniform] [--out profiles/locality_pp16-8.json]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402


def morton(x, y):
    r = np.zeros_like(x)
    for b in range(8):
        r |= (((x >> b) & 1) << (2 * b)) | (((y >> b) & 1) << (2 * b + 1))
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("problem", nargs="?", default="pp16-8")
    ap.add_argument("--steps", type=int, default=40000)
    ap.add_argument("--uniform", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    p = read_abmp(ROOT / "tests" / "golden" / f"{args.problem}.abmp.xz")
    n_tab, n_ev, _, n_cols, nnz = (int(v) for v in p["dims"])
    G, T, _, S, A = (int(v) for v in p["pp_meta"])
    M = sp.csc_matrix((np.ones(nnz, np.int32), p["nb_row"], p["nb_ptr"]), shape=(n_tab, n_cols))
    cpl = (M.T @ M).tocsr()          # coupled columns of every column, itself included
    deg = np.diff(cpl.indptr)
    if args.uniform:
        w = np.full(n_cols, 1.0 / n_cols)
        weighting = "uniform over the basis columns"
    else:
        from oracle.oracle import OracleChain, OracleProblem
        oc = OracleChain(OracleProblem(p))
        oc.seed_mt(7)
        oc.find_feasible()
        oc.step(args.steps // 4)
        j, _, _ = oc.step(args.steps, trace=True)
        w = np.bincount(j, minlength=n_cols) / float(j.size)
        weighting = f"{args.steps} proposals of an oracle chain (mt19937 seed 7) after {args.steps // 4} steps"

    def blocks_per_proposal(pos, gran):
        blk = pos[cpl.indices] // gran
        row = np.repeat(np.arange(n_cols), deg)
        key = np.unique(row.astype(np.int64) * (n_cols // gran + 2) + blk)
        per_col = np.bincount((key // (n_cols // gran + 2)).astype(np.int64), minlength=n_cols)
        return float((per_col * w).sum())

    ev = p["nb_col"].astype(np.int64)
    t, psi, a = ev // (S * A), (ev // A) % S, ev % A
    typ, cell = psi // (G * G), psi % (G * G)
    y, x = cell // G, cell % G
    orders = {"tableau order (as shipped)": np.arange(n_cols)}
    for name, keys in (("time, Morton(x,y), species, act", (a, typ, morton(x, y), t)),
                       ("time, y, x, species, act", (a, typ, x, y, t)),
                       ("Morton(x,y), time, species, act", (a, typ, t, morton(x, y)))):
        o = np.lexsort(keys)
        pos = np.empty(n_cols, np.int64)
        pos[o] = np.arange(n_cols)
        orders[name] = pos
    o = reverse_cuthill_mckee(cpl, symmetric_mode=True)
    pos = np.empty(n_cols, np.int64)
    pos[o] = np.arange(n_cols)
    orders["reverse Cuthill-McKee on the coupling graph"] = pos
    res = {"problem": args.problem, "basis_columns": n_cols, "weighting": weighting,
           "coupled_columns_per_proposal": float((deg * w).sum()), "orders": {}}
    for name, pos in orders.items():
        res["orders"][name] = {"sectors_32B (2 columns)": round(blocks_per_proposal(pos, 2), 2),
                               "lines_128B (8 columns)": round(blocks_per_proposal(pos, 8), 2),
                               "block_sums (32 columns)": round(blocks_per_proposal(pos, 32), 2)}
    print(json.dumps(res, indent=1))
    if args.out:
        Path(args.out).write_text(json.dumps(res, indent=1) + "\n")


if __name__ == "__main__":
    main()
