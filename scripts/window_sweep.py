#!/usr/bin/env python
"""Data-assimilation window sweep (BASELINE.json configs[4]): PredPrey 32x32, T in {4, 8, 16, 32}.

For every window length: chain-steps/s at the stationary state, ESS/s and R-hat of the GPU ensemble (8192 chains, each from its
own prior sample, burnt in; the reference's estimator over all chains, agentbasedmcmc_b200/diagnostics.py
`ensemble_effective_samples`), next to long chains of the UNMODIFIED reference on the same problem (`abm_ref posterior`,
reduced by `--reduce-reference`; they are run in the build container, their summary is committed as
profiles/r2_window_sweep_reference.json).  The mean synopsis of the two is the posterior cross-check; the relaxation of the GPU
ensemble during its burn-in is logged so that convergence can be judged.

    python scripts/window_sweep.py --reduce-reference /tmp/post/post32-4.abmp:4 /tmp/post/post32-8.abmp:8 ...   (CPU, no GPU)
    python scripts/window_sweep.py [--T 4 8 16 32] [--chains 8192] [--out gpurun_out/window_sweep.json]            (GPU)

Problems: tests/golden/pp32-{4,8,16,32}.abmp.xz (reference generator, seed 1234; T = 32 at density 0.02, see
tests/golden/make_golden.py).  A measurement script, not a test.
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
from agentbasedmcmc_b200 import diagnostics as D  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp  # noqa: E402

REF_JSON = ROOT / "profiles" / "r2_window_sweep_reference.json"
# T -> (burn-in Metropolis steps per chain, feasible samples per chain of the ESS window)
PLAN = {4: (1500000, 400000), 8: (4000000, 800000), 16: (7000000, 1600000), 32: (12000000, 100000)}


def reduce_reference(specs):
    out = json.loads(REF_JSON.read_text()) if REF_JSON.exists() else {}
    for spec in specs:
        path, T = spec.rsplit(":", 1)
        r = read_abmp(path)
        meta = r["meta"]
        n_t, n_s, burn, batch, thin, d, S = (int(meta[k]) for k in range(7))
        series = r["series_thin"].reshape(n_t, -1, d).astype(np.float64)
        half = series.shape[1] // 2
        ms = D.MultiChainStats([D.ChainStats.from_series(series[c, h * half:(h + 1) * half], n_lags=200, max_lag_proportion=0.5)
                                for c in range(n_t) for h in range(2)])
        n_eff = ms.effective_samples(reference_integer_division=False)
        tm = r["timing"].reshape(n_t, 4)
        bm = r["batch_means"].reshape(n_t, -1, d)
        big = max(1, bm.shape[1] // 10)
        bb = bm[:, : (bm.shape[1] // big) * big].reshape(n_t, -1, big, d).mean(axis=2).reshape(-1, d)
        ctr = r["counters"].reshape(n_t, 8).sum(0)
        prop, acc = ctr[:4], ctr[4:]
        sps = float(tm[:, 1].sum() / (n_t * n_s))
        tau = n_t * n_s / n_eff
        out[str(T)] = {"chains": n_t, "samples_per_chain": n_s, "burnin_samples": burn,
                       "synopsis_mean": (r["syn_sum"].reshape(n_t, d).sum(0) / (n_t * n_s)).tolist(),
                       "synopsis_se": (bb.std(axis=0, ddof=1) / np.sqrt(bb.shape[0])).tolist(),
                       "samples_per_effective_sample": tau.tolist(), "rhat": ms.potential_scale_reduction(reference_integer_division=False).tolist(),
                       "steps_per_sample": sps, "steps_per_effective_sample": float(tau.max() * sps),
                       "acceptance": float(acc.sum() / prop.sum()),
                       "infeasible_fraction": float((prop[0] + acc[2] + prop[1] - acc[1]) / prop.sum()),
                       "chain_steps_per_s_per_thread_in_the_build_container": float((tm[:, 1] / tm[:, 0]).mean()),
                       "burnin_curve_synopsis0": r["burnin_batch_means"].reshape(n_t, -1, d).mean(axis=0)[:, 0].tolist()}
        print(T, json.dumps({k: out[str(T)][k] for k in ("synopsis_mean", "synopsis_se", "samples_per_effective_sample", "rhat", "steps_per_sample")}))
    REF_JSON.write_text(json.dumps(out, indent=1))


def gpu_arm(T, n_chains, burnin, ess_samples, mh, reps):
    from agentbasedmcmc_b200 import sampler as S
    abmp = read_abmp(ROOT / "tests" / "golden" / f"pp32-{T}.abmp.xz")
    t0 = time.time()
    P = S.Problem(abmp)
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=777, prior_init=True)
    setup = time.time() - t0
    relax, done, interval = [], 0, max(250000, burnin // 16)
    while done < burnin:   # burn-in in sampling launches, so that the relaxation of the ensemble mean is on record
        k = min(interval, burnin - done)
        smp.reset_statistics()
        smp.sample(10 ** 12, max_steps=k)
        end, ss, sq, ns = smp.statistics()
        done += k
        relax.append({"steps": done, "synopsis_mean": (ss.sum(0) / max(ns.sum(), 1)).tolist(), "chains_without_sample": float((ns == 0).mean())})
        print(T, relax[-1], flush=True)
    rates = []
    for _ in range(reps):
        smp.step(mh)
        rates.append(n_chains * mh / (smp.last_launch_ms() * 1e-3))
    c0 = smp.counters().sum(0)
    thin = max(1, ess_samples // 400)
    smp.reset_statistics()
    smp.record_series(n_chains, thin, ess_samples // thin + 1)
    smp.synchronize()
    t0 = time.perf_counter()
    smp.sample(ess_samples, max_steps=20 * ess_samples)
    smp.synchronize()
    ess_s = time.perf_counter() - t0
    end, ss, sq, ns = smp.statistics()
    c = smp.counters().sum(0) - c0
    full = ns >= ess_samples
    k_pts = ess_samples // thin
    series = smp.series()[full, :k_pts].astype(np.float64)
    r = D.ensemble_effective_samples(series, thin, ns[full], ss[full], sq[full], group=8)
    half = k_pts // 2
    drift = (series[:, half:].mean(axis=(0, 1)) - series[:, :half].mean(axis=(0, 1))) / np.sqrt(r["var_plus"])
    prop, acc = c[:4], c[4:]
    cm = ss[full] / ns[full, None]
    return {"chain_steps_per_s": float(np.median(rates)), "setup_seconds": round(setup, 1), "chains": n_chains,
            "chain_state_gb": round(P.info.chain_state_bytes * n_chains / 1e9, 2), "n_cols": P.n_cols,
            "burnin_steps": burnin, "relaxation": relax, "ess_per_s": float(r["n_eff"].min() / ess_s),
            "samples_per_effective_sample": r["tau"].tolist(), "feasible_samples_per_chain": ess_samples, "chains_with_full_window": int(full.sum()),
            "ess_seconds": round(ess_s, 2), "rhat": r["rhat"].tolist(), "rhat_pooled8": r["rhat_grouped"].tolist(),
            "drift_second_minus_first_half_in_sd": drift.tolist(),
            "synopsis_mean": (ss[full].sum(0) / ns[full].sum()).tolist(), "synopsis_se": (cm.std(axis=0, ddof=1) / np.sqrt(full.sum())).tolist(),
            "acceptance": float(acc.sum() / prop.sum()), "infeasible_fraction": float((prop[0] + acc[2] + prop[1] - acc[1]) / prop.sum())}


def main():
    if "--reduce-reference" in sys.argv:
        return reduce_reference(sys.argv[sys.argv.index("--reduce-reference") + 1:])
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, nargs="+", default=[4, 8, 16, 32])
    ap.add_argument("--chains", type=int, default=8192)
    ap.add_argument("--mh-steps", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "window_sweep.json"))
    args = ap.parse_args()
    ref = json.loads(REF_JSON.read_text()) if REF_JSON.exists() else {}
    rows = []
    for T in args.T:
        burnin, ess_samples = PLAN[T]
        row = {"workload": f"PredPrey 32x32, T={T}", "gpu": gpu_arm(T, args.chains, burnin, ess_samples, args.mh_steps, args.reps),
               "reference_cpu": ref.get(str(T))}
        if row["reference_cpu"]:
            g, c = row["gpu"], row["reference_cpu"]
            row["synopsis_z"] = ((np.array(g["synopsis_mean"]) - np.array(c["synopsis_mean"])) /
                                 np.sqrt(np.array(g["synopsis_se"]) ** 2 + np.array(c["synopsis_se"]) ** 2)).tolist()
        rows.append(row)
        print(json.dumps({k: v for k, v in row.items() if k != "gpu"} | {"gpu": {k: v for k, v in row["gpu"].items() if k != "relaxation"}}), flush=True)
        Path(args.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
