#!/usr/bin/env python
"""Paper-scale posterior fixture from the UNMODIFIED reference: long reference CPU chains on the 32x32x16 problem with the
paper's protocol (FiguresForPaper.h:40-104: every chain from its own prior sample, 1e6 feasible samples of burn-in, then the
samples mapped by the reference's own ModelState(traj,T,T) and synopsis()).  `abm_ref posterior` does the sampling
(oracle/ref/abm_ref.cpp); this script reduces its output to what the tests and bench.py need:

  syn_mean / syn_se [d]       posterior mean of the synopsis statistics and its standard error (batch means, 2e6-sample batches)
  endstate_mean / _se [S]     posterior mean occupancy of every agent state at t = T (standard error: spread of the chain means)
  agents_mean / _se           total agents at T
  tau_int / n_eff / rhat [d]  the reference's estimators (MultiChainStats.h:81-181) on the half-chains, from every 200th sample
  chain_means [chains][d], burnin_curve [batches][d] (mean synopsis per 1e5 burn-in samples: the relaxation from the start)
  acceptance, infeasible_fraction, steps_per_sample, meta

    make -C oracle ref && python tests/golden/make_posterior_golden.py [--threads 6] [--samples 20000000] [--runs 2]
    python tests/golden/make_posterior_golden.py --from run_a.abmp run_b.abmp       (reduce existing `abm_ref posterior` outputs)

Two runs of 6 chains x 2e7 samples take ~35 minutes each on 6 cores.
"""
import lzma
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from agentbasedmcmc_b200 import diagnostics as D  # noqa: E402
from agentbasedmcmc_b200.abmp import read_abmp, write_abmp  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "abm_ref"
NAME = "pp32-16"


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    if "--from" in sys.argv:
        paths = [Path(a) for a in sys.argv[sys.argv.index("--from") + 1:]]
    else:
        threads, samples, runs = arg("--threads", 6), arg("--samples", 20000000), arg("--runs", 2)
        tmp = Path(tempfile.mkdtemp())
        prob = tmp / f"{NAME}.abmp"
        prob.write_bytes(lzma.decompress((HERE / f"{NAME}.abmp.xz").read_bytes()))
        paths = []
        for r in range(runs):
            out = tmp / f"posterior_{r}.abmp"
            subprocess.run([str(REF), "posterior", str(prob), str(threads), str(samples), "1000000", "100000", "200", str(out), str(100 * r)],
                           check=True)
            paths.append(out)
    runs = [read_abmp(p) for p in paths]
    meta = runs[0]["meta"]
    n_s, burn, batch, thin, d, S = (int(meta[k]) for k in (1, 2, 3, 4, 5, 6))
    n_t = [int(r["meta"][0]) for r in runs]
    chains = sum(n_t)
    cat = lambda key, shape: np.concatenate([r[key].reshape((t,) + shape) for r, t in zip(runs, n_t)])   # noqa: E731
    syn_sum, syn_sq = cat("syn_sum", (d,)), cat("syn_sumsq", (d,))
    bm = cat("batch_means", (n_s // batch, d))
    burn_bm = cat("burnin_batch_means", (burn // batch, d))
    end = cat("endstate_sum", (S,)) / n_s
    series = cat("series_thin", (-1, d)).astype(np.float64)
    timing, counters = cat("timing", (4,)), cat("counters", (8,))
    big = 20                                                       # batches of 20 x 1e5 = 2e6 samples: ~10 autocorrelation times
    bb = bm[:, : (bm.shape[1] // big) * big].reshape(chains, -1, big, d).mean(axis=2).reshape(-1, d)
    half = series.shape[1] // 2
    seqs = [series[c, h * half:(h + 1) * half] for c in range(chains) for h in range(2)]   # the reference's half-chains
    ms = D.MultiChainStats([D.ChainStats.from_series(x, n_lags=200, max_lag_proportion=0.5) for x in seqs])
    n_eff = ms.effective_samples(reference_integer_division=False)   # in units of the thinned series' sequences x their length ...
    n_eff_samples = n_eff                                            # ... which is a COUNT of effective samples, whatever the thinning
    prop, acc = counters[:, :4].sum(0), counters[:, 4:].sum(0)
    agents = end.sum(axis=1)
    out = {
        "syn_mean": syn_sum.sum(0) / (chains * n_s), "syn_se": bb.std(axis=0, ddof=1) / np.sqrt(bb.shape[0]),
        "syn_sd": np.sqrt((syn_sq.sum(0) - syn_sum.sum(0) ** 2 / (chains * n_s)) / (chains * n_s - 1)),
        "endstate_mean": end.mean(axis=0), "endstate_se": end.std(axis=0, ddof=1) / np.sqrt(chains),
        "agents_mean": np.array([agents.mean()]), "agents_se": np.array([agents.std(ddof=1) / np.sqrt(chains)]),
        "tau_int": chains * n_s / n_eff_samples, "n_eff": n_eff_samples,
        "rhat": ms.potential_scale_reduction(reference_integer_division=False),
        "chain_means": syn_sum / n_s, "burnin_curve": burn_bm.mean(axis=0),
        "acceptance": np.array([acc.sum() / prop.sum()]),
        "infeasible_fraction": np.array([(prop[0] + acc[2] + prop[1] - acc[1]) / prop.sum()]),
        "steps_per_sample": np.array([timing[:, 1].sum() / (chains * n_s)]),
        "steps_per_second_per_thread": np.array([(timing[:, 1] / timing[:, 0]).mean()]),
        "meta": np.array([chains, n_s, burn, batch, thin, d, S], np.float64),
    }
    write_abmp(HERE / f"posterior_{NAME}.abmp.xz", out)
    for k in ("syn_mean", "syn_se", "tau_int", "n_eff", "rhat", "agents_mean", "agents_se", "acceptance", "infeasible_fraction", "steps_per_sample"):
        print(k, np.round(out[k], 5))


if __name__ == "__main__":
    main()
