#!/usr/bin/env python
"""Benchmark of the SparseBasisSampler hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:       one process per GPU (torchrun for N>1).  A bench "step" = one launch advancing every chain of this
            rank by --mh-steps Metropolis steps (SparseBasisSampler.h:221-246).  `value` = chain-steps/s summed
            over ranks with the chain state resident in HBM; `e2e` = the same through the host-facing C-ABI
            sequence (step + statistics read-back into host buffers) timed with the host clock.
reference:  the unmodified reference sampler (oracle/_ref/abm_ref, built from /root/reference by
            `make -C oracle ref`) on all host cores, one chain per thread, same problem instance.
"""
from __future__ import annotations

import argparse
import json
import lzma
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# workload name -> (fixture, chains per GPU, algorithmic bytes per chain-step at reference data widths,
# SURVEY.md section 8d / BASELINE.md section 3.4)
WORKLOADS = {
    "pp32-16": ("pp32-16.abmp.xz", 8192, 6314.0, "PredPrey 32x32 grid, 16 timesteps, 8192 chains/GPU (paper scale, 65536 chains over 8 GPUs)"),
    "pp16-8": ("pp16-8.abmp.xz", 4096, 5329.0, "PredPrey 16x16 grid, 8 timesteps, 4096 chains/GPU"),
    "pp8-4": ("pp8-4.abmp.xz", 1024, 3567.0, "PredPrey 8x8 grid, 4 timesteps, 1024 chains/GPU"),
}
# workload -> (burn-in Metropolis steps per chain before anything is timed, feasible samples per chain of the ESS window)
STATIONARY = {"pp32-16": (8000000, 1600000), "pp16-8": (1500000, 300000), "pp8-4": (200000, 100000)}
METRIC = "chain-steps/sec, 32x32 PredPrey (SparseBasisSampler Metropolis steps, all chains, all GPUs)"
UNIT = "chain-steps/s"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows)}


def fixture_to_tempfile(name: str) -> str:
    raw = lzma.decompress((ROOT / "tests" / "golden" / name).read_bytes())
    f = tempfile.NamedTemporaryFile(prefix="abmp_", suffix=".abmp", delete=False)
    f.write(raw)
    f.close()
    return f.name


def run_reference_binary(fixture: str, threads: int, samples_per_step: int, warmup: int, steps: int):
    exe = ROOT / "oracle" / "_ref" / "abm_ref"
    if not exe.exists():
        return None
    path = fixture_to_tempfile(fixture)
    try:
        out = subprocess.run([str(exe), "bench", path, str(threads), str(samples_per_step), str(warmup), str(steps)],
                             capture_output=True, text=True, timeout=1500)
        if out.returncode != 0:
            return None
        return json.loads(out.stdout.strip().splitlines()[-1])
    finally:
        os.unlink(path)


def reference_tau(workload: str):
    """Metropolis steps per effective sample of the stationary reference chain (worst synopsis statistic): its autocorrelation
    time in feasible samples (reference estimator) x its steps per feasible sample, both from the long reference run behind the
    paper-scale posterior fixture (tests/golden/make_posterior_golden.py); None when there is no such fixture."""
    p = ROOT / "tests" / "golden" / f"posterior_{workload}.abmp.xz"
    if not p.exists():
        return None
    from agentbasedmcmc_b200.abmp import read_abmp
    f = read_abmp(p)
    return float(f["tau_int"].max() * f["steps_per_sample"][0])


def cpu_baseline(workload: str, threads: int, seconds_budget: float = 20.0):
    """Reference CPU sampler on this box's host cores on a bounded sample of the same workload."""
    fixture = WORKLOADS[workload][0]
    # ~13 us/step/core and 1-4 steps per feasible sample: size the sample for ~seconds_budget of timed work
    samples = {"pp32-16": 60000, "pp16-8": 400000, "pp8-4": 600000}[workload]
    samples = int(samples * seconds_budget / 20.0)
    res = run_reference_binary(fixture, threads, samples, 0, 1)
    if res is not None:
        v = res["steps_per_s"][0]
        return {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{threads} chains (one per host thread) x {samples} feasible samples = {res['chain_steps'][0]} Metropolis steps, "
                          f"unmodified reference SparseBasisSampler (oracle/_ref/abm_ref), every thread timed on its own loop, "
                          f"rates summed, {res['step_seconds'][0]:.1f}s"}
    # the C restatement (single thread) when the compiled reference is not on the box
    from agentbasedmcmc_b200.abmp import read_abmp
    from oracle.oracle import OracleChain, OracleProblem
    p = read_abmp(ROOT / "tests" / "golden" / fixture)
    oc = OracleChain(OracleProblem(p))
    oc.seed_mt(1)
    oc.find_feasible()
    n = {"pp32-16": 400000, "pp16-8": 1500000, "pp8-4": 2000000}[workload]
    t0 = time.time()
    oc.step(n)
    dt = time.time() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 chain x {n} Metropolis steps, oracle/sbs_oracle.c, {dt:.1f}s timed"}


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    fixture, _, _, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    samples = {"pp32-16": 20000, "pp16-8": 100000, "pp8-4": 200000}[args.workload]
    res = run_reference_binary(fixture, threads, samples, args.warmup, args.steps)
    ess = None
    if res is None:
        cb = cpu_baseline(args.workload, 1)
        value, ms, kind, cores, sample = cb["value"], None, cb["kind"], cb["cores"], cb["sample"]
    else:
        # every thread times its own loop and a round's rate is the sum of the per-thread rates (abm_ref bench), so a chain that
        # needs many steps per feasible sample does not hold the others' clocks
        value = float(np.mean(res["steps_per_s"]))
        ms = 1e3 * float(np.mean(res["step_seconds"]))
        tot_steps = sum(res["chain_steps"])
        kind, cores = "reference", threads
        sample = (f"{threads} chains (one per host thread), {samples} feasible samples per chain per step, {tot_steps} Metropolis steps "
                  f"timed, per-thread rates summed")
        tau = reference_tau(args.workload)
        if tau:
            ess = {"value": value / tau, "unit": "effective samples/s (worst synopsis statistic)", "derived": True,
                   "how": f"{value:.0f} chain-steps/s measured here / {tau:.0f} Metropolis steps per effective sample of the stationary "
                          f"reference chain (its autocorrelation time in feasible samples x its steps per feasible sample, from the long "
                          f"reference run behind tests/golden/posterior_{args.workload}.abmp.xz); the chains timed here are still "
                          f"relaxing from their start, a short CPU run cannot measure an autocorrelation time of ~1e5 samples"}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": desc, "host_threads": cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "ess": ess}
    print(json.dumps(line))
    return 0


def bench_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line: anything libraries print on fd 1 meanwhile (the NCCL version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from agentbasedmcmc_b200 import diagnostics as D
    from agentbasedmcmc_b200 import ensemble as E
    from agentbasedmcmc_b200 import sampler as S
    from agentbasedmcmc_b200.abmp import read_abmp

    fixture, default_chains, bytes_per_step, desc = WORKLOADS[args.workload]
    n_chains = args.chains or default_chains
    mh = args.mh_steps
    abmp = read_abmp(ROOT / "tests" / "golden" / fixture)
    stream = torch.cuda.current_stream()
    t_setup = time.time()
    P = S.Problem(abmp, device=local_rank)
    # every chain starts from its own prior sample, as FiguresForPaper.h:56 / :162 do (drawn on the device)
    smp = S.SparseBasisSampler(P, n_chains=n_chains, seed=args.seed, first_chain_id=rank * n_chains, stream=stream.cuda_stream,
                               prior_init=True)
    setup_s = time.time() - t_setup
    # the library's own NCCL communicator (abm_comm_*): the ensemble reduction goes through the C-ABI, as a C++ host's would
    comm = E.Comm(local_rank) if world > 1 else None
    n_syn = P.n_synopsis

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_launches(n_launch):
        """n_launch launches of `mh` Metropolis steps, CUDA events on the launch stream; (total ms, per-launch ms)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_launch + 1)]
        ev[0].record(stream)
        for k in range(n_launch):
            smp.step(mh)
            ev[k + 1].record(stream)
        barrier()
        return ev[0].elapsed_time(ev[-1]), [ev[k].elapsed_time(ev[k + 1]) for k in range(n_launch)]

    def counters_delta(prev):
        c = smp.counters().sum(0)
        d = c - prev
        prop, acc = d[:4], d[4:]
        tot = max(int(prop.sum()), 1)
        return c, float(acc.sum()) / tot, float(prop[0] + acc[2] + prop[1] - acc[1]) / tot   # (counters, acceptance, infeasible fraction)

    # ---- throughput straight after the feasibility search (the chains are far from the posterior: more agents, more infeasible
    # states), kept for comparison with the stationary figure below
    for _ in range(3):
        smp.step(mh)
    barrier()
    c0 = smp.counters().sum(0)
    ms_post_init, _ = timed_launches(5)
    c0, acc_post_init, inf_post_init = counters_delta(c0)
    value_post_init = float(n_chains) * mh * 5 / (ms_post_init * 1e-3)

    # ---- burn-in to the stationary state (the reference burns 1e6 feasible samples ~ 3.3e6 steps in, FiguresForPaper.h:89;
    # the ensemble mean of the synopsis is flat to within 2 % from ~6e6 steps, profiles/r2_relaxation_*.json)
    t_burn = time.time()
    done = 0
    while done < args.burnin_steps:
        k = min(500000, args.burnin_steps - done)
        smp.step(k)
        done += k
    barrier()
    burn_s = time.time() - t_burn

    # ---- ESS/s and R-hat at the stationary state (the second half of BASELINE.json's metric): every chain takes the same
    # number of feasible samples (operator() semantics); the reference's estimator (MultiChainStats::effectiveSamples,
    # diagnostics/MultiChainStats.h:81-102) with its lag grid (200 strided lags to half the sequence, FiguresForPaper.h:89-90),
    # the variogram averaged over ALL chains from every stride-th sample, var+ from the exact per-chain moments.
    ess = None
    if args.ess_samples > 0:
        try:
            n_s = args.ess_samples
            thin = max(1, n_s // 400)
            smp.reset_statistics()
            smp.record_series(n_chains, thin, n_s // thin + 1)
            barrier()
            t0 = time.perf_counter()
            smp.sample(n_s, max_steps=12 * n_s)      # bounded: a chain deep in the infeasible region must not stall the batch
            barrier()
            ess_s = time.perf_counter() - t0
            k_pts = n_s // thin
            n_lags = max(2, min(k_pts, int(0.5 * n_s / thin) + 1))
            # the ensemble's figures, over ALL chains of ALL GPUs, reduced on the devices: exact variogram sums and the additive
            # moments, both summed over the ranks by NCCL inside the library
            vsum, m_full = smp.series_variogram(k_pts, n_lags, n_s, comm)
            exact, real = smp.moments(comm)
            summ = E.summarize_moments(exact, real, n_syn)
            V = vsum.astype(np.float64) / (max(m_full, 1) * (k_pts - np.arange(n_lags))[:, None])
            n_eff_all, rho = D.effective_samples_from_variogram(V, summ.var_plus, summ.n_samples / max(summ.n_chains, 1), summ.n_chains, thin)
            # rank 0's own chains: R-hat of pooled sequences and the drift between the halves of the window
            end, ss, sq, ns = smp.statistics()
            ser = smp.series()
            smp.record_series(0, 1, 0)               # stop recording
            full = ns >= n_s
            series = ser[full, :k_pts].astype(np.float64)
            r = D.ensemble_effective_samples(series, thin, ns[full], ss[full], sq[full], group=8)
            half = k_pts // 2
            drift = (series[:, half:].mean(axis=(0, 1)) - series[:, :half].mean(axis=(0, 1))) / np.sqrt(r["var_plus"])
            te_ = torch.tensor([ess_s], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te_, op=dist.ReduceOp.MAX)   # the slowest rank sets the time
            ess_s_all = float(te_.item())
            ess = {"value": float(n_eff_all.min() / ess_s_all), "unit": "effective samples/s (worst synopsis statistic, all chains, all GPUs)",
                   "n_eff": [round(float(x), 1) for x in n_eff_all], "chains": int(summ.n_chains), "chains_with_full_window": int(m_full),
                   "feasible_samples_per_chain": n_s, "seconds": round(ess_s_all, 2),
                   "samples_per_effective_sample": [round(float(x)) for x in summ.n_samples / n_eff_all],
                   "rhat": [round(float(x), 4) for x in summ.rhat],
                   "rhat_note": "per-chain sequences of all GPUs, reference formula; for stationary chains R-hat^2 = 1 + 1/(effective samples per sequence)",
                   "rhat_pooled8": [round(float(x), 4) for x in r["rhat_grouped"]],
                   "drift_second_minus_first_half_in_sd": [round(float(x), 4) for x in drift],
                   "synopsis_mean": [round(float(x), 4) for x in summ.synopsis_mean],
                   "estimator": "MultiChainStats::effectiveSamples (reference); variogram (abm_chains_series_variogram, exact integer sums on the "
                                "devices) over all chains at stride %d, lags to n/2; var+ from the device-reduced moments; both summed over the "
                                "GPUs by NCCL inside the library" % thin,
                   "scope": "all chains of all GPUs; rhat_pooled8 and the drift: rank 0's chains"}
        except Exception as exc:   # the ESS estimate must never take the benchmark line down
            ess = {"value": None, "note": f"ESS estimate failed: {exc!r}"}
    barrier()

    # ---- device-resident throughput at the stationary state: W warm-up launches, then exactly K timed launches.  The ESS window
    # stops every chain AT a feasible state (operator() semantics); 2e5 steps let the feasible / infeasible mix settle again.
    if args.ess_samples > 0:
        smp.step(200000)
    for _ in range(args.warmup):
        smp.step(mh)
    barrier()
    c0 = smp.counters().sum(0)
    with ClockSampler(local_rank) as clocks:
        total_ms, launch_ms = timed_launches(args.steps)
    c0, acc_stat, inf_stat = counters_delta(c0)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    steps_done = float(world) * n_chains * mh * args.steps
    value = steps_done / (total_ms * 1e-3)

    # ---- end to end through the host-facing API: a sampling launch (steps + on-device statistics), the ensemble moments
    # reduced on the device, summed over the GPUs by NCCL inside the library and copied to host buffers -- every step, host clock.
    # Launch k+1 is queued behind the reduction of launch k, so the device never waits for the host.
    smp.reset_statistics()
    smp.enqueue_moments(comm)   # untimed: the communicator's first collective sets up its connections
    smp.wait_moments()
    barrier()
    t0 = time.perf_counter()
    summary = None
    d2h = 0
    smp.sample(10 ** 12, max_steps=mh)          # mh Metropolis steps per chain, statistics of every feasible state
    smp.enqueue_moments(comm)                   # device reduction + NCCL all-reduce + copies into pinned host memory, asynchronous
    for k in range(args.steps):
        if k + 1 < args.steps:
            smp.sample(10 ** 12, max_steps=mh)
        exact, real = smp.wait_moments()        # the ensemble's additive moments after launch k, in host buffers
        if k + 1 < args.steps:
            smp.enqueue_moments(comm)
        d2h = exact.nbytes + real.nbytes
        summary = E.summarize_moments(exact, real, n_syn)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = steps_done / float(te.item())

    if rank == 0:
        peak, peak_src = measured_peaks()
        kernel_ms = float(np.mean(launch_ms))
        achieved = n_chains * mh * bytes_per_step / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "dram_bytes_per_chain_step.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            if args.workload in tj:   # ncu DRAM bytes per chain-step (measured at 8192 chains) x chain-steps per launch
                traffic = tj[args.workload] * n_chains * mh
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "chains_per_gpu": n_chains, "mh_steps_per_launch": mh, "rng": "philox4x32-10", "initial_states": "one prior sample per chain, drawn on the device",
                       "l2_policy": f"per-GPU chain state {P.info.chain_state_bytes * n_chains / 1e9:.1f} GB >> 126 MB L2 (inputs larger than L2)",
                       "problem": "reference generator seed 1234, kappa 10 (tests/golden)", "setup_seconds": round(setup_s, 1),
                       "state": f"stationary: timed after {args.burnin_steps} burn-in Metropolis steps per chain ({burn_s:.0f} s) and the ESS window "
                                f"({args.ess_samples} feasible samples per chain)",
                       "acceptance": round(acc_stat, 4), "infeasible_fraction": round(inf_stat, 4),
                       "post_init": {"value": value_post_init, "acceptance": round(acc_post_init, 4), "infeasible_fraction": round(inf_post_init, 4),
                                     "what": "this rank's chain-steps/s straight after the feasibility search, before the burn-in"}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_chain_step": bytes_per_step, "kernel": "abm::stepKernel<MODE_MCMC>",
                         "kernel_ms_per_launch": kernel_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                    "what": "abm_chains_sample (steps + on-device end-state/synopsis statistics) + abm_chains_enqueue_moments / "
                            "abm_chains_wait_moments (device reduction of the ensemble moments, NCCL all-reduce inside the library, "
                            "copy into host buffers), host clock, every step; launch k+1 is queued behind the reduction of launch k; "
                            "a step has no per-step input to copy in: its inputs are kernel arguments, the problem and the chain "
                            "states are uploaded once, at construction (h2d_bytes_per_step is 0 by the nature of the workload)"},
            "ensemble": {"chains": summary.n_chains, "feasible_samples": summary.n_samples,
                         "mean_agents_at_T": float(summary.end_state_mean.sum()),
                         "synopsis_mean": [round(float(x), 4) for x in summary.synopsis_mean]},
            "ess": ess,
            # per bench step of the timed region: refreshBlockedTopKernel (0.16 ms) + stepKernel<MODE_MCMC> (profiles/*_bench_launches.csv)
            "gpu_launches": 2 * args.steps,
            "clocks": clocks.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.workload, os.cpu_count() or 1)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pp32-16", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--mh-steps", type=int, default=2000, help="Metropolis steps per chain per launch")
    ap.add_argument("--seed", type=int, default=20221)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--burnin-steps", type=int, default=-1,
                    help="Metropolis steps per chain before anything is timed (the ensemble is stationary from ~6e6 at 32x32x16)")
    ap.add_argument("--ess-samples", type=int, default=-1,
                    help="feasible samples per chain of the ESS / R-hat window (0 = skip); ~1 minute per 1e6 at 32x32x16")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.burnin_steps < 0:
        args.burnin_steps = STATIONARY[args.workload][0]
    if args.ess_samples < 0:
        args.ess_samples = STATIONARY[args.workload][1]
    if args.impl == "reference":
        return bench_reference(args)
    return bench_ours(args)


if __name__ == "__main__":
    sys.exit(main())
