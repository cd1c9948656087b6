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


def cpu_baseline(workload: str, threads: int, seconds_budget: float = 20.0):
    """Reference CPU sampler on this box's host cores on a bounded sample of the same workload."""
    fixture = WORKLOADS[workload][0]
    # ~13 us/step/core and 1-4 steps per feasible sample: size the sample for ~seconds_budget of timed work
    samples = {"pp32-16": 60000, "pp16-8": 400000, "pp8-4": 600000}[workload]
    samples = int(samples * seconds_budget / 20.0)
    res = run_reference_binary(fixture, threads, samples, 0, 1)
    if res is not None:
        v = res["chain_steps"][0] / res["step_seconds"][0]
        return {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{threads} chains (one per host thread) x {samples} feasible samples = {res['chain_steps'][0]} Metropolis steps, "
                          f"unmodified reference SparseBasisSampler (oracle/_ref/abm_ref), {res['step_seconds'][0]:.1f}s timed"}
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
    if res is None:
        cb = cpu_baseline(args.workload, 1)
        value, ms, kind, cores, sample = cb["value"], None, cb["kind"], cb["cores"], cb["sample"]
    else:
        tot_steps, tot_sec = sum(res["chain_steps"]), sum(res["step_seconds"])
        value, ms = tot_steps / tot_sec, 1e3 * tot_sec / len(res["step_seconds"])
        kind, cores = "reference", threads
        sample = f"{threads} chains (one per host thread), {samples} feasible samples per chain per step, {tot_steps} Metropolis steps timed"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": desc, "host_threads": cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up launches, then exactly K timed launches
    for _ in range(args.warmup):
        smp.step(mh)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        ev[0].record(stream)
        for k in range(args.steps):
            smp.step(mh)
            ev[k + 1].record(stream)
        barrier()
    launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    steps_done = float(world) * n_chains * mh * args.steps
    value = steps_done / (total_ms * 1e-3)

    # ---- end to end through the host-facing API: step + statistics into host buffers, host clock
    from agentbasedmcmc_b200 import ensemble as E
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    summary = None
    # launch k+1 is queued behind the read-back of launch k, so the device never waits for the host: stream order is
    # sample 0, read-back 0, sample 1, read-back 1, ...; the host collects read-back k while sample k+1 runs
    smp.sample(10 ** 12, max_steps=mh)          # mh Metropolis steps per chain, statistics of every feasible state
    smp.enqueue_statistics()                    # ensemble reduction + device -> pinned host copies, asynchronous
    for k in range(args.steps):
        if k + 1 < args.steps:
            smp.sample(10 ** 12, max_steps=mh)
        end, ss, sq, ns, c = smp.wait_statistics()   # statistics and counters after launch k, in host buffers
        if k + 1 < args.steps:
            smp.enqueue_statistics()
        d2h = end.nbytes + ss.nbytes + sq.nbytes + ns.nbytes + c.nbytes
        packed = E.pack_local(end, ss, sq, ns, c)
        if world > 1:
            packed = packed.cuda()
        summary = E.allreduce_summary(packed, end.size, ss.shape[1])   # NCCL sum over ranks (world > 1)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = steps_done / float(te.item())

    # ---- ESS/s (the second half of BASELINE.json's metric): reference formula (MultiChainStats::effectiveSamples,
    # diagnostics/MultiChainStats.h:81-102) on the synopsis series of the first `ess_chains` chains of rank 0, each a
    # sequence; chains are i.i.d., so the ensemble's effective sample size is the per-chain one times the chain count.
    ess = None
    if args.ess_samples > 0 and rank == 0:
        try:
            from agentbasedmcmc_b200 import diagnostics as D
            k = min(args.ess_chains, n_chains)
            burn = 0
            while burn < args.ess_burnin:       # relaxation from the first feasible state takes ~3e6 steps at 32x32x16
                smp.step(min(250000, args.ess_burnin - burn))
                burn += 250000
            smp.reset_statistics()
            smp.record_series(k, 1, args.ess_samples)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            smp.sample(args.ess_samples, max_steps=4 * args.ess_samples)   # bounded: a chain deep in the infeasible region must not stall the batch
            smp.synchronize()
            ess_s = time.perf_counter() - t0
            ns_k = smp.statistics()[3][:k]
            n_rec = int(min(np.percentile(ns_k, 25), args.ess_samples))      # sequences must be equally long: keep the upper 75 %
            use = [c for c in range(k) if ns_k[c] >= n_rec]
            if n_rec >= 400 and len(use) >= 4:
                series = smp.series().astype(np.float64)[use, :n_rec]
                ms = D.MultiChainStats([D.ChainStats.from_series(x, n_lags=200, max_lag_proportion=0.5) for x in series])
                per_chain = ms.effective_samples(reference_integer_division=False) / len(use)
                ess = {"value": float(per_chain.min() * n_chains * world / ess_s),
                       "unit": "effective samples/s (worst synopsis statistic, all chains)",
                       "n_eff_per_chain": [round(float(x), 2) for x in per_chain], "feasible_samples_per_chain": n_rec,
                       "sequences_analysed": len(use), "seconds": round(ess_s, 3), "burnin_steps_per_chain": args.ess_burnin,
                       "rhat": [round(float(x), 4) for x in ms.potential_scale_reduction(reference_integer_division=False)]}
            else:
                ess = {"value": None, "note": f"too few feasible samples in {4 * args.ess_samples} steps (25th percentile {n_rec})"}
        except Exception as exc:   # the ESS estimate must never take the benchmark line down
            ess = {"value": None, "note": f"ESS estimate failed: {exc!r}"}
    barrier()

    st = smp.stats
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
                       "acceptance": round(st.totalAccepted() / max(st.nSamples(), 1), 4),
                       "infeasible_fraction": round(st.nInfeasibleSamples() / max(st.nSamples(), 1), 4)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_chain_step": bytes_per_step, "kernel": "abm::stepKernel<MODE_MCMC>",
                         "kernel_ms_per_launch": kernel_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                    "what": "abm_chains_sample (steps + on-device end-state/synopsis statistics) + abm_chains_enqueue_statistics / "
                            "abm_chains_wait_statistics (ensemble reduction, statistics and counters into host buffers) + ensemble "
                            "all-reduce, host clock, every step; launch k+1 is queued behind the read-back of launch k; inputs of a "
                            "step are kernel arguments only (the problem and the chain states are uploaded once, at construction)"},
            "ensemble": {"chains": summary.n_chains, "feasible_samples": summary.n_samples,
                         "mean_agents_at_T": float(summary.end_state_mean.sum()),
                         "rhat_synopsis": [round(float(x), 4) for x in summary.rhat]},
            "ess": ess,
            # per bench step of the timed region: refreshBlockedTopKernel (0.16 ms) + stepKernel<MODE_MCMC> (profiles/*_bench_launches.csv)
            "gpu_launches": 2 * args.steps,
            "clocks": clocks.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.workload, os.cpu_count() or 1)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
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
    ap.add_argument("--ess-samples", type=int, default=0,
                    help="feasible samples per chain for the ESS/s estimate (0 = skip; needs --ess-burnin steps first, ~2 more minutes)")
    ap.add_argument("--ess-burnin", type=int, default=3000000, help="Metropolis steps per chain before the ESS estimate")
    ap.add_argument("--ess-chains", type=int, default=64, help="chains whose synopsis series are analysed for ESS/s")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return bench_reference(args)
    return bench_ours(args)


if __name__ == "__main__":
    sys.exit(main())
