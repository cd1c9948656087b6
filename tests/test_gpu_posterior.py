"""Paper-scale posterior parity (north_star: "posterior marginals must match the reference's long-run CPU chain within a stated
Monte-Carlo tolerance") at the configuration the benchmark is quoted on: PredPrey 32x32, 16 timesteps.

Reference side: tests/golden/posterior_pp32-16.abmp.xz, made by tests/golden/make_posterior_golden.py from the UNMODIFIED
reference -- 12 chains x 2e7 feasible samples after 1e6 of burn-in, each from its own prior sample (the protocol of
FiguresForPaper.h:40-104), statistics by the reference's own ModelState(traj,T,T) and synopsis(); standard errors from batch
means (synopsis) and from the spread of the chain means (end state).
GPU side: 2 x 1024 chains from their own prior samples, 1e7 Metropolis steps of burn-in, then 2e6 steps of sampling; the
ensemble relaxes from the first feasible states (3x too many agents) with a time constant of ~1e6 steps and is flat to within
its own noise from ~8e6 (profiles/r2_relaxation_pp32-16_4096x2.json); standard error from the spread of the chain means.

Stated tolerance: every synopsis mean, and the mean number of agents at T, within 5 combined standard errors; the end-state
occupancies within 3 x the combined r.m.s. standard error (r.m.s. over the 2048 agent states; the GPU's own from the difference of
its two half-ensembles); acceptance rate and infeasible fraction within 0.01 of the reference chains'.  Takes ~3.5 minutes of GPU time: the relaxation is a property of
the chain, not of the implementation (the reference needs the same number of steps per chain)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_paper_scale_posterior_matches_reference_chain():
    import torch
    from agentbasedmcmc_b200 import sampler as S
    ref = load_golden("posterior_pp32-16.abmp.xz")
    P = S.Problem(load_golden("pp32-16.abmp.xz"))
    # two independent half-ensembles side by side on their own streams (same GPU time as one): their difference measures the
    # GPU estimate's own standard error per agent state
    m, halves = 1024, []
    streams = [torch.cuda.Stream() for _ in range(2)]
    for h in range(2):
        halves.append(S.SparseBasisSampler(P, n_chains=m, seed=20260 + h, first_chain_id=h * m, stream=streams[h].cuda_stream, prior_init=True))
    for _ in range(20):
        for smp in halves:
            smp.step(500000)                     # 1e7 steps of burn-in
    c0 = [smp.counters().sum(0) for smp in halves]
    for smp in halves:
        smp.reset_statistics()
        smp.sample(10 ** 12, max_steps=2000000)
    st = [smp.statistics() for smp in halves]
    c = sum(smp.counters().sum(0) - c00 for smp, c00 in zip(halves, c0))
    end = st[0][0] + st[1][0]
    ss, sq, ns = (np.concatenate([a[k], b[k]]) for k, (a, b) in zip((1, 2, 3), [(st[0], st[1])] * 3))
    assert np.all(ns > 1000)
    cm = ss / ns[:, None]                        # chain means of the synopsis statistics
    mean, se = ss.sum(0) / ns.sum(), cm.std(axis=0, ddof=1) / np.sqrt(2 * m)
    z = (mean - ref["syn_mean"]) / np.sqrt(se ** 2 + ref["syn_se"] ** 2)
    print("synopsis mean GPU", mean, "+-", se, " reference", ref["syn_mean"], "+-", ref["syn_se"], " z", z)
    assert np.all(np.abs(z) < 5.0)
    occ_h = [s_[0] / s_[3].sum() for s_ in st]
    agents_h = [o.sum() for o in occ_h]
    agents, agents_se = end.sum() / ns.sum(), abs(agents_h[0] - agents_h[1]) / 2.0
    za = (agents - ref["agents_mean"][0]) / np.sqrt(ref["agents_se"][0] ** 2 + max(agents_se, 4.0 * se[0]) ** 2)
    print("agents at T GPU", agents, "+-", agents_se, " reference", ref["agents_mean"][0], "+-", ref["agents_se"][0], " z", za)
    assert abs(za) < 5.0
    occ = end / ns.sum()
    rms = np.sqrt(((occ - ref["endstate_mean"]) ** 2).mean())
    rms_se_ref = np.sqrt((ref["endstate_se"] ** 2).mean())
    rms_se_gpu = np.sqrt(((occ_h[0] - occ_h[1]) ** 2).mean()) / 2.0      # s.e. of the mean of the two halves
    print("end-state occupancy: r.m.s. difference", rms, " r.m.s. standard errors: reference", rms_se_ref, " GPU", rms_se_gpu)
    assert rms < 3.0 * np.sqrt(rms_se_ref ** 2 + rms_se_gpu ** 2)
    prop, acc = c[:4], c[4:]
    acceptance, infeasible = acc.sum() / prop.sum(), (prop[0] + acc[2] + prop[1] - acc[1]) / prop.sum()
    print("acceptance", acceptance, ref["acceptance"][0], " infeasible fraction", infeasible, ref["infeasible_fraction"][0])
    assert abs(acceptance - ref["acceptance"][0]) < 0.01 and abs(infeasible - ref["infeasible_fraction"][0]) < 0.01
