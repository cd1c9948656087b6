"""The multi-GPU path on CPU: chains are partitioned over ranks and only additive sufficient statistics are
all-reduced.  world_size-2 gloo processes must reproduce the single-process ensemble summary."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agentbasedmcmc_b200 import ensemble as E
from agentbasedmcmc_b200.diagnostics import rhat_from_moments


def _fake_stats(seed, n_chains, S=32, d=3):
    rng = np.random.default_rng(seed)
    n = rng.integers(500, 600, size=n_chains).astype(np.int64)
    x = [rng.normal(5.0 + 0.1 * rng.normal(), 2.0, size=(int(k), d)).round() for k in n]
    ss = np.stack([a.sum(0) for a in x]).astype(np.int64)
    sq = np.stack([(a * a).sum(0) for a in x]).astype(np.int64)
    end = rng.integers(0, 1000, size=S).astype(np.int64)
    ctr = rng.integers(0, 10000, size=(n_chains, 8)).astype(np.int64)
    return end, ss, sq, n, ctr


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    end, ss, sq, n, ctr = _fake_stats(100 + rank, 40)
    summ = E.allreduce_summary(E.pack_local(end, ss, sq, n, ctr), end.size, ss.shape[1])
    if rank == 0:
        out.put((summ.n_chains, summ.n_samples, summ.end_state_mean, summ.rhat, summ.W, summ.B, summ.counters))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = [_fake_stats(100 + r, 40) for r in range(2)]
    end = sum(p[0] for p in parts)
    ss, sq, n = (np.concatenate([p[k] for p in parts]) for k in (1, 2, 3))
    ctr = np.concatenate([p[4] for p in parts]).sum(0)
    rhat, W, B, _ = rhat_from_moments(n, ss, sq)
    assert got[0] == 80 and got[1] == int(n.sum())
    assert np.allclose(got[2], end / n.sum())
    assert np.allclose(got[3], rhat, rtol=1e-9) and np.allclose(got[4], W, rtol=1e-9) and np.allclose(got[5], B, rtol=1e-7)
    assert np.array_equal(got[6], ctr)


def test_single_process_summary_without_process_group():
    end, ss, sq, n, ctr = _fake_stats(3, 10)
    s = E.allreduce_summary(E.pack_local(end, ss, sq, n, ctr), end.size, ss.shape[1])
    assert s.n_chains == 10 and np.allclose(s.synopsis_mean, ss.sum(0) / n.sum())


def test_summarize_moments_equals_the_packed_route():
    """The vectors abm_chains_get_moments returns (layout of include/abmcmc_b200.h, restated here on synthetic per-chain
    tables) give the same W, B, var+, R-hat and marginals as the per-chain tables packed on the host."""
    import numpy as np
    from agentbasedmcmc_b200 import ensemble as E
    m, d, S = 37, 3, 32
    end, ss, sq, n, ctr = _fake_stats(7, m, S, d)
    n = n.copy()
    n[3], ss[3], sq[3] = 1, ss[3] // 500, sq[3] // 500        # a chain with a single sample contributes to no moment
    ok = n > 1
    exact = np.zeros(26 + S, np.int64)
    exact[0], exact[1], exact[2:10] = ok.sum(), n.sum(), ctr.sum(0)
    exact[10:10 + d], exact[18:18 + d], exact[26:] = ss.sum(0), sq.sum(0), end
    real = np.zeros(24)
    mean = ss[ok] * (1.0 / n[ok, None])
    var = (sq[ok] - ss[ok].astype(np.float64) * ss[ok] * (1.0 / n[ok, None])) * (1.0 / (n[ok, None] - 1))
    real[0:d], real[8:8 + d], real[16:16 + d] = mean.sum(0), (mean * mean).sum(0), var.sum(0)
    a = E.summarize_moments(exact, real, d)
    b = E.summarize(E.pack_local(end, ss, sq, n, ctr), S, d)
    assert a.n_chains == b.n_chains and a.n_samples == b.n_samples and np.array_equal(a.counters, b.counters)
    for x, y in ((a.W, b.W), (a.B, b.B), (a.var_plus, b.var_plus), (a.rhat, b.rhat), (a.end_state_mean, b.end_state_mean),
                 (a.synopsis_mean, b.synopsis_mean)):
        assert np.allclose(x, y, rtol=1e-10)


def test_summarize_moments_without_synopsis_statistics():
    """Agents described by tables have end-state statistics only (no nested-square synopsis): the summary of their moments is the
    end-state means and the counters, with empty diagnostics."""
    from agentbasedmcmc_b200 import _capi, ensemble as E
    S = 4
    exact = np.zeros(_capi.MOMENTS_EXACT_HEAD + S, np.int64)
    exact[0], exact[1] = 10, 1000
    exact[2:10] = [100, 50, 60, 300, 80, 40, 40, 250]
    exact[_capi.MOMENTS_EXACT_HEAD:] = [500, 200, 10, 900]
    r = E.summarize_moments(exact, np.zeros(_capi.MOMENTS_REAL), 0)
    assert r.n_chains == 10 and r.n_samples == 1000
    assert np.allclose(r.end_state_mean, [0.5, 0.2, 0.01, 0.9])
    assert r.rhat.size == 0 and r.synopsis_mean.size == 0
    assert np.array_equal(r.counters, exact[2:10])
