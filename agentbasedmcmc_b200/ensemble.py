"""Ensemble statistics across chains and across GPUs.

Chains are independent, so the only exchange step of the whole sampler is a SUM all-reduce of a few KB of sufficient
statistics per reporting interval: the end-state occupancy sums, the sample count, and -- per synopsis statistic -- the
sums over chains of the chain mean, its square, and the chain variance, from which W, B, var+ and R-hat follow
(diagnostics/MultiChainStats.h:141-181).  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


@dataclass
class EnsembleSummary:
    n_chains: int
    n_samples: int                 # feasible samples over all chains
    end_state_mean: np.ndarray     # [S] posterior mean occupancy of every agent state at t = T
    synopsis_mean: np.ndarray      # [d]
    W: np.ndarray
    B: np.ndarray
    var_plus: np.ndarray
    rhat: np.ndarray
    counters: np.ndarray           # [8] MCMCStatistics nProposals / nAccepted summed over all chains


class Comm:
    """abm_comm: the library's own NCCL communicator over the ranks of a torch.distributed group (the 128-byte NCCL id is
    made on rank 0 and broadcast through the group, gloo or nccl) -- the multi-GPU path a C++ host uses as well."""

    def __init__(self, device: int, group=None):
        import ctypes as C
        from . import _capi
        L = _capi.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        idbuf = np.zeros(_capi.COMM_ID_BYTES, np.uint8)
        if rank == 0:
            _capi.check(L.abm_comm_unique_id(idbuf.ctypes.data_as(C.POINTER(C.c_uint8))))
        t = torch.from_numpy(idbuf)
        if dist.get_backend(group) == "nccl":
            t = t.cuda(device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idbuf = t.cpu().numpy().copy()
        h = C.c_void_p()
        _capi.check(L.abm_comm_create(world, rank, idbuf.ctypes.data_as(C.POINTER(C.c_uint8)), device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            from . import _capi
            _capi.load().abm_comm_destroy(self.h)
            self.h = None

    __del__ = close


def summarize_moments(exact, real, n_syn: int) -> EnsembleSummary:
    """EnsembleSummary from the additive vectors of abm_chains_get_moments / abm_chains_allreduce_moments
    (include/abmcmc_b200.h): W, B, var+, R-hat as diagnostics/MultiChainStats.h:141-181 define them."""
    exact = np.asarray(exact)
    real = np.asarray(real, dtype=np.float64)
    m, n_tot = float(exact[0]), float(exact[1])
    counters = exact[2:10].astype(np.int64)
    ssum = exact[10:10 + n_syn].astype(np.float64)
    end = exact[26:].astype(np.float64)
    sm, sm2, sv = real[0:n_syn], real[8:8 + n_syn], real[16:16 + n_syn]
    nbar = n_tot / max(m, 1.0)
    W = sv / max(m, 1.0)
    grand = sm / max(m, 1.0)
    B = nbar / max(m - 1.0, 1.0) * (sm2 - m * grand * grand)
    var_plus = W * (nbar - 1.0) / max(nbar, 1.0) + B / max(nbar, 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        rhat = np.sqrt(var_plus / W)
    return EnsembleSummary(int(m), int(n_tot), end / max(n_tot, 1.0), ssum / max(n_tot, 1.0), W, B, var_plus, rhat, counters)


def pack_local(end_state_sum, syn_sum, syn_sumsq, n_samples, counters) -> torch.Tensor:
    """Per-rank float64 vector [S + 2 + 4 d + 8] that is additive across ranks.  Inputs are this rank's chains."""
    end = np.asarray(end_state_sum, dtype=np.float64).ravel()
    s = np.asarray(syn_sum, dtype=np.float64)
    sq = np.asarray(syn_sumsq, dtype=np.float64)
    n = np.asarray(n_samples, dtype=np.float64)
    ok = n > 1
    nn = np.where(ok, n, 1.0)[:, None]
    mean = np.where(ok[:, None], s / nn, 0.0)
    var = np.where(ok[:, None], (sq - s * s / nn) / np.maximum(nn - 1.0, 1.0), 0.0)
    head = np.array([float(ok.sum()), n.sum()])
    ctr = np.asarray(counters, dtype=np.float64).reshape(-1, 8).sum(0)
    return torch.from_numpy(np.concatenate([end, head, mean.sum(0), (mean * mean).sum(0), var.sum(0), s.sum(0), ctr]))


def summarize(packed: torch.Tensor, n_states: int, n_syn: int) -> EnsembleSummary:
    v = packed.detach().cpu().numpy()
    end, rest = v[:n_states], v[n_states:]
    m, n_tot = rest[0], rest[1]
    sm, sm2, sv, ssum = (rest[2 + k * n_syn: 2 + (k + 1) * n_syn] for k in range(4))
    counters = rest[2 + 4 * n_syn:]
    nbar = n_tot / max(m, 1.0)
    W = sv / max(m, 1.0)
    grand = sm / max(m, 1.0)
    B = nbar / max(m - 1.0, 1.0) * (sm2 - m * grand * grand)
    var_plus = W * (nbar - 1.0) / nbar + B / nbar
    with np.errstate(divide="ignore", invalid="ignore"):
        rhat = np.sqrt(var_plus / W)
    return EnsembleSummary(int(m), int(n_tot), end / max(n_tot, 1.0), ssum / max(n_tot, 1.0), W, B, var_plus, rhat,
                           counters.astype(np.int64))


def allreduce_summary(packed: torch.Tensor, n_states: int, n_syn: int, group=None) -> EnsembleSummary:
    """SUM all-reduce of the packed statistics over the process group (no-op when not initialised), then summary."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return summarize(packed, n_states, n_syn)


def sampler_summary(smp, group=None, device=None) -> EnsembleSummary:
    """Statistics of a `SparseBasisSampler` (all its chains), all-reduced over `group`."""
    end, ss, sq, ns = smp.statistics()
    packed = pack_local(end, ss, sq, ns, smp.counters())
    if device is not None:
        packed = packed.to(device)
    return allreduce_summary(packed, end.size, ss.shape[1], group)
