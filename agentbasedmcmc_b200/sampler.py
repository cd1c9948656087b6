"""Host-side mirror of the reference's sampler interface on top of the C-ABI.

``SparseBasisSampler`` keeps the reference names (SparseBasisSampler.h:31-101): ``X``, ``delta``, ``currentDE``,
``currentInfeasibility``, ``currentLogImportance``, ``stats``, ``nRows()``, ``nCols()``, ``__call__`` (= operator()),
``findInitialFeasibleSolution`` -- but one object drives ``n_chains`` independent chains on one GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import AbmError, ProblemDesc, ProblemInfo, check


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _problem_desc(abmp: dict, with_importance: bool = True, sum_tree: str = "blocked"):
    """abm_problem_desc of an ABMP dict (and the arrays it points into, which must outlive the call that uses it)."""
    k = {}
    for name in ("tab_L", "tab_U", "tab_F", "nb_col", "nb_ptr", "nb_row", "nb_val", "fac_id", "ut_lo", "ut_off"):
        k[name] = np.ascontiguousarray(abmp[name], dtype=np.int32)
    for name in ("ut_val", "imp_lgamma", "imp_logratio"):
        k[name] = np.ascontiguousarray(abmp.get(name, np.zeros(24)), dtype=np.float64)
    d = ProblemDesc()
    d.n_tab_rows, d.n_nonbasic = int(abmp["dims"][0]), int(abmp["dims"][3])
    for name in ("tab_L", "tab_U", "tab_F", "nb_col", "nb_ptr", "nb_row", "nb_val", "fac_id", "ut_lo", "ut_off"):
        setattr(d, name, _ptr(k[name], C.c_int32))
    d.n_tables = k["ut_lo"].size
    d.ut_val = _ptr(k["ut_val"], C.c_double)
    d.kappa = float(abmp["kappa"][0])
    use_imp = with_importance and "pp_meta" in abmp
    d.grid_size = int(abmp["pp_meta"][0]) if use_imp else 0
    d.n_timesteps = int(abmp["pp_meta"][1]) if use_imp else 0
    if with_importance and "ag_meta" in abmp:   # an agent type other than PredPrey, by tables (reference_flatten.hpp)
        d.n_agent_states, d.n_acts, d.n_timesteps, d.n_agent_classes = (int(v) for v in abmp["ag_meta"][:4])
        for name, field in (("ag_class", "agent_class"), ("ag_infl", "agent_influencers"), ("ag_nb_ptr", "agent_nb_ptr"),
                            ("ag_nb", "agent_nb"), ("ag_cq_ptr", "agent_cq_ptr"), ("ag_cq", "agent_cq")):
            k[name] = np.ascontiguousarray(abmp[name], dtype=np.int32)
            if k[name].size == 0:   # an empty list still needs an address
                k[name] = np.zeros(1, np.int32)
            setattr(d, field, _ptr(k[name], C.c_int32))
    d.imp_lgamma = _ptr(k["imp_lgamma"], C.c_double)
    d.imp_logratio = _ptr(k["imp_logratio"], C.c_double)
    d.sum_tree = {"blocked": 0, "reference": 1}[sum_tree]
    return d, k


def validate_problem(abmp: dict, with_importance: bool = True, sum_tree: str = "blocked") -> ProblemInfo:
    """The host half of `Problem(...)` alone (abm_problem_validate): every check and the sizes, no device needed."""
    d, keep = _problem_desc(abmp, with_importance, sum_tree)
    info = ProblemInfo()
    check(_capi.load().abm_problem_validate(C.byref(d), C.byref(info)))
    del keep
    return info


class Problem:
    """Shared read-only structure of a sampler (tableau, bounds, factor tables, importance tables) on one GPU.

    Built from an ABMP dict (``abmp.read_abmp``), i.e. from what the reference's constructor reads out of its
    ``TableauNormMinimiser`` + factor closures (SparseBasisSampler.h:107-158).
    """

    def __init__(self, abmp: dict, device: int = 0, with_importance: bool = True, sum_tree: str = "blocked"):
        """sum_tree: "blocked" (32-ary tree, default) or "reference" (MutableCategoricalArray's binary tree, node for
        node); see ABM_TREE_* in include/abmcmc_b200.h."""
        L = _capi.load()
        self.abmp = abmp
        self.device = device
        d, self._keep = _problem_desc(abmp, with_importance, sum_tree)
        self.sum_tree = sum_tree
        h = C.c_void_p()
        check(L.abm_problem_create(C.byref(d), device, C.byref(h)))
        self.h = h
        info = ProblemInfo()
        check(L.abm_problem_get_info(self.h, C.byref(info)))
        self.info = info
        self.n_rows, self.n_cols, self.n_events = info.n_rows, info.n_cols, info.n_events
        self.grid_size, self.n_timesteps = d.grid_size, d.n_timesteps
        # agent states per timestep (ModelState's length): PredPrey's two species on the grid, or the table agent's domain
        self.n_agent_states = d.n_agent_states if d.grid_size == 0 else 2 * d.grid_size ** 2
        self.n_synopsis = L.abm_problem_n_synopsis(self.h)

    def close(self):
        if getattr(self, "h", None):
            _capi.load().abm_problem_destroy(self.h)
            self.h = None

    __del__ = close


@dataclass
class MCMCStatistics:
    """MCMCStatistics (MCMCStatistics.h/.cpp), summed over the chains it was built from."""
    nProposals: np.ndarray  # [2][2] indexed [startFeasible][proposalFeasible]
    nAccepted: np.ndarray

    def nSamples(self):
        return int(self.nProposals.sum())

    def totalAccepted(self):
        return int(self.nAccepted.sum())

    def nRejected(self, a, b):
        return int(self.nProposals[a, b] - self.nAccepted[a, b])

    def nInfeasibleSamples(self):  # MCMCStatistics.cpp:36-38
        return int(self.nProposals[0, 0] + self.nAccepted[1, 0] + self.nRejected(0, 1))

    def nFeasibleSamples(self):  # MCMCStatistics.cpp:40-42
        return int(self.nProposals[1, 1] + self.nAccepted[0, 1] + self.nRejected(1, 0))


class SparseBasisSampler:
    """``n_chains`` SparseBasisSampler chains on one GPU (reference: one object == one chain)."""

    def __init__(self, problem: Problem, initial_states=None, n_chains: int | None = None, seed: int = 0,
                 first_chain_id: int = 0, stream=None, find_feasible: bool = True, prior_init: bool = False,
                 return_prior_sample: bool = False):
        """initial_states: [n_chains][n_events] int8 (or one row), or None.  prior_init=True draws one initial
        trajectory per chain from the prior on the device (`prior.nextSample(false)`, FiguresForPaper.h:56)."""
        L = _capi.load()
        self.problem = problem
        if initial_states is not None:
            initial_states = np.ascontiguousarray(initial_states, dtype=np.int8)
            if initial_states.ndim == 1:
                initial_states = initial_states[None, :]
            if n_chains is None:
                n_chains = initial_states.shape[0]
            if initial_states.shape[0] != n_chains:
                raise ValueError("initial_states must have one row per chain")
        if n_chains is None:
            n_chains = 1
        self.n_chains = n_chains
        h = C.c_void_p()
        check(L.abm_chains_create(problem.h, n_chains, seed, first_chain_id, C.c_void_p(stream or 0), C.byref(h)))
        self.h = h
        self.prior_sample = None
        if prior_init:
            a = problem.abmp
            pmf = np.ascontiguousarray(a["pp_actpmf"], dtype=np.float64)
            if return_prior_sample:
                self.prior_sample = np.zeros((n_chains, problem.n_events), np.int8)
            check(L.abm_chains_init_from_prior(self.h, _ptr(pmf, C.c_double), float(a["pp_params"][0]), float(a["pp_params"][1]),
                                               seed ^ 0x5DEECE66D, _ptr(self.prior_sample, C.c_int8),
                                               int(a["start_kind"][0]) if "start_kind" in a else 0))
        else:
            n_init = 0 if initial_states is None else initial_states.shape[1]
            check(L.abm_chains_init(self.h, _ptr(initial_states, C.c_int8), n_init))
        self.init_iterations = None
        if find_feasible:
            self.findInitialFeasibleSolution()

    def close(self):
        if getattr(self, "h", None):
            _capi.load().abm_chains_destroy(self.h)
            self.h = None

    __del__ = close

    def nRows(self):
        return self.problem.n_rows

    def nCols(self):
        return self.problem.n_cols

    # SparseBasisSampler.h:254-263 (+ :183-184)
    def findInitialFeasibleSolution(self, max_iterations: int = 0, uniforms=None):
        it = np.zeros(self.n_chains, np.int64)
        n_u = 0
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(self.n_chains, -1)
            n_u = uniforms.shape[1]
        check(_capi.load().abm_chains_find_feasible(self.h, max_iterations, _ptr(uniforms, C.c_double), n_u,
                                                    _ptr(it, C.c_int64)))
        self.init_iterations = it
        return it

    def step(self, n_steps: int):
        """n_steps Metropolis steps on every chain (asynchronous)."""
        check(_capi.load().abm_chains_step(self.h, n_steps))

    def step_traced(self, n_steps: int, uniforms=None, proposals=None):
        """Steps with injected randomness; returns (j, accepted, proposalInfeasibility), each [n_chains][n_steps]."""
        n = self.n_chains
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(n, 2 * n_steps)
        if proposals is not None:
            proposals = np.ascontiguousarray(proposals, dtype=np.int32).reshape(n, n_steps)
        j = np.empty((n, n_steps), np.int32)
        acc = np.empty((n, n_steps), np.uint8)
        inf = np.empty((n, n_steps), np.int32)
        check(_capi.load().abm_chains_step_traced(self.h, n_steps, _ptr(uniforms, C.c_double), _ptr(proposals, C.c_int32),
                                                  _ptr(j, C.c_int32), _ptr(acc, C.c_uint8), _ptr(inf, C.c_int32)))
        return j, acc, inf

    # operator() for every chain + the statistics pipeline of FiguresForPaper.h:93-104
    def sample(self, n_samples: int, max_steps: int = 0):
        """Steps every chain until it has visited n_samples further feasible states (asynchronous)."""
        check(_capi.load().abm_chains_sample(self.h, n_samples, max_steps))

    def statistics(self):
        """(end_state_sum[S] summed over chains, syn_sum[n_chains][d], syn_sumsq[n_chains][d], n_samples[n_chains])."""
        P = self.problem
        end = np.zeros(P.n_agent_states, np.int64)
        ss = np.zeros((self.n_chains, P.n_synopsis), np.int64)
        sq = np.zeros((self.n_chains, P.n_synopsis), np.int64)
        ns = np.zeros(self.n_chains, np.int64)
        check(_capi.load().abm_chains_get_statistics(self.h, _ptr(end, C.c_int64), _ptr(ss, C.c_int64), _ptr(sq, C.c_int64),
                                                     _ptr(ns, C.c_int64), 0))
        return end, ss, sq, ns

    def enqueue_statistics(self):
        """Orders the statistics read-back after everything launched so far and returns (abm_chains_enqueue_statistics);
        launches issued afterwards keep running while `wait_statistics` collects it."""
        check(_capi.load().abm_chains_enqueue_statistics(self.h))

    def wait_statistics(self):
        """(end_state_sum, syn_sum, syn_sumsq, n_samples, counters[n_chains][8]) of the last `enqueue_statistics`."""
        P = self.problem
        end = np.zeros(P.n_agent_states, np.int64)
        ss = np.zeros((self.n_chains, P.n_synopsis), np.int64)
        sq = np.zeros((self.n_chains, P.n_synopsis), np.int64)
        ns = np.zeros(self.n_chains, np.int64)
        cnt = np.zeros((self.n_chains, 8), np.int64)
        check(_capi.load().abm_chains_wait_statistics(self.h, _ptr(end, C.c_int64), _ptr(ss, C.c_int64), _ptr(sq, C.c_int64),
                                                      _ptr(ns, C.c_int64), _ptr(cnt, C.c_int64)))
        return end, ss, sq, ns, cnt

    def sample_injected(self, n_steps: int, uniforms):
        """`sample` for exactly n_steps steps with the uniforms injected ([n_chains][2*n_steps]); statistics parity runs."""
        u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(self.n_chains, 2 * n_steps)
        check(_capi.load().abm_chains_sample_injected(self.h, n_steps, _ptr(u, C.c_double)))

    # ---- ensemble moments reduced on the device (abm_chains_get_moments & co., include/abmcmc_b200.h)
    def _moment_buffers(self):
        S = self.problem.n_agent_states
        return np.zeros(_capi.MOMENTS_EXACT_HEAD + S, np.int64), np.zeros(_capi.MOMENTS_REAL, np.float64)

    def moments(self, comm=None):
        """(exact int64[26 + S], real float64[24]) of this GPU's chains, summed over the ranks of `comm` (ensemble.Comm)."""
        ex, re = self._moment_buffers()
        check(_capi.load().abm_chains_allreduce_moments(self.h, comm.h if comm is not None else None, _ptr(ex, C.c_int64),
                                                        _ptr(re, C.c_double)))
        return ex, re

    def enqueue_moments(self, comm=None):
        check(_capi.load().abm_chains_enqueue_moments(self.h, comm.h if comm is not None else None))

    def wait_moments(self):
        ex, re = self._moment_buffers()
        check(_capi.load().abm_chains_wait_moments(self.h, _ptr(ex, C.c_int64), _ptr(re, C.c_double)))
        return ex, re

    def statistics_into(self, end_ptr: int, syn_sum_ptr: int, syn_sumsq_ptr: int, n_samples_ptr: int):
        """Same, into DEVICE buffers (raw pointers, e.g. torch tensors' data_ptr()) for an NCCL all-reduce."""
        vp = lambda p: C.cast(C.c_void_p(p), C.POINTER(C.c_int64)) if p else None
        check(_capi.load().abm_chains_get_statistics(self.h, vp(end_ptr), vp(syn_sum_ptr), vp(syn_sumsq_ptr), vp(n_samples_ptr), 1))

    def reset_statistics(self):
        check(_capi.load().abm_chains_reset_statistics(self.h))

    def record_series(self, n_chains: int, thin: int, capacity: int):
        check(_capi.load().abm_chains_record_series(self.h, n_chains, thin, capacity))
        self._series_shape = (n_chains, capacity, self.problem.n_synopsis)

    def series(self):
        out = np.zeros(self._series_shape, np.int32)
        check(_capi.load().abm_chains_get_series(self.h, _ptr(out, C.c_int32)))
        return out

    def rebuild_tree(self):
        check(_capi.load().abm_chains_rebuild_tree(self.h))

    def tree_drift(self, chain: int = 0):
        """(root, max error of the sums of 32 / 1024 / 32768 leaves against the exact sums of their leaves); blocked tree."""
        out = np.zeros(4, np.float64)
        check(_capi.load().abm_chains_tree_drift(self.h, chain, _ptr(out, C.c_double)))
        return out

    def series_variogram(self, n_points: int, n_lags: int, min_samples: int, comm=None):
        """(sums int64 [n_lags][n_synopsis], contributing chains) of abm_chains_series_variogram, summed over `comm`'s ranks."""
        sums = np.zeros((n_lags, self.problem.n_synopsis), np.int64)
        n = C.c_int64()
        check(_capi.load().abm_chains_series_variogram(self.h, comm.h if comm is not None else None, n_points, n_lags, min_samples,
                                                       _ptr(sums, C.c_int64), C.byref(n)))
        return sums, int(n.value)

    def last_launch_ms(self) -> float:
        """Device time of the last step / step_traced / sample launch (waits for it)."""
        ms = C.c_float()
        check(_capi.load().abm_chains_last_launch_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def synchronize(self):
        check(_capi.load().abm_chains_synchronize(self.h))

    def state(self, chain: int = 0):
        """(X, delta, currentDE, tree, currentLogImportance, currentInfeasibility) of one chain."""
        P = self.problem
        X = np.empty(P.n_rows, np.int32)
        delta = np.empty(P.n_cols, np.int8)
        DE = np.empty(P.n_cols, np.float64)
        tree = np.empty(P.n_cols, np.float64)
        li = C.c_double()
        inf = C.c_int32()
        check(_capi.load().abm_chains_get_state(self.h, chain, _ptr(X, C.c_int32), _ptr(delta, C.c_int8), _ptr(DE, C.c_double),
                                                _ptr(tree, C.c_double), C.byref(li), C.byref(inf)))
        return X, delta, DE, tree, li.value, inf.value

    @property
    def X(self):
        return self.state(0)[0]

    def counters(self):
        out = np.zeros((self.n_chains, 8), np.int64)
        check(_capi.load().abm_chains_get_counters(self.h, _ptr(out, C.c_int64)))
        return out

    @property
    def stats(self):
        c = self.counters().sum(axis=0)
        return MCMCStatistics(c[:4].reshape(2, 2), c[4:].reshape(2, 2))

    def infeasibility(self):
        out = np.zeros(self.n_chains, np.int32)
        check(_capi.load().abm_chains_get_infeasibility(self.h, _ptr(out, C.c_int32)))
        return out

    def trajectories(self):
        out = np.zeros((self.n_chains, self.problem.n_events), np.int8)
        check(_capi.load().abm_chains_get_trajectories(self.h, _ptr(out, C.c_int8)))
        return out


def leaf_probabilities(tree: np.ndarray) -> np.ndarray:
    """MutableCategoricalArray::get for every index (MutableCategoricalArray.h:78, .cpp:30-39): tree[i] minus the
    descendant sum accumulated in the reference's order (offsets 1, 2, 4, ... while the offset bit of i is clear)."""
    n = tree.size
    idx = np.arange(n)
    s = np.zeros(n)
    off = 1
    while off < n:
        m = ((idx & (2 * off - 1)) == 0) & (idx + off < n)
        s[m] = s[m] + tree[idx[m] + off]
        off <<= 1
    return tree - s
