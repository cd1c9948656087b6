"""ctypes wrapper around oracle/_build/libsbs_oracle.so (the C restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "_build" / "libsbs_oracle.so"
    src = _HERE / "sbs_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s", "_build/libsbs_oracle.so"], check=True,
                       capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        vp, i32p, i8p, f64p, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int8), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.orc_problem_create.restype = vp
        L.orc_problem_create.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, i32p, i32p, i32p, i32p, i32p, i32p, i32p, f64p,
                                         C.c_double, C.c_int, C.c_int, f64p, f64p]
        L.orc_problem_destroy.argtypes = [vp]
        L.orc_problem_set_agent.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p, i32p, i32p, i32p, f64p, f64p]
        for f in ("orc_problem_nrows", "orc_problem_ncols", "orc_problem_nnz"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [vp]
        L.orc_chain_create.restype = vp
        L.orc_chain_create.argtypes = [vp, i8p, C.c_int]
        L.orc_chain_destroy.argtypes = [vp]
        L.orc_chain_seed_mt.argtypes = [vp, C.c_uint32]
        L.orc_chain_seed_philox.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_chain_set_injected.argtypes = [vp, f64p, C.c_long]
        L.orc_chain_set_injected_proposals.argtypes = [vp, i32p]
        L.orc_chain_philox_iter.restype = C.c_uint64
        L.orc_chain_philox_iter.argtypes = [vp]
        L.orc_chain_find_feasible.restype = C.c_long
        L.orc_chain_find_feasible.argtypes = [vp, C.c_long]
        L.orc_chain_step.argtypes = [vp, C.c_long, i32p, u8p, i32p]
        L.orc_chain_next_sample.restype = C.c_long
        L.orc_chain_next_sample.argtypes = [vp]
        L.orc_chain_get_X.argtypes = [vp, i32p]
        L.orc_chain_get_delta.argtypes = [vp, i8p]
        for f in ("orc_chain_get_DE", "orc_chain_get_tree", "orc_chain_get_leaf"):
            getattr(L, f).argtypes = [vp, f64p]
        L.orc_chain_sum.restype = C.c_double
        L.orc_chain_sum.argtypes = [vp]
        L.orc_chain_logimp.restype = C.c_double
        L.orc_chain_logimp.argtypes = [vp]
        L.orc_chain_infeasibility.restype = C.c_int
        L.orc_chain_infeasibility.argtypes = [vp]
        L.orc_chain_get_counters.argtypes = [vp, C.POINTER(C.c_int64)]
        L.orc_chain_hash_X.restype = C.c_uint64
        L.orc_chain_hash_X.argtypes = [vp]
        L.orc_chain_end_state.argtypes = [vp, i32p]
        L.orc_chain_sample_stats.restype = C.c_long
        L.orc_chain_sample_stats.argtypes = [vp, C.c_long, C.c_long, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                             C.c_long, i32p, C.c_long, C.c_long]
        L.orc_mt_uniforms.argtypes = [C.c_uint32, C.c_long, f64p]
        L.orc_philox_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_long, f64p]
        _LIB = L
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class OracleProblem:
    """Flat sampler problem built from an ABMP dict (see agentbasedmcmc_b200/abmp.py)."""

    def __init__(self, abmp: dict, with_importance: bool = True):
        L = lib()
        self.abmp = abmp
        d = abmp["dims"]
        keep = {}
        for k in ("tab_L", "tab_U", "tab_F", "nb_col", "nb_ptr", "nb_row", "nb_val", "fac_id", "ut_lo", "ut_off"):
            keep[k] = np.ascontiguousarray(abmp[k], dtype=np.int32)
        for k in ("ut_val", "imp_lgamma", "imp_logratio"):
            keep[k] = np.ascontiguousarray(abmp[k], dtype=np.float64)
        self._keep = keep
        G, T = (int(abmp["pp_meta"][0]), int(abmp["pp_meta"][1])) if with_importance and "pp_meta" in abmp else (0, 0)
        self.G, self.T = G, T
        i32 = C.c_int32
        self.h = L.orc_problem_create(int(d[0]), int(d[3]), _p(keep["tab_L"], i32), _p(keep["tab_U"], i32), _p(keep["tab_F"], i32),
                                      _p(keep["nb_col"], i32), _p(keep["nb_ptr"], i32), _p(keep["nb_row"], i32), _p(keep["nb_val"], i32),
                                      _p(keep["fac_id"], i32), _p(keep["ut_lo"], i32), _p(keep["ut_off"], i32), _p(keep["ut_val"], C.c_double),
                                      float(abmp["kappa"][0]), G, T, _p(keep["imp_lgamma"], C.c_double), _p(keep["imp_logratio"], C.c_double))
        self.n_rows = L.orc_problem_nrows(self.h)
        self.n_cols = L.orc_problem_ncols(self.h)
        self.n_events = int(abmp["pp_meta"][2]) if "pp_meta" in abmp else 0
        self.n_agent_states = 2 * G * G
        if with_importance and "ag_meta" in abmp:   # an agent type other than PredPrey, by tables
            S, A, T, n_cls = (int(v) for v in abmp["ag_meta"][:4])
            for k in ("ag_class", "ag_infl", "ag_nb_ptr", "ag_nb", "ag_cq_ptr", "ag_cq"):
                keep[k] = np.ascontiguousarray(abmp[k], dtype=np.int32)
                if keep[k].size == 0:
                    keep[k] = np.zeros(1, np.int32)
            L.orc_problem_set_agent(self.h, S, A, T, n_cls, *(_p(keep[k], i32) for k in ("ag_class", "ag_infl", "ag_nb_ptr", "ag_nb",
                                                                                          "ag_cq_ptr", "ag_cq")),
                                    _p(keep["imp_lgamma"], C.c_double), _p(keep["imp_logratio"], C.c_double))
            self.T, self.n_events, self.n_agent_states = T, T * S * A, S

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_problem_destroy(self.h)
            self.h = None


class OracleChain:
    def __init__(self, problem: OracleProblem, init_state=None):
        self.problem = problem
        if init_state is None:
            init_state = np.zeros(0, dtype=np.int8)
        self._init = np.ascontiguousarray(init_state, dtype=np.int8)
        self.h = lib().orc_chain_create(problem.h, _p(self._init, C.c_int8), self._init.size)
        self._inj = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_chain_destroy(self.h)
            self.h = None

    def seed_mt(self, seed):
        lib().orc_chain_seed_mt(self.h, seed)

    def seed_philox(self, seed, chain_id, it=0):
        lib().orc_chain_seed_philox(self.h, seed, chain_id, it)

    def set_injected(self, u, proposals=None):
        self._inj = np.ascontiguousarray(u, dtype=np.float64)
        lib().orc_chain_set_injected(self.h, _p(self._inj, C.c_double), self._inj.size)
        if proposals is not None:
            self._injj = np.ascontiguousarray(proposals, dtype=np.int32)
            lib().orc_chain_set_injected_proposals(self.h, _p(self._injj, C.c_int32))

    def find_feasible(self, max_iter=0):
        return lib().orc_chain_find_feasible(self.h, max_iter)

    def step(self, n, trace=False):
        if not trace:
            lib().orc_chain_step(self.h, n, None, None, None)
            return None
        j = np.empty(n, np.int32)
        acc = np.empty(n, np.uint8)
        inf = np.empty(n, np.int32)
        lib().orc_chain_step(self.h, n, _p(j, C.c_int32), _p(acc, C.c_uint8), _p(inf, C.c_int32))
        return j, acc, inf

    def next_sample(self):
        return lib().orc_chain_next_sample(self.h)

    def sample_stats(self, n_samples, max_steps, end_sum, syn_sum, syn_sumsq, sample_base=0, series=None, thin=1):
        """Accumulates into the int64 arrays; series: int32 [cap][8] (columns beyond n_synopsis unused)."""
        cap = 0 if series is None else series.shape[0]
        return lib().orc_chain_sample_stats(self.h, n_samples, max_steps, _p(end_sum, C.c_int64), _p(syn_sum, C.c_int64),
                                            _p(syn_sumsq, C.c_int64), sample_base,
                                            None if series is None else _p(series, C.c_int32), cap, thin)

    @property
    def X(self):
        out = np.empty(self.problem.n_rows, np.int32)
        lib().orc_chain_get_X(self.h, _p(out, C.c_int32))
        return out

    @property
    def delta(self):
        out = np.empty(self.problem.n_cols, np.int8)
        lib().orc_chain_get_delta(self.h, _p(out, C.c_int8))
        return out

    def _f64(self, fn):
        out = np.empty(self.problem.n_cols, np.float64)
        getattr(lib(), fn)(self.h, _p(out, C.c_double))
        return out

    DE = property(lambda self: self._f64("orc_chain_get_DE"))
    tree = property(lambda self: self._f64("orc_chain_get_tree"))
    leaf = property(lambda self: self._f64("orc_chain_get_leaf"))
    sum = property(lambda self: lib().orc_chain_sum(self.h))
    log_importance = property(lambda self: lib().orc_chain_logimp(self.h))
    infeasibility = property(lambda self: lib().orc_chain_infeasibility(self.h))
    hash_X = property(lambda self: lib().orc_chain_hash_X(self.h))

    @property
    def counters(self):
        out = np.empty(8, np.int64)
        lib().orc_chain_get_counters(self.h, _p(out, C.c_int64))
        return out

    @property
    def end_state(self):
        out = np.empty(self.problem.n_agent_states, np.int32)
        lib().orc_chain_end_state(self.h, _p(out, C.c_int32))
        return out


def mt_uniforms(seed, n):
    out = np.empty(n, np.float64)
    lib().orc_mt_uniforms(seed, n, _p(out, C.c_double))
    return out


def philox_uniforms(seed, chain_id, iter0, n_iter):
    out = np.empty(2 * n_iter, np.float64)
    lib().orc_philox_uniforms(seed, chain_id, iter0, n_iter, _p(out, C.c_double))
    return out
