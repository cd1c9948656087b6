"""ctypes binding of include/abmcmc_b200.h (libabmcmc_b200.so, built in-tree by ``build.py``).

There is no fallback: if the CUDA library is missing or no device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

_PKG = Path(__file__).resolve().parent
# ABM_LIB names an experimental build of the same library inside the package directory (kernel A/B measurements)
LIB_PATH = _PKG / os.environ.get("ABM_LIB", "libabmcmc_b200.so")

SYMBOLS = [
    "abm_last_error", "abm_version", "abm_device_count", "abm_problem_create", "abm_problem_validate", "abm_problem_destroy",
    "abm_problem_get_info", "abm_problem_n_synopsis", "abm_chains_create", "abm_chains_destroy", "abm_chains_init",
    "abm_chains_init_from_prior", "abm_chains_find_feasible", "abm_chains_step", "abm_chains_step_traced", "abm_chains_sample",
    "abm_chains_synchronize", "abm_chains_get_state", "abm_chains_get_counters", "abm_chains_get_infeasibility",
    "abm_chains_get_trajectories", "abm_chains_get_statistics", "abm_chains_enqueue_statistics", "abm_chains_wait_statistics", "abm_chains_reset_statistics",
    "abm_chains_record_series", "abm_chains_get_series",
    "abm_chains_sample_injected", "abm_comm_unique_id", "abm_comm_create", "abm_comm_from_nccl", "abm_comm_destroy",
    "abm_chains_get_moments", "abm_chains_allreduce_moments", "abm_chains_enqueue_moments", "abm_chains_wait_moments",
    "abm_chains_rebuild_tree", "abm_chains_tree_drift", "abm_chains_series_variogram",
    "abm_chains_last_launch_ms", "abm_basis_minimise", "abm_basis_destroy", "abm_basis_get_dims", "abm_basis_get", "abm_basis_get_rows",
]


MOMENTS_EXACT_HEAD, MOMENTS_REAL, COMM_ID_BYTES = 26, 24, 128   # ABM_MOMENTS_EXACT_HEAD, ABM_MOMENTS_REAL, ABM_COMM_ID_BYTES


class AbmError(RuntimeError):
    pass


class ProblemDesc(C.Structure):
    _fields_ = [
        ("n_tab_rows", C.c_int32), ("n_nonbasic", C.c_int32),
        ("tab_L", C.POINTER(C.c_int32)), ("tab_U", C.POINTER(C.c_int32)), ("tab_F", C.POINTER(C.c_int32)),
        ("nb_col", C.POINTER(C.c_int32)), ("nb_ptr", C.POINTER(C.c_int32)), ("nb_row", C.POINTER(C.c_int32)),
        ("nb_val", C.POINTER(C.c_int32)), ("fac_id", C.POINTER(C.c_int32)),
        ("n_tables", C.c_int32), ("ut_lo", C.POINTER(C.c_int32)), ("ut_off", C.POINTER(C.c_int32)),
        ("ut_val", C.POINTER(C.c_double)), ("kappa", C.c_double),
        ("grid_size", C.c_int32), ("n_timesteps", C.c_int32),
        ("imp_lgamma", C.POINTER(C.c_double)), ("imp_logratio", C.POINTER(C.c_double)),
        ("sum_tree", C.c_int32),
        ("n_agent_states", C.c_int32), ("n_acts", C.c_int32), ("n_agent_classes", C.c_int32),
        ("agent_class", C.POINTER(C.c_int32)), ("agent_influencers", C.POINTER(C.c_int32)),
        ("agent_nb_ptr", C.POINTER(C.c_int32)), ("agent_nb", C.POINTER(C.c_int32)),
        ("agent_cq_ptr", C.POINTER(C.c_int32)), ("agent_cq", C.POINTER(C.c_int32)),
    ]


class ProblemInfo(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int32), ("n_cols", C.c_int32), ("nnz", C.c_int32), ("n_events", C.c_int32),
        ("max_col_nnz", C.c_int32), ("max_coupled", C.c_int32), ("tree_depth", C.c_int32), ("n_states", C.c_int32),
        ("chain_state_bytes", C.c_int64), ("shared_bytes", C.c_int64),
    ]


_lib = None


def load():
    """Loads the CUDA library; raises AbmError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise AbmError(f"{LIB_PATH} is missing: run `python -m agentbasedmcmc_b200.build` (there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
    i8p, i32p, i64p, u8p, f64p = (C.POINTER(t) for t in (C.c_int8, C.c_int32, C.c_int64, C.c_uint8, C.c_double))
    L.abm_last_error.restype = C.c_char_p
    L.abm_problem_create.argtypes = [C.POINTER(ProblemDesc), C.c_int, vpp]
    L.abm_problem_validate.argtypes = [C.POINTER(ProblemDesc), C.POINTER(ProblemInfo)]
    L.abm_problem_destroy.argtypes = [vp]
    L.abm_problem_get_info.argtypes = [vp, C.POINTER(ProblemInfo)]
    L.abm_problem_n_synopsis.argtypes = [vp]
    L.abm_chains_create.argtypes = [vp, C.c_int32, C.c_uint64, C.c_uint64, vp, vpp]
    L.abm_chains_destroy.argtypes = [vp]
    L.abm_chains_init.argtypes = [vp, i8p, C.c_int32]
    L.abm_chains_init_from_prior.argtypes = [vp, f64p, C.c_double, C.c_double, C.c_uint64, i8p, C.c_int32]
    L.abm_chains_find_feasible.argtypes = [vp, C.c_int64, f64p, C.c_int64, i64p]
    L.abm_chains_step.argtypes = [vp, C.c_int64]
    L.abm_chains_step_traced.argtypes = [vp, C.c_int64, f64p, i32p, i32p, u8p, i32p]
    L.abm_chains_sample.argtypes = [vp, C.c_int64, C.c_int64]
    L.abm_chains_synchronize.argtypes = [vp]
    L.abm_chains_get_state.argtypes = [vp, C.c_int32, i32p, i8p, f64p, f64p, f64p, i32p]
    L.abm_chains_get_counters.argtypes = [vp, i64p]
    L.abm_chains_get_infeasibility.argtypes = [vp, i32p]
    L.abm_chains_get_trajectories.argtypes = [vp, i8p]
    L.abm_chains_get_statistics.argtypes = [vp, i64p, i64p, i64p, i64p, C.c_int]
    L.abm_chains_enqueue_statistics.argtypes = [vp]
    L.abm_chains_wait_statistics.argtypes = [vp, i64p, i64p, i64p, i64p, i64p]
    L.abm_chains_reset_statistics.argtypes = [vp]
    L.abm_chains_record_series.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    L.abm_chains_get_series.argtypes = [vp, i32p]
    L.abm_chains_last_launch_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.abm_chains_sample_injected.argtypes = [vp, C.c_int64, f64p]
    L.abm_chains_series_variogram.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int64, i64p, i64p]
    L.abm_chains_rebuild_tree.argtypes = [vp]
    L.abm_chains_tree_drift.argtypes = [vp, C.c_int32, f64p]
    L.abm_comm_unique_id.argtypes = [u8p]
    L.abm_comm_create.argtypes = [C.c_int32, C.c_int32, u8p, C.c_int, vpp]
    L.abm_comm_from_nccl.argtypes = [vp, vpp]
    L.abm_comm_destroy.argtypes = [vp]
    L.abm_chains_get_moments.argtypes = [vp, i64p, f64p]
    L.abm_chains_allreduce_moments.argtypes = [vp, vp, i64p, f64p]
    L.abm_chains_enqueue_moments.argtypes = [vp, vp]
    L.abm_chains_wait_moments.argtypes = [vp, i64p, f64p]
    L.abm_basis_minimise.argtypes = [C.c_int32, i32p, i32p, i32p, i32p, i32p, vpp]
    L.abm_basis_destroy.argtypes = [vp]
    L.abm_basis_get_dims.argtypes = [vp, i32p, i32p, i32p, i32p, i64p, i64p]
    L.abm_basis_get.argtypes = [vp, i32p, i32p, i32p, i32p, u8p, i32p, i32p, i32p, i32p]
    L.abm_basis_get_rows.argtypes = [vp, i64p, i32p, i32p]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise AbmError(f"abmcmc_b200 error {rc}: {load().abm_last_error().decode()}")
