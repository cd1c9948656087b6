"""The host half of abm_problem_create (validation, record layout, sizes) through abm_problem_validate: no device needed.
Structural numbers are the ones SURVEY.md section 8 measured on the reference's own problems (seed 1234)."""
import numpy as np
import pytest

from conftest import load_golden


def _S():
    from agentbasedmcmc_b200 import sampler
    return sampler


# name -> (sampler rows, basis columns, non-zeros, max column nnz, tree depth)   SURVEY.md section 8
STRUCTURE = {"pp8-4": (3456, 2658, 13738, 17, 12), "pp16-8": (27136, 20788, 112552, 17, 15),
             "pp32-16": (215040, 164219, 908996, 17, 18)}


def max_coupled(p):
    """Most columns any column shares a row with, itself included: what sizes the per-chain shared-memory lists."""
    import scipy.sparse as sp
    n_tab, _, _, n_cols, nnz = (int(v) for v in p["dims"])
    M = sp.csc_matrix((np.ones(nnz, np.int32), p["nb_row"], p["nb_ptr"]), shape=(n_tab, n_cols))
    return int(np.diff((M.T @ M).tocsr().indptr).max())


@pytest.mark.parametrize("name", sorted(STRUCTURE))
def test_structure_matches_the_survey(name):
    S = _S()
    info = S.validate_problem(load_golden(f"{name}.abmp.xz"))
    rows, cols, nnz, col_nnz, depth = STRUCTURE[name]
    assert (info.n_rows, info.n_cols, info.nnz, info.max_col_nnz, info.tree_depth) == (rows, cols, nnz, col_nnz, depth)
    assert info.max_coupled == max_coupled(load_golden(f"{name}.abmp.xz"))
    g, t = (int(v) for v in load_golden(f"{name}.abmp.xz")["pp_meta"][:2])
    assert info.n_states == t * 2 * g * g and info.n_events == 6 * info.n_states
    # per-chain HBM state: int8 X (8 bytes per agent state + the other rows), 16-byte column records + delta byte, block sums
    assert info.chain_state_bytes >= 17 * cols + 8 * info.n_states + (rows - info.n_events)
    assert info.chain_state_bytes < 20 * cols + 8 * info.n_states + rows + 65536
    assert info.shared_bytes > 24 * nnz


def test_both_tree_storages_and_no_importance_validate():
    S = _S()
    p = load_golden("pp8-4.abmp.xz")
    a, b = S.validate_problem(p, sum_tree="blocked"), S.validate_problem(p, sum_tree="reference")
    assert (a.n_rows, a.n_cols, a.max_coupled) == (b.n_rows, b.n_cols, b.max_coupled)
    c = S.validate_problem(p, with_importance=False)
    assert c.n_states == 0 and c.n_cols == a.n_cols


def test_validation_experiment_problem_is_accepted():
    """3x3 grid (not a power of two), Poisson start (unbounded rows), a tableau entry of magnitude 2 (Experiments.cpp:49-65)."""
    S = _S()
    p = load_golden("exp_pp3-2-single.abmp.xz")
    info = S.validate_problem(p)
    assert info.n_cols == int(p["dims"][3]) and info.n_states == 2 * 18


def test_bad_problems_are_rejected_on_the_host():
    S = _S()
    p = load_golden("pp8-4.abmp.xz")

    def broken(**changes):
        q = dict(p)
        for k, f in changes.items():
            q[k] = p[k].copy()
            f(q[k])
        return q

    def set0(v):
        def f(a):
            a[0] = v
        return f

    with pytest.raises(S.AbmError, match="at most 4"):             # entry magnitude beyond what a proposal tabulates
        S.validate_problem(broken(nb_val=set0(5)))
    with pytest.raises(S.AbmError, match="non-zero"):
        S.validate_problem(broken(nb_val=set0(0)))
    with pytest.raises(S.AbmError, match="ascending"):             # column entries must be ascending by row
        S.validate_problem(broken(nb_row=lambda a: a.__setitem__(slice(0, 2), a[[1, 0]])))
    with pytest.raises(S.AbmError):                                # a factor table that no longer covers its row's range
        S.validate_problem(broken(ut_off=lambda a: a.__setitem__(slice(1, None), np.maximum(a[1:] - 1, 0))))
    with pytest.raises(S.AbmError, match="int8"):                  # a row whose reachable values leave int8
        S.validate_problem(broken(tab_F=lambda a: a.__setitem__(int(np.flatnonzero(p["tab_L"] != p["tab_U"])[0]), 1000)))
    with pytest.raises(S.AbmError, match="bad factor id"):
        S.validate_problem(broken(fac_id=lambda a: a.__setitem__(int(np.flatnonzero(p["tab_L"] != p["tab_U"])[0]), 10 ** 6)))


def test_malformed_descriptors_are_rejected_not_dereferenced():
    """abm_problem_validate on descriptors with missing arrays or inconsistent pointers returns ABM_ERR_INVALID instead of
    crashing: null arrays, nb_ptr not starting at 0 or decreasing, no factor tables, ut_off not increasing."""
    import ctypes as C
    from agentbasedmcmc_b200 import _capi
    from agentbasedmcmc_b200.sampler import _problem_desc
    S = _S()
    L = _capi.load()
    p = load_golden("pp8-4.abmp.xz")

    def status(mutate):
        d, keep = _problem_desc(p)
        mutate(d, keep)
        info = _capi.ProblemInfo()
        rc = L.abm_problem_validate(C.byref(d), C.byref(info))
        del keep
        return rc

    null32 = C.POINTER(C.c_int32)()
    assert status(lambda d, k: None) == 0
    for field in ("tab_L", "tab_U", "tab_F", "nb_col", "nb_ptr", "nb_row", "nb_val", "fac_id", "ut_lo", "ut_off"):
        assert status(lambda d, k, f=field: setattr(d, f, null32)) == -1, field
    assert status(lambda d, k: setattr(d, "ut_val", C.POINTER(C.c_double)())) == -1
    assert status(lambda d, k: setattr(d, "n_tables", 0)) == -1
    assert status(lambda d, k: k["nb_ptr"].__setitem__(0, 1)) == -1
    assert status(lambda d, k: k["nb_ptr"].__setitem__(5, int(k["nb_ptr"][4]) - 1)) == -1
    assert status(lambda d, k: k["ut_off"].__setitem__(1, 0)) == -1
    assert L.abm_problem_validate(None, None) == -1


def test_table_agent_descriptor_is_validated():
    """Agents described by tables (abm_problem_desc::agent_*): the reference's toy experiments validate, malformed tables are
    refused on the host."""
    S = _S()
    p = dict(load_golden("toy_catmouse.abmp.xz"))
    info = S.validate_problem(p)
    assert info.n_states == 8 and info.n_events == 16   # T = 2, CatMouseAgent: 4 agent states x 2 acts
    for name in ("toy_binomial", "toy_catmouse-assim", "pp8-4-astable", "exp_pp3-2-single-astable"):
        S.validate_problem(load_golden(f"{name}.abmp.xz"))
    for key, bad in (("ag_class", np.array([0, 0, 0, 7], np.int32)), ("ag_infl", np.full(16, 9, np.int32)),
                     ("ag_nb", np.array([2, 40], np.int32)), ("ag_cq", np.full(8, -3, np.int32)),
                     ("ag_nb_ptr", np.array([0, 2, 1, 2, 2], np.int32)), ("ag_meta", np.array([4, 9, 2, 1], np.int32))):
        q = dict(p)
        q[key] = bad
        with pytest.raises(S.AbmError):
            S.validate_problem(q)
