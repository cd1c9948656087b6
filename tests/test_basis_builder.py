"""The own basis builder (`abm_basis_minimise`, agentbasedmcmc_b200/csrc/abm_basis.cpp) against the UNMODIFIED reference
TableauNormMinimiser (TableauNormMinimiser.h:201-334).  Host code only: runs without a GPU.

Fixtures: tests/golden/constraints_*.abmp.xz hold `posterior.constraints` exactly as the reference assembles them
(`abm_ref constraints`, tests/golden/make_golden.py); the matching problem fixtures hold the tableau the reference's own
minimiser made of them.  Everything is integer: the comparison is exact, entry for entry."""
import time

import numpy as np
import pytest

from conftest import load_golden

KEYS = ("dims", "tab_L", "tab_U", "tab_F", "col_basic", "nb_col", "nb_ptr", "nb_row", "nb_val")


# exp_pp3-2-single: the reference's validation experiment (Experiments.cpp:49-65); its tableau has an entry of magnitude 2
@pytest.mark.parametrize("name", ["pp8-4", "pp16-8", "pp16-8-poisson", "pp32-16", "exp_pp3-2-single"])
def test_basis_equals_reference_tableau(name):
    from agentbasedmcmc_b200 import basis
    c, p = load_golden(f"constraints_{name}.abmp.xz"), load_golden(f"{name}.abmp.xz")
    t0 = time.time()
    b = basis.minimise(c["con_ptr"], c["con_idx"], c["con_val"], c["con_lower"], c["con_upper"])
    dt = time.time() - t0
    for k in KEYS:
        assert np.array_equal(b[k], p[k]), k
    # the reference needed gen_seconds for the whole problem (dominated by TableauNormMinimiser: 211-320 s at 32x32x16)
    if "gen_seconds" in p:
        assert dt < max(5.0, 0.1 * float(p["gen_seconds"][0]))
    # structure of the result: every equality row reduced (its basic variable is a column), one auxiliary per inequality
    basic_rows = b["basis"][(b["tab_L"] == b["tab_U"])]
    assert np.all(basic_rows >= 0) and np.all(basic_rows < b["dims"][1]) and np.unique(basic_rows).size == basic_rows.size
    assert np.all(b["col_basic"][basic_rows] == 1) and int(b["col_basic"].sum()) == basic_rows.size
    aux = b["basis"][(b["tab_L"] != b["tab_U"])]
    assert np.array_equal(aux, -np.arange(1, aux.size + 1))


def test_basis_rows_satisfy_the_constraints():
    """Size-independent property: every reduced row is a linear combination of the original equalities, so any X that
    satisfies the original constraints satisfies  sum_j M_ij X_j + F_i = (auxiliary or 0)  row by row.  Checked with the
    real trajectory the problem was generated from (a feasible point of the polyhedron)."""
    from agentbasedmcmc_b200 import basis
    c, p = load_golden("constraints_pp16-8.abmp.xz"), load_golden("pp16-8.abmp.xz")
    b = basis.minimise(c["con_ptr"], c["con_idx"], c["con_val"], c["con_lower"], c["con_upper"])
    n_cols = int(b["dims"][1])
    x = np.zeros(n_cols, np.int64)
    x[p["rt_idx"]] = p["rt_val"]
    # original constraints hold for the real trajectory
    lhs = np.add.reduceat(c["con_val"].astype(np.int64) * x[c["con_idx"]], c["con_ptr"][:-1])
    assert np.all(lhs >= c["con_lower"]) and np.all(lhs <= c["con_upper"])
    # reduced rows: equality rows give basic = -(sum over the other columns) - F ... i.e. the full row dotted with X plus F is 0;
    # inequality rows give a value inside [L, U]
    rp, rc, rv = b["row_ptr"], b["row_col"], b["row_val"].astype(np.int64)
    val = np.add.reduceat(rv * x[rc], rp[:-1]) + b["tab_F"]
    eq = b["tab_L"] == b["tab_U"]
    assert np.all(val[eq] == 0)
    assert np.all(val[~eq] >= b["tab_L"][~eq]) and np.all(val[~eq] <= b["tab_U"][~eq])


def test_basis_rejects_bad_input():
    from agentbasedmcmc_b200 import basis
    from agentbasedmcmc_b200._capi import AbmError
    with pytest.raises(AbmError):
        basis.minimise([0, 1], [-3], [1], [0], [0])
    with pytest.raises(AbmError):   # an equality whose only coefficient is not a unit: no admissible pivot
        basis.minimise([0, 1], [0], [2], [4], [4])


@pytest.mark.parametrize("grid,T,seed,start", [(3, 2, 1, "bernoulli"), (3, 3, 2, "poisson"), (4, 3, 3, "bernoulli"), (4, 4, 4, "poisson"),
                                               (8, 2, 5, "bernoulli"), (8, 3, 6, "poisson"), (8, 4, 7, "bernoulli"), (16, 2, 8, "poisson")])
def test_basis_equals_reference_on_fresh_problems(grid, T, seed, start, tmp_path):
    """Differential run against the unmodified reference on problems that are not fixtures: the reference's generator
    (`abm_ref gen`: PredPreyProblem's constructor, other seeds, grids, horizons and both start states) builds the problem and
    its tableau; the own builder must reproduce that tableau from the exported constraints, entry for entry."""
    import subprocess
    from pathlib import Path
    from agentbasedmcmc_b200 import basis
    from agentbasedmcmc_b200.abmp import read_abmp
    ref = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "abm_ref"
    if not ref.exists():
        pytest.skip("oracle/_ref/abm_ref not built (needs /root/reference at build time)")
    prob, cons = tmp_path / "p.abmp", tmp_path / "c.abmp"
    dens = "0.1" if grid <= 4 else "0.05"
    subprocess.run([str(ref), "gen", str(grid), str(T), str(seed), "10.0", str(prob), dens, dens, "0.1", "1.0", start], check=True,
                   capture_output=True, timeout=600)
    subprocess.run([str(ref), "constraints", str(prob), str(cons), "0"], check=True, capture_output=True, timeout=600)
    p, c = read_abmp(prob), read_abmp(cons)
    b = basis.minimise(c["con_ptr"], c["con_idx"], c["con_val"], c["con_lower"], c["con_upper"])
    for k in KEYS:
        assert np.array_equal(b[k], p[k]), k
