"""Host-side basis builder: the reference's TableauNormMinimiser (TableauNormMinimiser.h:201-334) behind the C-ABI
(`abm_basis_*`, include/abmcmc_b200.h).  Same pivots in the same order as the reference, hence the same basis; the
Markowitz search visits eligible rows / columns only, which removes the reference's quadratic cost."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import check, load


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def minimise(con_ptr, con_idx, con_val, lower, upper) -> dict:
    """Constraints (CSR, int32) -> the tableau arrays of an ABMP problem file (dims, tab_L, tab_U, tab_F, col_basic,
    nb_col, nb_ptr, nb_row, nb_val; layout of include/abmcmc_b200/reference_flatten.hpp) plus `basis` and the full rows."""
    L = load()
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (con_ptr, con_idx, con_val, lower, upper)]
    n = arrs[3].size
    h = C.c_void_p()
    check(L.abm_basis_minimise(n, *(_p(a, C.c_int32) for a in arrs), C.byref(h)))
    try:
        nr, nc, na, nb = (C.c_int32() for _ in range(4))
        nnz, nnzr = C.c_int64(), C.c_int64()
        check(L.abm_basis_get_dims(h, C.byref(nr), C.byref(nc), C.byref(na), C.byref(nb), C.byref(nnz), C.byref(nnzr)))
        out = {"tab_L": np.zeros(nr.value, np.int32), "tab_U": np.zeros(nr.value, np.int32), "tab_F": np.zeros(nr.value, np.int32),
               "basis": np.zeros(nr.value, np.int32), "col_basic": np.zeros(nc.value, np.uint8), "nb_col": np.zeros(nb.value, np.int32),
               "nb_ptr": np.zeros(nb.value + 1, np.int32), "nb_row": np.zeros(nnz.value, np.int32), "nb_val": np.zeros(nnz.value, np.int32)}
        check(L.abm_basis_get(h, _p(out["tab_L"], C.c_int32), _p(out["tab_U"], C.c_int32), _p(out["tab_F"], C.c_int32),
                              _p(out["basis"], C.c_int32), _p(out["col_basic"], C.c_uint8), _p(out["nb_col"], C.c_int32),
                              _p(out["nb_ptr"], C.c_int32), _p(out["nb_row"], C.c_int32), _p(out["nb_val"], C.c_int32)))
        out["row_ptr"] = np.zeros(nr.value + 1, np.int64)
        out["row_col"] = np.zeros(nnzr.value, np.int32)
        out["row_val"] = np.zeros(nnzr.value, np.int32)
        check(L.abm_basis_get_rows(h, _p(out["row_ptr"], C.c_int64), _p(out["row_col"], C.c_int32), _p(out["row_val"], C.c_int32)))
        out["dims"] = np.array([nr.value, nc.value, na.value, nb.value, nnz.value], np.int32)
    finally:
        L.abm_basis_destroy(h)
    return out
