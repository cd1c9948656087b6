// Drop-in for the reference's basis construction, for the reference's own host code.
//
// Compiled TOGETHER WITH the reference headers (it names their types, defines none): include TableauNormMinimiser.h first.
//     TableauNormMinimiser<T> tableau = abmcmc::minimiseBasis(posterior.constraints);
// replaces
//     TableauNormMinimiser<T> tableau(posterior.constraints);            // TableauNormMinimiser.h:201-246, runs
//     tableau.findMinimalBasis();                                        // findMinimalBasis (:248-256) itself
// (PredPreyProblem.h:52-55) and yields the same object: rows, cols, isBasic, basis, nAuxiliaryVars, F, L, U are equal member
// by member (oracle/ref/basis_check.cpp compares them).  Only the two work lists of the elimination (colsBySparsity /
// rowsBySparsity) are left empty: nothing reads them once the basis exists.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../abmcmc_b200.h"

namespace abmcmc {

template <class T>
TableauNormMinimiser<T> minimiseBasis(const ConvexPolyhedron<T> &problem) {
    std::vector<int32_t> ptr{0}, idx, val, lo, up;
    for (const Constraint<T> &c : problem) {
        for (int k = 0; k < c.coefficients.sparseSize(); ++k) {
            idx.push_back(c.coefficients.indices[k]);
            val.push_back(int32_t(c.coefficients.values[k]));
            if (T(val.back()) != c.coefficients.values[k]) throw std::runtime_error("minimiseBasis: non-integer coefficient");
        }
        ptr.push_back(int32_t(idx.size()));
        lo.push_back(int32_t(c.lowerBound));
        up.push_back(int32_t(c.upperBound));
    }
    abm_basis *b = nullptr;
    if (abm_basis_minimise(int32_t(lo.size()), ptr.data(), idx.data(), val.data(), lo.data(), up.data(), &b) != ABM_OK)
        throw std::runtime_error(std::string("abm_basis_minimise: ") + abm_last_error());
    int32_t nRows = 0, nCols = 0, nAux = 0, nNonBasic = 0;
    int64_t nnzNb = 0, nnzRows = 0;
    abm_basis_get_dims(b, &nRows, &nCols, &nAux, &nNonBasic, &nnzNb, &nnzRows);
    std::vector<int32_t> L(nRows), U(nRows), F(nRows), basis(nRows), nbCol(nNonBasic), nbPtr(nNonBasic + 1), nbRow(nnzNb), nbVal(nnzNb);
    std::vector<uint8_t> isBasic(nCols);
    std::vector<int64_t> rowPtr(nRows + 1);
    std::vector<int32_t> rowCol(nnzRows), rowVal(nnzRows);
    abm_basis_get(b, L.data(), U.data(), F.data(), basis.data(), isBasic.data(), nbCol.data(), nbPtr.data(), nbRow.data(), nbVal.data());
    abm_basis_get_rows(b, rowPtr.data(), rowCol.data(), rowVal.data());
    abm_basis_destroy(b);

    TableauNormMinimiser<T> t;
    t.nAuxiliaryVars = nAux;
    t.rows.resize(nRows);
    t.cols.resize(nCols);
    t.basis.assign(basis.begin(), basis.end());
    t.F.assign(F.begin(), F.end());
    t.L.assign(L.begin(), L.end());
    t.U.assign(U.begin(), U.end());
    for (int i = 0; i < nRows; ++i) {
        t.rows[i].isActive = false;   // every equality row has been reduced (inactivateRow, :404-409)
        for (int64_t k = rowPtr[i]; k < rowPtr[i + 1]; ++k) t.rows[i][rowCol[k]] = T(rowVal[k]);
    }
    for (int j = 0; j < nCols; ++j) t.cols[j].isBasic = isBasic[j] != 0;
    for (int i = 0; i < nRows; ++i)
        if (L[i] == U[i] && basis[i] >= 0 && basis[i] < nCols) t.cols[basis[i]].insert(i);   // setColBasic (:396-402)
    for (int c = 0; c < nNonBasic; ++c)
        for (int k = nbPtr[c]; k < nbPtr[c + 1]; ++k) t.cols[nbCol[c]].insert(nbRow[k]);
    return t;
}

}  // namespace abmcmc
