// Host-side basis builder: what the reference's TableauNormMinimiser does before the sampler can start
// (TableauNormMinimiser.h:201-334, paths relative to /root/reference/src), behind the C-ABI of include/abmcmc_b200.h.
//
// The reference turns the constraints L <= C X <= U into a tableau with one auxiliary variable per inequality and
// eliminates every equality row by pivoting, choosing pivots with a bounded Markowitz search over rows and columns
// bucketed by their number of non-zeros.  Its result depends on every tie-break: the order of the bucket lists (entries
// are pushed to the FRONT, :340-341 / :357-358), the ascending iteration of std::map rows and std::set columns, the
// strict comparisons of the search, and the early exits after `maxVectorsToTest` candidates.  All of that is kept here,
// so the basis is the reference's, entry for entry.
//
// What is different is the cost.  The reference's search walks the bucket lists in full and skips, one by one, every
// column that no longer has an active row with a unit entry (sparsestRowInCol() == -1, :425-437) -- after a few thousand
// pivots that is most columns, which makes the whole elimination quadratic (211 s at 32x32x16, 1094 s at 32x32x32 here).
// A column's or row's eligibility can only change when the pivot touches it, and a pivot re-inserts everything it
// touches at the front of its bucket; so a second set of lists that holds only the eligible entries, updated at exactly
// those moments, preserves the relative order and lets the search visit candidates only.  Rows and columns are sorted
// vectors instead of red-black trees.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "abmcmc_b200.h"

int abmSetError(int code, const std::string &msg);   // abm_capi.cu: records the last-error string, returns code

namespace {

// Buckets of intrusive doubly-linked lists: item -> bucket; push_front / erase in O(1); `size` mirrors the
// std::vector<std::list<int>>::size() the reference's search bounds itself with (grown on insertion, trailing empty
// buckets popped on removal, TableauNormMinimiser.h:338-368).
struct Buckets {
    std::vector<int> head, prev, next, bucketOf, count;
    int size = 0;
    explicit Buckets(int nItems) : prev(nItems, -1), next(nItems, -1), bucketOf(nItems, -1) {}
    void pushFront(int item, int k) {
        if ((int)head.size() <= k) { head.resize(k + 1, -1); count.resize(k + 1, 0); }
        if (size <= k) size = k + 1;
        prev[item] = -1;
        next[item] = head[k];
        if (head[k] >= 0) prev[head[k]] = item;
        head[k] = item;
        bucketOf[item] = k;
        ++count[k];
    }
    void erase(int item) {
        const int k = bucketOf[item];
        if (prev[item] >= 0) next[prev[item]] = next[item]; else head[k] = next[item];
        if (next[item] >= 0) prev[next[item]] = prev[item];
        bucketOf[item] = -1;
        --count[k];
    }
    bool contains(int item) const { return bucketOf[item] >= 0; }
    void popEmptyBack() { while (size > 0 && count[size - 1] == 0) --size; }
    int first(int k) const { return k < (int)head.size() ? head[k] : -1; }
};

typedef std::pair<int, int> Entry;   // (index, value)

struct Tableau {
    std::vector<std::vector<Entry>> rows;   // by row: (column, value), ascending column  (Row = std::map<int,T>, :80)
    std::vector<std::vector<int>> cols;     // by column: non-reduced rows, ascending      (Column = std::set<int>, :62)
    std::vector<char> rowActive, colBasic;
    std::vector<int> F, L, U, basis;
    int nAux = 0;
    Buckets colsBySparsity{0}, rowsBySparsity{0};   // every non-basic column / active row (defines `size`)
    Buckets liveCols{0}, liveRows{0};               // those of them the search can use, in the same relative order

    static std::vector<Entry>::iterator findIn(std::vector<Entry> &r, int j) {
        return std::lower_bound(r.begin(), r.end(), j, [](const Entry &e, int key) { return e.first < key; });
    }
    int at(int i, int j) {
        auto it = findIn(rows[i], j);
        return it->second;   // callers only ask for entries that exist (std::map::at would throw otherwise)
    }
    static void insertSorted(std::vector<int> &v, int x) {
        auto it = std::lower_bound(v.begin(), v.end(), x);
        if (it == v.end() || *it != x) v.insert(it, x);
    }
    static void eraseSorted(std::vector<int> &v, int x) {
        auto it = std::lower_bound(v.begin(), v.end(), x);
        if (it != v.end() && *it == x) v.erase(it);
    }

    // sparsestRowInCol (:425-437) / sparsestColInRow (:411-423): first strict minimum in ascending index order
    int sparsestRowInCol(int j) {
        int best = INT_MAX, mini = -1;
        for (int i : cols[j]) {
            const int sz = (int)rows[i].size();
            if (sz < best && rowActive[i] && std::abs(at(i, j)) == 1) { best = sz; mini = i; }
        }
        return mini;
    }
    int sparsestColInRow(int i) {
        int best = INT_MAX, minj = -1;
        for (const Entry &e : rows[i]) {
            const int sz = (int)cols[e.first].size();
            if (sz < best && !colBasic[e.first] && std::abs(e.second) == 1) { best = sz; minj = e.first; }
        }
        return minj;
    }

    void addCol(int j) {   // addColSparsityEntry (:354-360)
        const int k = (int)cols[j].size();
        colsBySparsity.pushFront(j, k);
        if (sparsestRowInCol(j) != -1) liveCols.pushFront(j, k);
    }
    void removeCol(int j) {   // removeColSparsityEntry (:362-368)
        colsBySparsity.erase(j);
        colsBySparsity.popEmptyBack();
        if (liveCols.contains(j)) liveCols.erase(j);
    }
    void addRow(int i) {   // addRowSparsityEntry (:338-343)
        const int k = (int)rows[i].size();
        rowsBySparsity.pushFront(i, k);
        if (sparsestColInRow(i) != -1) liveRows.pushFront(i, k);
    }
    void removeRow(int i) {   // removeRowSparsityEntry (:345-351)
        rowsBySparsity.erase(i);
        rowsBySparsity.popEmptyBack();
        if (liveRows.contains(i)) liveRows.erase(i);
    }

    // Pivot selection.  Semantics (and only the semantics) of TableauNormMinimiser::findMarkowitzPivot (:252-285): candidates
    // are visited by increasing sparsity k -- first the columns with k entries, each paired with its sparsest eligible row,
    // then the rows with k entries, each paired with its sparsest eligible column -- the candidate with the strictly smallest
    // Markowitz count (r - 1)(c - 1) so far is kept, and the search stops at once when that count cannot be beaten any more
    // at this sparsity ((k - 1)^2), or after ten candidates (:57).  The visiting order is what makes the basis the
    // reference's, so it is kept; the bucket lists hold eligible vectors only, which is where the speed comes from.
    std::pair<int, int> findMarkowitzPivot() {
        constexpr int kCandidateLimit = 10;
        std::pair<int, int> chosen(-1, -1);
        int chosenCount = INT_MAX, seen = 0;
        bool done = false;
        // returns true when the search is over
        auto consider = [&](int row, int col, int k) {
            const int count = ((int)rows[row].size() - 1) * ((int)cols[col].size() - 1);
            if (count < chosenCount) {
                chosen = std::make_pair(row, col);
                if (count <= (k - 1) * (k - 1)) return true;
                chosenCount = count;
            }
            return ++seen >= kCandidateLimit;
        };
        for (int k = 1; !done && k < colsBySparsity.size && k < rowsBySparsity.size; ++k) {
            for (int col = liveCols.first(k); !done && col >= 0; col = liveCols.next[col]) {
                const int row = sparsestRowInCol(col);
                if (row != -1) done = consider(row, col, k);
            }
            for (int row = liveRows.first(k); !done && row >= 0; row = liveRows.next[row]) {
                const int col = sparsestColInRow(row);
                if (col != -1) done = consider(row, col, k);
            }
        }
        return chosen;
    }

    // One elimination step (what TableauNormMinimiser::pivot does, :288-335): the pivot row is scaled so that its pivot entry
    // becomes -1, then added to every other row of the pivot column so that the column keeps its pivot entry only.  Entries are
    // integers and pivots have unit magnitude (ABM::occupation_type in the reference), so the scale factor is exactly -1 or +1.
    void pivot(int prow, int pcol) {
        const int pivotEntry = at(prow, pcol);
        for (const Entry &e : rows[prow]) removeCol(e.first);
        const int scale = (int)(-1.0 / pivotEntry);
        for (Entry &e : rows[prow]) e.second *= scale;
        F[prow] = (int)(F[prow] * (-1.0 / pivotEntry));
        const std::vector<Entry> &src = rows[prow];
        for (int i : cols[pcol]) {
            if (i == prow) continue;
            if (rowActive[i]) removeRow(i);
            const int factor = at(i, pcol);
            std::vector<Entry> &dst = rows[i];
            for (const Entry &pe : src) {
                const int j = pe.first;
                if (j == pcol) continue;
                auto it = findIn(dst, j);
                if (it == dst.end() || it->first != j) {          // the row gains an entry in column j
                    dst.insert(it, Entry(j, factor * pe.second));
                    insertSorted(cols[j], i);
                } else {
                    it->second += factor * pe.second;
                    if (it->second == 0) {                        // the entry cancels
                        eraseSorted(cols[j], i);
                        dst.erase(it);
                    }
                }
            }
            F[i] += factor * F[prow];
        }
        for (int i : cols[pcol]) {
            if (i == prow) continue;
            std::vector<Entry> &dst = rows[i];
            auto it = findIn(dst, pcol);
            if (it != dst.end() && it->first == pcol) dst.erase(it);
        }
        // the column becomes basic before the rows are re-listed: a row's eligibility looks at colBasic, and the reference's
        // rows no longer contain pcol at that point either
        colBasic[pcol] = 1;
        for (int i : cols[pcol])
            if (i != prow && rowActive[i]) addRow(i);
        cols[pcol].assign(1, prow);   // setColBasic (:396-402)
        basis[prow] = pcol;
        // the pivot row stops being active before its columns are re-listed: their eligibility must not count it, exactly as
        // the reference's later sparsestRowInCol() calls would not
        removeRow(prow);              // inactivateRow (:404-409)
        rowActive[prow] = 0;
        for (const Entry &pe : rows[prow]) {
            if (pe.first == pcol) continue;
            eraseSorted(cols[pe.first], prow);   // column lists hold the rows that are still to be reduced
            addCol(pe.first);
        }
    }
};

}  // namespace

struct abm_basis {
    Tableau t;
    long long nPivots = 0;
};

extern "C" {

int abm_basis_minimise(int32_t nConstraints, const int32_t *conPtr, const int32_t *conIdx, const int32_t *conVal,
                       const int32_t *lower, const int32_t *upper, abm_basis **out) {
    if (!out || nConstraints <= 0 || !conPtr || !conIdx || !conVal || !lower || !upper) return abmSetError(ABM_ERR_INVALID, "bad argument");
    *out = nullptr;
    abm_basis *b = new abm_basis();
    Tableau &t = b->t;
    // TableauNormMinimiser(const ConvexPolyhedron&) (:201-246)
    t.rows.resize(nConstraints);
    t.rowActive.assign(nConstraints, 0);
    t.F.assign(nConstraints, 0);
    t.L.assign(nConstraints, 0);
    t.U.assign(nConstraints, 0);
    t.basis.assign(nConstraints, 0);
    t.rowsBySparsity = Buckets(nConstraints);
    t.liveRows = Buckets(nConstraints);
    int maxCol = 0;
    for (int i = 0; i < nConstraints; ++i) {
        std::vector<Entry> &r = t.rows[i];
        for (int k = conPtr[i]; k < conPtr[i + 1]; ++k) {   // Row::Row (:371-376): map[index] = value, later duplicates win
            if (conIdx[k] < 0) { delete b; return abmSetError(ABM_ERR_INVALID, "negative variable index"); }
            auto it = Tableau::findIn(r, conIdx[k]);
            if (it != r.end() && it->first == conIdx[k]) it->second = conVal[k]; else r.insert(it, Entry(conIdx[k], conVal[k]));
        }
        if (r.empty()) { delete b; return abmSetError(ABM_ERR_INVALID, "constraint without coefficients"); }
        maxCol = std::max(maxCol, r.back().first);
    }
    const int nCols = maxCol + 1;
    t.cols.resize(nCols);
    t.colBasic.assign(nCols, 0);
    t.colsBySparsity = Buckets(nCols);
    t.liveCols = Buckets(nCols);
    for (int i = 0; i < nConstraints; ++i)
        for (const Entry &e : t.rows[i]) t.cols[e.first].push_back(i);   // ascending i
    for (int i = 0; i < nConstraints; ++i) {
        const bool active = upper[i] == lower[i];
        t.rowActive[i] = active ? 1 : 0;
        if (active) {
            t.basis[i] = INT_MAX;      // fixed variable
            t.F[i] = -lower[i];        // equality constraints become D X - E = 0
        } else {
            t.basis[i] = -(++t.nAux);
            t.L[i] = lower[i];
            t.U[i] = upper[i];
        }
    }
    // the reference lists the active rows while it reads the constraints (before any column exists, so eligibility is not
    // looked at there); listing them afterwards in the same order builds the same lists
    for (int i = 0; i < nConstraints; ++i)
        if (t.rowActive[i]) t.addRow(i);
    for (int j = 0; j < nCols; ++j) t.addCol(j);
    // findMinimalBasis (:248-256)
    while (t.rowsBySparsity.size > 0) {
        const std::pair<int, int> pv = t.findMarkowitzPivot();
        if (pv.first < 0) {
            delete b;
            return abmSetError(ABM_ERR_INVALID, "no unit pivot left for an equality constraint (the reference would fail here too)");
        }
        t.pivot(pv.first, pv.second);
        ++b->nPivots;
    }
    *out = b;
    return ABM_OK;
}

int abm_basis_destroy(abm_basis *b) {
    delete b;
    return ABM_OK;
}

int abm_basis_get_dims(const abm_basis *b, int32_t *nRows, int32_t *nCols, int32_t *nAux, int32_t *nNonBasic, int64_t *nnzNonBasic,
                       int64_t *nnzRows) {
    if (!b) return abmSetError(ABM_ERR_INVALID, "null basis");
    const Tableau &t = b->t;
    int nb = 0;
    int64_t nnz = 0, nnzR = 0;
    for (size_t j = 0; j < t.cols.size(); ++j)
        if (!t.colBasic[j]) { ++nb; nnz += (int64_t)t.cols[j].size(); }
    for (const auto &r : t.rows) nnzR += (int64_t)r.size();
    if (nRows) *nRows = (int32_t)t.rows.size();
    if (nCols) *nCols = (int32_t)t.cols.size();
    if (nAux) *nAux = t.nAux;
    if (nNonBasic) *nNonBasic = nb;
    if (nnzNonBasic) *nnzNonBasic = nnz;
    if (nnzRows) *nnzRows = nnzR;
    return ABM_OK;
}

int abm_basis_get(const abm_basis *b, int32_t *L, int32_t *U, int32_t *F, int32_t *basis, uint8_t *colIsBasic, int32_t *nbCol,
                  int32_t *nbPtr, int32_t *nbRow, int32_t *nbVal) {
    if (!b) return abmSetError(ABM_ERR_INVALID, "null basis");
    const Tableau &t = b->t;
    const size_t nR = t.rows.size(), nC = t.cols.size();
    for (size_t i = 0; i < nR; ++i) {
        if (L) L[i] = t.L[i];
        if (U) U[i] = t.U[i];
        if (F) F[i] = t.F[i];
        if (basis) basis[i] = t.basis[i];
    }
    int32_t c = 0, k = 0;
    if (nbPtr) nbPtr[0] = 0;
    for (size_t j = 0; j < nC; ++j) {
        if (colIsBasic) colIsBasic[j] = (uint8_t)t.colBasic[j];
        if (t.colBasic[j]) continue;
        if (nbCol) nbCol[c] = (int32_t)j;
        for (int i : t.cols[j]) {
            if (nbRow) nbRow[k] = i;
            if (nbVal) {
                auto it = std::lower_bound(t.rows[i].begin(), t.rows[i].end(), (int)j, [](const Entry &e, int key) { return e.first < key; });
                nbVal[k] = it->second;
            }
            ++k;
        }
        ++c;
        if (nbPtr) nbPtr[c] = k;
    }
    return ABM_OK;
}

int abm_basis_get_rows(const abm_basis *b, int64_t *rowPtr, int32_t *colIdx, int32_t *val) {
    if (!b || !rowPtr) return abmSetError(ABM_ERR_INVALID, "null argument");
    const Tableau &t = b->t;
    int64_t k = 0;
    rowPtr[0] = 0;
    for (size_t i = 0; i < t.rows.size(); ++i) {
        for (const Entry &e : t.rows[i]) {
            if (colIdx) colIdx[k] = e.first;
            if (val) val[k] = e.second;
            ++k;
        }
        rowPtr[i + 1] = k;
    }
    return ABM_OK;
}

}  // extern "C"
