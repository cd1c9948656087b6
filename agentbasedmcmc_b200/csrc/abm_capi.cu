// C-ABI implementation (include/abmcmc_b200.h): host-side problem builder, device memory management and
// kernel launches.  No CPU compute fallback exists: without a CUDA device every entry point that would
// compute returns ABM_ERR_CUDA.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "abmcmc_b200.h"
#include "abm_kernels.cuh"

using namespace abm;

int abmSetError(int code, const std::string &msg);   // shared with abm_basis.cpp

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
}  // namespace

int abmSetError(int code, const std::string &msg) { return fail(code, msg); }

namespace {

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ABM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));         \
    } while (0)

struct DeviceArena {
    std::vector<void *> ptrs;
    size_t bytes = 0;
    ~DeviceArena() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <class T>
    cudaError_t alloc(T **out, size_t count) {
        void *p = nullptr;
        size_t b = std::max<size_t>(count, 1) * sizeof(T);
        cudaError_t e = cudaMalloc(&p, b);
        if (e != cudaSuccess) return e;
        ptrs.push_back(p);
        bytes += b;
        *out = static_cast<T *>(p);
        return cudaSuccess;
    }
    template <class T>
    cudaError_t upload(const T **out, const std::vector<T> &v) {
        T *p = nullptr;
        cudaError_t e = alloc(&p, v.size());
        if (e != cudaSuccess) return e;
        if (!v.empty()) e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
        *out = p;
        return e;
    }
};

}  // namespace

struct abm_problem {
    int device = 0;
    DevProblem d{};
    abm_problem_info info{};
    DeviceArena arena;
    std::vector<int32_t> xoff;   // physical Xb offset of each sampler row (host copy for read-back)
};

struct abm_chains {
    int32_t *series = nullptr;   // thinned synopsis series of the first nSeriesChains chains
    int nSeriesChains = 0, seriesCap = 0, seriesThin = 1;
    const abm_problem *p = nullptr;
    DevChains d{};
    DeviceArena arena;
    int32_t n = 0;
    uint64_t seed = 0, firstChain = 0;
    cudaStream_t stream = nullptr;
    bool initialised = false, feasible = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // around the last step / sample launch (abm_chains_last_launch_ms)
    bool timed = false;
    // abm_chains_enqueue_statistics / abm_chains_wait_statistics: pinned staging + the event after its copies
    int64_t *stage = nullptr;
    size_t stageWords = 0;
    cudaEvent_t evStats = nullptr;
    bool statsPending = false;
    // abm_chains_enqueue_moments / abm_chains_wait_moments
    long long *momExact = nullptr;   // device [kMomExactHead + S]; d.endStateSum points at its tail
    double *momReal = nullptr;       // device [kMomReal]
    int64_t *momStage = nullptr;     // pinned: exact, then real
    cudaEvent_t evMom = nullptr;
    bool momPending = false;
    int launchesSinceRebuild = 0;    // blocked tree: step / sample launches since the sums of 32 leaves were last rebuilt from the leaves
};

// CUDA events on the chains' stream around a step kernel, so a caller can time the kernel itself
static cudaError_t markLaunch(abm_chains *c, bool begin) {
    if (!c->ev0) {
        cudaError_t e = cudaEventCreate(&c->ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
        if (e != cudaSuccess) return e;
    }
    if (!begin) c->timed = true;
    return cudaEventRecord(begin ? c->ev0 : c->ev1, c->stream);
}

extern "C" {

static int zeroStatistics(abm_chains *c);

const char *abm_last_error(void) { return g_err.c_str(); }
int abm_version(void) { return 2; }   // 2: abm_problem_desc gained the agent_* tables
int abm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// HBM bytes abm_chains_create allocates per chain (same expressions)
static int64_t chainStateBytes(const DevProblem &P) {
    const int64_t ns = std::max(P.nSynopsis, 1), S = std::max(P.S, 1);
    return int64_t(P.xStride)                                   // Xb
           + int64_t(P.nCols) * (sizeof(double2) + 1)            // colRec, delta
           + int64_t(P.t1Len + P.t2Len + 32) * sizeof(double)    // tier1, tier2, tier3
           + int64_t(P.sCap) * sizeof(double2)                   // scratch
           + 8 + 4 + 8 + 8 * 8                                   // logImp, infeas, iter, counters
           + ns * (8 + 8 + 4) + 8 + 8                            // synSum, synSumSq, synNow, nSamples, launchSamples
           + S * int64_t(sizeof(EndRec));                        // endRec
}

// ------------------------------------------------------------------------------------------------
// hostOnly != null: the host half only (validation, record layout, sizes) -- abm_problem_validate; nothing touches a device
static int problemCreateImpl(const abm_problem_desc *desc, int device, abm_problem **out, abm_problem_info *hostOnly) {
    if (!desc || (!out && !hostOnly)) return fail(ABM_ERR_INVALID, "null argument");
    if (out) *out = nullptr;
    if (!hostOnly) {
        if (abm_device_count() <= device || device < 0) return fail(ABM_ERR_CUDA, "no such CUDA device (there is no CPU fallback)");
        CUDA_TRY(cudaSetDevice(device));
        if (const char *e = getenv("ABM_L2_FETCH")) CUDA_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(atoi(e))));   // experiment: 32 / 64 / 128
    }
    const int nTab = desc->n_tab_rows, nCols = desc->n_nonbasic;
    if (nTab <= 0 || nCols <= 0) return fail(ABM_ERR_INVALID, "empty problem");
    if (!desc->tab_L || !desc->tab_U || !desc->tab_F || !desc->nb_col || !desc->nb_ptr || !desc->fac_id || !desc->ut_lo ||
        !desc->ut_off || !desc->ut_val)
        return fail(ABM_ERR_INVALID, "null array in the problem description");
    if (desc->nb_ptr[0] != 0) return fail(ABM_ERR_INVALID, "nb_ptr must start at 0");
    for (int j = 0; j < nCols; ++j)
        if (desc->nb_ptr[j + 1] < desc->nb_ptr[j]) return fail(ABM_ERR_INVALID, "nb_ptr must be non-decreasing");
    if (desc->nb_ptr[nCols] > 0 && (!desc->nb_row || !desc->nb_val)) return fail(ABM_ERR_INVALID, "null array in the problem description");
    if (desc->n_tables <= 0) return fail(ABM_ERR_INVALID, "no factor tables");
    if (desc->ut_off[0] != 0) return fail(ABM_ERR_INVALID, "ut_off must start at 0");
    for (int k = 0; k < desc->n_tables; ++k)
        if (desc->ut_off[k + 1] <= desc->ut_off[k]) return fail(ABM_ERR_INVALID, "ut_off must be increasing (no empty factor table)");

    // sampler rows = tableau rows that are not equalities, in order (SparseBasisSampler.h:128-140)
    std::vector<int32_t> rowOf(nTab, -1);
    int nRows = 0;
    for (int i = 0; i < nTab; ++i)
        if (desc->tab_L[i] != desc->tab_U[i]) rowOf[i] = nRows++;
    const int nnz = desc->nb_ptr[nCols];
    const int G = desc->grid_size, T = desc->n_timesteps;
    // the agent: PredPrey by its grid geometry, or any other agent type by tables (abmcmc_b200.h); either way the host works
    // from the same four tables: class, influencers, neighbours(), consequences()
    const bool tableAgent = G == 0 && desc->n_agent_states > 0;
    if (G < 0 || G > 4096 || desc->n_agent_states < 0 || desc->n_agent_states > (1 << 22) || T < 0 || T > (1 << 16))
        return fail(ABM_ERR_INVALID, "agent description out of range (grid size, agent states or timesteps)");
    const int A = tableAgent ? desc->n_acts : kActs;
    const int S = tableAgent ? desc->n_agent_states : 2 * G * G, nStates = T * S, nEvents = nStates * A;
    const int nCls = tableAgent ? desc->n_agent_classes : 2;
    std::vector<int32_t> agCls(std::max(S, 0)), agInfl(size_t(std::max(S, 0)) * 4, -1), agNbPtr(std::max(S, 0) + 1, 0), agNb,
        agCqPtr(size_t(std::max(S, 0)) * A + 1, 0), agCq;
    if (tableAgent) {
        if (A < 1 || A > kActs || T < 1 || nCls < 1 || nCls > 1024) return fail(ABM_ERR_INVALID, "agent tables: 1..6 acts, at least one timestep and class");
        if (!desc->agent_class || !desc->agent_influencers || !desc->agent_nb_ptr || !desc->agent_cq_ptr)
            return fail(ABM_ERR_INVALID, "null array in the agent tables");
        if (desc->agent_nb_ptr[0] != 0 || desc->agent_cq_ptr[0] != 0) return fail(ABM_ERR_INVALID, "agent tables: CSR pointers must start at 0");
        for (int q = 0; q < S; ++q) {
            if (desc->agent_class[q] < 0 || desc->agent_class[q] >= nCls) return fail(ABM_ERR_INVALID, "agent tables: class out of range");
            agCls[q] = desc->agent_class[q];
            for (int k = 0; k < 4; ++k) {
                const int v = desc->agent_influencers[4 * q + k];
                if (v < -1 || v >= S) return fail(ABM_ERR_INVALID, "agent tables: influencer out of range");
                agInfl[4 * q + k] = v;
            }
            if (desc->agent_nb_ptr[q + 1] < desc->agent_nb_ptr[q]) return fail(ABM_ERR_INVALID, "agent tables: neighbour pointers must be non-decreasing");
        }
        for (int q = 0; q < S * A; ++q)
            if (desc->agent_cq_ptr[q + 1] < desc->agent_cq_ptr[q]) return fail(ABM_ERR_INVALID, "agent tables: consequence pointers must be non-decreasing");
        if ((desc->agent_nb_ptr[S] > 0 && !desc->agent_nb) || (desc->agent_cq_ptr[S * A] > 0 && !desc->agent_cq))
            return fail(ABM_ERR_INVALID, "null array in the agent tables");
        agNbPtr.assign(desc->agent_nb_ptr, desc->agent_nb_ptr + S + 1);
        agNb.assign(desc->agent_nb, desc->agent_nb + agNbPtr[S]);
        agCqPtr.assign(desc->agent_cq_ptr, desc->agent_cq_ptr + S * A + 1);
        agCq.assign(desc->agent_cq, desc->agent_cq + agCqPtr[S * A]);
        for (int v : agNb) if (v < 0 || v >= S) return fail(ABM_ERR_INVALID, "agent tables: neighbour out of range");
        for (int v : agCq) if (v < 0 || v >= S) return fail(ABM_ERR_INVALID, "agent tables: consequence out of range");
    } else if (G > 0) {
        const int GG = G * G;
        for (int psi = 0; psi < S; ++psi) {
            const int type = psi / GG, cell = psi % GG, cy = cell / G, cx = cell % G;
            const int ob = (1 - type) * GG, tb = type * GG;
            const int xl = (cx + G - 1) % G, xr = (cx + 1) % G, yu = (cy + 1) % G, yd = (cy + G - 1) % G;
            agCls[psi] = type;
            // PredPreyAgent.h:199-207 (neighbours: left, right, up, down of the other species) -- also what timestep() reads (:124-157)
            const int nb4[4] = {ob + xl + G * cy, ob + xr + G * cy, ob + cx + G * yu, ob + cx + G * yd};
            for (int k = 0; k < 4; ++k) { agInfl[4 * psi + k] = nb4[k]; agNb.push_back(nb4[k]); }
            agNbPtr[psi + 1] = int32_t(agNb.size());
            // PredPreyAgent::consequences (PredPreyAgent.h:98-115): die -> nothing; the four moves; give birth -> parent and child
            for (int a2 = 0; a2 < kActs; ++a2) {
                if (a2 == 1) agCq.push_back(tb + xl + G * cy);
                else if (a2 == 2) agCq.push_back(tb + xr + G * cy);
                else if (a2 == 3) agCq.push_back(tb + cx + G * yu);
                else if (a2 == 4) agCq.push_back(tb + cx + G * yd);
                else if (a2 == 5) { agCq.push_back(tb + cx + G * cy); agCq.push_back(tb + xr + G * cy); }
                agCqPtr[psi * kActs + a2 + 1] = int32_t(agCq.size());
            }
        }
    }
    if (G > 0 || tableAgent) {
        if (nEvents > nRows) return fail(ABM_ERR_INVALID, "the trajectory importance needs one row per event");
        for (int i = 0; i < nEvents; ++i)
            if (rowOf[i] != i) return fail(ABM_ERR_INVALID, "event rows must be the first, non-equality tableau rows");
        if (!desc->imp_lgamma || !desc->imp_logratio) return fail(ABM_ERR_INVALID, "missing importance tables");
    }

    std::vector<int32_t> colPtr(desc->nb_ptr, desc->nb_ptr + nCols + 1), colRow(nnz), colVal(nnz);
    std::vector<int32_t> rowPtr(nRows + 1, 0), rowCol(nnz), rowVal(nnz), rowF(nRows);
    std::vector<long> xmin(nRows), xmax(nRows);
    for (int i = 0; i < nTab; ++i)
        if (rowOf[i] >= 0) rowF[rowOf[i]] = desc->tab_F[i], xmin[rowOf[i]] = xmax[rowOf[i]] = desc->tab_F[i];
    int maxColNnz = 0, maxAbs = 1;
    constexpr int kMaxAbsEntry = 4;
    for (int j = 0; j < nCols; ++j) {
        maxColNnz = std::max(maxColNnz, colPtr[j + 1] - colPtr[j]);
        for (int k = colPtr[j]; k < colPtr[j + 1]; ++k) {
            const int r = desc->nb_row[k] >= 0 && desc->nb_row[k] < nTab ? rowOf[desc->nb_row[k]] : -1;
            if (r < 0) return fail(ABM_ERR_INVALID, "non-basic column has an entry in an equality row");
            if (k > colPtr[j] && colRow[k - 1] >= r) return fail(ABM_ERR_INVALID, "column entries must be ascending by row");
            colRow[k] = r;
            colVal[k] = desc->nb_val[k];
            // the paper's PredPrey bases have unit entries only; small problems (the reference's 3x3
            // validation experiment) do not.  The kernels tabulate, per row of a proposal, every move -maxAbs..maxAbs a
            // coupled column could add; the bound keeps that table small.
            if (colVal[k] == 0 || colVal[k] < -kMaxAbsEntry || colVal[k] > kMaxAbsEntry)
                return fail(ABM_ERR_INVALID, "tableau entries must be non-zero and at most 4 in magnitude");
            maxAbs = std::max(maxAbs, std::abs(colVal[k]));
            rowPtr[r + 1]++;
            (colVal[k] < 0 ? xmin[r] : xmax[r]) += colVal[k];
        }
    }
    for (int i = 0; i < nRows; ++i) {
        rowPtr[i + 1] += rowPtr[i];
        if (xmin[i] < -128 || xmax[i] > 127) return fail(ABM_ERR_INVALID, "row value range exceeds int8");
    }
    {
        std::vector<int32_t> fill(rowPtr.begin(), rowPtr.end() - 1);
        for (int j = 0; j < nCols; ++j)   // rows[i] gets its entries in ascending column order (:148-153)
            for (int k = colPtr[j]; k < colPtr[j + 1]; ++k) {
                rowCol[fill[colRow[k]]] = j;
                rowVal[fill[colRow[k]]++] = colVal[k];
            }
    }

    // physical layout of X: event rows 8 bytes per agent state, then the remaining rows
    // (with the kernels' table path -- grids that are not a power of two, table agents -- one more, always-zero state word follows:
    // the word an unused influencer slot reads)
    const bool pow2Grid = G > 0 && (G & (G - 1)) == 0;
    const int evBytes = (nStates + (nStates > 0 && !pow2Grid ? 1 : 0)) * kStateBytes;
    std::vector<int32_t> xoff(nRows);
    for (int i = 0; i < nRows; ++i)
        xoff[i] = i < nEvents ? (i / A) * kStateBytes + i % A : evBytes + (i - nEvents);
    const int xStride = ((evBytes + (nRows - nEvents)) + 15) & ~15;
    if (xStride >= (1 << 24)) return fail(ABM_ERR_INVALID, "too many rows");

    // factor tables
    int nVal = desc->ut_off[desc->n_tables];
    // kFacPad zeros on both sides: a proposal tabulates every move up to +-maxAbs of each of its rows, including moves no
    // coupled column can make, whose (unused) table look-ups may fall just outside the row's reachable range
    constexpr int kFacPad = 8;
    std::vector<double> facVal(size_t(nVal) + 2 * kFacPad, 0.0);
    std::copy(desc->ut_val, desc->ut_val + nVal, facVal.begin() + kFacPad);
    std::vector<int4> rowParam(nRows);
    for (int i = 0, b = 0; i < nTab; ++i) {
        if (rowOf[i] < 0) continue;
        const int id = desc->fac_id[b];
        if (id < 0 || id >= desc->n_tables) return fail(ABM_ERR_INVALID, "bad factor id");
        const int lo = desc->ut_lo[id], len = desc->ut_off[id + 1] - desc->ut_off[id];
        const long L = desc->tab_L[i], U = desc->tab_U[i];
        const long dlo = std::min(std::max(xmin[b], L), U), dhi = std::max(std::min(xmax[b], U), L);
        if (dlo < lo || dhi > lo + len - 1) return fail(ABM_ERR_INVALID, "factor table does not cover the reachable range of its row");
        rowParam[b] = make_int4(xoff[b], desc->tab_L[i], desc->tab_U[i], kFacPad + desc->ut_off[id] - lo);
        ++b;
    }

    // lanes per chain (the records do not depend on it)
    int W = 16;
    if (const char *e = getenv("ABM_GROUP_WIDTH")) W = atoi(e);
    if (W != 8 && W != 16 && W != 32) return fail(ABM_ERR_INVALID, "ABM_GROUP_WIDTH must be 8, 16 or 32");

    // energy tables: energy(i, x) (SparseBasisSampler.h:341-345) of every row class {L, U, factor table}, tabulated over the
    // values a step can ask for (the row's reachable range widened by the largest entry, for the moves of coupled columns),
    // with exactly the arithmetic the kernels used to do per evaluation: factor(clamp(x)) - kappa * distance, each operation
    // rounded once.  A row then carries one 16-bit table base and energy(i, x) is a single load.
    constexpr int kRowsMax = 24;   // rows per column a record can describe (5 bits of a pair code, 24..31 reserved)
    if (maxColNnz > kRowsMax) return fail(ABM_ERR_INVALID, "columns with more than 24 entries are not supported");
    std::vector<double> eTab(256, 0.0);
    std::vector<int32_t> rowEBase(nRows);
    {
        struct Cls { long L, U; int id; long lo, hi; std::vector<int> rows; };
        std::vector<Cls> classes;
        std::vector<int> clsOf(nRows);
        for (int i = 0, b = 0; i < nTab; ++i) {
            if (rowOf[i] < 0) continue;
            const long L = desc->tab_L[i], U = desc->tab_U[i];
            const int id = desc->fac_id[b];
            int c = -1;
            for (size_t q = 0; q < classes.size(); ++q)
                if (classes[q].L == L && classes[q].U == U && classes[q].id == id) { c = int(q); break; }
            if (c < 0) { classes.push_back({L, U, id, xmin[b], xmax[b], {}}); c = int(classes.size()) - 1; }
            classes[c].lo = std::min(classes[c].lo, xmin[b]);
            classes[c].hi = std::max(classes[c].hi, xmax[b]);
            clsOf[b] = c;
            ++b;
        }
        std::vector<int> base(classes.size());
        const long margin = maxAbs + 1;
        for (size_t q = 0; q < classes.size(); ++q) {
            const Cls &c = classes[q];
            const long clo = c.lo - margin, chi = c.hi + margin;
            const long b0 = long(eTab.size()) - clo;   // energy(x) = eTab[b0 + x]
            if (b0 < 0 || b0 + chi >= 65536 + 128) return fail(ABM_ERR_INVALID, "energy tables too large");
            base[q] = int(b0);
            const int lo = desc->ut_lo[c.id], len = desc->ut_off[c.id + 1] - desc->ut_off[c.id];
            for (long x = clo; x <= chi; ++x) {
                const long cl = std::max(c.L, std::min(x, c.U));
                const long dist = x < c.L ? c.L - x : (x > c.U ? x - c.U : 0);
                const long idx = std::min<long>(std::max<long>(cl - lo, 0), len - 1);   // outside the table: unreachable values only
                volatile double f = desc->ut_val[desc->ut_off[c.id] + idx];
                volatile double pen = desc->kappa * double(dist);
                volatile double e = f - pen;
                eTab.push_back(double(e));
            }
        }
        for (int b = 0; b < nRows; ++b) {
            rowEBase[b] = base[clsOf[b]];
            if (rowEBase[b] < 0 || rowEBase[b] > 65535) return fail(ABM_ERR_INVALID, "energy table base out of range");
        }
        eTab.insert(eTab.end(), 256, 0.0);
    }

    // column records (layout: abm_device.h)
    bool blockedTree = desc->sum_tree != ABM_TREE_REFERENCE;
    if (const char *e = getenv("ABM_TREE")) blockedTree = std::string(e) != "reference";
    std::vector<uint2> recIdx(nCols);
    std::vector<uint32_t> rec;
    rec.reserve((size_t)nnz * 12);
    int maxCoupled = 1, maxTasks = 1, maxSF = 0, maxTA = 1, maxTB = 1, maxTC = 1;
    struct Pair { int u, k, m; };
    std::vector<Pair> pairs;
    std::vector<int32_t> staleF, nodes, endFx;
    std::vector<uint32_t> slots, xtra;
    int nSyn = 0;
    for (int ps = G / 2; ps > 1; ps /= 2) ++nSyn;   // FiguresForPaper.h:246-249
    if (nSyn > 8) return fail(ABM_ERR_INVALID, "more than 8 synopsis statistics");
    std::vector<char> seen(std::max(nStates, 1), 0);
    auto pairCode = [](int k, int m) { return uint32_t(k | ((m > 0 ? m - 1 : 4 + (-m - 1)) << 5)); };   // 5 bits row, 3 bits entry
    size_t maxRecWords = 0;
    for (int j = 0; j < nCols; ++j) {
        while (rec.size() % 4) rec.push_back(0);
        const size_t recStart = rec.size();
        const int n = colPtr[j + 1] - colPtr[j];
        pairs.clear();
        pairs.push_back({j, -1, 0});   // the proposed column itself (changedDE[j] = -currentDE[j], :294)
        for (int k = 0; k < n; ++k) {
            const int i = colRow[colPtr[j] + k];
            for (int r = rowPtr[i]; r < rowPtr[i + 1]; ++r)
                if (rowCol[r] != j) pairs.push_back({rowCol[r], k, rowVal[r]});
        }
        std::stable_sort(pairs.begin(), pairs.end(), [](const Pair &a, const Pair &b) { return a.u < b.u || (a.u == b.u && a.k < b.k); });
        // distinct columns ascending, each with its shared rows in row order: one 32-bit slot per column
        // {u - uBase (16 bits) | first shared row | second shared row}; a column sharing more than two rows points into `xtra`
        slots.clear();
        xtra.clear();
        nodes.clear();
        const int uBase = pairs.front().u;
        // a column whose coupled columns span more than 16 bits of ids (64x64 grids) gets WIDE slots: the id in a word of its own
        const bool wide = pairs.back().u - uBase > 0xffff;
        for (size_t a2 = 0; a2 < pairs.size();) {
            size_t b2 = a2;
            while (b2 < pairs.size() && pairs[b2].u == pairs[a2].u) ++b2;
            const int u = pairs[a2].u, cnt = pairs[a2].k < 0 ? 0 : int(b2 - a2);
            if (wide) slots.push_back(uint32_t(u));
            uint32_t w = wide ? 0u : uint32_t(u - uBase);
            if (cnt <= 2) {
                w |= (cnt >= 1 ? pairCode(pairs[a2].k, pairs[a2].m) : 0xffu) << 16;
                w |= (cnt >= 2 ? pairCode(pairs[a2 + 1].k, pairs[a2 + 1].m) : 0xffu) << 24;
            } else {
                if (xtra.size() > 255) return fail(ABM_ERR_INVALID, "too many columns sharing more than two rows with one column");
                w |= 0xfeu << 16 | uint32_t(xtra.size()) << 24;
                // list: count, then the pair codes, four per word
                std::vector<uint32_t> bytes{uint32_t(cnt)};
                for (int t = 0; t < cnt; ++t) bytes.push_back(pairCode(pairs[a2 + t].k, pairs[a2 + t].m));
                while (bytes.size() % 4) bytes.push_back(0xff);
                for (size_t q = 0; q < bytes.size(); q += 4) xtra.push_back(bytes[q] | bytes[q + 1] << 8 | bytes[q + 2] << 16 | bytes[q + 3] << 24);
            }
            slots.push_back(w);
            if (!blockedTree) {   // the reference tree's commit lists the nodes it touches: u and its ancestors
                int a3 = u;
                nodes.push_back(a3);
                while (a3) { a3 &= a3 - 1; nodes.push_back(a3); }
            }
            a2 = b2;
        }
        const int C = int(slots.size()) / (wide ? 2 : 1);
        maxCoupled = std::max(maxCoupled, C);
        // distinct tree nodes a commit of this column touches (sizes the reference tree's task list; the blocked tree keeps none)
        std::sort(nodes.begin(), nodes.end());
        if (!blockedTree) {
            const int nn = int(std::unique(nodes.begin(), nodes.end()) - nodes.begin());
            int ta = 0, tb = 0, tc = 0;
            for (int q = 0; q < nn; ++q) {
                const int z = nodes[q] ? __builtin_ctz(nodes[q]) : 31;
                (z >= 10 ? ta : (z >= 5 ? tb : tc))++;
            }
            maxTasks = std::max(maxTasks, nn);
            maxTA = std::max(maxTA, ta); maxTB = std::max(maxTB, tb); maxTC = std::max(maxTC, tc);
        }
        // stale factors in TrajectoryImportance::perturb insertion order (TrajectoryImportance.h:41-53)
        staleF.clear();
        if (nStates > 0) {
            auto ins = [&](int sid) { if (!seen[sid]) { seen[sid] = 1; staleF.push_back(sid); } };
            for (int k = 0; k < n; ++k) {
                const int i = colRow[colPtr[j] + k];
                if (i >= nEvents) continue;
                const int sid = i / A, t = sid / S, psi = sid % S;
                ins(sid);
                for (int q = agNbPtr[psi]; q < agNbPtr[psi + 1]; ++q) ins(t * S + agNb[q]);
            }
            for (int sid : staleF) seen[sid] = 0;
        }
        if (staleF.size() > 255) return fail(ABM_ERR_INVALID, "too many stale factors for one column");
        maxSF = std::max(maxSF, int(staleF.size()));
        int nEv = 0;
        for (int k = 0; k < n; ++k) nEv += colRow[colPtr[j] + k] < nEvents;
        for (int k = 0; k < nEv; ++k)
            if (colRow[colPtr[j] + k] >= nEvents) return fail(ABM_ERR_INVALID, "internal: event rows are not leading");
        if (C > 2040) return fail(ABM_ERR_INVALID, "a column is coupled to more than 2040 others");
        bool touchesEnd = false;
        for (int k = 0; k < nEv; ++k) touchesEnd = touchesEnd || colRow[colPtr[j] + k] >= (T - 1) * S * A;
        // header (4 words), rows (2 words each), stale factors, slots, extra pair lists, end-state effects
        rec.push_back(uint32_t(n) | uint32_t(staleF.size()) << 8 | uint32_t(C) << 16);
        if (xtra.size() > 0x3fff) return fail(ABM_ERR_INVALID, "pair lists of one column too long");
        rec.push_back(uint32_t(nEv) | uint32_t(xtra.size()) << 8 | (wide ? 0x40000000u : 0u) | (touchesEnd ? 0x80000000u : 0u));
        rec.push_back(uint32_t(uBase));
        rec.push_back(0);
        for (int k = 0; k < n; ++k) {
            const int i = colRow[colPtr[j] + k];
            const int4 rp = rowParam[i];
            const long L8 = std::max<long>(-128, std::min<long>(127, rp.y)), U8 = std::max<long>(-128, std::min<long>(127, rp.z));
            rec.push_back(uint32_t(rp.x & 0xffffff) | uint32_t(colVal[colPtr[j] + k] & 0xff) << 24);
            rec.push_back(uint32_t(rowEBase[i]) | uint32_t(L8 & 0xff) << 16 | uint32_t(U8 & 0xff) << 24);
        }
        for (int sid : staleF) rec.push_back(uint32_t(sid));
        rec.insert(rec.end(), slots.begin(), slots.end());
        rec.insert(rec.end(), xtra.begin(), xtra.end());
        // static end-state effect of the column (MODE_SAMPLE), per unit of delta_j: which states of ModelState(traj,T,T) its
        // events of the last timestep feed (PredPreyAgent::consequences, PredPreyAgent.h:98-115) and what that adds to the
        // synopsis squares (FiguresForPaper.h:245-263)
        if (touchesEnd) {
            endFx.clear();
            int synD[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < nEv; ++k) {
                const int i = colRow[colPtr[j] + k];
                if (i < (T - 1) * S * A) continue;
                const int a2 = i % A, psi = (i / A) % S, m = colVal[colPtr[j] + k];
                for (int q = agCqPtr[psi * A + a2]; q < agCqPtr[psi * A + a2 + 1]; ++q) {
                    const int to = agCq[q];
                    endFx.push_back(int32_t(uint32_t(to) | (uint32_t(m & 0xff) << 24)));
                    if (G > 0) {
                        const int cell = to % (G * G), tx = cell % G, ty = cell / G;
                        for (int z = 0; z < nSyn; ++z) {
                            const int sz = G >> (z + 1), org = G - 2 * sz;
                            if (tx >= org && tx < org + sz && ty >= org && ty < org + sz) synD[z] += m;
                        }
                    }
                }
            }
            rec.push_back(uint32_t(endFx.size()));
            for (int w = 0; w < 2; ++w) {
                uint32_t pk = 0;
                for (int z = 0; z < 4; ++z) pk |= uint32_t(synD[4 * w + z] & 0xff) << (8 * z);
                rec.push_back(pk);
            }
            for (int32_t v : endFx) rec.push_back(uint32_t(v));
        }
        while (rec.size() % 4) rec.push_back(0);
        const size_t words = rec.size() - recStart;
        maxRecWords = std::max(maxRecWords, words);
        if (recStart / 4 > 0xfffffff0ull || words * 4 > 0xffffu * 16u) return fail(ABM_ERR_INVALID, "record table too large");
        recIdx[j] = make_uint2(uint32_t(recStart / 4), uint32_t(words * 4));   // {offset in 16-byte units, bytes (multiple of 16)}
    }
    rec.insert(rec.end(), 64, 0);   // speculative reads past a record's last list stay inside the table
    if (getenv("ABM_VERBOSE")) fprintf(stderr, "abm: %d column records, %.1f MB (+ %.1f MB index), longest %zu bytes; energy tables %zu entries\n", nCols,
                                       rec.size() * 4e-6, recIdx.size() * 8e-6, maxRecWords * 4, eTab.size());

    // [7] lgamma(occ+1), [2][2][6] log(pi/pibar), then the factor log-probability of TrajectoryImportance::refresh
    // (TrajectoryImportance.h:96-108) for every (type, neighbour flag, set of present acts), summed in the reference's order
    // (classes: the two agent types of PredPrey, or the table agent's)
    std::vector<double> impTab(7 + 24 + size_t(2 * nCls) * 64, 0.0);
    if (nStates > 0) {
        std::copy(desc->imp_lgamma, desc->imp_lgamma + 7, impTab.begin());
        if (!tableAgent) std::copy(desc->imp_logratio, desc->imp_logratio + 24, impTab.begin() + 7);
        for (int tn = 0; tn < 2 * nCls; ++tn)
            for (int bits = 0; bits < (1 << A); ++bits) {
                const int occ = __builtin_popcount(bits);
                volatile double pv = occ > 1 ? impTab[occ] : 0.0;   // volatile: one rounded add per act, no reassociation
                for (int a2 = 0; a2 < A; ++a2)
                    if ((bits >> a2) & 1) pv = pv + desc->imp_logratio[tn * A + a2];
                impTab[31 + tn * 64 + bits] = pv;
            }
    }
    // device tables: {class, 4 influencers} per agent state for the kernels' table path (grids that are not a power of two and
    // table agents), and for table agents the events feeding each state of the next timestep (byte offsets into a timestep of Xb)
    std::vector<int32_t> agentTab, endInPtr, endIn;
    int gShift = -1;
    for (int k = 0; k < 30; ++k) if (G == (1 << k)) gShift = k;
    if (nStates > 0 && gShift < 0) {
        agentTab.resize(size_t(S) * 5 + 3);
        for (int q = 0; q < S; ++q) {
            agentTab[5 * q] = agCls[q];
            for (int k = 0; k < 4; ++k) agentTab[5 * q + 1 + k] = agInfl[4 * q + k];
        }
    }
    if (tableAgent) {
        std::vector<std::vector<int32_t>> in(S);
        for (int psi = 0; psi < S; ++psi)
            for (int a2 = 0; a2 < A; ++a2)
                for (int q = agCqPtr[psi * A + a2]; q < agCqPtr[psi * A + a2 + 1]; ++q) in[agCq[q]].push_back(psi * kStateBytes + a2);
        endInPtr.push_back(0);
        for (int q = 0; q < S; ++q) {
            endIn.insert(endIn.end(), in[q].begin(), in[q].end());
            endInPtr.push_back(int32_t(endIn.size()));
        }
        endIn.push_back(0);
    }

    std::unique_ptr<abm_problem> owner(new abm_problem());
    abm_problem *p = owner.get();
    p->device = device;
    p->xoff = xoff;
    DevProblem &d = p->d;
    d.nRows = nRows; d.nCols = nCols; d.nEvents = nEvents; d.xStride = xStride;
    int depth = 0;
    while ((1 << depth) < nCols) ++depth;   // bit length of nCols-1
    if (nCols == 1) depth = 0;
    d.treeDepth = depth;
    d.t1Len = (nCols + 31) / 32;
    d.t2Len = (nCols + 1023) / 1024;
    d.topShift = std::max(10, ((depth - 1) / 5) * 5);   // the first descent round's levels
    d.nTop = ((nCols - 1) >> d.topShift) + 1;
    if (d.nTop > 32) return fail(ABM_ERR_INVALID, "internal: too many top tree nodes");
    d.evBytes = evBytes;
    d.G = G; d.S = S; d.T = T; d.nStates = nStates;
    d.blockedTree = blockedTree ? 1 : 0;
    if (d.blockedTree && nCols > (1 << 20)) return fail(ABM_ERR_INVALID, "blocked tree supports up to 2^20 columns");
    d.gShift = gShift;
    d.groupWidth = W;
    d.rowsCap = (maxColNnz + 7) & ~7;
    d.sCap = (maxCoupled + 7) & ~7;
    d.tCapA = (maxTA + 7) & ~7;
    d.tCapB = (maxTB + 7) & ~7;
    d.tCap = d.tCapA + d.tCapB + ((maxTC + 7) & ~7);
    if (d.blockedTree) d.tCap = d.tCapA = d.tCapB = 0;   // the blocked tree's commit keeps no task list
    d.maxAbs = maxAbs;
    d.twoM = 2 * maxAbs;
    d.scrCap = 0;
    d.nodeCap = d.blockedTree ? 0 : 32;
    d.t2Cap = 0;
    // the compile-time shared-memory layout (stepKernel<..., FIX>): blocked tree, 16 lanes per chain, unit entries, sizes within
    // its constants -- true of the paper's PredPrey problems at every size; ABM_FIX_LAYOUT=0 forces the run-time layout
    d.fixLayout = d.blockedTree && W == 16 && maxAbs == 1 && d.rowsCap <= kFixRowsCap && d.sCap <= kFixSCap && !ABM_T2_SMEM && !ABM_SMEM_SCRATCH &&
                  (!kFixWinWhole || maxRecWords <= size_t(kFixWinWords));   // a whole-record window: every record must fit
    if (const char *e = getenv("ABM_FIX_LAYOUT")) if (!atoi(e)) d.fixLayout = 0;
    if (d.fixLayout) { d.rowsCap = kFixRowsCap; d.sCap = kFixSCap; }
    if (d.blockedTree && !d.fixLayout) {
        // what 7 resident CTAs per SM leave to one chain (228 KB per SM, 1 KB reserved per CTA), after the fixed arrays
        const int chainsPerCta = kWarpsPerCta * (32 / W);
        [[maybe_unused]] const int budget = (((228 * 1024) / 7 - 1024) / chainsPerCta) & ~15;
#if ABM_T2_SMEM
        if (groupSmemBytes(d.rowsCap, d.sCap, d.tCap, 0, d.maxAbs, d.nodeCap, d.t2Len, 0) <= budget) d.t2Cap = d.t2Len;
#endif
#if ABM_SMEM_SCRATCH
        const int room = (budget - groupSmemBytes(d.rowsCap, d.sCap, d.tCap, 0, d.maxAbs, d.nodeCap, d.t2Cap, 0)) / 16;
        d.scrCap = std::max(0, std::min(room, d.sCap + W - 1) / W * W);
#endif
    }
    if (const char *e = getenv("ABM_T2_SMEM_OFF")) if (atoi(e)) d.t2Cap = 0;   // measurement switch
    if (maxTA > 1000 || maxTB > 1000 || maxTC > 1000) return fail(ABM_ERR_INVALID, "too many sum-tree nodes touched by one column");
    d.nSynopsis = nSyn;
    d.kappa = desc->kappa;
    std::vector<int32_t> colEvent(desc->nb_col, desc->nb_col + nCols);
    abm_problem_info &inf = p->info;
    inf.n_rows = nRows; inf.n_cols = nCols; inf.nnz = nnz; inf.n_events = nEvents;
    inf.max_col_nnz = maxColNnz; inf.max_coupled = maxCoupled; inf.tree_depth = depth; inf.n_states = nStates;
    inf.chain_state_bytes = chainStateBytes(d);
    if (hostOnly) {   // what the uploads below would occupy
        auto bytes = [](const auto &v) { return int64_t(v.size() * sizeof(v[0])); };
        inf.shared_bytes = bytes(recIdx) + bytes(rec) + bytes(facVal) + bytes(eTab) + bytes(impTab) + bytes(colPtr) + bytes(colRow) + bytes(colVal) +
                           bytes(rowPtr) + bytes(rowCol) + bytes(rowVal) + bytes(rowF) + bytes(rowParam) + bytes(colEvent) + bytes(agentTab) +
                           bytes(endInPtr) + bytes(endIn);
        *hostOnly = inf;
        return ABM_OK;
    }
    cudaError_t e = cudaSuccess;
    auto up = [&](auto **dst, const auto &v) { if (e == cudaSuccess) e = p->arena.upload(dst, v); };
    up(&d.recIdx, recIdx); up(&d.rec, rec); up(&d.facVal, facVal); up(&d.eTab, eTab); up(&d.impTab, impTab);
    if (!agentTab.empty()) up(&d.agentTab, agentTab);
    if (!endInPtr.empty()) { up(&d.endInPtr, endInPtr); up(&d.endIn, endIn); }
    up(&d.colPtr, colPtr); up(&d.colRow, colRow); up(&d.colVal, colVal);
    up(&d.rowPtr, rowPtr); up(&d.rowCol, rowCol); up(&d.rowVal, rowVal);
    up(&d.rowF, rowF); up(&d.rowParam, rowParam); up(&d.colEvent, colEvent);
    if (e != cudaSuccess) return fail(ABM_ERR_CUDA, std::string("problem upload: ") + cudaGetErrorString(e));
    inf.shared_bytes = int64_t(p->arena.bytes);
    *out = owner.release();
    return ABM_OK;
}

int abm_problem_create(const abm_problem_desc *desc, int device, abm_problem **out) {
    if (!out) return fail(ABM_ERR_INVALID, "null argument");
    return problemCreateImpl(desc, device, out, nullptr);
}

int abm_problem_validate(const abm_problem_desc *desc, abm_problem_info *info) {
    if (!info) return fail(ABM_ERR_INVALID, "null argument");
    return problemCreateImpl(desc, 0, nullptr, info);
}

int abm_problem_destroy(abm_problem *p) {
    if (!p) return ABM_OK;
    cudaSetDevice(p->device);
    delete p;
    return ABM_OK;
}

int abm_problem_get_info(const abm_problem *p, abm_problem_info *info) {
    if (!p || !info) return fail(ABM_ERR_INVALID, "null argument");
    *info = p->info;
    return ABM_OK;
}

int abm_problem_n_synopsis(const abm_problem *p) { return p ? p->d.nSynopsis : 0; }

// ------------------------------------------------------------------------------------------------
int abm_chains_create(const abm_problem *p, int32_t nChains, uint64_t seed, uint64_t firstChainId, void *stream,
                      abm_chains **out) {
    if (!p || !out || nChains <= 0) return fail(ABM_ERR_INVALID, "bad argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(p->device));
    abm_chains *c = new abm_chains();
    c->p = p;
    c->n = nChains;
    c->seed = seed;
    c->firstChain = firstChainId;
    c->stream = static_cast<cudaStream_t>(stream);
    const DevProblem &P = p->d;
    DevChains &d = c->d;
    d.nChains = nChains;
    const size_t n = size_t(nChains);
    cudaError_t e = cudaSuccess;
    auto al = [&](auto **dst, size_t count) { if (e == cudaSuccess) e = c->arena.alloc(dst, count); };
    al(&d.Xb, n * P.xStride); al(&d.colRec, n * P.nCols); al(&d.delta, n * P.nCols);
    al(&d.tier1, n * P.t1Len); al(&d.tier2, n * P.t2Len); al(&d.tier3, n * 32);
    al(&d.logImp, n); al(&d.infeas, n); al(&d.iter, n); al(&d.counters, n * 8);
    al(&d.scratch, n * P.sCap);
    al(&c->momExact, size_t(kMomExactHead) + size_t(std::max(P.S, 1))); al(&c->momReal, size_t(kMomReal));
    if (e == cudaSuccess) d.endStateSum = c->momExact + kMomExactHead;
    al(&d.synSum, n * std::max(P.nSynopsis, 1));
    al(&d.synSumSq, n * std::max(P.nSynopsis, 1)); al(&d.nSamples, n);
    al(&d.synNow, n * std::max(P.nSynopsis, 1)); al(&d.launchSamples, n);
    al(&d.endRec, n * std::max(P.S, 1));
    if (e != cudaSuccess) {
        delete c;
        return fail(ABM_ERR_CUDA, std::string("chain allocation: ") + cudaGetErrorString(e));
    }
    *out = c;
    return ABM_OK;
}

int abm_chains_destroy(abm_chains *c) {
    if (!c) return ABM_OK;
    cudaSetDevice(c->p->device);
    cudaStreamSynchronize(c->stream);
    if (c->series) cudaFree(c->series);
    if (c->ev0) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); }
    if (c->evStats) cudaEventDestroy(c->evStats);
    if (c->stage) cudaFreeHost(c->stage);
    if (c->evMom) cudaEventDestroy(c->evMom);
    if (c->momStage) cudaFreeHost(c->momStage);
    delete c;
    return ABM_OK;
}

int abm_chains_synchronize(abm_chains *c) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ABM_OK;
}

// constructor kernels on a device-resident initial trajectory (dInit may be null when nInit == 0)
static int initFromDevice(abm_chains *c, const int8_t *dInit, int32_t nInit) {
    const DevProblem &P = c->p->d;
    DevChains &d = c->d;
    const size_t n = size_t(c->n);
    CUDA_TRY(cudaMemsetAsync(d.Xb, 0, n * P.xStride, c->stream));
    CUDA_TRY(cudaMemsetAsync(d.infeas, 0, n * sizeof(int32_t), c->stream));
    CUDA_TRY(cudaMemsetAsync(d.logImp, 0, n * sizeof(double), c->stream));
    CUDA_TRY(cudaMemsetAsync(d.iter, 0, n * sizeof(unsigned long long), c->stream));
    CUDA_TRY(cudaMemsetAsync(d.counters, 0, n * 8 * sizeof(long long), c->stream));
    CUDA_TRY(cudaMemsetAsync(d.tier1, 0, n * P.t1Len * sizeof(double), c->stream));
    CUDA_TRY(cudaMemsetAsync(d.tier2, 0, n * P.t2Len * sizeof(double), c->stream));
    CUDA_TRY(cudaMemsetAsync(d.tier3, 0, n * 32 * sizeof(double), c->stream));
    { int rc = zeroStatistics(c); if (rc != ABM_OK) return rc; }
    const int threads = 256;
    dim3 gridC(std::min((P.nCols + threads - 1) / threads, 64), c->n), gridR(std::min((P.nRows + threads - 1) / threads, 64), c->n);
    initDeltaKernel<<<gridC, threads, 0, c->stream>>>(P, d, dInit, nInit);
    initRowsKernel<<<gridR, threads, 0, c->stream>>>(P, d);
    initColsKernel<<<gridC, threads, 0, c->stream>>>(P, d);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->initialised = true;
    c->feasible = false;
    return ABM_OK;
}

int abm_chains_init(abm_chains *c, const int8_t *initEvents, int32_t nInit) {
    if (!c || nInit < 0 || (nInit > 0 && !initEvents)) return fail(ABM_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const size_t n = size_t(c->n);
    DeviceArena tmp;   // freed on every return path
    int8_t *dInit = nullptr;
    if (nInit > 0) {
        CUDA_TRY(tmp.alloc(&dInit, n * nInit));
        CUDA_TRY(cudaMemcpyAsync(dInit, initEvents, n * nInit, cudaMemcpyHostToDevice, c->stream));
    }
    return initFromDevice(c, dInit, nInit);   // synchronises the stream before returning
}

int abm_chains_init_from_prior(abm_chains *c, const double *actPmf, double pPredator, double pPrey, uint64_t seed,
                               int8_t *eventsOut, int32_t startKind) {
    if (!c || !actPmf) return fail(ABM_ERR_INVALID, "null argument");
    if (startKind != ABM_START_BERNOULLI && startKind != ABM_START_POISSON) return fail(ABM_ERR_INVALID, "unknown start state kind");
    const DevProblem &P = c->p->d;
    if (P.G <= 0) return fail(ABM_ERR_INVALID, "prior sampling needs the PredPrey geometry");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const size_t n = size_t(c->n), bytes = n * size_t(P.nEvents);
    DeviceArena tmp;   // freed on every return path
    int8_t *dEvents = nullptr;
    double *dPmf = nullptr;
    CUDA_TRY(tmp.alloc(&dEvents, bytes));
    CUDA_TRY(tmp.alloc(&dPmf, size_t(24)));
    CUDA_TRY(cudaMemcpyAsync(dPmf, actPmf, 24 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const size_t smem = size_t(2) * P.S * sizeof(int);
    cudaError_t e = cudaFuncSetAttribute(priorSampleKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e == cudaSuccess) {
        priorSampleKernel<<<c->n, 128, smem, c->stream>>>(P, c->n, seed, c->firstChain, dPmf, pPredator, pPrey, startKind, dEvents);
        e = cudaGetLastError();
    }
    int rc = ABM_OK;
    if (e != cudaSuccess) rc = fail(ABM_ERR_CUDA, std::string("priorSampleKernel: ") + cudaGetErrorString(e));
    if (rc == ABM_OK && eventsOut) {
        e = cudaMemcpyAsync(eventsOut, dEvents, bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) rc = fail(ABM_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc == ABM_OK) rc = initFromDevice(c, dEvents, P.nEvents);
    else cudaStreamSynchronize(c->stream);   // nothing may still use the temporaries when the arena frees them
    return rc;
}

}  // extern "C" (templates below need C++ linkage)

template <int MODE, int W, bool BLK, bool FIX>
static cudaError_t launchStepWBF(const abm_chains *c, const RunParams &r) {
    const DevProblem &P = c->p->d;
    constexpr int NG = 32 / W;
    const int chainsPerCta = kWarpsPerCta * NG;
    const int grid = (c->n + chainsPerCta - 1) / chainsPerCta;
    const size_t smem = size_t(chainsPerCta) * (FIX ? groupSmemBytes(kFixRowsCap, kFixSCap, 0, kFixScrCap, 1, 0, 0, 0, kFixWinBytes)
                                                    : groupSmemBytes(P.rowsCap, P.sCap, P.tCap, P.scrCap, P.maxAbs, P.nodeCap, P.t2Cap, BLK ? 0 : P.sCap));
    auto kernel = stepKernel<MODE, W, BLK, FIX>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    static const int carveout = getenv("ABM_SMEM_CARVEOUT") ? atoi(getenv("ABM_SMEM_CARVEOUT")) : -1;   // measurement switch: percent of
    if (carveout >= 0) {                                                                              // the L1/shared array given to shared memory
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, kWarpsPerCta * 32, smem, c->stream>>>(P, c->d, r);
    return cudaGetLastError();
}

template <int MODE, int W, bool BLK>
static cudaError_t launchStepWB(const abm_chains *c, const RunParams &r) {
    if constexpr (BLK && W == 16) {
        if (c->p->d.fixLayout) return launchStepWBF<MODE, W, BLK, true>(c, r);
    }
    return launchStepWBF<MODE, W, BLK, false>(c, r);
}

constexpr int kTier1RebuildEvery = 256;   // launches between exact rebuilds of the blocked tree's sums of 32 leaves

static void rebuildTier1(const abm_chains *c) {
    const DevProblem &P = c->p->d;
    rebuildTier1Kernel<<<dim3(std::min((P.t1Len + 127) / 128, 64), c->n), 128, 0, c->stream>>>(P, c->d);
}

template <int MODE, int W>
static cudaError_t launchStepW(const abm_chains *c, const RunParams &r) {
    if (!c->p->d.blockedTree) return launchStepWB<MODE, W, false>(c, r);
    // before every launch the two upper levels are recomputed from the level below; every kTier1RebuildEvery-th launch the
    // lowest sums are recomputed from the leaves as well, so no level carries incremental rounding for long
    if (MODE != MODE_INIT && ++const_cast<abm_chains *>(c)->launchesSinceRebuild >= kTier1RebuildEvery) {
        const_cast<abm_chains *>(c)->launchesSinceRebuild = 0;
        rebuildTier1(c);
    }
    refreshBlockedTopKernel<<<(c->n + 3) / 4, 128, 0, c->stream>>>(c->p->d, c->d);
    return launchStepWB<MODE, W, true>(c, r);
}

template <int MODE>
static cudaError_t launchStep(const abm_chains *c, const RunParams &r) {
    switch (c->p->d.groupWidth) {
        case 8: return launchStepW<MODE, 8>(c, r);
        case 32: return launchStepW<MODE, 32>(c, r);
        default: return launchStepW<MODE, 16>(c, r);
    }
}

extern "C" {

constexpr int64_t kMaxStepsPerLaunch = 1ll << 30;   // per-launch step and sample counters are 32-bit registers

static RunParams baseParams(const abm_chains *c) {
    RunParams r{};
    r.seed = c->seed;
    r.firstChain = c->firstChain;
    return r;
}

int abm_chains_find_feasible(abm_chains *c, int64_t maxIterations, const double *uniforms, int64_t nUniforms,
                             int64_t *iterationsOut) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    if (!c->initialised) return fail(ABM_ERR_STATE, "abm_chains_init has not been called");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const DevProblem &P = c->p->d;
    const size_t n = size_t(c->n);
    if (maxIterations <= 0) maxIterations = INT64_MAX;
    DeviceArena tmp;   // freed on every return path (each launch below is followed by a stream synchronisation)
    double *dU = nullptr;
    long long *dIters = nullptr;   // per-call iteration counts: the kernel ADDS to them across the launches of this call only
    CUDA_TRY(tmp.alloc(&dIters, n));
    CUDA_TRY(cudaMemsetAsync(dIters, 0, n * sizeof(long long), c->stream));
    if (uniforms) {
        if (nUniforms <= 0) return fail(ABM_ERR_INVALID, "uniforms without a count");
        CUDA_TRY(tmp.alloc(&dU, n * size_t(nUniforms)));
        CUDA_TRY(cudaMemcpyAsync(dU, uniforms, n * nUniforms * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        maxIterations = std::min<int64_t>(maxIterations, nUniforms);
    }
    std::vector<int32_t> inf(n);
    int64_t done = 0;
    bool allFeasible = false;
    // The search ends when the SLOWEST chain is feasible (32x32x16: 2.4e6 iterations against a mean of 4.9e5), and a chain's
    // search is sequential, so the tail is a handful of chains running at the latency of one iteration.  With few chains
    // left, a whole warp per chain (32 lanes: half the rounds per iteration) has the lower latency (7.9 against 11.3 us per
    // iteration at 256 chains), so the tail of a large ensemble's search runs on the 32-lane kernel; the records do not depend
    // on the lanes per chain.  ABM_INIT_TAIL = 0 / 1 switches it off / on for every size (tests); ABM_GROUP_WIDTH overrides.
    constexpr int64_t kTailChains = 2048;
    const int tailMode = getenv("ABM_INIT_TAIL") ? atoi(getenv("ABM_INIT_TAIL")) : -1;
    const bool widthForced = getenv("ABM_GROUP_WIDTH") != nullptr;
    int64_t unfinished = int64_t(n);
    while (done < maxIterations) {
        const int64_t chunk = std::min<int64_t>(maxIterations - done, 1 << 16);
        RunParams r = baseParams(c);
        r.nSteps = chunk;
        r.injU = dU;
        r.uStride = nUniforms;
        r.uOffset = done;   // valid because with injected uniforms all chains advance in lock step per launch
        r.itersOut = dIters;
        const bool wide = !widthForced && (tailMode == 1 || (tailMode != 0 && int64_t(n) > kTailChains && unfinished <= kTailChains));
        if (wide) CUDA_TRY((launchStepW<MODE_INIT, 32>(c, r)));
        else CUDA_TRY(launchStep<MODE_INIT>(c, r));
        CUDA_TRY(cudaMemcpyAsync(inf.data(), c->d.infeas, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        done += chunk;
        unfinished = std::count_if(inf.begin(), inf.end(), [](int32_t v) { return v != 0; });
        allFeasible = unfinished == 0;
        if (allFeasible) break;
    }
    if (iterationsOut) {
        std::vector<long long> it(n);
        CUDA_TRY(cudaMemcpy(it.data(), dIters, n * sizeof(long long), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < n; ++k) iterationsOut[k] = it[k];
    }
    if (!allFeasible) return fail(ABM_ERR_NOT_FEASIBLE, "feasibility search hit its iteration limit");
    setStateKernel<<<(c->n + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, 0, c->stream>>>(P, c->d);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->feasible = true;
    return ABM_OK;
}

int abm_chains_step(abm_chains *c, int64_t nSteps) {
    if (!c || nSteps < 0) return fail(ABM_ERR_INVALID, "bad argument");
    if (!c->feasible) return fail(ABM_ERR_STATE, "chains have no feasible start (call abm_chains_find_feasible)");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(markLaunch(c, true));
    // the kernel counts a launch's steps in 32-bit registers (the per-chain totals are 64-bit): longer requests are split
    for (int64_t done = 0; done < nSteps || (nSteps == 0 && done == 0);) {
        RunParams r = baseParams(c);
        r.nSteps = std::min<int64_t>(nSteps - done, kMaxStepsPerLaunch);
        CUDA_TRY(launchStep<MODE_MCMC>(c, r));
        done += std::max<int64_t>(r.nSteps, 1);
    }
    CUDA_TRY(markLaunch(c, false));
    return ABM_OK;
}

int abm_chains_last_launch_ms(abm_chains *c, float *ms) {
    if (!c || !ms) return fail(ABM_ERR_INVALID, "null argument");
    if (!c->timed) return fail(ABM_ERR_STATE, "no step has been launched yet");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return ABM_OK;
}

int abm_chains_step_traced(abm_chains *c, int64_t nSteps, const double *uniforms, const int32_t *proposals,
                           int32_t *outJ, uint8_t *outAcc, int32_t *outInf) {
    if (!c || nSteps < 0) return fail(ABM_ERR_INVALID, "bad argument");
    if (nSteps > kMaxStepsPerLaunch) return fail(ABM_ERR_INVALID, "at most 2^30 traced steps per call");
    if (!c->feasible) return fail(ABM_ERR_STATE, "chains have no feasible start (call abm_chains_find_feasible)");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const size_t n = size_t(c->n), ns = size_t(nSteps);
    DeviceArena tmp;
    RunParams r = baseParams(c);
    r.nSteps = nSteps;
    double *dU = nullptr;
    int32_t *dJ = nullptr, *dOutJ = nullptr, *dOutInf = nullptr;
    uint8_t *dOutAcc = nullptr;
    if (uniforms) {
        CUDA_TRY(tmp.alloc(&dU, n * ns * 2));
        CUDA_TRY(cudaMemcpyAsync(dU, uniforms, n * ns * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        r.injU = dU;
        r.uStride = 2 * nSteps;
    }
    if (proposals) {
        CUDA_TRY(tmp.alloc(&dJ, n * ns));
        CUDA_TRY(cudaMemcpyAsync(dJ, proposals, n * ns * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        r.injJ = dJ;
    }
    CUDA_TRY(tmp.alloc(&dOutJ, n * ns));
    CUDA_TRY(tmp.alloc(&dOutAcc, n * ns));
    CUDA_TRY(tmp.alloc(&dOutInf, n * ns));
    r.outJ = dOutJ; r.outAcc = dOutAcc; r.outInf = dOutInf;
    CUDA_TRY(markLaunch(c, true));
    CUDA_TRY(launchStep<MODE_MCMC>(c, r));
    CUDA_TRY(markLaunch(c, false));
    if (outJ) CUDA_TRY(cudaMemcpyAsync(outJ, dOutJ, n * ns * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (outAcc) CUDA_TRY(cudaMemcpyAsync(outAcc, dOutAcc, n * ns, cudaMemcpyDeviceToHost, c->stream));
    if (outInf) CUDA_TRY(cudaMemcpyAsync(outInf, dOutInf, n * ns * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ABM_OK;
}

static int zeroStatistics(abm_chains *c) {
    const DevProblem &P = c->p->d;
    const size_t n = size_t(c->n), ns = size_t(std::max(P.nSynopsis, 1));
    CUDA_TRY(cudaMemsetAsync(c->d.endRec, 0, n * std::max(P.S, 1) * sizeof(EndRec), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d.synSum, 0, n * ns * sizeof(long long), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d.synSumSq, 0, n * ns * sizeof(long long), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d.nSamples, 0, n * sizeof(long long), c->stream));
    if (c->series)
        CUDA_TRY(cudaMemsetAsync(c->series, 0, size_t(c->nSeriesChains) * c->seriesCap * ns * sizeof(int32_t), c->stream));
    return ABM_OK;
}

int abm_chains_sample(abm_chains *c, int64_t nSamples, int64_t maxSteps) {
    if (!c || nSamples < 0 || (nSamples == 0 && maxSteps <= 0)) return fail(ABM_ERR_INVALID, "bad argument");
    if (!c->feasible) return fail(ABM_ERR_STATE, "chains have no feasible start (call abm_chains_find_feasible)");
    if (c->p->d.nStates <= 0) return fail(ABM_ERR_INVALID, "sample statistics need an agent description (PredPrey grid or agent tables)");
    CUDA_TRY(cudaSetDevice(c->p->device));
    // per-launch step and sample counts live in 32-bit registers: one call is one launch of at most 2^30 steps / samples
    if (nSamples > kMaxStepsPerLaunch && (maxSteps <= 0 || maxSteps > kMaxStepsPerLaunch))
        return fail(ABM_ERR_INVALID, "at most 2^30 samples or steps per abm_chains_sample call");
    if (maxSteps > kMaxStepsPerLaunch) return fail(ABM_ERR_INVALID, "at most 2^30 steps per abm_chains_sample call");
    RunParams r = baseParams(c);
    r.nSteps = maxSteps > 0 ? maxSteps : kMaxStepsPerLaunch;
    r.nSamples = nSamples;
    r.series = c->series;
    r.nSeriesChains = c->nSeriesChains;
    r.seriesCap = c->seriesCap;
    r.seriesThin = c->seriesThin;
    const DevProblem &P = c->p->d;
    sampleBeginKernel<<<(c->n + 3) / 4, 128, 0, c->stream>>>(P, c->d);
    CUDA_TRY(cudaGetLastError());
    // without a quota (n_samples == 0 or beyond what max_steps can yield) every chain simply runs max_steps steps
    CUDA_TRY(markLaunch(c, true));
    if (maxSteps > 0 && (nSamples == 0 || nSamples >= maxSteps)) CUDA_TRY(launchStep<MODE_SAMPLE_FIXED>(c, r));
    else CUDA_TRY(launchStep<MODE_SAMPLE>(c, r));
    CUDA_TRY(markLaunch(c, false));
    sampleEndKernel<<<c->n, 256, 0, c->stream>>>(P, c->d);
    CUDA_TRY(cudaGetLastError());
    return ABM_OK;
}

int abm_chains_record_series(abm_chains *c, int32_t nChains, int32_t thin, int32_t capacity) {
    if (!c || nChains < 0 || nChains > c->n || thin < 1 || capacity < 0) return fail(ABM_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->series) { cudaFree(c->series); c->series = nullptr; }
    c->nSeriesChains = nChains; c->seriesCap = capacity; c->seriesThin = thin;
    const size_t count = size_t(nChains) * capacity * std::max(c->p->d.nSynopsis, 1);
    if (count) {
        CUDA_TRY(cudaMalloc(&c->series, count * sizeof(int32_t)));
        CUDA_TRY(cudaMemset(c->series, 0, count * sizeof(int32_t)));
    }
    return ABM_OK;
}

int abm_chains_get_series(abm_chains *c, int32_t *out) {
    if (!c || !out) return fail(ABM_ERR_INVALID, "null argument");
    if (!c->series) return fail(ABM_ERR_STATE, "no series is being recorded");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const size_t count = size_t(c->nSeriesChains) * c->seriesCap * std::max(c->p->d.nSynopsis, 1);
    CUDA_TRY(cudaMemcpy(out, c->series, count * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return ABM_OK;
}

// ------------------------------------------------------------------------------------------------
int abm_chains_get_state(abm_chains *c, int32_t chain, int32_t *X, int8_t *delta, double *currentDE, double *tree,
                         double *logImportance, int32_t *infeasibility) {
    if (!c || chain < 0 || chain >= c->n) return fail(ABM_ERR_INVALID, "bad chain index");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const DevProblem &P = c->p->d;
    const size_t ch = size_t(chain);
    if (X) {
        std::vector<int8_t> xb(P.xStride);
        CUDA_TRY(cudaMemcpy(xb.data(), c->d.Xb + ch * P.xStride, P.xStride, cudaMemcpyDeviceToHost));
        for (int i = 0; i < P.nRows; ++i) X[i] = xb[c->p->xoff[i]];
    }
    if (delta && !P.blockedTree) CUDA_TRY(cudaMemcpy(delta, c->d.delta + ch * P.nCols, P.nCols, cudaMemcpyDeviceToHost));
    if (currentDE || tree || (delta && P.blockedTree)) {
        std::vector<double2> cr(P.nCols);
        std::vector<double> t1(P.t1Len), t2(P.t2Len);
        CUDA_TRY(cudaMemcpy(cr.data(), c->d.colRec + ch * P.nCols, sizeof(double2) * P.nCols, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(t1.data(), c->d.tier1 + ch * P.t1Len, sizeof(double) * P.t1Len, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(t2.data(), c->d.tier2 + ch * P.t2Len, sizeof(double) * P.t2Len, cudaMemcpyDeviceToHost));
        for (int j = 0; j < P.nCols; ++j) {
            if (currentDE) currentDE[j] = cr[j].x;
            if (delta && P.blockedTree) delta[j] = std::signbit(cr[j].y) ? -1 : 1;   // the blocked tree's leaf carries delta_j in its sign
            if (tree && !P.blockedTree) tree[j] = (j & 31) ? cr[j].y : ((j & 1023) ? t1[j >> 5] : t2[j >> 10]);
        }
        if (tree && P.blockedTree) {
            // the blocked tree stores leaves; present them in the reference's node layout (node i = sum of the leaves in
            // [i, i + lowbit(i)), MutableCategoricalArray.h:9-24), each node summed directly from its leaves
            for (int j = 0; j < P.nCols; ++j) {
                const int span = j ? (j & -j) : P.nCols;
                double v = 0.0;
                for (int k = std::min(j + span, P.nCols) - 1; k >= j; --k) v += std::fabs(cr[k].y);
                tree[j] = v;
            }
        }
    }
    if (logImportance) CUDA_TRY(cudaMemcpy(logImportance, c->d.logImp + ch, sizeof(double), cudaMemcpyDeviceToHost));
    if (infeasibility) CUDA_TRY(cudaMemcpy(infeasibility, c->d.infeas + ch, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return ABM_OK;
}

int abm_chains_get_counters(abm_chains *c, abm_mcmc_counters *perChain) {
    if (!c || !perChain) return fail(ABM_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    static_assert(sizeof(abm_mcmc_counters) == 64, "counter layout");
    CUDA_TRY(cudaMemcpy(perChain, c->d.counters, size_t(c->n) * 64, cudaMemcpyDeviceToHost));
    return ABM_OK;
}

int abm_chains_get_infeasibility(abm_chains *c, int32_t *perChain) {
    if (!c || !perChain) return fail(ABM_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(perChain, c->d.infeas, size_t(c->n) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return ABM_OK;
}

int abm_chains_get_trajectories(abm_chains *c, int8_t *events) {
    if (!c || !events) return fail(ABM_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const DevProblem &P = c->p->d;
    std::vector<int8_t> xb(size_t(c->n) * P.xStride);
    CUDA_TRY(cudaMemcpy(xb.data(), c->d.Xb, xb.size(), cudaMemcpyDeviceToHost));
    for (int ch = 0; ch < c->n; ++ch)
        for (int e = 0; e < P.nEvents; ++e)
            events[size_t(ch) * P.nEvents + e] = xb[size_t(ch) * P.xStride + c->p->xoff[e]];
    return ABM_OK;
}

int abm_chains_get_statistics(abm_chains *c, int64_t *endStateSum, int64_t *synSum, int64_t *synSumSq, int64_t *nSamples,
                              int devicePtrs) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const DevProblem &P = c->p->d;
    const size_t n = size_t(c->n), ns = size_t(P.nSynopsis);
    const cudaMemcpyKind kind = devicePtrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (endStateSum && P.S > 0) {
        CUDA_TRY(cudaMemsetAsync(c->d.endStateSum, 0, size_t(P.S) * sizeof(long long), c->stream));
        reduceEndStateKernel<<<dim3((P.S + 127) / 128, (c->n + kReduceSlice - 1) / kReduceSlice), 128, 0, c->stream>>>(P, c->d);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(endStateSum, c->d.endStateSum, size_t(P.S) * sizeof(int64_t), kind, c->stream));
    }
    if (synSum && ns) CUDA_TRY(cudaMemcpyAsync(synSum, c->d.synSum, n * ns * sizeof(int64_t), kind, c->stream));
    if (synSumSq && ns) CUDA_TRY(cudaMemcpyAsync(synSumSq, c->d.synSumSq, n * ns * sizeof(int64_t), kind, c->stream));
    if (nSamples) CUDA_TRY(cudaMemcpyAsync(nSamples, c->d.nSamples, n * sizeof(int64_t), kind, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ABM_OK;
}

// Statistics read-back that does not stall the stream: enqueue = ensemble reduction + device-to-host copies into pinned
// staging owned by the handle, ordered after everything launched so far; wait = block until those copies have landed (later
// launches keep running) and hand the values to the caller's buffers.  Layout of the staging: end_state_sum[S],
// syn_sum[n][d], syn_sumsq[n][d], n_samples[n], counters[n][8].
int abm_chains_enqueue_statistics(abm_chains *c) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const DevProblem &P = c->p->d;
    const size_t n = size_t(c->n), ns = size_t(P.nSynopsis), S = size_t(std::max(P.S, 0));
    const size_t words = S + 2 * n * ns + n + 8 * n;
    if (!c->stage || c->stageWords != words) {
        if (c->stage) CUDA_TRY(cudaFreeHost(c->stage));
        c->stage = nullptr;
        CUDA_TRY(cudaMallocHost(&c->stage, words * sizeof(int64_t)));
        c->stageWords = words;
    }
    if (!c->evStats) CUDA_TRY(cudaEventCreateWithFlags(&c->evStats, cudaEventDisableTiming));
    int64_t *o = c->stage;
    if (S > 0) {
        CUDA_TRY(cudaMemsetAsync(c->d.endStateSum, 0, S * sizeof(long long), c->stream));
        reduceEndStateKernel<<<dim3((P.S + 127) / 128, (c->n + kReduceSlice - 1) / kReduceSlice), 128, 0, c->stream>>>(P, c->d);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(o, c->d.endStateSum, S * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    }
    o += S;
    if (ns) {
        CUDA_TRY(cudaMemcpyAsync(o, c->d.synSum, n * ns * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(o + n * ns, c->d.synSumSq, n * ns * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    }
    o += 2 * n * ns;
    CUDA_TRY(cudaMemcpyAsync(o, c->d.nSamples, n * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(o + n, c->d.counters, n * 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaEventRecord(c->evStats, c->stream));
    c->statsPending = true;
    return ABM_OK;
}

int abm_chains_wait_statistics(abm_chains *c, int64_t *endStateSum, int64_t *synSum, int64_t *synSumSq, int64_t *nSamples,
                               abm_mcmc_counters *counters) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    if (!c->statsPending) return fail(ABM_ERR_STATE, "abm_chains_enqueue_statistics has not been called");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaEventSynchronize(c->evStats));
    c->statsPending = false;
    const DevProblem &P = c->p->d;
    const size_t n = size_t(c->n), ns = size_t(P.nSynopsis), S = size_t(std::max(P.S, 0));
    const int64_t *o = c->stage;
    if (endStateSum) memcpy(endStateSum, o, S * sizeof(int64_t));
    o += S;
    if (synSum) memcpy(synSum, o, n * ns * sizeof(int64_t));
    if (synSumSq) memcpy(synSumSq, o + n * ns, n * ns * sizeof(int64_t));
    o += 2 * n * ns;
    if (nSamples) memcpy(nSamples, o, n * sizeof(int64_t));
    if (counters) memcpy(counters, o + n, n * 8 * sizeof(int64_t));
    return ABM_OK;
}

int abm_chains_reset_statistics(abm_chains *c) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    CUDA_TRY(cudaSetDevice(c->p->device));
    int rc = zeroStatistics(c);
    if (rc != ABM_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ABM_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// NCCL, loaded on first use: the library must load (and single-GPU use must work) on a host without NCCL, and no NCCL type
// appears in the C ABI.  Only the handful of entry points the ensemble reduction needs.
struct Id128 { char bytes[ABM_COMM_ID_BYTES]; };   // ncclUniqueId (nccl.h: struct { char internal[128]; })
namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128 /* ncclUniqueId, by value */, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};
constexpr int kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

NcclApi *ncclApi() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("ABM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = "NCCL library not found (libnccl.so.2; set ABM_NCCL_LIB)"; return; }
        auto sym = [&](const char *n) { void *f = dlsym(api.lib, n); if (!f) api.error = std::string("missing NCCL symbol ") + n; return f; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return &api;
}
#define NCCL_TRY(api, expr)                                                                                  \
    do {                                                                                                     \
        int r_ = (expr);                                                                                     \
        if (r_ != 0) return fail(ABM_ERR_CUDA, std::string(#expr) + ": " + (api)->GetErrorString(r_));       \
    } while (0)
}  // namespace

struct abm_comm {
    void *nccl = nullptr;   // ncclComm_t
    bool owned = false;
    int device = -1;
};

extern "C" {

int abm_comm_unique_id(uint8_t id[ABM_COMM_ID_BYTES]) {
    if (!id) return fail(ABM_ERR_INVALID, "null argument");
    NcclApi *a = ncclApi();
    if (!a->error.empty()) return fail(ABM_ERR_CUDA, a->error);
    NCCL_TRY(a, a->GetUniqueId(id));
    return ABM_OK;
}

int abm_comm_create(int32_t nRanks, int32_t rank, const uint8_t id[ABM_COMM_ID_BYTES], int device, abm_comm **out) {
    if (!out || !id || nRanks <= 0 || rank < 0 || rank >= nRanks) return fail(ABM_ERR_INVALID, "bad argument");
    *out = nullptr;
    NcclApi *a = ncclApi();
    if (!a->error.empty()) return fail(ABM_ERR_CUDA, a->error);
    if (abm_device_count() <= device || device < 0) return fail(ABM_ERR_CUDA, "no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    Id128 uid;
    memcpy(uid.bytes, id, ABM_COMM_ID_BYTES);
    void *comm = nullptr;
    NCCL_TRY(a, a->CommInitRank(&comm, nRanks, uid, rank));
    abm_comm *c = new abm_comm();
    c->nccl = comm;
    c->owned = true;
    c->device = device;
    *out = c;
    return ABM_OK;
}

int abm_comm_from_nccl(void *ncclComm, abm_comm **out) {
    if (!out || !ncclComm) return fail(ABM_ERR_INVALID, "null argument");
    NcclApi *a = ncclApi();
    if (!a->error.empty()) return fail(ABM_ERR_CUDA, a->error);
    abm_comm *c = new abm_comm();
    c->nccl = ncclComm;
    *out = c;
    return ABM_OK;
}

int abm_comm_destroy(abm_comm *comm) {
    if (!comm) return ABM_OK;
    if (comm->owned && comm->nccl) {
        if (comm->device >= 0) cudaSetDevice(comm->device);
        ncclApi()->CommDestroy(comm->nccl);
    }
    delete comm;
    return ABM_OK;
}

// device reduction of this GPU's chains (+ NCCL sum over the ranks of comm), then the copies into pinned staging
int abm_chains_enqueue_moments(abm_chains *c, abm_comm *comm) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const DevProblem &P = c->p->d;
    const size_t S = size_t(std::max(P.S, 0)), nExact = kMomExactHead + S;
    if (!c->momStage) CUDA_TRY(cudaMallocHost(&c->momStage, nExact * sizeof(int64_t) + kMomReal * sizeof(double)));
    if (!c->evMom) CUDA_TRY(cudaEventCreateWithFlags(&c->evMom, cudaEventDisableTiming));
    if (S > 0) {
        CUDA_TRY(cudaMemsetAsync(c->d.endStateSum, 0, S * sizeof(long long), c->stream));
        reduceEndStateKernel<<<dim3((P.S + 127) / 128, (c->n + kReduceSlice - 1) / kReduceSlice), 128, 0, c->stream>>>(P, c->d);
        CUDA_TRY(cudaGetLastError());
    }
    momentsKernel<<<1, 256, 0, c->stream>>>(P, c->d, c->momExact, c->momReal);
    CUDA_TRY(cudaGetLastError());
    if (comm && comm->nccl) {
        NcclApi *a = ncclApi();
        NCCL_TRY(a, a->GroupStart());
        NCCL_TRY(a, a->AllReduce(c->momExact, c->momExact, nExact, kNcclInt64, kNcclSum, comm->nccl, c->stream));
        NCCL_TRY(a, a->AllReduce(c->momReal, c->momReal, kMomReal, kNcclFloat64, kNcclSum, comm->nccl, c->stream));
        NCCL_TRY(a, a->GroupEnd());
    }
    CUDA_TRY(cudaMemcpyAsync(c->momStage, c->momExact, nExact * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->momStage + nExact, c->momReal, kMomReal * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaEventRecord(c->evMom, c->stream));
    c->momPending = true;
    return ABM_OK;
}

int abm_chains_wait_moments(abm_chains *c, int64_t *exact, double *real) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    if (!c->momPending) return fail(ABM_ERR_STATE, "abm_chains_enqueue_moments has not been called");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaEventSynchronize(c->evMom));
    c->momPending = false;
    const size_t nExact = kMomExactHead + size_t(std::max(c->p->d.S, 0));
    if (exact) memcpy(exact, c->momStage, nExact * sizeof(int64_t));
    if (real) memcpy(real, c->momStage + nExact, kMomReal * sizeof(double));
    return ABM_OK;
}

int abm_chains_allreduce_moments(abm_chains *c, abm_comm *comm, int64_t *exact, double *real) {
    int rc = abm_chains_enqueue_moments(c, comm);
    if (rc != ABM_OK) return rc;
    return abm_chains_wait_moments(c, exact, real);
}

int abm_chains_get_moments(abm_chains *c, int64_t *exact, double *real) { return abm_chains_allreduce_moments(c, nullptr, exact, real); }

// Blocked tree maintenance and its measurement (DESIGN.md "Two storages").
int abm_chains_rebuild_tree(abm_chains *c) {
    if (!c) return fail(ABM_ERR_INVALID, "null chains");
    if (!c->p->d.blockedTree) return ABM_OK;   // the reference tree is the reference's, drift included
    CUDA_TRY(cudaSetDevice(c->p->device));
    rebuildTier1(c);
    refreshBlockedTopKernel<<<(c->n + 3) / 4, 128, 0, c->stream>>>(c->p->d, c->d);
    CUDA_TRY(cudaGetLastError());
    c->launchesSinceRebuild = 0;
    return ABM_OK;
}

int abm_chains_tree_drift(abm_chains *c, int32_t chain, double out[4]) {
    if (!c || !out || chain < 0 || chain >= c->n) return fail(ABM_ERR_INVALID, "bad argument");
    const DevProblem &P = c->p->d;
    if (!P.blockedTree) return fail(ABM_ERR_INVALID, "the blocked sum tree is not in use");
    CUDA_TRY(cudaSetDevice(c->p->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const size_t ch = size_t(chain);
    std::vector<double2> cr(P.nCols);
    std::vector<double> t1(P.t1Len), t2(P.t2Len), t3(32);
    CUDA_TRY(cudaMemcpy(cr.data(), c->d.colRec + ch * P.nCols, sizeof(double2) * P.nCols, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(t1.data(), c->d.tier1 + ch * P.t1Len, sizeof(double) * P.t1Len, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(t2.data(), c->d.tier2 + ch * P.t2Len, sizeof(double) * P.t2Len, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(t3.data(), c->d.tier3 + ch * 32, sizeof(double) * 32, cudaMemcpyDeviceToHost));
    auto exact = [&](int lo, int hi) {   // leaves [lo, hi) summed in long double
        long double v = 0.0L;
        for (int k = lo; k < std::min(hi, P.nCols); ++k) v += std::fabs(cr[k].y);
        return v;
    };
    const long double root = exact(0, P.nCols);
    double e1 = 0.0, e2 = 0.0, e3 = 0.0;
    for (int b = 0; b < P.t1Len; ++b) e1 = std::max(e1, double(fabsl(t1[b] - exact(32 * b, 32 * b + 32))));
    for (int b = 0; b < P.t2Len; ++b) e2 = std::max(e2, double(fabsl(t2[b] - exact(1024 * b, 1024 * b + 1024))));
    for (int b = 0; b * 32768 < P.nCols; ++b) e3 = std::max(e3, double(fabsl(t3[b] - exact(32768 * b, 32768 * b + 32768))));
    out[0] = double(root); out[1] = e1; out[2] = e2; out[3] = e3;
    return ABM_OK;
}

// Variogram sums of the recorded series over all chains of all ranks (exact integers), see variogramKernel
int abm_chains_series_variogram(abm_chains *c, abm_comm *comm, int32_t nPoints, int32_t nLags, int64_t minSamples, int64_t *sums,
                                int64_t *nChainsOut) {
    if (!c || !sums || nPoints <= 0 || nLags <= 0 || nLags > nPoints) return fail(ABM_ERR_INVALID, "bad argument");
    if (!c->series || nPoints > c->seriesCap) return fail(ABM_ERR_STATE, "no series of that length is being recorded");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const int nSyn = c->p->d.nSynopsis;
    DeviceArena tmp;
    long long *dOut = nullptr;
    const size_t words = size_t(nLags) * 8 + 1;
    CUDA_TRY(tmp.alloc(&dOut, words));
    CUDA_TRY(cudaMemsetAsync(dOut, 0, words * sizeof(long long), c->stream));
    variogramKernel<<<dim3(nLags, (c->nSeriesChains + kVarioSlice - 1) / kVarioSlice), 128, 0, c->stream>>>(
        c->d, c->series, c->seriesCap, c->nSeriesChains, nSyn, nPoints, minSamples, dOut, dOut + size_t(nLags) * 8);
    CUDA_TRY(cudaGetLastError());
    if (comm && comm->nccl) {
        NcclApi *a = ncclApi();
        NCCL_TRY(a, a->AllReduce(dOut, dOut, words, kNcclInt64, kNcclSum, comm->nccl, c->stream));
    }
    std::vector<long long> host(words);
    CUDA_TRY(cudaMemcpyAsync(host.data(), dOut, words * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int j = 0; j < nLags; ++j)
        for (int k = 0; k < nSyn; ++k) sums[size_t(j) * nSyn + k] = host[size_t(j) * 8 + k];
    if (nChainsOut) *nChainsOut = host[size_t(nLags) * 8];
    return ABM_OK;
}

// abm_chains_sample with the uniforms injected: exactly nSteps steps, statistics after every step that ends feasible
int abm_chains_sample_injected(abm_chains *c, int64_t nSteps, const double *uniforms) {
    if (!c || nSteps <= 0 || !uniforms) return fail(ABM_ERR_INVALID, "bad argument");
    if (nSteps > INT32_MAX) return fail(ABM_ERR_INVALID, "at most 2^31-1 steps per launch");
    if (!c->feasible) return fail(ABM_ERR_STATE, "chains have no feasible start (call abm_chains_find_feasible)");
    if (c->p->d.nStates <= 0) return fail(ABM_ERR_INVALID, "sample statistics need an agent description (PredPrey grid or agent tables)");
    CUDA_TRY(cudaSetDevice(c->p->device));
    const size_t n = size_t(c->n), ns = size_t(nSteps);
    DeviceArena tmp;
    double *dU = nullptr;
    CUDA_TRY(tmp.alloc(&dU, n * ns * 2));
    CUDA_TRY(cudaMemcpyAsync(dU, uniforms, n * ns * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RunParams r = baseParams(c);
    r.nSteps = nSteps;
    r.injU = dU;
    r.uStride = 2 * nSteps;
    const DevProblem &P = c->p->d;
    sampleBeginKernel<<<(c->n + 3) / 4, 128, 0, c->stream>>>(P, c->d);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(launchStep<MODE_SAMPLE_FIXED>(c, r));
    sampleEndKernel<<<c->n, 256, 0, c->stream>>>(P, c->d);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ABM_OK;
}

}  // extern "C"
