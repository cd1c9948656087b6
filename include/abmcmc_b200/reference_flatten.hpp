// Host-side flattening boundary: reference C++ objects  ->  ABMP flat arrays.
//
// This header is compiled TOGETHER WITH the reference's own headers (it names their types but
// defines none of them).  It does what SparseBasisSampler's constructor does when it copies the
// tableau (SparseBasisSampler.h:121-158) but writes plain arrays instead of std::vector<SparseVec>,
// and it tabulates every std::function factor by *calling it* on the host, so the doubles the GPU
// sees are the reference's own bits.
//
// Required before including: TableauNormMinimiser.h, TrajectoryImportance.h, agents/PredPreyAgent.h (and the headers of any
// other agent type handed to flattenAgentImportance)
// (include order as in the reference's main.cpp: Trajectory.h, Forecast.h, Likelihood.h,
// PredPreyProblem.h, SparseBasisSampler.h).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

#include "abmp_io.hpp"

namespace abmcmc {

// Reduced tableau + tabulated factors.  `tableau` is the reference's TableauNormMinimiser<T>
// (TableauNormMinimiser.h:108-125), `factors` one closure per tableau row (constraint order).
template <class TABLEAU, class T>
void flattenTableau(AbmpFile &out, const TABLEAU &tableau, const std::vector<std::function<double(T)>> &factors,
                    double kappa) {
    const int nTabRows = int(tableau.rows.size());
    const int nTabCols = int(tableau.cols.size());
    std::vector<int32_t> L(nTabRows), U(nTabRows), F(nTabRows);
    for (int i = 0; i < nTabRows; ++i) {
        L[i] = int32_t(tableau.L[i]);
        U[i] = int32_t(tableau.U[i]);
        F[i] = int32_t(tableau.F[i]);
    }
    std::vector<uint8_t> isBasic(nTabCols);
    std::vector<int32_t> nbCol, nbPtr{0}, nbRow, nbVal;
    // reachable range of every tableau row: F_i + sum of negative / positive entries (X_N in {0,1})
    std::vector<long> xmin(F.begin(), F.end()), xmax(F.begin(), F.end());
    for (int j = 0; j < nTabCols; ++j) {
        isBasic[j] = tableau.cols[j].isBasic ? 1 : 0;
        if (tableau.cols[j].isBasic) continue;
        nbCol.push_back(j);
        for (int i : tableau.cols[j]) {  // std::set: ascending tableau row (TableauNormMinimiser.h:73)
            T v = tableau.rows[i].at(j);
            if (double(v) != std::floor(double(v))) throw std::runtime_error("flatten: non-integer tableau entry");
            nbRow.push_back(i);
            nbVal.push_back(int32_t(v));
            if (v < 0) xmin[i] += long(v); else xmax[i] += long(v);
        }
        nbPtr.push_back(int32_t(nbRow.size()));
    }
    out.i32["dims"] = {nTabRows, nTabCols, int32_t(tableau.nAuxiliaryVars), int32_t(nbCol.size()), int32_t(nbRow.size())};
    out.i32["tab_L"] = L;
    out.i32["tab_U"] = U;
    out.i32["tab_F"] = F;
    out.u8["col_basic"] = isBasic;
    out.i32["nb_col"] = nbCol;
    out.i32["nb_ptr"] = nbPtr;
    out.i32["nb_row"] = nbRow;
    out.i32["nb_val"] = nbVal;
    out.f64["kappa"] = {kappa};

    // Factor tables, de-duplicated by (lo, values) bit pattern.  Row i's table covers
    // [lo_i, hi_i] = the reachable range clamped into [L_i, U_i]; energy() clamps x into [L_i,U_i]
    // before it evaluates the factor (SparseBasisSampler.h:341-345), so every lookup lands inside.
    std::vector<int32_t> facId, utLo, utOff{0};
    std::vector<double> utVal;
    std::map<std::vector<uint64_t>, int32_t> seen;
    for (int i = 0; i < nTabRows; ++i) {
        if (tableau.isEqualityConstraint(i)) continue;  // dropped by the sampler (SparseBasisSampler.h:129-131)
        long lo = std::min<long>(std::max<long>(xmin[i], L[i]), U[i]);
        long hi = std::max<long>(std::min<long>(xmax[i], U[i]), L[i]);
        if (hi < lo) hi = lo;
        if (hi - lo > 4096) throw std::runtime_error("flatten: factor domain too wide to tabulate");
        std::vector<uint64_t> key;
        key.push_back(uint64_t(lo));
        std::vector<double> vals;
        for (long x = lo; x <= hi; ++x) {
            double v = factors[i](T(x));
            uint64_t bits;
            std::memcpy(&bits, &v, 8);
            key.push_back(bits);
            vals.push_back(v);
        }
        auto it = seen.find(key);
        if (it == seen.end()) {
            it = seen.emplace(key, int32_t(utLo.size())).first;
            utLo.push_back(int32_t(lo));
            utVal.insert(utVal.end(), vals.begin(), vals.end());
            utOff.push_back(int32_t(utVal.size()));
        }
        facId.push_back(it->second);
    }
    out.i32["fac_id"] = facId;
    out.i32["ut_lo"] = utLo;
    out.i32["ut_off"] = utOff;
    out.f64["ut_val"] = utVal;
}

// PredPrey importance tables: every double TrajectoryImportance::refresh can add for a PredPrey
// agent (TrajectoryImportance.h:98-108), obtained by calling the reference's timestep() /
// marginalTimestep() (PredPreyAgent.h:124-182) for each (type, "has opposite-species neighbour").
template <class AGENT>
void flattenPredPreyImportance(AbmpFile &out, int gridSize, int nTimesteps) {
    std::vector<double> lg(AGENT::actDomainSize() + 1, 0.0);
    for (int occ = 0; occ <= AGENT::actDomainSize(); ++occ) lg[occ] = (occ > 1) ? lgamma(occ + 1.0) : 0.0;
    std::vector<double> ratio(2 * 2 * AGENT::actDomainSize(), 0.0), actPmf(2 * 2 * AGENT::actDomainSize(), 0.0);
    for (int type = 0; type < 2; ++type) {
        AGENT agent(0, 0, typename AGENT::Type(type));
        for (int flag = 0; flag < 2; ++flag) {
            ModelState<AGENT> others;
            if (flag) others[AGENT(agent.xRight(), agent.yPosition(), typename AGENT::Type(1 - type))] = 1;
            std::vector<double> pmf = agent.timestep(others);
            for (int act = 0; act < AGENT::actDomainSize(); ++act) {
                double r = 0.0;  // "+= 0.0" is what skipping the term amounts to (TrajectoryImportance.h:105)
                if (pmf[act] > 0.0) r = log(pmf[act] / agent.marginalTimestep(act));
                ratio[(type * 2 + flag) * AGENT::actDomainSize() + act] = r;
                actPmf[(type * 2 + flag) * AGENT::actDomainSize() + act] = pmf[act];   // for prior sampling (Forecast.h:102-150)
            }
        }
    }
    out.i32["pp_meta"] = {gridSize, nTimesteps, nTimesteps * AGENT::domainSize() * AGENT::actDomainSize(),
                          AGENT::domainSize(), AGENT::actDomainSize()};
    out.f64["imp_lgamma"] = lg;
    out.f64["imp_logratio"] = ratio;
    out.f64["pp_actpmf"] = actPmf;
}

// Any other agent type (CatMouseAgent, BinomialAgent<N>, ...): TrajectoryImportance<AGENT> as tables, every entry obtained by
// calling the agent's own functions.  The form the kernels evaluate is the one all agents of the reference have: timestep()
// of an agent state reads the other states only through "is any of up to four given states occupied" (PredPreyAgent.h:124-157,
// CatMouseAgent.cpp:8-24, BinomialAgent.h:33).  Those states -- the influencers -- are found by probing timestep() with one
// other state occupied at a time; that the pmf then depends on nothing else is verified over every combination of occupied
// influencers and occupation numbers, and anything else is refused here rather than mis-sampled on the device.
//   ag_meta {S, A, T, nClasses}; ag_class [S]; ag_infl [S][4] (-1 unused); ag_nb_ptr/ag_nb: neighbours() in its own order
//   (the factors TrajectoryImportance::perturb marks stale, :48-51); ag_cq_ptr/ag_cq over (state, act): consequences(act);
//   imp_lgamma [7]; imp_logratio [nClasses][2][A] = log(pmf/marginal) by class, "any influencer occupied", act (:105).
template <class AGENT>
void flattenAgentImportance(AbmpFile &out, int nTimesteps) {
    const int S = AGENT::domainSize(), A = AGENT::actDomainSize();
    if (A < 1 || A > 6) throw std::runtime_error("flatten: agents with 1..6 acts are supported");
    std::vector<double> lg(7, 0.0);
    for (int occ = 0; occ <= 6; ++occ) lg[occ] = (occ > 1) ? lgamma(occ + 1.0) : 0.0;   // TrajectoryImportance.h:96
    auto same = [](const std::vector<double> &a, const std::vector<double> &b) {
        return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0;
    };
    std::vector<int32_t> cls(S), infl(size_t(S) * 4, -1), nbPtr{0}, nb, cqPtr{0}, cq;
    std::vector<double> ratio;   // [class][flag][act]
    std::map<std::vector<uint64_t>, int32_t> classes;
    std::vector<std::vector<int>> inflOf(S);
    for (int psi = 0; psi < S; ++psi) {
        AGENT agent(psi);
        const std::vector<double> base = agent.timestep(ModelState<AGENT>());
        if (int(base.size()) != A) throw std::runtime_error("flatten: timestep() must return one probability per act");
        for (int q = 0; q < S; ++q) {
            ModelState<AGENT> others;
            others[q] = 1;
            if (!same(agent.timestep(others), base)) inflOf[psi].push_back(q);
        }
        const std::vector<int> &in = inflOf[psi];
        if (in.size() > 4) throw std::runtime_error("flatten: an agent state whose timestep() reads more than four other states");
        std::vector<double> flagged = base;
        if (!in.empty()) {
            ModelState<AGENT> others;
            others[in[0]] = 1;
            flagged = agent.timestep(others);
        }
        for (int occ = 1; occ <= A; ++occ) {
            for (unsigned mask = 1; mask < (1u << in.size()); ++mask) {
                ModelState<AGENT> others;
                for (size_t k = 0; k < in.size(); ++k)
                    if ((mask >> k) & 1u) others[in[k]] = occ;
                if (!same(agent.timestep(others), flagged))
                    throw std::runtime_error("flatten: timestep() depends on more than whether any influencer state is occupied");
            }
            ModelState<AGENT> rest, all;   // everything that is NOT an influencer occupied: no effect, alone or together
            for (int q = 0; q < S; ++q) {
                const bool isIn = std::find(in.begin(), in.end(), q) != in.end();
                if (!isIn) rest[q] = occ;
                all[q] = occ;
            }
            if (!same(agent.timestep(rest), base) || (!in.empty() && !same(agent.timestep(all), flagged)))
                throw std::runtime_error("flatten: timestep() depends on more than whether any influencer state is occupied");
        }
        std::vector<double> rows(size_t(2) * A, 0.0);
        std::vector<uint64_t> key;
        for (int flag = 0; flag < 2; ++flag)
            for (int act = 0; act < A; ++act) {
                const double pa = (flag ? flagged : base)[act];
                double r = 0.0;   // a skipped term adds nothing (TrajectoryImportance.h:104)
                if (pa > 0.0) r = log(pa / agent.marginalTimestep(act));
                rows[flag * A + act] = r;
                uint64_t bits;
                std::memcpy(&bits, &r, 8);
                key.push_back(bits);
            }
        auto it = classes.find(key);
        if (it == classes.end()) {
            it = classes.emplace(key, int32_t(classes.size())).first;
            ratio.insert(ratio.end(), rows.begin(), rows.end());
        }
        cls[psi] = it->second;
        for (size_t k = 0; k < in.size(); ++k) infl[4 * psi + k] = in[k];
        for (const AGENT &n : agent.neighbours()) nb.push_back(int(n));
        nbPtr.push_back(int32_t(nb.size()));
        for (int act = 0; act < A; ++act) {
            for (const AGENT &c : agent.consequences(act)) cq.push_back(int(c));
            cqPtr.push_back(int32_t(cq.size()));
        }
    }
    // a state that influences another one's pmf must list it among its neighbours(), or the reference's own factor cache would go
    // stale unnoticed (TrajectoryImportance.h:48-51) and the two implementations would no longer mean the same thing
    for (int psi = 0; psi < S; ++psi)
        for (int q : inflOf[psi])
            if (std::find(nb.begin() + nbPtr[q], nb.begin() + nbPtr[q + 1], psi) == nb.begin() + nbPtr[q + 1])
                throw std::runtime_error("flatten: neighbours() does not cover the states whose timestep() an agent state influences");
    out.i32["ag_meta"] = {S, A, nTimesteps, int32_t(classes.size())};
    out.i32["ag_class"] = cls;
    out.i32["ag_infl"] = infl;
    out.i32["ag_nb_ptr"] = nbPtr;
    out.i32["ag_nb"] = nb;
    out.i32["ag_cq_ptr"] = cqPtr;
    out.i32["ag_cq"] = cq;
    out.f64["imp_lgamma"] = lg;
    out.f64["imp_logratio"] = ratio;
}

}  // namespace abmcmc
