// GpuSparseBasisSampler: header-only C++ drop-in for the reference's SparseBasisSampler<T>
// (src/SparseBasisSampler.h:31-101) on top of the C-ABI in abmcmc_b200.h.
//
// Compile this header TOGETHER WITH the reference's headers (include them first, in the order of its main.cpp:
// Trajectory.h, Forecast.h, Likelihood.h, PredPreyProblem.h, SparseBasisSampler.h) and link libabmcmc_b200.so.
// The constructor keeps the reference signature: the CPU code still builds the basis (TableauNormMinimiser) and the
// factor closures; this class flattens them (reference_flatten.hpp), uploads them and runs the chain on the GPU.
//
// Agent types: PredPreyAgent<G> (G in ABMCMC_PREDPREY_GRIDS) by its grid geometry; CatMouseAgent and BinomialAgent<2..8> by
// tables when agents/CatMouseAgent.h / agents/BinomialAgent.h were included BEFORE this header; any further agent type listed in
// ABMCMC_TABLE_AGENTS (comma-separated) -- see flattenAgentImportance in reference_flatten.hpp for the form an agent must have.
//
// Two ways to drive it:
//   * operator()()  -- exact drop-in: one feasible sample per call, consuming the reference's own generator
//     Random::gen exactly like SparseBasisSampler does (one generate_canonical double per feasibility iteration, two
//     per Metropolis step), so a program that swaps the class keeps its results sample for sample.
//   * batched(nChains, seed) / step / sample / statistics -- thousands of chains on device-side Philox streams.
#pragma once
#include <cstdint>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "abmcmc_b200.h"
#include "abmcmc_b200/abmp_io.hpp"
#include "abmcmc_b200/reference_flatten.hpp"

namespace abmcmc {

namespace detail {
inline void check(int rc, const char *what) {
    if (rc != ABM_OK) throw std::runtime_error(std::string(what) + ": " + abm_last_error());
}

// Finds G such that the importance object is TrajectoryImportance<PredPreyAgent<G>> and fills the PredPrey tables.
template <int... GS>
struct PredPreyGrids {};

template <class T>
bool detectPredPrey(PerturbableFunction<T, double> *, AbmpFile &, PredPreyGrids<>) { return false; }

template <class T, int G, int... REST>
bool detectPredPrey(PerturbableFunction<T, double> *f, AbmpFile &out, PredPreyGrids<G, REST...>) {
    if (auto *imp = dynamic_cast<TrajectoryImportance<PredPreyAgent<G>> *>(f)) {
        flattenPredPreyImportance<PredPreyAgent<G>>(out, G, imp->nTimesteps());
        return true;
    }
    return detectPredPrey<T>(f, out, PredPreyGrids<REST...>{});
}

// Any other agent type, by tables (flattenAgentImportance): the candidates are the agent classes of the reference whose
// headers the program has included, plus whatever ABMCMC_TABLE_AGENTS names (a comma-separated list of agent types).
template <class... AGENTS>
struct TableAgents {};

template <class T>
bool detectTableAgent(PerturbableFunction<T, double> *, AbmpFile &, TableAgents<>) { return false; }

template <class T, class AGENT, class... REST>
bool detectTableAgent(PerturbableFunction<T, double> *f, AbmpFile &out, TableAgents<AGENT, REST...>) {
    if (auto *imp = dynamic_cast<TrajectoryImportance<AGENT> *>(f)) {
        flattenAgentImportance<AGENT>(out, imp->nTimesteps());
        return true;
    }
    return detectTableAgent<T>(f, out, TableAgents<REST...>{});
}

template <class T>
bool detectKnownTableAgents(PerturbableFunction<T, double> *f, AbmpFile &out) {
    bool found = false;
#ifdef GLPKTEST_CATMOUSEAGENT_H   // agents/CatMouseAgent.h
    found = found || detectTableAgent<T>(f, out, TableAgents<CatMouseAgent>{});
#endif
#ifdef GLPKTEST_BINOMIALAGENT_H   // agents/BinomialAgent.h
    found = found || detectTableAgent<T>(f, out, TableAgents<BinomialAgent<2>, BinomialAgent<3>, BinomialAgent<4>, BinomialAgent<5>,
                                                             BinomialAgent<6>, BinomialAgent<7>, BinomialAgent<8>>{});
#endif
#ifdef ABMCMC_TABLE_AGENTS
    found = found || detectTableAgent<T>(f, out, TableAgents<ABMCMC_TABLE_AGENTS>{});
#endif
    return found;
}

// libstdc++ generate_canonical<double,53> on a 32-bit engine, as Random::nextDouble() /
// MutableCategoricalArray::uniformDist produce it (Random.h:21-23, MutableCategoricalArray.cpp:8)
template <class GEN>
double canonical(GEN &gen) { return std::generate_canonical<double, 53>(gen); }
}  // namespace detail

#ifndef ABMCMC_PREDPREY_GRIDS
#define ABMCMC_PREDPREY_GRIDS 3, 4, 8, 10, 16, 32, 64
#endif

template <class T>
class GpuSparseBasisSampler {
public:
    std::vector<T> X;          // SparseBasisSampler::X (:44), refreshed by operator()
    MCMCStatistics stats;      // SparseBasisSampler::stats (:54)
    double kappa;
    T currentInfeasibility = 0;   // (:49)

    // SparseBasisSampler.h:69
    GpuSparseBasisSampler(const TableauNormMinimiser<T> &tableau, const std::vector<std::function<double(T)>> &tableauFactors,
                          std::unique_ptr<PerturbableFunction<T, double>> &&importanceFunc, const std::vector<T> &initialState,
                          double Kappa, int device = 0, int sumTree = ABM_TREE_BLOCKED)
        : kappa(Kappa) {
        AbmpFile f;
        flattenTableau<TableauNormMinimiser<T>, T>(f, tableau, tableauFactors, Kappa);
        const bool pp = detail::detectPredPrey<T>(importanceFunc.get(), f, detail::PredPreyGrids<ABMCMC_PREDPREY_GRIDS>{});
        if (!pp && !detail::detectKnownTableAgents<T>(importanceFunc.get(), f))
            throw std::runtime_error("GpuSparseBasisSampler: the importance function must be TrajectoryImportance<AGENT> for PredPreyAgent<G> "
                                     "(G in ABMCMC_PREDPREY_GRIDS), CatMouseAgent, BinomialAgent<2..8> or a type listed in ABMCMC_TABLE_AGENTS");
        abm_problem_desc d{};
        const auto &dims = f.I("dims");
        d.n_tab_rows = dims[0];
        d.n_nonbasic = dims[3];
        d.tab_L = f.I("tab_L").data(); d.tab_U = f.I("tab_U").data(); d.tab_F = f.I("tab_F").data();
        d.nb_col = f.I("nb_col").data(); d.nb_ptr = f.I("nb_ptr").data(); d.nb_row = f.I("nb_row").data(); d.nb_val = f.I("nb_val").data();
        d.fac_id = f.I("fac_id").data();
        d.n_tables = int32_t(f.I("ut_lo").size());
        d.ut_lo = f.I("ut_lo").data(); d.ut_off = f.I("ut_off").data(); d.ut_val = f.D("ut_val").data();
        d.kappa = Kappa;
        if (pp) {
            d.grid_size = f.I("pp_meta")[0];
            d.n_timesteps = f.I("pp_meta")[1];
        } else {
            const auto &meta = f.I("ag_meta");
            d.n_agent_states = meta[0]; d.n_acts = meta[1]; d.n_timesteps = meta[2]; d.n_agent_classes = meta[3];
            d.agent_class = f.I("ag_class").data(); d.agent_influencers = f.I("ag_infl").data();
            d.agent_nb_ptr = f.I("ag_nb_ptr").data(); d.agent_nb = f.I("ag_nb").data();
            d.agent_cq_ptr = f.I("ag_cq_ptr").data(); d.agent_cq = f.I("ag_cq").data();
        }
        d.imp_lgamma = f.D("imp_lgamma").data();
        d.imp_logratio = f.D("imp_logratio").data();
        d.sum_tree = sumTree;
        detail::check(abm_problem_create(&d, device, &problem_), "abm_problem_create");
        abm_problem_info info;
        detail::check(abm_problem_get_info(problem_, &info), "abm_problem_get_info");
        nRows_ = info.n_rows;
        nCols_ = info.n_cols;
        init_.assign(initialState.begin(), initialState.end());
        if (pp) actPmf_ = f.D("pp_actpmf");
        // the reference constructs one chain here, including findInitialFeasibleSolution() on Random::gen (:181)
        createChains(1, 0, /*exact=*/true);
    }

    // SparseBasisSampler.h:71,73
    GpuSparseBasisSampler(const WeightedFactoredConvexDistribution<T> &distribution, const std::vector<T> &initialState, double Kappa)
        : GpuSparseBasisSampler(TableauNormMinimiser<T>(distribution.constraints), distribution.factors,
                                distribution.perturbableFunctionFactory(), initialState, Kappa) {}
    GpuSparseBasisSampler(const WeightedFactoredConvexDistribution<T> &distribution, double Kappa)
        : GpuSparseBasisSampler(distribution, std::vector<T>(), Kappa) {}

    GpuSparseBasisSampler(const GpuSparseBasisSampler &) = delete;
    GpuSparseBasisSampler &operator=(const GpuSparseBasisSampler &) = delete;
    ~GpuSparseBasisSampler() {
        if (chains_) abm_chains_destroy(chains_);
        if (problem_) abm_problem_destroy(problem_);
    }

    int nRows() const { return nRows_; }
    int nCols() const { return nCols_; }

    // operator() (:75,212-214): the next FEASIBLE sample; the returned reference stays valid until the next call.
    const std::vector<T> &operator()() {
        // one Metropolis step per device call, each fed the two doubles the reference would draw from Random::gen
        // (column draw, then accept test); stops at the first feasible state like the do-while at :221-247.
        // This is the exact, sample-for-sample mode; throughput comes from the batched interface below.
        // abm_chains_step_traced reads [n_chains][2] uniforms and writes [n_chains] flags: the buffers below are one chain's
        if (nChains_ != 1)
            throw std::logic_error("GpuSparseBasisSampler::operator(): the exact mode drives one chain; after batched() / "
                                   "batchedFromPrior() use step() / sample() / statistics");
        double u[2];
        int32_t inf = 0;
        uint8_t acc = 0;
        do {
            u[0] = detail::canonical(Random::gen);
            u[1] = detail::canonical(Random::gen);
            detail::check(abm_chains_step_traced(chains_, 1, u, nullptr, nullptr, &acc, &inf), "abm_chains_step_traced");
            if (acc) currentInfeasibility = inf;
        } while (currentInfeasibility > 0);
        refresh();
        return X;
    }

    // ---- batched use: nChains independent chains from the same initial state on Philox streams (seed, chain)
    void batched(int nChains, uint64_t seed) { createChains(nChains, seed, /*exact=*/false); }
    // ... or each chain from its own prior sample drawn on the device (what FiguresForPaper.h:56 gives every chain);
    // pPredator / pPrey are the start-state probabilities (Bernoulli) or rates (Poisson) of the problem (PredPreyProblem.h:36-37)
    void batchedFromPrior(int nChains, uint64_t seed, double pPredator, double pPrey, int startKind = ABM_START_BERNOULLI) {
        if (actPmf_.empty()) throw std::logic_error("GpuSparseBasisSampler::batchedFromPrior: the device-side prior is PredPrey's; use batched()");
        if (chains_) { abm_chains_destroy(chains_); chains_ = nullptr; }
        nChains_ = nChains;
        detail::check(abm_chains_create(problem_, nChains, seed, 0, nullptr, &chains_), "abm_chains_create");
        detail::check(abm_chains_init_from_prior(chains_, actPmf_.data(), pPredator, pPrey, seed ^ 0x5DEECE66Dull, nullptr, startKind), "abm_chains_init_from_prior");
        detail::check(abm_chains_find_feasible(chains_, 0, nullptr, 0, nullptr), "abm_chains_find_feasible");
        refresh();
    }
    void step(int64_t nSteps) { detail::check(abm_chains_step(chains_, nSteps), "abm_chains_step"); }
    void sample(int64_t nSamples, int64_t maxSteps = 0) { detail::check(abm_chains_sample(chains_, nSamples, maxSteps), "abm_chains_sample"); }
    void synchronize() { detail::check(abm_chains_synchronize(chains_), "abm_chains_synchronize"); }
    // statistics of the batched chains without stalling them (what the statistics thread of FiguresForPaper.h:86-104 does beside
    // the sampler thread): enqueue after a sample() call, keep launching, collect with waitStatistics (any pointer may be null)
    void enqueueStatistics() { detail::check(abm_chains_enqueue_statistics(chains_), "abm_chains_enqueue_statistics"); }
    void waitStatistics(int64_t *endStateSum, int64_t *synSum, int64_t *synSumSq, int64_t *nSamples, abm_mcmc_counters *counters) {
        detail::check(abm_chains_wait_statistics(chains_, endStateSum, synSum, synSumSq, nSamples, counters), "abm_chains_wait_statistics");
    }
    // uniforms drawn ahead per device call of the exact-mode feasibility search (tests shrink it to cross many blocks)
    static int64_t &feasibilityBlock() { static int64_t b = 1 << 14; return b; }
    abm_chains *chains() { return chains_; }
    abm_problem *problem() { return problem_; }

private:
    abm_problem *problem_ = nullptr;
    abm_chains *chains_ = nullptr;
    int nRows_ = 0, nCols_ = 0, nChains_ = 0;
    std::vector<int8_t> init_;
    std::vector<double> actPmf_;

    void createChains(int nChains, uint64_t seed, bool exact) {
        if (chains_) { abm_chains_destroy(chains_); chains_ = nullptr; }
        nChains_ = nChains;
        detail::check(abm_chains_create(problem_, nChains, seed, 0, nullptr, &chains_), "abm_chains_create");
        std::vector<int8_t> init(size_t(nChains) * init_.size());
        for (int c = 0; c < nChains; ++c) std::copy(init_.begin(), init_.end(), init.begin() + size_t(c) * init_.size());
        detail::check(abm_chains_init(chains_, init.empty() ? nullptr : init.data(), int32_t(init_.size())), "abm_chains_init");
        if (!exact) {
            detail::check(abm_chains_find_feasible(chains_, 0, nullptr, 0, nullptr), "abm_chains_find_feasible");
        } else {
            // findInitialFeasibleSolution (:254-263) on Random::gen: one uniform per iteration.  Blocks are drawn from a
            // copy; afterwards the real generator is advanced by exactly the iterations the search used.
            // abm_chains_find_feasible reports the iterations of THAT call (its device counter is zeroed per call), so after a
            // block that ends feasible the generator is the block's start advanced by the block's own iterations
            const int64_t block = feasibilityBlock();
            std::vector<double> u(block);
            int rc;
            do {
                std::mt19937 saved = Random::gen;
                for (auto &x : u) x = detail::canonical(Random::gen);
                int64_t iters = 0;
                rc = abm_chains_find_feasible(chains_, block, u.data(), block, &iters);
                if (rc != ABM_OK && rc != ABM_ERR_NOT_FEASIBLE) detail::check(rc, "abm_chains_find_feasible");
                if (rc == ABM_OK) {
                    Random::gen = saved;
                    Random::gen.discard(2ull * uint64_t(iters));
                }
            } while (rc == ABM_ERR_NOT_FEASIBLE);
        }
        refresh();
    }

    void refresh() {
        std::vector<int32_t> x(nRows_);
        detail::check(abm_chains_get_state(chains_, 0, x.data(), nullptr, nullptr, nullptr, nullptr, nullptr), "abm_chains_get_state");
        X.assign(x.begin(), x.end());
        std::vector<int32_t> inf0(nChains_);
        detail::check(abm_chains_get_infeasibility(chains_, inf0.data()), "abm_chains_get_infeasibility");
        currentInfeasibility = inf0[0];
        std::vector<abm_mcmc_counters> c(nChains_);
        detail::check(abm_chains_get_counters(chains_, c.data()), "abm_chains_get_counters");
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                int64_t p = 0, q = 0;
                for (const auto &k : c) { p += k.n_proposals[a][b]; q += k.n_accepted[a][b]; }
                stats.nProposals[a][b] = int(p);
                stats.nAccepted[a][b] = int(q);
            }
    }
};

}  // namespace abmcmc
