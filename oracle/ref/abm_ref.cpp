// abm_ref: oracle harness around the UNMODIFIED reference implementation.
//
// TEST INFRASTRUCTURE ONLY.  This translation unit is compiled against the reference headers
// where they lie under /root/reference/src (see oracle/Makefile); no reference source is copied
// into this repository.  It exists to (1) export reference-built problems as ABMP files,
// (2) emit golden step traces of the reference SparseBasisSampler, (3) time the reference
// sampler on host cores (bench.py --impl reference / cpu_baseline), (4) run long reference
// chains for posterior marginals.  Nothing in the product path may link or execute it.
//
// Include order mirrors the reference's main.cpp -> Experiments.h (the headers are
// include-order dependent).
#include <boost/serialization/access.hpp>
#include <chrono>
#include <future>
#include <thread>

#include "Trajectory.h"
#include "Forecast.h"
#include "Likelihood.h"
#include "PredPreyProblem.h"
#include "SparseBasisSampler.h"
// FiguresForPaper.h (synopsis(), :245-263) and diagnostics/MultiChainStats.h pull in the vendored gnuplot-iostream, which needs
// Boost.Iostreams; its include guard is pre-defined and a no-op Gnuplot class stands in
#define GNUPLOT_IOSTREAM_H
#include <cfloat>
#include <tuple>
struct Gnuplot {
    template <class T> Gnuplot &operator<<(const T &) { return *this; }
    template <class T> void send1d(const T &) {}
    template <class T> void send2d(const T &) {}
};
#include "FiguresForPaper.h"

#include "abmcmc_b200/abmp_io.hpp"
#include "abmcmc_b200/reference_flatten.hpp"
#ifdef ABM_WITH_BASIS_ADAPTOR   // the abm_ref_fast build (oracle/Makefile `fast`): `gen ... fastbasis` builds the basis with the
#include "abmcmc_b200/basis_adaptor.hpp"   // drop-in (proven equal to the reference's, tests/test_basis_*.py) for sizes where
#endif                                     // the reference's own TableauNormMinimiser takes hours (64x64x16)

using abmcmc::AbmpFile;
typedef ABM::occupation_type occ_t;

// ---------------------------------------------------------------------------------------------
// Step-by-step access to the reference sampler.  step() is the body of the do-loop of
// nextSampleWithInfeasibleImportance() (SparseBasisSampler.h:221-246) re-issued call by call on the
// reference's own Proposal / perturb / getLogValue / calcRatioOfSums / applyProposal, so that a
// trace can be recorded per Metropolis step; `selfcheck` proves it is state-identical to the
// unmodified operator().
template <class T>
struct TracedSampler : public SparseBasisSampler<T> {
    using Base = SparseBasisSampler<T>;
    using Base::Base;

    struct StepRecord {
        int j;
        bool accepted;
        T infeasibility;  // of the proposal
    };

    StepRecord step() {
        typename Base::Proposal proposal(*this);
        this->importanceFunc->perturb(this->X, proposal.changedXIndices);
        for (size_t nzi = 0; nzi < proposal.changedXIndices.size(); ++nzi)
            std::swap(this->X[proposal.changedXIndices[nzi]], proposal.changedXValues[nzi]);
        double proposedLogImportance = this->importanceFunc->getLogValue(this->X);
        double ratioOfSums = proposal.calcRatioOfSums(this->basisDistribution);
        double acceptance = exp(proposedLogImportance - this->currentLogImportance) * ratioOfSums;
        StepRecord rec{proposal.proposedj, false, proposal.infeasibility};
        if (Random::nextDouble() < acceptance) {
            this->stats.addSample(true, this->currentInfeasibility == 0, proposal.infeasibility == 0);
            this->currentLogImportance = proposedLogImportance;
            this->applyProposal(proposal, false);
            rec.accepted = true;
        } else {
            for (size_t nzi = 0; nzi < proposal.changedXIndices.size(); ++nzi)
                this->X[proposal.changedXIndices[nzi]] = proposal.changedXValues[nzi];
            this->importanceFunc->undo();
            this->stats.addSample(false, this->currentInfeasibility == 0, proposal.infeasibility == 0);
        }
        return rec;
    }
};

static uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static uint64_t hashX(const std::vector<occ_t> &X) {  // sum_i X_i * mix64(i)  (mod 2^64)
    uint64_t h = 0;
    for (size_t i = 0; i < X.size(); ++i) h += uint64_t(int64_t(X[i])) * mix64(i);
    return h;
}
static uint64_t bitsOf(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
}

// ---------------------------------------------------------------------------------------------
// Start-state prior of a problem: the Bernoulli one PredPreyProblem hard-wires (PredPreyProblem.h:64-66), or the Poisson
// alternative the reference keeps commented out next to it (:60-62, PoissonStartState.h:18-31) with the same per-type
// rates.  startKind: 0 = Bernoulli, 1 = Poisson.
template <int G>
Prior<PredPreyAgent<G>> makePrior(int T, int startKind, double pPred, double pPrey) {
    typedef PredPreyAgent<G> Agent;
    auto rate = [pPred, pPrey](Agent agent) { return agent.type() == Agent::PREDATOR ? pPred : pPrey; };
    if (startKind == 1) return Prior<Agent>(T, PoissonStartState<Agent>(rate));
    return Prior<Agent>(T, BernoulliStartState<Agent>(rate));
}

template <int G>
AbmpFile exportProblem(const PredPreyProblem<G> &problem, int seed, double pMakeObs, double pObsIfPresent, int startKind = 0) {
    typedef PredPreyAgent<G> Agent;
    AbmpFile f;
    abmcmc::flattenTableau<TableauNormMinimiser<occ_t>, occ_t>(f, problem.tableau, problem.posterior.factors, problem.kappa);
    abmcmc::flattenPredPreyImportance<Agent>(f, G, problem.nTimesteps());
    f.f64["pp_params"] = {problem.pPredator, problem.pPrey, pMakeObs, pObsIfPresent};
    f.i32["gen_seed"] = {seed};
    f.i32["start_kind"] = {startKind};
    std::vector<int32_t> rtIdx, rtVal;
    const std::vector<occ_t> &rt = problem.realTrajectory;
    for (size_t e = 0; e < rt.size(); ++e)
        if (rt[e] != 0) {
            rtIdx.push_back(int32_t(e));
            rtVal.push_back(rt[e]);
        }
    f.i32["rt_idx"] = rtIdx;
    f.i32["rt_val"] = rtVal;
    std::vector<int32_t> ot, oa, on;
    std::vector<double> op;
    for (const auto &o : problem.likelihood.observations) {
        ot.push_back(o.state.time);
        oa.push_back(int(o.state.agent));
        on.push_back(o.lowerBound);
        op.push_back(o.pObserveIfPresent);
    }
    f.i32["obs_time"] = ot;
    f.i32["obs_agent"] = oa;
    f.i32["obs_n"] = on;
    f.f64["obs_p"] = op;
    return f;
}

// Rebuild a reference PredPreyProblem<G> from an ABMP file: the same field-by-field assembly the
// reference's own load() does (PredPreyProblem.h:92-99), with the tableau filled from the flat
// arrays instead of a Boost archive.
template <int G>
PredPreyProblem<G> importProblem(const AbmpFile &f) {
    typedef PredPreyAgent<G> Agent;
    PredPreyProblem<G> p;
    const auto &meta = f.I("pp_meta");
    if (meta[0] != G) throw std::runtime_error("importProblem: grid size mismatch");
    int T = meta[1];
    p.pPredator = f.D("pp_params")[0];
    p.pPrey = f.D("pp_params")[1];
    p.kappa = f.D("kappa")[0];
    p.realTrajectory = Trajectory<Agent>(T);
    const auto &ri = f.I("rt_idx"), &rv = f.I("rt_val");
    for (size_t k = 0; k < ri.size(); ++k) p.realTrajectory[ri[k]] = rv[k];
    std::vector<AgentStateObservation<Agent>> observations;
    const auto &ot = f.I("obs_time"), &oa = f.I("obs_agent"), &on = f.I("obs_n");
    const auto &op = f.D("obs_p");
    for (size_t k = 0; k < ot.size(); ++k)
        observations.emplace_back(State<Agent>(ot[k], Agent(oa[k])), on[k], op[k]);
    const int startKind = f.i32.count("start_kind") ? f.I("start_kind")[0] : 0;
    p.prior = makePrior<G>(T, startKind, p.pPredator, p.pPrey);
    p.likelihood = Likelihood<Agent>(observations);
    p.posterior = p.likelihood * p.prior;

    const auto &dims = f.I("dims");
    int nTabRows = dims[0], nTabCols = dims[1];
    auto &tab = p.tableau;
    tab.rows.resize(nTabRows);
    tab.cols.resize(nTabCols);
    tab.nAuxiliaryVars = dims[2];
    tab.L.assign(f.I("tab_L").begin(), f.I("tab_L").end());
    tab.U.assign(f.I("tab_U").begin(), f.I("tab_U").end());
    tab.F.assign(f.I("tab_F").begin(), f.I("tab_F").end());
    const auto &cb = f.u8.at("col_basic");
    for (int j = 0; j < nTabCols; ++j) tab.cols[j].isBasic = cb[j] != 0;
    const auto &nbCol = f.I("nb_col"), &nbPtr = f.I("nb_ptr"), &nbRow = f.I("nb_row"), &nbVal = f.I("nb_val");
    for (size_t c = 0; c < nbCol.size(); ++c)
        for (int k = nbPtr[c]; k < nbPtr[c + 1]; ++k) {
            tab.cols[nbCol[c]].insert(nbRow[k]);
            tab.rows[nbRow[k]][nbCol[c]] = nbVal[k];
        }
    if ((int)p.posterior.factors.size() != nTabRows) throw std::runtime_error("importProblem: factor count mismatch");
    return p;
}

template <int G>
std::vector<occ_t> priorSample(const PredPreyProblem<G> &problem) {
    return problem.prior.nextSample(false);  // FiguresForPaper.h:56,162
}

// ---------------------------------------------------------------------------------------------
template <int G>
int cmdGen(int T, int seed, double kappa, const std::string &out, double pPred, double pPrey, double pMakeObs,
           double pObsIfPresent, int startKind, bool fastBasis) {
    auto t0 = std::chrono::steady_clock::now();
    Random::gen.seed(seed);  // FiguresForPaper.h:33
    // the member-by-member assembly of PredPreyProblem's constructor (PredPreyProblem.h:46-56), so that the start state can
    // be either of the two the reference provides; with startKind 0 it draws the same random numbers in the same order
    PredPreyProblem<G> problem;
    problem.pPredator = pPred;
    problem.pPrey = pPrey;
    problem.kappa = kappa;
    problem.prior = makePrior<G>(T, startKind, pPred, pPrey);
    problem.realTrajectory = problem.prior.nextSample(true);
    problem.likelihood = Likelihood<PredPreyAgent<G>>(problem.realTrajectory, pMakeObs, pObsIfPresent);
    problem.posterior = problem.likelihood * problem.prior;
    if (fastBasis) {
#ifdef ABM_WITH_BASIS_ADAPTOR
        problem.tableau = abmcmc::minimiseBasis(problem.posterior.constraints);
#else
        throw std::runtime_error("gen: fastbasis needs the abm_ref_fast build");
#endif
    } else {
        problem.tableau = TableauNormMinimiser<occ_t>(problem.posterior.constraints);
        problem.tableau.findMinimalBasis();
    }
    auto t1 = std::chrono::steady_clock::now();
    AbmpFile f = exportProblem<G>(problem, seed, pMakeObs, pObsIfPresent, startKind);
    f.f64["gen_seconds"] = {std::chrono::duration<double>(t1 - t0).count()};
    f.write(out);
    const auto &d = f.I("dims");
    fprintf(stderr, "gen G=%d T=%d: tableau %d x %d, nAux=%d nNonBasic=%d nnz=%d, %zu unique factor tables, built in %.2fs\n",
            G, T, d[0], d[1], d[2], d[3], d[4], f.I("ut_lo").size(), f.D("gen_seconds")[0]);
    return 0;
}

// Writes nChains prior samples (event vectors, int8, [nChains][nEvents]) drawn as FiguresForPaper.h:56.
template <int G>
int cmdInit(const AbmpFile &f, int nChains, int seed, const std::string &out) {
    PredPreyProblem<G> problem = importProblem<G>(f);
    Random::gen.seed(seed);
    FILE *fo = fopen(out.c_str(), "wb");
    if (!fo) throw std::runtime_error("cannot open output");
    for (int c = 0; c < nChains; ++c) {
        std::vector<occ_t> s = priorSample<G>(problem);
        std::vector<int8_t> b(s.begin(), s.end());
        fwrite(b.data(), 1, b.size(), fo);
    }
    fclose(fo);
    return 0;
}

struct SilenceStdout {  // the reference prints progress lines to std::cout from its constructors
    std::streambuf *old;
    std::ostringstream sink;
    SilenceStdout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~SilenceStdout() { std::cout.rdbuf(old); }
};

// Golden trace.  Stream layout (little-endian), all through one AbmpFile for convenience:
//   init_state i8[nEvents]; post-ctor: init_iters, X0 i32[nRows], delta0 i8[nCols], DE0 f64[nCols],
//   leaf0 f64[nCols] (basisDistribution.get), sum0, logImp0;
//   per step: tr_j i32, tr_acc u8, tr_inf i32 (proposal infeasibility);
//   checkpoints every `every` steps: ck_hash (two i32 per u64), ck_logimp f64, ck_sum f64, ck_inf i32;
//   final: X, delta, DE, leaf, counters.
// (the body works for any agent type: `synopsis` maps an end state to the statistics of FiguresForPaper.h:245-263 where the
// agent has them -- PredPrey -- and to nothing otherwise)
template <class Agent, class SYNOPSIS>
int traceSampler(const TableauNormMinimiser<occ_t> &tableau, const std::vector<std::function<double(occ_t)>> &factors,
                 std::unique_ptr<PerturbableFunction<occ_t, double>> importance, const std::vector<occ_t> &init, double kappa, int T,
                 SYNOPSIS synopsis, int initSeed, int chainSeed, long nSteps, int every, const std::string &out) {
    Random::gen.seed(chainSeed);
    std::unique_ptr<TracedSampler<occ_t>> sp;
    std::string ctorLog;
    {
        SilenceStdout quiet;
        sp.reset(new TracedSampler<occ_t>(tableau, factors, std::move(importance), init, kappa));
        ctorLog = quiet.sink.str();
    }
    auto &s = *sp;
    AbmpFile t;
    int initIters = 0;
    sscanf(ctorLog.c_str(), "Found feasible solution in %d", &initIters);
    t.i32["seeds"] = {initSeed, chainSeed};
    t.i32["init_iters"] = {initIters};
    t.i8["init_state"] = std::vector<int8_t>(init.begin(), init.end());
    auto snapshot = [&](const std::string &sfx) {
        t.i32["X" + sfx] = std::vector<int32_t>(s.X.begin(), s.X.end());
        t.i8["delta" + sfx] = std::vector<int8_t>(s.delta.begin(), s.delta.end());
        t.f64["DE" + sfx] = s.currentDE;
        std::vector<double> leaf(s.nCols());
        for (int j = 0; j < s.nCols(); ++j) leaf[j] = s.basisDistribution.get(j);
        t.f64["leaf" + sfx] = leaf;
        t.f64["scal" + sfx] = {s.basisDistribution.sum(), s.currentLogImportance};
        t.i32["inf" + sfx] = {int32_t(s.currentInfeasibility)};
    };
    snapshot("0");
    std::vector<int32_t> trJ, trInf, ckHash, ckInf, ckEnd, ckSyn;
    std::vector<uint8_t> trAcc;
    std::vector<double> ckLogImp, ckSum;
    trJ.reserve(nSteps);
    // what the reference's statistics pipeline sees (FiguresForPaper.h:93-104): every state the do-while of
    // nextSampleWithInfeasibleImportance (SparseBasisSampler.h:247) hands out, i.e. the state after every step that ends
    // feasible, mapped by TrajectoryToModelState(T,T) (ModelState.h:27-36) and synopsis() (FiguresForPaper.h:245-263) --
    // the reference's own code -- and summed as Sum<ModelState> / MeanAndVariance do
    std::vector<double> smpEnd(Agent::domainSize(), 0.0);
    MeanAndVariance smpSyn;
    long nFeasible = 0;
    for (long k = 0; k < nSteps; ++k) {
        auto r = s.step();
        trJ.push_back(r.j);
        trAcc.push_back(r.accepted ? 1 : 0);
        trInf.push_back(int32_t(r.infeasibility));
        if (s.currentInfeasibility == 0) {
            ModelState<Agent> endState(s.X, T, T);
            for (int q = 0; q < Agent::domainSize(); ++q) smpEnd[q] += endState[q];
            smpSyn(synopsis(endState));
            ++nFeasible;
        }
        if ((k + 1) % every == 0 || k + 1 == nSteps) {
            uint64_t h = hashX(s.X);
            ckHash.push_back(int32_t(uint32_t(h)));
            ckHash.push_back(int32_t(uint32_t(h >> 32)));
            ckLogImp.push_back(s.currentLogImportance);
            ckSum.push_back(s.basisDistribution.sum());
            ckInf.push_back(int32_t(s.currentInfeasibility));
            ModelState<Agent> endState(s.X, T, T);
            for (int q = 0; q < Agent::domainSize(); ++q) ckEnd.push_back(int32_t(endState[q]));
            for (double v : synopsis(endState)) ckSyn.push_back(int32_t(v));
        }
    }
    t.i32["ck_endstate"] = ckEnd;        // [checkpoints][S]  ModelState(X,T,T)
    t.i32["ck_synopsis"] = ckSyn;        // [checkpoints][nSynopsis]
    t.f64["smp_endstate_sum"] = smpEnd;  // [S] over the feasible states
    t.f64["smp_syn_sum"] = std::vector<double>(std::begin(smpSyn.sum), std::end(smpSyn.sum));
    t.f64["smp_syn_sumsq"] = std::vector<double>(std::begin(smpSyn.sumOfSquares), std::end(smpSyn.sumOfSquares));
    t.i32["smp_n"] = {int32_t(nFeasible)};
    t.i32["tr_j"] = trJ;
    t.u8["tr_acc"] = trAcc;
    t.i32["tr_inf"] = trInf;
    t.i32["ck_every"] = {every};
    t.i32["ck_hash"] = ckHash;
    t.f64["ck_logimp"] = ckLogImp;
    t.f64["ck_sum"] = ckSum;
    t.i32["ck_inf"] = ckInf;
    snapshot("N");
    std::vector<int32_t> ctr;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) ctr.push_back(s.stats.nProposals[a][b]);
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) ctr.push_back(s.stats.nAccepted[a][b]);
    t.i32["counters"] = ctr;
    t.write(out);
    fprintf(stderr, "trace: init_iters=%d steps=%ld accepted=%d feasible-ending=%d\n", initIters, nSteps,
            s.stats.totalAccepted(), s.stats.nFeasibleSamples());
    return 0;
}

template <int G>
int cmdTrace(const AbmpFile &f, int initSeed, int chainSeed, long nSteps, int every, const std::string &out) {
    PredPreyProblem<G> problem = importProblem<G>(f);
    Random::gen.seed(initSeed);
    std::vector<occ_t> init = priorSample<G>(problem);
    return traceSampler<PredPreyAgent<G>>(problem.tableau, problem.posterior.factors, problem.posterior.perturbableFunctionFactory(), init,
                                          problem.kappa, problem.nTimesteps(),
                                          [](const ModelState<PredPreyAgent<G>> &e) { return FiguresForPaper<G, 1>::synopsis(e); }, initSeed,
                                          chainSeed, nSteps, every, out);
}

// Proves TracedSampler::step() == the unmodified operator(): two samplers, same seeds, one driven
// by operator(), the other by step() until feasible; all state compared after every sample.
template <int G>
int cmdSelfcheck(const AbmpFile &f, int nSamples) {
    PredPreyProblem<G> problem = importProblem<G>(f);
    Random::gen.seed(7);
    std::vector<occ_t> init = priorSample<G>(problem);
    std::mt19937 genA, genB;
    SilenceStdout quiet;
    Random::gen.seed(99);
    SparseBasisSampler<occ_t> a(problem.tableau, problem.posterior.factors, problem.posterior.perturbableFunctionFactory(), init, problem.kappa);
    genA = Random::gen;
    Random::gen.seed(99);
    TracedSampler<occ_t> b(problem.tableau, problem.posterior.factors, problem.posterior.perturbableFunctionFactory(), init, problem.kappa);
    genB = Random::gen;
    for (int n = 0; n < nSamples; ++n) {
        Random::gen = genA;
        a();
        genA = Random::gen;
        Random::gen = genB;
        do { b.step(); } while (b.currentInfeasibility > 0);
        genB = Random::gen;
        bool same = a.X == b.X && a.delta == b.delta && a.currentDE == b.currentDE &&
                    bitsOf(a.currentLogImportance) == bitsOf(b.currentLogImportance) &&
                    bitsOf(a.basisDistribution.sum()) == bitsOf(b.basisDistribution.sum()) &&
                    a.stats.nSamples() == b.stats.nSamples() && a.stats.totalAccepted() == b.stats.totalAccepted() &&
                    genA == genB;
        if (!same) {
            fprintf(stderr, "selfcheck: MISMATCH at sample %d\n", n);
            return 1;
        }
    }
    fprintf(stderr, "selfcheck: traced step() identical to operator() over %d samples (%d steps)\n", nSamples, a.stats.nSamples());
    return 0;
}

struct TimingResult {
    long steps = 0, feasible = 0, accepted = 0;
    double seconds = 0;
};

// sampleTiming protocol (FiguresForPaper.h:151-178) on one thread: construct from a prior sample,
// nSamples/4 burn-in calls, then time nSamples calls of sampler(); steps from sampler.stats.
template <int G>
TimingResult timeOneChain(const PredPreyProblem<G> &problem, std::vector<occ_t> init, int threadSeed, long nSamples, double maxSeconds) {
    Random::gen.seed(threadSeed);  // thread-local generator (Random.h:15)
    SparseBasisSampler<occ_t> sampler(problem.tableau, problem.posterior.factors,
                                      problem.posterior.perturbableFunctionFactory(), init, problem.kappa);
    for (long b = 0; b < nSamples / 4; ++b) sampler();
    long steps0 = sampler.stats.nSamples(), acc0 = sampler.stats.totalAccepted();
    auto t0 = std::chrono::steady_clock::now();
    long done = 0;
    for (; done < nSamples; ++done) {
        sampler();
        if ((done & 1023) == 1023 && maxSeconds > 0 &&
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > maxSeconds) { ++done; break; }
    }
    auto t1 = std::chrono::steady_clock::now();
    TimingResult r;
    r.steps = sampler.stats.nSamples() - steps0;
    r.accepted = sampler.stats.totalAccepted() - acc0;
    r.feasible = done;
    r.seconds = std::chrono::duration<double>(t1 - t0).count();
    return r;
}

template <int G>
int cmdTime(const AbmpFile &f, int nThreads, long nSamples, double maxSeconds) {
    PredPreyProblem<G> problem = importProblem<G>(f);
    Random::gen.seed(4321);
    std::vector<std::vector<occ_t>> inits;
    for (int t = 0; t < nThreads; ++t) inits.push_back(priorSample<G>(problem));
    std::vector<std::future<TimingResult>> fut;
    auto w0 = std::chrono::steady_clock::now();
    {
        SilenceStdout quiet;
        for (int t = 0; t < nThreads; ++t)  // one independent chain per host thread (FiguresForPaper.h:54-57)
            fut.push_back(std::async(std::launch::async, [&problem, &inits, t, nSamples, maxSeconds]() {
                return timeOneChain<G>(problem, inits[t], 1000 + t, nSamples, maxSeconds);
            }));
        for (auto &x : fut) x.wait();
    }
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    long steps = 0, feas = 0, acc = 0;
    double maxSec = 0, rate = 0;
    for (auto &x : fut) {
        TimingResult r = x.get();
        steps += r.steps;
        feas += r.feasible;
        acc += r.accepted;
        maxSec = std::max(maxSec, r.seconds);
        rate += r.steps / r.seconds;
    }
    printf("{\"impl\": \"reference\", \"grid\": %d, \"threads\": %d, \"steps\": %ld, \"feasible_samples\": %ld, \"accepted\": %ld, "
           "\"timed_seconds_max\": %.6f, \"steps_per_s\": %.3f, \"wall_seconds_incl_setup\": %.3f}\n",
           G, nThreads, steps, feas, acc, maxSec, rate, wall);
    return 0;
}


// bench.py --impl reference: nThreads persistent reference samplers (one chain per host thread, as
// FiguresForPaper.h:54-57), advanced in `warmup + steps` rounds of `samplesPerStep` calls of sampler() each.
// Prints one JSON object with the wall time and Metropolis-step count of every timed round.
template <int G>
int cmdBench(const AbmpFile &f, int nThreads, long samplesPerStep, int warmup, int steps) {
    PredPreyProblem<G> problem = importProblem<G>(f);
    Random::gen.seed(4321);
    std::vector<std::vector<occ_t>> inits;
    for (int t = 0; t < nThreads; ++t) inits.push_back(priorSample<G>(problem));
    std::vector<std::unique_ptr<SparseBasisSampler<occ_t>>> samplers(nThreads);
    std::vector<std::mt19937> gens(nThreads);
    auto c0 = std::chrono::steady_clock::now();
    {
        SilenceStdout quiet;
        std::vector<std::future<void>> fut;
        for (int t = 0; t < nThreads; ++t)
            fut.push_back(std::async(std::launch::async, [&, t]() {
                Random::gen.seed(1000 + t);
                samplers[t].reset(new SparseBasisSampler<occ_t>(problem.tableau, problem.posterior.factors,
                                                                problem.posterior.perturbableFunctionFactory(), inits[t], problem.kappa));
                gens[t] = Random::gen;
            }));
        for (auto &x : fut) x.wait();
    }
    double setup = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();
    // Every thread times its own loop: a round's Metropolis-step rate is the SUM of the per-thread rates (steps per feasible
    // sample vary widely between chains, so joining the threads and dividing by the slowest would under-report the CPU).
    std::vector<double> secs, rates, sampleRates;
    std::vector<long> stepCounts;
    for (int r = 0; r < warmup + steps; ++r) {
        std::vector<double> tSec(nThreads, 0.0);
        std::vector<long> tSteps(nThreads, 0);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::future<void>> fut;
        for (int t = 0; t < nThreads; ++t)
            fut.push_back(std::async(std::launch::async, [&, t]() {
                Random::gen = gens[t];
                const long before = samplers[t]->stats.nSamples();
                auto a = std::chrono::steady_clock::now();
                for (long k = 0; k < samplesPerStep; ++k) (*samplers[t])();
                tSec[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
                tSteps[t] = samplers[t]->stats.nSamples() - before;
                gens[t] = Random::gen;
            }));
        for (auto &x : fut) x.wait();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r >= warmup) {
            double rate = 0.0, srate = 0.0;
            long st = 0;
            for (int t = 0; t < nThreads; ++t) { rate += tSteps[t] / tSec[t]; srate += samplesPerStep / tSec[t]; st += tSteps[t]; }
            secs.push_back(dt);
            rates.push_back(rate);
            sampleRates.push_back(srate);
            stepCounts.push_back(st);
        }
    }
    printf("{\"impl\": \"reference\", \"grid\": %d, \"threads\": %d, \"samples_per_step\": %ld, \"setup_seconds\": %.3f, \"step_seconds\": [", G, nThreads, samplesPerStep, setup);
    for (size_t k = 0; k < secs.size(); ++k) printf("%s%.6f", k ? ", " : "", secs[k]);
    printf("], \"chain_steps\": [");
    for (size_t k = 0; k < stepCounts.size(); ++k) printf("%s%ld", k ? ", " : "", stepCounts[k]);
    printf("], \"steps_per_s\": [");
    for (size_t k = 0; k < rates.size(); ++k) printf("%s%.3f", k ? ", " : "", rates[k]);
    printf("], \"samples_per_s\": [");
    for (size_t k = 0; k < sampleRates.size(); ++k) printf("%s%.3f", k ? ", " : "", sampleRates[k]);
    printf("]}\n");
    return 0;
}

// Long reference chains for posterior marginals: per thread one chain, `burnin` + nSamples calls of
// sampler(); accumulates the end-state occupancy (ModelState(traj,T,T), ModelState.h:27-36) and the
// whole-trajectory event marginals.  Output: ABMP with endstate_sum i64-as-f64[S], event_sum f64[nEvents],
// n_samples.
template <int G>
int cmdMarginals(const AbmpFile &f, int nThreads, long nSamples, long burnin, const std::string &out) {
    typedef PredPreyAgent<G> Agent;
    PredPreyProblem<G> problem = importProblem<G>(f);
    int T = problem.nTimesteps();
    Random::gen.seed(8765);
    std::vector<std::vector<occ_t>> inits;
    for (int t = 0; t < nThreads; ++t) inits.push_back(priorSample<G>(problem));
    struct Acc { std::vector<double> end, ev; long n = 0, steps = 0; };
    std::vector<std::future<Acc>> fut;
    {
        SilenceStdout quiet;
        for (int t = 0; t < nThreads; ++t)
            fut.push_back(std::async(std::launch::async, [&problem, &inits, t, T, nSamples, burnin]() {
                Random::gen.seed(5000 + t);
                SparseBasisSampler<occ_t> sampler(problem.tableau, problem.posterior.factors,
                                                  problem.posterior.perturbableFunctionFactory(), inits[t], problem.kappa);
                for (long b = 0; b < burnin; ++b) sampler();
                Acc a;
                int nEvents = T * Agent::domainSize() * Agent::actDomainSize();
                a.end.assign(Agent::domainSize(), 0.0);
                a.ev.assign(nEvents, 0.0);
                long s0 = sampler.stats.nSamples();
                for (long n = 0; n < nSamples; ++n) {
                    const std::vector<occ_t> &x = sampler();
                    ModelState<Agent> endState(x, T, T);
                    for (int k = 0; k < Agent::domainSize(); ++k) a.end[k] += endState[k];
                    for (int e = 0; e < nEvents; ++e) a.ev[e] += x[e];
                }
                a.n = nSamples;
                a.steps = sampler.stats.nSamples() - s0;
                return a;
            }));
        for (auto &x : fut) x.wait();
    }
    AbmpFile o;
    std::vector<double> end, ev;
    double n = 0, steps = 0;
    std::vector<double> perChainEnd;
    for (auto &x : fut) {
        Acc a = x.get();
        if (end.empty()) { end.assign(a.end.size(), 0.0); ev.assign(a.ev.size(), 0.0); }
        for (size_t k = 0; k < end.size(); ++k) end[k] += a.end[k];
        for (size_t k = 0; k < ev.size(); ++k) ev[k] += a.ev[k];
        for (size_t k = 0; k < a.end.size(); ++k) perChainEnd.push_back(a.end[k] / a.n);
        n += a.n;
        steps += a.steps;
    }
    o.f64["endstate_sum"] = end;
    o.f64["event_sum"] = ev;
    o.f64["perchain_endstate_mean"] = perChainEnd;
    o.f64["n_samples"] = {n, steps, double(nThreads)};
    o.write(out);
    return 0;
}

// Synopsis series of long reference chains (the window sweep's CPU arm): per thread one chain from its own prior sample,
// `burnin` calls of sampler(), then nSamples calls, each mapped to the synopsis of its end state exactly as
// FiguresForPaper.h:93-104 does (TrajectoryToModelState(T,T) then synopsis(), :245-263: agents inside nested squares of
// side G/2, G/4, ... > 1 along the diagonal).  Output: series i32[nThreads][nSamples][nSyn], seconds of the sampling part
// (max over threads), Metropolis steps.
template <int G>
int cmdSeries(const AbmpFile &f, int nThreads, long nSamples, long burnin, const std::string &out) {
    typedef PredPreyAgent<G> Agent;
    PredPreyProblem<G> problem = importProblem<G>(f);
    const int T = problem.nTimesteps();
    int nSyn = 0;
    for (int ps = G / 2; ps > 1; ps /= 2) ++nSyn;
    Random::gen.seed(2468);
    std::vector<std::vector<occ_t>> inits;
    for (int t = 0; t < nThreads; ++t) inits.push_back(priorSample<G>(problem));
    struct Res { std::vector<int32_t> series; double seconds = 0; long steps = 0; };
    std::vector<std::future<Res>> fut;
    {
        SilenceStdout quiet;
        for (int t = 0; t < nThreads; ++t)
            fut.push_back(std::async(std::launch::async, [&problem, &inits, t, T, nSyn, nSamples, burnin]() {
                Random::gen.seed(7000 + t);
                SparseBasisSampler<occ_t> sampler(problem.tableau, problem.posterior.factors,
                                                  problem.posterior.perturbableFunctionFactory(), inits[t], problem.kappa);
                for (long b = 0; b < burnin; ++b) sampler();
                Res r;
                r.series.reserve(size_t(nSamples) * nSyn);
                const long s0 = sampler.stats.nSamples();
                auto t0 = std::chrono::steady_clock::now();
                for (long n = 0; n < nSamples; ++n) {
                    const std::vector<occ_t> &x = sampler();
                    ModelState<Agent> endState(x, T, T);
                    int origin = 0;
                    for (int ps = G / 2; ps > 1; ps /= 2) {
                        double occ = 0.0;
                        for (int xx = 0; xx < ps; ++xx)
                            for (int yy = 0; yy < ps; ++yy)
                                occ += endState[Agent(origin + xx, origin + yy, Agent::PREDATOR)] +
                                       endState[Agent(origin + xx, origin + yy, Agent::PREY)];
                        r.series.push_back(int32_t(occ));
                        origin += ps;
                    }
                }
                r.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                r.steps = sampler.stats.nSamples() - s0;
                return r;
            }));
        for (auto &x : fut) x.wait();
    }
    AbmpFile o;
    std::vector<int32_t> all;
    double sec = 0, steps = 0;
    for (auto &x : fut) {
        Res r = x.get();
        all.insert(all.end(), r.series.begin(), r.series.end());
        sec = std::max(sec, r.seconds);
        steps += r.steps;
    }
    o.i32["series"] = all;
    o.i32["shape"] = {nThreads, int32_t(nSamples), nSyn};
    o.f64["timing"] = {sec, steps, double(burnin)};
    o.write(out);
    return 0;
}

// Long reference chains for the paper-scale posterior fixture (protocol: FiguresForPaper.h:40-104): per thread one chain from
// its own prior sample (:56), `burnin` calls of sampler() (Drop, :93), then nSamples calls, each mapped by the reference's
// own TrajectoryToModelState(T,T) (ModelState.h:27-36) and synopsis() (FiguresForPaper.h:245-263).  Kept per chain: the
// MeanAndVariance sums, the end-state sum, means over consecutive batches of `batch` samples (burn-in batches separately:
// the relaxation from the first feasible state), and every `thin`-th synopsis sample.  Sampling part timed per thread.
template <int G>
int cmdPosterior(const AbmpFile &f, int nThreads, long nSamples, long burnin, long batch, long thin, const std::string &out, int seedBase) {
    typedef PredPreyAgent<G> Agent;
    PredPreyProblem<G> problem = importProblem<G>(f);
    const int T = problem.nTimesteps();
    const int S = Agent::domainSize();
    int nSyn = 0;
    for (int ps = G / 2; ps > 1; ps /= 2) ++nSyn;
    Random::gen.seed(97531 + seedBase);
    std::vector<std::vector<occ_t>> inits;
    for (int t = 0; t < nThreads; ++t) inits.push_back(priorSample<G>(problem));
    struct Res {
        std::vector<double> synSum, synSumSq, end, batchMeans, burnMeans;
        std::vector<int32_t> series;
        double seconds = 0, burnSeconds = 0;
        long steps = 0, burnSteps = 0, initIters = 0;
        std::vector<long> counters;
    };
    std::vector<std::future<Res>> fut;
    {
        SilenceStdout quiet;
        for (int t = 0; t < nThreads; ++t)
            fut.push_back(std::async(std::launch::async, [&problem, &inits, t, T, S, nSyn, nSamples, burnin, batch, thin, seedBase]() {
                Random::gen.seed(9000 + seedBase + t);
                SparseBasisSampler<occ_t> sampler(problem.tableau, problem.posterior.factors,
                                                  problem.posterior.perturbableFunctionFactory(), inits[t], problem.kappa);
                Res r;
                r.synSum.assign(nSyn, 0.0); r.synSumSq.assign(nSyn, 0.0); r.end.assign(S, 0.0);
                std::vector<double> bacc(nSyn, 0.0);
                auto tb = std::chrono::steady_clock::now();
                for (long b = 0; b < burnin; ++b) {
                    const std::vector<occ_t> &x = sampler();
                    if (batch > 0) {   // relaxation curve only: the burn-in samples enter no statistic
                        if (b % 16 == 0) {   // the end state of every 16th burn-in sample is enough for a batch mean
                            ModelState<Agent> endState(x, T, T);
                            std::valarray<double> syn = FiguresForPaper<G, 1>::synopsis(endState);
                            for (int d = 0; d < nSyn; ++d) bacc[d] += syn[d];
                        }
                        if ((b + 1) % batch == 0) {
                            for (int d = 0; d < nSyn; ++d) { r.burnMeans.push_back(bacc[d] / (batch / 16)); bacc[d] = 0.0; }
                        }
                    }
                }
                r.burnSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count();
                r.burnSteps = sampler.stats.nSamples();
                std::fill(bacc.begin(), bacc.end(), 0.0);
                const long s0 = sampler.stats.nSamples();
                MeanAndVariance mv;
                auto t0 = std::chrono::steady_clock::now();
                for (long n = 0; n < nSamples; ++n) {
                    const std::vector<occ_t> &x = sampler();
                    ModelState<Agent> endState(x, T, T);
                    std::valarray<double> syn = FiguresForPaper<G, 1>::synopsis(endState);
                    mv(syn);
                    for (int q = 0; q < S; ++q) r.end[q] += endState[q];
                    for (int d = 0; d < nSyn; ++d) bacc[d] += syn[d];
                    if (batch > 0 && (n + 1) % batch == 0)
                        for (int d = 0; d < nSyn; ++d) { r.batchMeans.push_back(bacc[d] / batch); bacc[d] = 0.0; }
                    if (thin > 0 && n % thin == 0)
                        for (int d = 0; d < nSyn; ++d) r.series.push_back(int32_t(syn[d]));
                }
                r.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                r.steps = sampler.stats.nSamples() - s0;
                for (int d = 0; d < nSyn; ++d) { r.synSum[d] = mv.sum[d]; r.synSumSq[d] = mv.sumOfSquares[d]; }
                for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) r.counters.push_back(sampler.stats.nProposals[a][b]);
                for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) r.counters.push_back(sampler.stats.nAccepted[a][b]);
                return r;
            }));
        for (auto &x : fut) x.wait();
    }
    AbmpFile o;
    std::vector<double> synSum, synSumSq, end, batchMeans, burnMeans, timing, counters;
    std::vector<int32_t> series;
    for (auto &x : fut) {
        Res r = x.get();
        synSum.insert(synSum.end(), r.synSum.begin(), r.synSum.end());
        synSumSq.insert(synSumSq.end(), r.synSumSq.begin(), r.synSumSq.end());
        end.insert(end.end(), r.end.begin(), r.end.end());
        batchMeans.insert(batchMeans.end(), r.batchMeans.begin(), r.batchMeans.end());
        burnMeans.insert(burnMeans.end(), r.burnMeans.begin(), r.burnMeans.end());
        series.insert(series.end(), r.series.begin(), r.series.end());
        timing.push_back(r.seconds); timing.push_back(double(r.steps)); timing.push_back(r.burnSeconds); timing.push_back(double(r.burnSteps));
        for (long c : r.counters) counters.push_back(double(c));
    }
    o.f64["syn_sum"] = synSum;            // [nThreads][nSyn]
    o.f64["syn_sumsq"] = synSumSq;
    o.f64["endstate_sum"] = end;          // [nThreads][S]
    o.f64["batch_means"] = batchMeans;    // [nThreads][nSamples / batch][nSyn]
    o.f64["burnin_batch_means"] = burnMeans;   // [nThreads][burnin / batch][nSyn]
    o.i32["series_thin"] = series;        // [nThreads][ceil(nSamples / thin)][nSyn]
    o.f64["timing"] = timing;             // [nThreads]{sampling seconds, sampling Metropolis steps, burn-in seconds, burn-in steps}
    o.f64["counters"] = counters;         // [nThreads][8] nProposals[2][2], nAccepted[2][2] (whole chain)
    o.f64["meta"] = {double(nThreads), double(nSamples), double(burnin), double(batch), double(thin), double(nSyn), double(S), double(seedBase)};
    o.write(out);
    return 0;
}

// The reference's own diagnostics (diagnostics/ChainStats.h, MultiChainStats.h) on a stored synopsis series
// (i32 "series" [nSeq][n][d], as written by `series`): every sequence becomes one ChainStats exactly as
// FiguresForPaper::startStatsThread builds them (nLags = 200, maxLagProportion = 0.5, :89-90), then
// MultiChainStats::{W, B, varPlus, potentialScaleReduction, effectiveSamples}.  Prints one JSON line; pins diagnostics.py.
static int cmdRefStats(const AbmpFile &f, int split) {
    const auto &sh = f.I("shape");
    const int nSeq = sh[0], n = sh[1], d = sh[2];
    const auto &ser = f.I("series");
    MultiChainStats ms(0.0, "refstats");
    const int parts = split ? 2 : 1, len = n / parts;
    for (int c = 0; c < nSeq; ++c)
        for (int h = 0; h < parts; ++h) {
            std::valarray<std::valarray<double>> samples(std::valarray<double>(0.0, d), len);
            for (int i = 0; i < len; ++i)
                for (int k = 0; k < d; ++k) samples[i][k] = ser[(size_t(c) * n + size_t(h) * len + i) * d + k];
            std::vector<ChainStats> one;
            one.emplace_back(std::move(samples), 200, 0.5, std::valarray<double>(0.0, 1), MCMCStatistics());
            ms += std::move(one);
        }
    auto pr = [](const char *name, const std::valarray<double> &v, bool last = false) {
        printf("\"%s\": [", name);
        for (size_t k = 0; k < v.size(); ++k) printf("%s%.17g", k ? ", " : "", v[k]);
        printf("]%s", last ? "" : ", ");
    };
    printf("{\"n_sequences\": %d, \"n_samples_per_sequence\": %d, \"vario_stride\": %d, ", ms.nChains(), ms.nSamplesPerChain(), ms.front().varioStride);
    pr("W", ms.W()); pr("B", ms.B()); pr("var_plus", ms.varPlus()); pr("rhat", ms.potentialScaleReduction());
    pr("mean", ms.mean()); pr("n_eff", ms.effectiveSamples(), true);
    printf("}\n");
    return 0;
}

// The input of TableauNormMinimiser: posterior.constraints (a ConvexPolyhedron, one Constraint per factor:
// lowerBound <= coefficients . X <= upperBound), exported as CSR in the reference's own entry order, plus the tableau the
// unmodified TableauNormMinimiser makes of it (timed).  Feeds the parity test of the own basis builder (abm_basis_*).
template <int G>
int cmdConstraints(const AbmpFile &f, const std::string &out, int rebuild) {
    PredPreyProblem<G> problem = importProblem<G>(f);
    AbmpFile o;
    std::vector<int32_t> ptr{0}, idx, val, lo, up;
    for (const auto &c : problem.posterior.constraints) {
        for (int k = 0; k < c.coefficients.sparseSize(); ++k) {
            idx.push_back(c.coefficients.indices[k]);
            val.push_back(int32_t(c.coefficients.values[k]));
        }
        ptr.push_back(int32_t(idx.size()));
        lo.push_back(int32_t(c.lowerBound));
        up.push_back(int32_t(c.upperBound));
    }
    o.i32["con_ptr"] = ptr;
    o.i32["con_idx"] = idx;
    o.i32["con_val"] = val;
    o.i32["con_lower"] = lo;
    o.i32["con_upper"] = up;
    if (rebuild) {   // the reference's own constructor + findMinimalBasis, timed, for the speed comparison
        SilenceStdout quiet;
        auto t0 = std::chrono::steady_clock::now();
        TableauNormMinimiser<occ_t> tab(problem.posterior.constraints);
        o.f64["reference_seconds"] = {std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()};
        if (tab.rows.size() != problem.tableau.rows.size()) throw std::runtime_error("constraints: rebuilt tableau differs");
    }
    o.write(out);
    fprintf(stderr, "constraints: %zu constraints, %zu coefficients\n", lo.size(), idx.size());
    return 0;
}

// The reference's own validation experiment for this path on PredPrey (Experiments::PredPreySingleObservation,
// Experiments.cpp:49-65, through doSingleObservationExperiment, :129-168): 3x3 grid, 2 timesteps, Poisson start state with
// rates 0.1 (predator) / 0.2 (prey), one observation "one prey at (1,1) at time 1", kappa = 3; the sampler starts from the
// empty trajectory (SparseBasisSampler(posterior, kappa), SparseBasisSampler.h:199-206).  Writes the problem as an ABMP
// file and, to `margOut`, the end-state means ModelState(sample, T, T) of (a) the reference sampler after nBurnin samples,
// over nSamples, and (b) the reference RejectionSampler over nRejection samples -- the two things the experiment prints
// and compares (RMS error, :163-167).
#include "RejectionSampler.h"
static int cmdSingleObservationExperiment(const std::string &out, const std::string &margOut, long nSamples, long nBurnin,
                                          long nRejection, int seed) {
    constexpr int G = 3, T = 2;
    typedef PredPreyAgent<G> Agent;
    const double pPredator = 0.1, pPrey = 2.0 * pPredator, kappa = 3.0;
    Random::gen.seed(seed);
    PredPreyProblem<G> problem;
    problem.pPredator = pPredator;
    problem.pPrey = pPrey;
    problem.kappa = kappa;
    problem.prior = makePrior<G>(T, 1, pPredator, pPrey);
    problem.realTrajectory = Trajectory<Agent>(T);
    problem.likelihood = Likelihood<Agent>(AgentStateObservation<Agent>(State<Agent>(1, Agent(1, 1, Agent::PREY)), 1, 1.0));
    problem.posterior = problem.likelihood * problem.prior;
    {
        SilenceStdout quiet;
        problem.tableau = TableauNormMinimiser<occ_t>(problem.posterior.constraints);   // the constructor runs findMinimalBasis()
    }
    AbmpFile f = exportProblem<G>(problem, seed, 0.0, 1.0, 1);
    f.write(out);

    std::vector<double> mcmc(Agent::domainSize(), 0.0), rej(Agent::domainSize(), 0.0);
    long steps = 0;
    {
        SilenceStdout quiet;
        SparseBasisSampler<occ_t> sampler(problem.tableau, problem.posterior.factors, problem.posterior.perturbableFunctionFactory(),
                                          std::vector<occ_t>(), kappa);
        for (long b = 0; b < nBurnin; ++b) sampler();
        const long s0 = sampler.stats.nSamples();
        for (long n = 0; n < nSamples; ++n) {
            ModelState<Agent> endState(sampler(), T, T);
            for (int k = 0; k < Agent::domainSize(); ++k) mcmc[k] += endState[k];
        }
        steps = sampler.stats.nSamples() - s0;
    }
    RejectionSampler<Trajectory<Agent>> rejectionSampler(problem.prior, problem.likelihood);
    for (long n = 0; n < nRejection; ++n) {
        ModelState<Agent> endState(rejectionSampler(), T, T);
        for (int k = 0; k < Agent::domainSize(); ++k) rej[k] += endState[k];
    }
    double rms = 0.0;
    for (int k = 0; k < Agent::domainSize(); ++k) {
        mcmc[k] /= nSamples;
        rej[k] /= nRejection;
        rms += (mcmc[k] - rej[k]) * (mcmc[k] - rej[k]);
    }
    rms = sqrt(rms / Agent::domainSize());
    AbmpFile o;
    o.f64["mcmc_endstate_mean"] = mcmc;
    o.f64["rejection_endstate_mean"] = rej;
    o.f64["counts"] = {double(nSamples), double(nBurnin), double(nRejection), double(steps), rms};
    o.write(margOut);
    fprintf(stderr, "single-observation experiment: %ld samples (%ld Metropolis steps), %ld rejection samples, RMS error %.5f\n",
            nSamples, steps, nRejection, rms);
    return 0;
}

// The reference's validation experiments on its OTHER agent types (Experiments.cpp:21-47, 70-125): BinomialAgent<3> with one noisy
// observation (closed-form answer :32), CatMouseAgent with one observation, CatMouseAgent assimilation of a random trajectory
// (checked against the exhaustive ExactSolver, :113-124).  Writes the problem as an ABMP file whose agent is described by tables
// (flattenAgentImportance), together with the answers the experiment compares against -- end-state means of the reference's
// RejectionSampler and, where the support is enumerable, the exact posterior over trajectories -- and a golden step trace of the
// reference sampler started, as in the experiments, from the empty trajectory (SparseBasisSampler(posterior, kappa)).
#include "agents/BinomialAgent.h"
#include "agents/CatMouseAgent.h"
#include "BernoulliStartState.h"
#include "ExactSolver.h"
#include "test_agents.h"
template <class Agent>
static int toyExperiment(Prior<Agent> &prior, Likelihood<Agent> &likelihood, int T, double kappa, int chainSeed, long nSteps,
                         int every, long nRejection, const std::string &problemOut, const std::string &traceOut) {
    auto posterior = likelihood * prior;
    std::unique_ptr<TableauNormMinimiser<occ_t>> tableau;
    {
        SilenceStdout quiet;
        tableau.reset(new TableauNormMinimiser<occ_t>(posterior.constraints));   // the constructor runs findMinimalBasis()
    }
    AbmpFile f;
    abmcmc::flattenTableau<TableauNormMinimiser<occ_t>, occ_t>(f, *tableau, posterior.factors, kappa);
    abmcmc::flattenAgentImportance<Agent>(f, T);
    const int S = Agent::domainSize();
    std::vector<double> rej(S, 0.0), exact(S, 0.0), exactP;
    std::vector<int8_t> exactTraj;
    if (nRejection > 0) {
        RejectionSampler<Trajectory<Agent>> rejectionSampler(prior, likelihood);
        for (long n = 0; n < nRejection; ++n) {
            ModelState<Agent> endState(rejectionSampler(), T, T);
            for (int k = 0; k < S; ++k) rej[k] += endState[k];
        }
        for (double &v : rej) v /= double(nRejection);
    }
    if (nRejection > 0) {   // (the exhaustive enumeration takes a minute on the larger toy problems: fixtures only)
        ExactSolver<Agent> solver(posterior);
        for (const auto &tp : solver.pmf) {
            ModelState<Agent> endState(tp.first, T, T);
            for (int k = 0; k < S; ++k) exact[k] += tp.second * endState[k];
            exactTraj.insert(exactTraj.end(), tp.first.begin(), tp.first.end());
            exactP.push_back(tp.second);
        }
    }
    f.f64["rejection_endstate_mean"] = rej;
    f.f64["exact_endstate_mean"] = exact;      // over the binary trajectories ExactSolver enumerates (BinarySolutionSet)
    f.i8["exact_traj"] = exactTraj;            // [n][nEvents]
    f.f64["exact_p"] = exactP;
    f.f64["counts"] = {double(nRejection)};
    f.write(problemOut);
    const auto &d = f.I("dims");
    fprintf(stderr, "toy: tableau %d x %d, nNonBasic=%d nnz=%d, %d agent classes, %zu trajectories in the exact support\n", d[0], d[1], d[3],
            d[4], f.I("ag_meta")[3], exactP.size());
    return traceSampler<Agent>(*tableau, posterior.factors, posterior.perturbableFunctionFactory(), std::vector<occ_t>(), kappa, T,
                               [](const ModelState<Agent> &) { return std::valarray<double>(); }, 0, chainSeed, nSteps, every, traceOut);
}

static int cmdToy(const std::string &kind, int seed, int chainSeed, long nSteps, int every, long nRejection, const std::string &problemOut,
                  const std::string &traceOut) {
    Random::gen.seed(seed);
    if (kind == "binomial") {          // Experiments::BinomialAgentSingleObservation (:21-33)
        constexpr int T = 2;
        typedef BinomialAgent<T + 1> Agent;
        Prior<Agent> prior(T, BernoulliStartState<Agent>({1.0, 0.1, 0.0}));
        Likelihood<Agent> likelihood(AgentStateObservation<Agent>(State<Agent>(1, 0), 1, 0.9));
        return toyExperiment<Agent>(prior, likelihood, T, 1.0, chainSeed, nSteps, every, nRejection, problemOut, traceOut);
    }
    if (kind == "catmouse") {          // Experiments::CatMouseSingleObservation (:36-46)
        constexpr int T = 2;
        typedef CatMouseAgent Agent;
        Prior<Agent> prior(T, BernoulliStartState<Agent>({0.9, 0.1, 0.1, 0.9}));
        Likelihood<Agent> likelihood(AgentStateObservation<Agent>(State<Agent>(1, Agent(Agent::CAT, Agent::LEFT)), 1, 1.0));
        return toyExperiment<Agent>(prior, likelihood, T, 1.75, chainSeed, nSteps, every, nRejection, problemOut, traceOut);
    }
    if (kind == "catmouse-assim") {    // Experiments::CatMouseAssimilation (:70-125)
        constexpr int T = 2;
        typedef CatMouseAgent Agent;
        Prior<Agent> prior(T, PoissonStartState<Agent>({0.5, 0.5, 0.3, 0.3}));
        Trajectory<Agent> realTrajectory = prior.nextSample(true);
        Likelihood<Agent> likelihood(realTrajectory, 0.2, 1.0);
        return toyExperiment<Agent>(prior, likelihood, T, 1.25, chainSeed, nSteps, every, nRejection, problemOut, traceOut);
    }
    if (kind == "ring") {              // own test agent with non-trivial importance factors (oracle/ref/test_agents.h)
        constexpr int T = 3;
        typedef RingAgent<5> Agent;
        Prior<Agent> prior(T, BernoulliStartState<Agent>({0.5, 0.3, 0.5, 0.3, 0.5}));
        Likelihood<Agent> likelihood(AgentStateObservation<Agent>(State<Agent>(1, 2), 1, 1.0));
        return toyExperiment<Agent>(prior, likelihood, T, 1.5, chainSeed, nSteps, every, nRejection, problemOut, traceOut);
    }
    if (kind == "crowded-ring") {      // an agent the tables cannot describe: must be refused by flattenAgentImportance
        constexpr int T = 2;
        typedef CrowdedRingAgent<4> Agent;
        Prior<Agent> prior(T, BernoulliStartState<Agent>({0.5, 0.3, 0.5, 0.3}));
        Likelihood<Agent> likelihood(AgentStateObservation<Agent>(State<Agent>(1, 2), 1, 1.0));
        return toyExperiment<Agent>(prior, likelihood, T, 1.5, chainSeed, nSteps, every, nRejection, problemOut, traceOut);
    }
    fprintf(stderr, "toy: unknown experiment %s\n", kind.c_str());
    return 2;
}

// A PredPrey problem re-described as a table agent: pp_meta dropped, the ag_* tables of flattenAgentImportance<PredPreyAgent<G>>
// (found by probing timestep()) in its place.  The oracle and the GPU library must produce the very same chain from either
// description -- the check that the table path means what the geometry path means, with PredPrey's non-trivial log-ratios.
template <int G>
int cmdAsTable(AbmpFile f, const std::string &out) {
    const int T = f.I("pp_meta")[1];
    f.i32.erase("pp_meta");
    f.f64.erase("pp_actpmf");
    abmcmc::flattenAgentImportance<PredPreyAgent<G>>(f, T);
    f.write(out);
    fprintf(stderr, "astable: %d agent states, %d classes\n", f.I("ag_meta")[0], f.I("ag_meta")[3]);
    return 0;
}

#define DISPATCH_G(G_, CALL)                                   \
    switch (G_) {                                              \
        case 3: { constexpr int G = 3; return CALL; }          \
        case 4: { constexpr int G = 4; return CALL; }          \
        case 8: { constexpr int G = 8; return CALL; }          \
        case 16: { constexpr int G = 16; return CALL; }        \
        case 32: { constexpr int G = 32; return CALL; }        \
        case 64: { constexpr int G = 64; return CALL; }        \
        default: fprintf(stderr, "unsupported grid size %d\n", G_); return 2; \
    }

static int run(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr,
                "usage:\n"
                "  abm_ref gen G T seed kappa out.abmp [pPred pPrey pMakeObs pObsIfPresent [bernoulli|poisson [fastbasis]]]\n"
                "  abm_ref init in.abmp nChains seed out.i8\n"
                "  abm_ref trace in.abmp initSeed chainSeed nSteps checkpointEvery out.abmp\n"
                "  abm_ref selfcheck in.abmp nSamples\n"
                "  abm_ref time in.abmp nThreads nSamplesPerThread [maxSeconds]\n"
                "  abm_ref bench in.abmp nThreads samplesPerStep warmup steps\n"
                "  abm_ref marginals in.abmp nThreads nSamples burnin out.abmp\n"
                "  abm_ref series in.abmp nThreads nSamples burnin out.abmp\n"
                "  abm_ref posterior in.abmp nThreads nSamples burnin batch thin out.abmp [seedBase]\n"
                "  abm_ref refstats series.abmp [splitHalves]\n"
                "  abm_ref constraints in.abmp out.abmp [rebuild: 1 = also time the reference TableauNormMinimiser on them]\n"
                "  abm_ref astable in.abmp out.abmp\n"
                "  abm_ref toy binomial|catmouse|catmouse-assim|ring|crowded-ring seed chainSeed nSteps checkpointEvery nRejection problem.abmp trace.abmp\n");
        return 2;
    }
    std::string cmd = argv[1];
    if (cmd == "actpmf") {   // the act pmf table of PredPreyAgent::timestep by (type, has-neighbour, act); grid independent
        AbmpFile f;
        abmcmc::flattenPredPreyImportance<PredPreyAgent<8>>(f, 8, 1);
        for (double v : f.D("pp_actpmf")) printf("%.17g\n", v);
        return 0;
    }
    if (cmd == "gen") {
        int g = atoi(argv[2]), T = atoi(argv[3]), seed = atoi(argv[4]);
        double kappa = atof(argv[5]);
        std::string out = argv[6];
        double pPred = argc > 7 ? atof(argv[7]) : 0.05, pPrey = argc > 8 ? atof(argv[8]) : 0.05;
        double pObs = argc > 9 ? atof(argv[9]) : 0.05, pObsIf = argc > 10 ? atof(argv[10]) : 1.0;
        const int startKind = argc > 11 && std::string(argv[11]) == "poisson" ? 1 : 0;
        const bool fastBasis = argc > 12 && std::string(argv[12]) == "fastbasis";
        DISPATCH_G(g, cmdGen<G>(T, seed, kappa, out, pPred, pPrey, pObs, pObsIf, startKind, fastBasis));
    }
    if (cmd == "toy")   // toy binomial|catmouse|catmouse-assim|ring|crowded-ring seed chainSeed nSteps checkpointEvery nRejection problem.abmp trace.abmp
        return cmdToy(argv[2], atoi(argv[3]), atoi(argv[4]), atol(argv[5]), atoi(argv[6]), atol(argv[7]), argv[8], argv[9]);
    if (cmd == "experiment3")   // experiment3 problem.abmp marginals.abmp nSamples nBurnin nRejection seed
        return cmdSingleObservationExperiment(argv[2], argv[3], atol(argv[4]), atol(argv[5]), atol(argv[6]), atoi(argv[7]));
    AbmpFile f = AbmpFile::read(argv[2]);
    if (cmd == "refstats") return cmdRefStats(f, argc > 3 ? atoi(argv[3]) : 0);   // refstats series.abmp [split each sequence in halves]
    int g = f.I("pp_meta")[0];
    if (cmd == "astable") { DISPATCH_G(g, cmdAsTable<G>(f, argv[3])); }   // astable in.abmp out.abmp
    if (cmd == "init") { DISPATCH_G(g, cmdInit<G>(f, atoi(argv[3]), atoi(argv[4]), argv[5])); }
    if (cmd == "trace") { DISPATCH_G(g, cmdTrace<G>(f, atoi(argv[3]), atoi(argv[4]), atol(argv[5]), atoi(argv[6]), argv[7])); }
    if (cmd == "selfcheck") { DISPATCH_G(g, cmdSelfcheck<G>(f, atoi(argv[3]))); }
    if (cmd == "time") { DISPATCH_G(g, cmdTime<G>(f, atoi(argv[3]), atol(argv[4]), argc > 5 ? atof(argv[5]) : 0.0)); }
    if (cmd == "bench") { DISPATCH_G(g, cmdBench<G>(f, atoi(argv[3]), atol(argv[4]), atoi(argv[5]), atoi(argv[6]))); }
    if (cmd == "marginals") { DISPATCH_G(g, cmdMarginals<G>(f, atoi(argv[3]), atol(argv[4]), atol(argv[5]), argv[6])); }
    if (cmd == "constraints") { DISPATCH_G(g, cmdConstraints<G>(f, argv[3], argc > 4 ? atoi(argv[4]) : 0)); }
    if (cmd == "posterior") {   // posterior in.abmp nThreads nSamples burnin batch thin out.abmp [seedBase]
        DISPATCH_G(g, cmdPosterior<G>(f, atoi(argv[3]), atol(argv[4]), atol(argv[5]), atol(argv[6]), atol(argv[7]), argv[8], argc > 9 ? atoi(argv[9]) : 0));
    }
    if (cmd == "series") { DISPATCH_G(g, cmdSeries<G>(f, atoi(argv[3]), atol(argv[4]), atol(argv[5]), argv[6])); }
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}

int main(int argc, char **argv) {
    try {
        return run(argc, argv);
    } catch (const std::exception &e) {
        fprintf(stderr, "abm_ref: %s\n", e.what());
        return 1;
    } catch (const char *e) {
        fprintf(stderr, "abm_ref: %s\n", e);
        return 1;
    }
}
