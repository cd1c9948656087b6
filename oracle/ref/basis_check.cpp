// basis_check: the reference's host code building its posterior constraints, then the basis twice -- once with the
// UNMODIFIED TableauNormMinimiser, once with abmcmc::minimiseBasis (include/abmcmc_b200/basis_adaptor.hpp over the C-ABI)
// -- and a member-by-member comparison of the two objects.  TEST INFRASTRUCTURE ONLY; compiled against the reference
// headers where they lie (oracle/Makefile `basis`), nothing of the reference is copied.
//     basis_check G T [poisson]
#include <boost/serialization/access.hpp>
#include <chrono>

#include "Trajectory.h"
#include "Forecast.h"
#include "Likelihood.h"
#include "PredPreyProblem.h"
#include "SparseBasisSampler.h"

#include "abmcmc_b200/basis_adaptor.hpp"

typedef ABM::occupation_type occ_t;

template <int G>
int run(int T, bool poisson) {
    typedef PredPreyAgent<G> Agent;
    Random::gen.seed(1234);
    auto rate = [](Agent agent) { return 0.05; };
    Prior<Agent> prior = poisson ? Prior<Agent>(T, PoissonStartState<Agent>(rate)) : Prior<Agent>(T, BernoulliStartState<Agent>(rate));
    Trajectory<Agent> real = prior.nextSample(true);
    Likelihood<Agent> likelihood(real, 0.05, 1.0);
    WeightedFactoredConvexDistribution<occ_t> posterior = likelihood * prior;

    std::streambuf *old = std::cout.rdbuf(nullptr);   // the reference prints norms from its constructor
    auto t0 = std::chrono::steady_clock::now();
    TableauNormMinimiser<occ_t> ref(posterior.constraints);
    auto t1 = std::chrono::steady_clock::now();
    std::cout.rdbuf(old);
    TableauNormMinimiser<occ_t> own = abmcmc::minimiseBasis(posterior.constraints);
    auto t2 = std::chrono::steady_clock::now();

    bool same = ref.rows.size() == own.rows.size() && ref.cols.size() == own.cols.size() && ref.nAuxiliaryVars == own.nAuxiliaryVars &&
                ref.basis == own.basis && ref.F == own.F && ref.L == own.L && ref.U == own.U && ref.nNonBasic() == own.nNonBasic();
    for (size_t i = 0; same && i < ref.rows.size(); ++i)
        same = static_cast<const std::map<int, occ_t> &>(ref.rows[i]) == static_cast<const std::map<int, occ_t> &>(own.rows[i]) &&
               ref.rows[i].isActive == own.rows[i].isActive;
    for (size_t j = 0; same && j < ref.cols.size(); ++j)
        same = static_cast<const std::set<int> &>(ref.cols[j]) == static_cast<const std::set<int> &>(own.cols[j]) &&
               ref.cols[j].isBasic == own.cols[j].isBasic;
    printf("%s: G=%d T=%d %s start, %zu x %zu tableau, reference %.3fs, abm_basis %.3fs\n", same ? "OK" : "FAIL", G, T,
           poisson ? "Poisson" : "Bernoulli", ref.rows.size(), ref.cols.size(), std::chrono::duration<double>(t1 - t0).count(),
           std::chrono::duration<double>(t2 - t1).count());
    return same ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: basis_check G T [poisson]\n"); return 2; }
    const int g = atoi(argv[1]), T = atoi(argv[2]);
    const bool poisson = argc > 3 && std::string(argv[3]) == "poisson";
    try {
        switch (g) {
            case 4: return run<4>(T, poisson);
            case 8: return run<8>(T, poisson);
            case 16: return run<16>(T, poisson);
            case 32: return run<32>(T, poisson);
            default: fprintf(stderr, "unsupported grid size\n"); return 2;
        }
    } catch (const std::exception &e) {
        printf("FAIL: %s\n", e.what());
        return 1;
    } catch (const char *e) {
        printf("FAIL: %s\n", e);
        return 1;
    }
}
