// Agent types that exist only to test the table description of an agent (include/abmcmc_b200/reference_flatten.hpp
// flattenAgentImportance) against the UNMODIFIED reference machinery (Prior, Likelihood, TrajectoryImportance<AGENT>,
// SparseBasisSampler) with importance factors that are not trivial.  TEST INFRASTRUCTURE ONLY; they implement the reference's
// agent concept (agents/*.h: domainSize, actDomainSize, timestep, marginalTimestep, consequences, constraints, neighbours).
#pragma once
#include <ostream>
#include <vector>

#include "ABM.h"
#include "Constraint.h"
#include "ModelState.h"

// N sites on a ring, acts {STAY, STEP to the next site}.  How likely an agent steps depends on whether the site ahead is
// occupied and on the parity of its own site: two classes of log-ratio rows, all of them non-zero, one influencer per state.
template <int N>
class RingAgent {
public:
    typedef int Act;
    int stateId;

    static constexpr int domainSize() { return N; }
    static constexpr int actDomainSize() { return 2; }

    RingAgent(int id) : stateId(id) {}
    operator int() const { return stateId; }
    int ahead() const { return (stateId + 1) % N; }

    std::vector<double> timestep(const ModelState<RingAgent<N>> &others) const {
        const bool blocked = others[ahead()] >= 1;
        if (stateId % 2 == 0) return blocked ? std::vector<double>{0.8, 0.2} : std::vector<double>{0.3, 0.7};
        return blocked ? std::vector<double>{0.6, 0.4} : std::vector<double>{0.1, 0.9};
    }
    double marginalTimestep(Act act) const { return act == 0 ? 0.45 : 0.55; }
    std::vector<RingAgent<N>> consequences(Act act) const { return {RingAgent<N>(act == 1 ? ahead() : stateId)}; }
    std::vector<Constraint<ABM::occupation_type>> constraints(int, Act) const { return {}; }
    // the site behind reads this one: its factor goes stale when this state changes (TrajectoryImportance.h:48-51)
    std::vector<RingAgent<N>> neighbours() { return {RingAgent<N>((stateId + N - 1) % N)}; }

    friend std::ostream &operator<<(std::ostream &out, const RingAgent<N> &agent) { return out << "ring:" << agent.stateId; }
};

// The same ring, but stepping depends on whether at least TWO agents sit on the site ahead: not of the form the tables
// describe (presence of an influencer); flattenAgentImportance must refuse it.
template <int N>
class CrowdedRingAgent : public RingAgent<N> {
public:
    using RingAgent<N>::RingAgent;
    std::vector<double> timestep(const ModelState<CrowdedRingAgent<N>> &others) const {
        return others[this->ahead()] >= 2 ? std::vector<double>{0.8, 0.2} : std::vector<double>{0.3, 0.7};
    }
    std::vector<CrowdedRingAgent<N>> consequences(typename RingAgent<N>::Act act) const {
        return {CrowdedRingAgent<N>(act == 1 ? this->ahead() : this->stateId)};
    }
    std::vector<CrowdedRingAgent<N>> neighbours() { return {CrowdedRingAgent<N>((this->stateId + N - 1) % N)}; }
};
