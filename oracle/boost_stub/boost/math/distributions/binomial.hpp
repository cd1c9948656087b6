// Stand-in for Boost.Math's binomial pdf (used by AgentStateObservation.h:47 only when
// pObserveIfPresent < 1).  Restated through lgamma; NOT bit-identical to Boost.Math
// ("parity unpinned" at this boundary).  Exact for p in {0,1}.
#pragma once
#include <cmath>
namespace boost { namespace math {
struct binomial { double n, p; binomial(double n_, double p_) : n(n_), p(p_) {} };
inline double pdf(const binomial &d, double k) {
    if (k < 0 || k > d.n) return 0.0;
    if (d.p == 0.0) return k == 0 ? 1.0 : 0.0;
    if (d.p == 1.0) return k == d.n ? 1.0 : 0.0;
    return std::exp(std::lgamma(d.n + 1) - std::lgamma(k + 1) - std::lgamma(d.n - k + 1)
                    + k * std::log(d.p) + (d.n - k) * std::log1p(-d.p));
}
} }
