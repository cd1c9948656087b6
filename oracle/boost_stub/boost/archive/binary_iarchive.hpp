// Stand-in: the oracle harness never reads boost archives (it uses its own flat problem format).
#pragma once
#include "../serialization/access.hpp"
namespace boost { namespace archive {
struct binary_iarchive {
    template<class S> explicit binary_iarchive(S &) {}
    template<class T> binary_iarchive &operator>>(T &) { throw std::runtime_error("boost stub: binary_iarchive unavailable"); }
    template<class T> binary_iarchive &operator&(T &) { throw std::runtime_error("boost stub: binary_iarchive unavailable"); }
};
} }
