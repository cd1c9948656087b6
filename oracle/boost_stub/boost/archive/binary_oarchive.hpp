#pragma once
#include "../serialization/access.hpp"
namespace boost { namespace archive {
struct binary_oarchive {
    template<class S> explicit binary_oarchive(S &) {}
    template<class T> binary_oarchive &operator<<(const T &) { throw std::runtime_error("boost stub: binary_oarchive unavailable"); }
    template<class T> binary_oarchive &operator&(const T &) { throw std::runtime_error("boost stub: binary_oarchive unavailable"); }
};
} }
