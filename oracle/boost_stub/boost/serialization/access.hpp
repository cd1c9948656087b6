// Minimal stand-in for <boost/serialization/access.hpp>.
// TEST INFRASTRUCTURE ONLY: lets the unmodified reference headers under /root/reference/src
// compile in a container without Boost.  It also pulls in the std headers the reference
// receives transitively from the real Boost headers.
#pragma once
#include <climits>
#include <cmath>
#include <cfloat>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <memory>
#include <functional>
#include <valarray>
#include <limits>
#include <stdexcept>
#include <string>
namespace boost { namespace serialization { class access {}; } }
