#pragma once
#include "access.hpp"
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
