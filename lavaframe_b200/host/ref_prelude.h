// Force-included in front of the reference's unchanged host sources: split_bvh.cpp:79,104-105 calls an
// unqualified isnan(), which MSVC provides globally and libstdc++ does not.
#pragma once
#include <cmath>
using std::isnan;
