// Stand-in for boost/math/special_functions/sign.hpp -- TEST INFRASTRUCTURE ONLY (see sinc.hpp).
#pragma once
#include <cmath>
namespace boost { namespace math {
template <typename T> inline int sign(const T& z) { return (z == 0) ? 0 : (std::signbit(z) ? -1 : 1); }
template <typename T> inline int signbit(const T& z) { return std::signbit(z) ? 1 : 0; }
template <typename T> inline T copysign(const T& x, const T& y) { return std::copysign(x, y); }
}}  // namespace boost::math
