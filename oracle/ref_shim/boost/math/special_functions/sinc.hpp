// Stand-in for boost/math/special_functions/sinc.hpp (Boost is not installed in the build container) -- TEST
// INFRASTRUCTURE ONLY, used to compile the unmodified reference sources into oracle/_ref.
// sinc_pi(x) = sin(x)/x with the Taylor expansion near zero, as Boost.Math documents it: below the fourth root of
// epsilon the series 1 - x^2/6 + x^4/120 (terms added only while they are above epsilon).
#pragma once
#include <cmath>
#include <limits>
namespace boost { namespace math {
template <typename T> inline T sinc_pi(const T x) {
  const T eps = std::numeric_limits<T>::epsilon();
  const T t2 = std::sqrt(eps), t4 = std::sqrt(t2);
  if (std::fabs(x) >= t4) return std::sin(x) / x;
  T r = T(1);
  if (std::fabs(x) >= eps) {
    const T x2 = x * x;
    r -= x2 / T(6);
    if (std::fabs(x) >= t2) r += (x2 * x2) / T(120);
  }
  return r;
}
}}  // namespace boost::math
