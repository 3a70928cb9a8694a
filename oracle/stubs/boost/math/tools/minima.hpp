// TEST INFRASTRUCTURE ONLY, NOT Boost.  brent_find_minima is only reached with optimize_hyper = TRUE,
// which none of the reference's R callers pass (R/signal.R, R/difference.R) and which is outside the
// hot path; the stub refuses instead of guessing Boost's exact iteration.
#ifndef ORACLE_STUB_BOOST_MINIMA
#define ORACLE_STUB_BOOST_MINIMA
#include <stdexcept>
#include <utility>
#include <boost/cstdint.hpp>
namespace boost { namespace math { namespace tools {
template <class F, class T>
std::pair<T, T> brent_find_minima(F, T, T, int, boost::uintmax_t&) {
  throw std::runtime_error("stub boost: brent_find_minima (optimize_hyper) is not available in the pinning build");
}
}}}
#endif
