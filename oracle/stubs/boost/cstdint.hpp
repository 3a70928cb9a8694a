// TEST INFRASTRUCTURE ONLY, NOT Boost: the one typedef BetaBinomialRateDistribution.hpp names.
#ifndef ORACLE_STUB_BOOST_CSTDINT
#define ORACLE_STUB_BOOST_CSTDINT
#include <cstdint>
namespace boost { typedef std::uintmax_t uintmax_t; }
#endif
