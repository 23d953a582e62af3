// yk/random.hpp -- the reference re-implements mt19937 / uniform_real_distribution as constexpr code
// (yk/random.hpp:43-287) that is bit-identical to libstdc++'s (SURVEY.md probe C.7); the host mirror names the
// standard ones.  canonical(gen) is the one draw every reference material makes: generate_canonical<double, 53>,
// two engine outputs per double (random.hpp:161-183).
#pragma once
#include <random>
namespace yk {
using mt19937 = std::mt19937;
template <class T> using uniform_real_distribution = std::uniform_real_distribution<T>;
template <class Gen>
inline double canonical(Gen& gen) { return std::generate_canonical<double, 53>(gen); }
}  // namespace yk
