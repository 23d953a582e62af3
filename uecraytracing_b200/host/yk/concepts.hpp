// yk/concepts.hpp -- host mirror of concepts::arithmetic (yk/concepts.hpp:10-11)
#pragma once
#include <type_traits>
namespace yk::concepts {
template <class T>
concept arithmetic = std::is_arithmetic_v<T>;
}  // namespace yk::concepts
