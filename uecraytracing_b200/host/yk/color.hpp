// yk/color.hpp -- host mirror of color3<T> (yk/color.hpp:12-120)
#pragma once
#include <algorithm>
#include <cstdint>

namespace yk {
template <class T>
struct color3 {
  using value_type = T;
  T r{}, g{}, b{};
  constexpr color3() = default;
  constexpr color3(T r_, T g_, T b_) : r(r_), g(g_), b(b_) {}
  template <class U> constexpr color3<U> to() const { return {static_cast<U>(r), static_cast<U>(g), static_cast<U>(b)}; }
  constexpr color3 clamped(T lo, T hi) const { return {std::clamp(r, lo, hi), std::clamp(g, lo, hi), std::clamp(b, lo, hi)}; }
  constexpr bool operator==(const color3&) const = default;
  constexpr color3& operator+=(const color3& o) { r += o.r; g += o.g; b += o.b; return *this; }
  constexpr color3& operator*=(const color3& o) { r *= o.r; g *= o.g; b *= o.b; return *this; }
};
template <class T> constexpr color3<T> operator+(const color3<T>& a, const color3<T>& b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }
template <class T> constexpr color3<T> operator*(const color3<T>& a, const color3<T>& b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }
template <class T> constexpr color3<T> operator*(const color3<T>& a, T s) { return {a.r * s, a.g * s, a.b * s}; }
template <class T> constexpr color3<T> operator*(T s, const color3<T>& a) { return a * s; }
template <class T> constexpr color3<T> operator/(const color3<T>& a, T s) { return {a.r / s, a.g / s, a.b / s}; }
using color3b = color3<std::uint8_t>;
using color3d = color3<double>;
}  // namespace yk
