// yk/vec3.hpp -- host mirror of the reference's vec3/pos3 value types (yk/vec3.hpp:14-207):
// same spellings (vec3<T,Tag>, pos3, world_tag/default_tag, dot/cross/reflect, length,
// length_squared, normalized, near_zero) so scene-building code written against the
// reference compiles unchanged.  Host only; the device works on flattened float arrays.
#pragma once
#include <cmath>
#include <type_traits>

namespace yk {

struct tag_base {};
struct default_tag : tag_base {};
struct world_tag : default_tag {};

template <class T, class Tag = default_tag>
struct vec3 {
  T x{}, y{}, z{};
  constexpr vec3() = default;
  constexpr vec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  template <class Tag2>
  constexpr vec3(const vec3<T, Tag2>& o) : x(o.x), y(o.y), z(o.z) {}   // tags are phantom: freely convertible

  template <class U>
  constexpr vec3<U, Tag> to() const { return {static_cast<U>(x), static_cast<U>(y), static_cast<U>(z)}; }

  // the reference tests x twice and never z (vec3.hpp:76-80); the quirk is part of the behaviour
  constexpr bool near_zero() const {
    constexpr double s = 1e-8;
    auto f = [](T v) { return v > 0 ? v : -v; };
    return (f(x) < s) && (f(y) < s) && (f(x) < s);
  }
  constexpr bool operator==(const vec3&) const = default;
  constexpr vec3 operator-() const { return {-x, -y, -z}; }
  constexpr vec3& operator+=(const vec3& r) { x += r.x; y += r.y; z += r.z; return *this; }
  constexpr vec3& operator-=(const vec3& r) { x -= r.x; y -= r.y; z -= r.z; return *this; }
  constexpr vec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
  constexpr vec3& operator/=(T s) { x /= s; y /= s; z /= s; return *this; }
  constexpr T dot(const vec3& r) const { return x * r.x + y * r.y + z * r.z; }
  constexpr vec3 cross(const vec3& r) const { return {y * r.z - z * r.y, z * r.x - x * r.z, x * r.y - y * r.x}; }
  T length() const { return std::sqrt(x * x + y * y + z * z); }
  constexpr T length_squared() const { return x * x + y * y + z * z; }
  vec3& normalize() { return *this /= length(); }
  vec3 normalized() const { return vec3(*this) /= length(); }
};

template <class T, class A, class B> constexpr vec3<T, A> operator+(const vec3<T, A>& l, const vec3<T, B>& r) { return {l.x + r.x, l.y + r.y, l.z + r.z}; }
template <class T, class A, class B> constexpr vec3<T, A> operator-(const vec3<T, A>& l, const vec3<T, B>& r) { return {l.x - r.x, l.y - r.y, l.z - r.z}; }
template <class T, class A> constexpr vec3<T, A> operator*(const vec3<T, A>& v, T s) { return {v.x * s, v.y * s, v.z * s}; }
template <class T, class A> constexpr vec3<T, A> operator*(T s, const vec3<T, A>& v) { return v * s; }
template <class T, class A> constexpr vec3<T, A> operator/(const vec3<T, A>& v, T s) { return {v.x / s, v.y / s, v.z / s}; }
template <class T, class A, class B> constexpr T dot(const vec3<T, A>& l, const vec3<T, B>& r) { return l.x * r.x + l.y * r.y + l.z * r.z; }
template <class T, class A> constexpr vec3<T, A> cross(const vec3<T, A>& l, const vec3<T, A>& r) { return l.cross(r); }
template <class T, class A> constexpr vec3<T, A> reflect(const vec3<T, A>& v, const vec3<T, A>& n) { return v - n * (2 * dot(v, n)); }

using vec3d = vec3<double>;
template <class T, class Tag> using pos3 = vec3<T, Tag>;

}  // namespace yk
