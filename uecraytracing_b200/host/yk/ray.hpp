// yk/ray.hpp -- host mirror of ray<T> (yk/ray.hpp:11-24): origin, direction, at(scalar), and the
// deduction guide that lets `ray(rec.p, direction)` name ray<T> as the reference's materials do.
#pragma once
#include "concepts.hpp"
#include "vec3.hpp"

namespace yk {

template <concepts::arithmetic T>
struct ray {
  pos3<T, world_tag> origin;
  vec3<T> direction;
  template <concepts::arithmetic U>
  constexpr pos3<T, world_tag> at(U scalar) const { return origin + direction * static_cast<T>(scalar); }
};
template <concepts::arithmetic T>
ray(pos3<T, world_tag>, vec3<T>) -> ray<T>;

}  // namespace yk
