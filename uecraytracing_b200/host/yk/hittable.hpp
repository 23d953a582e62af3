// yk/hittable.hpp -- host mirror of hit_record<T> and the CRTP hittable_interface (yk/hittable.hpp:16-47).
// hit() / scatter() forward to the derived type's hit_impl / scatter_impl exactly as upstream; the derived
// types of this mirror (sphere, hittable_list) run both on the GPU through the C ABI.
#pragma once
#include <cstddef>
#include <optional>
#include <random>
#include <type_traits>
#include <utility>

#include "color.hpp"
#include "concepts.hpp"
#include "ray.hpp"
#include "vec3.hpp"

namespace yk {

template <concepts::arithmetic T>
struct hit_record {
  pos3<T, world_tag> p;
  vec3<T> normal;
  T t{};
  std::size_t id{};
  bool front_face{};
  constexpr void set_face_normal(const ray<T>& r, const vec3<T>& outward_normal) {   // hittable.hpp:23-27
    front_face = dot(r.direction, outward_normal) < 0;
    normal = front_face ? outward_normal : -outward_normal;
  }
};

template <concepts::arithmetic T, class Derived>
struct hittable_interface {
  std::optional<hit_record<T>> hit(const ray<T>& r, T t_min, T t_max) const {
    return static_cast<const Derived*>(this)->hit_impl(r, t_min, t_max);
  }
  // optional of { color3<U> attenuation, ray<T> scattered }; nullopt = absorbed
  template <concepts::arithmetic U, std::uniform_random_bit_generator Gen>
  std::optional<std::pair<color3<U>, ray<T>>> scatter(const ray<T>& r, const hit_record<T>& rec, Gen& gen) const {
    return static_cast<const Derived*>(this)->template scatter_impl<U>(r, rec, gen);
  }
};

namespace concepts {
template <class H, class T>
concept hittable = arithmetic<T> && std::is_base_of_v<hittable_interface<T, H>, H>;
}  // namespace concepts

}  // namespace yk
