// yk/sphere.hpp -- host mirror of sphere<T, M> (yk/sphere.hpp:16-58): centre, radius, material, the deduction
// guide, and the hittable_interface hooks.  hit_impl sends the ray to the GPU closest-hit kernel as a one-sphere
// scene (uecrt_primary_hits, FP64-resolved in the reference's operation order); scatter_impl forwards to the
// material like upstream.  Scenes are normally built by hittable_list::add, which flattens spheres into arrays.
#pragma once
#include <optional>
#include <stdexcept>
#include <string>

#include "../../../include/uecrt.h"
#include "hittable.hpp"
#include "material.hpp"

namespace yk {

template <concepts::arithmetic T, concepts::material M>
struct sphere : public hittable_interface<T, sphere<T, M>> {
  pos3<T, world_tag> center;
  T radius;
  M material;
  constexpr sphere(pos3<T, world_tag> c, T r, M m) : center(c), radius(r), material(m) {}

  std::optional<hit_record<T>> hit_impl(const ray<T>& r, T t_min, T t_max) const {
    const uecrt_float4 g{(float)center.x, (float)center.y, (float)center.z, (float)radius};
    const std::uint32_t mt = M::kind;
    const uecrt_float4 mp{(float)material.albedo.r, (float)material.albedo.g, (float)material.albedo.b, (float)material.param()};
    const uecrt_scene sc{1, &g, &mt, &mp};
    const float o[3] = {(float)r.origin.x, (float)r.origin.y, (float)r.origin.z};
    const float d[3] = {(float)r.direction.x, (float)r.direction.y, (float)r.direction.z};
    double t = 0, n[3] = {0, 0, 0};
    std::int32_t id = -1;
    std::uint8_t ff = 0;
    if (uecrt_primary_hits(&sc, 1, o, d, (double)t_min, (double)t_max, 1, &t, n, &id, &ff) != UECRT_OK)
      throw std::runtime_error(std::string("uecrt_primary_hits: ") + uecrt_last_error());
    if (id < 0) return std::nullopt;
    hit_record<T> rec;
    rec.t = (T)t;
    rec.p = r.at(rec.t);
    rec.normal = vec3<T>((T)n[0], (T)n[1], (T)n[2]);
    rec.front_face = ff != 0;
    return rec;
  }
  template <concepts::arithmetic U, std::uniform_random_bit_generator Gen>
  std::optional<std::pair<color3<U>, ray<T>>> scatter_impl(const ray<T>& r, const hit_record<T>& rec, Gen& gen) const {
    return material.template scatter<T>(r, rec, gen);
  }
};
template <concepts::arithmetic T, concepts::material M>
sphere(pos3<T, world_tag>, T, M) -> sphere<T, M>;

}  // namespace yk
