// yk/material.hpp -- host material types with the reference's spellings (yk/material.hpp:44-76:
// lambertian<U>(albedo), metal<U>(albedo)), plus the RTIOW extensions north_star names: metal fuzz and
// dielectric(index_of_refraction).  hittable_list::add flattens them to {type, albedo, param}.
// scatter(r_in, rec, gen) has the reference's signature and draws from `gen` exactly what the reference
// draws (material.hpp:33-36, 50-59: three canonical doubles per lambertian bounce, none for metal), but the
// arithmetic runs on the GPU: the uniforms go to uecrt_scatter (csrc/materials.cuh), the same device function
// K1 shades with.  There is no host scatter code in the product.
#pragma once
#include <cstdint>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../../include/uecrt.h"
#include "color.hpp"
#include "hittable.hpp"
#include "random.hpp"
#include "ray.hpp"

namespace yk {
namespace concepts {
template <class T>
concept material = true;     // material.hpp:22-23: any type is accepted
}  // namespace concepts

namespace detail {
// one scatter event on the device with caller-supplied uniforms in [0, 1)
template <class U, class T>
inline std::optional<std::pair<color3<U>, ray<T>>> gpu_scatter(std::uint32_t kind, const color3<U>& albedo, U param, const ray<T>& r_in,
                                                               const hit_record<T>& rec, const float (&un)[4]) {
  const uecrt_float4 mp{(float)albedo.r, (float)albedo.g, (float)albedo.b, (float)param};
  const float d[3] = {(float)r_in.direction.x, (float)r_in.direction.y, (float)r_in.direction.z};
  const float p[3] = {(float)rec.p.x, (float)rec.p.y, (float)rec.p.z};
  const float n[3] = {(float)rec.normal.x, (float)rec.normal.y, (float)rec.normal.z};
  const std::uint8_t ff = rec.front_face ? 1 : 0;
  std::uint8_t ok = 0;
  float att[3], oo[3], dd[3];
  if (uecrt_scatter(1, &kind, &mp, d, p, n, &ff, un, &ok, att, oo, dd) != UECRT_OK)
    throw std::runtime_error(std::string("uecrt_scatter: ") + uecrt_last_error());
  if (!ok) return std::nullopt;
  return std::make_pair(color3<U>((U)att[0], (U)att[1], (U)att[2]), ray<T>{rec.p, vec3<T>((T)dd[0], (T)dd[1], (T)dd[2])});
}
}  // namespace detail

template <concepts::arithmetic U>
struct lambertian {
  color3<U> albedo;
  constexpr lambertian(const color3<U>& a) : albedo(a) {}
  static constexpr std::uint32_t kind = UECRT_MAT_LAMBERTIAN;
  constexpr U param() const { return U(0); }
  template <concepts::arithmetic T, std::uniform_random_bit_generator Gen>
  std::optional<std::pair<color3<U>, ray<T>>> scatter(const ray<T>& r_in, const hit_record<T>& rec, Gen& gen) const {
    float un[4] = {0, 0, 0, 0};
    for (int i = 0; i < 3; ++i) un[i] = (float)canonical(gen);     // vec3::random(gen, -1, 1): x, y, z (vec3.hpp:134-142)
    return detail::gpu_scatter<U, T>(kind, albedo, param(), r_in, rec, un);
  }
};

template <concepts::arithmetic U>
struct metal {
  color3<U> albedo;
  U fuzz;
  constexpr metal(const color3<U>& a, U f = U(0)) : albedo(a), fuzz(f < U(1) ? f : U(1)) {}
  static constexpr std::uint32_t kind = UECRT_MAT_METAL;
  constexpr U param() const { return fuzz; }
  template <concepts::arithmetic T, std::uniform_random_bit_generator Gen>
  std::optional<std::pair<color3<U>, ray<T>>> scatter(const ray<T>& r_in, const hit_record<T>& rec, Gen& gen) const {
    float un[4] = {0, 0, 0, 0};
    if (fuzz > U(0))                                               // the reference's metal draws nothing (material.hpp:67-75)
      for (int i = 0; i < 4; ++i) un[i] = (float)canonical(gen);  // random_in_unit_sphere: direction x, y, z, then the radius
    return detail::gpu_scatter<U, T>(kind, albedo, param(), r_in, rec, un);
  }
};

template <concepts::arithmetic U>
struct dielectric {
  color3<U> albedo{U(1), U(1), U(1)};
  U ir;
  constexpr explicit dielectric(U index_of_refraction) : ir(index_of_refraction) {}
  static constexpr std::uint32_t kind = UECRT_MAT_DIELECTRIC;
  constexpr U param() const { return ir; }
  template <concepts::arithmetic T, std::uniform_random_bit_generator Gen>
  std::optional<std::pair<color3<U>, ray<T>>> scatter(const ray<T>& r_in, const hit_record<T>& rec, Gen& gen) const {
    const float un[4] = {(float)canonical(gen), 0, 0, 0};          // one draw: reflect or refract (Schlick)
    return detail::gpu_scatter<U, T>(kind, albedo, param(), r_in, rec, un);
  }
};
}  // namespace yk
