// yk/material.hpp -- host material descriptors with the reference's spellings
// (yk/material.hpp:44-76: lambertian<U>(albedo), metal<U>(albedo)), plus the RTIOW
// extensions north_star names: metal fuzz and dielectric(index_of_refraction).
// scatter() itself runs on the GPU (csrc/materials.cuh); these types only carry parameters
// into hittable_list::add, which flattens them to {type, albedo, param}.
#pragma once
#include "../../../include/uecrt.h"
#include "color.hpp"

namespace yk {
template <class U>
struct lambertian {
  color3<U> albedo;
  constexpr lambertian(const color3<U>& a) : albedo(a) {}
  static constexpr std::uint32_t kind = UECRT_MAT_LAMBERTIAN;
  constexpr U param() const { return U(0); }
};
template <class U>
struct metal {
  color3<U> albedo;
  U fuzz;
  constexpr metal(const color3<U>& a, U f = U(0)) : albedo(a), fuzz(f < U(1) ? f : U(1)) {}
  static constexpr std::uint32_t kind = UECRT_MAT_METAL;
  constexpr U param() const { return fuzz; }
};
template <class U>
struct dielectric {
  color3<U> albedo{U(1), U(1), U(1)};
  U ir;
  constexpr explicit dielectric(U index_of_refraction) : ir(index_of_refraction) {}
  static constexpr std::uint32_t kind = UECRT_MAT_DIELECTRIC;
  constexpr U param() const { return ir; }
};
}  // namespace yk
