// yk/hittable_list.hpp, yk/sphere.hpp, yk/hittable.hpp mirror -- host scene container.
//
// The reference's hittable_list<T, Hs...> is a compile-time std::tuple whose add() returns a
// NEW type (yk/hittable_list.hpp:18-30); 488 elements take 11 minutes to compile and 10k is
// infeasible.  This mirror keeps the call shape -- hittable_list<T>{}.add(sphere(
// pos3<T,world_tag>(..), radius, material)).add(..) -- but is runtime-sized: add() appends the
// flattened record {centre, radius, material kind, albedo, param} in insertion order, which IS
// hit_record::id (hittable_list.hpp:41) and fixes the tie rule.  hit() forwards single rays to
// the GPU closest-hit kernel (uecrt_primary_hits, exact mode); there is no host intersection
// code in the product.
#pragma once
#include <cstddef>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/uecrt.h"
#include "material.hpp"
#include "vec3.hpp"

namespace yk {

template <class T>
struct hit_record {   // yk/hittable.hpp:16-28
  pos3<T, world_tag> p;
  vec3<T> normal;
  T t{};
  std::size_t id{};
  bool front_face{};
};

template <class T, class M>
struct sphere {       // yk/sphere.hpp:16-23
  pos3<T, world_tag> center;
  T radius;
  M material;
  constexpr sphere(pos3<T, world_tag> c, T r, M m) : center(c), radius(r), material(m) {}
};
template <class T, class M> sphere(pos3<T, world_tag>, T, M) -> sphere<T, M>;

template <class T>
struct hittable_list {
  std::vector<uecrt_float4> geom;
  std::vector<std::uint32_t> mat_type;
  std::vector<uecrt_float4> mat_param;

  template <class M, class... More>
  [[nodiscard]] hittable_list add(const sphere<T, M>& s, const More&... more) const {
    hittable_list out = *this;
    out.push(s);
    (out.push(more), ...);
    return out;
  }
  template <class M>
  void push(const sphere<T, M>& s) {
    geom.push_back({(float)s.center.x, (float)s.center.y, (float)s.center.z, (float)s.radius});
    mat_type.push_back(M::kind);
    mat_param.push_back({(float)s.material.albedo.r, (float)s.material.albedo.g, (float)s.material.albedo.b, (float)s.material.param()});
  }
  std::size_t size() const { return geom.size(); }
  uecrt_scene flatten() const {
    return uecrt_scene{(std::uint32_t)geom.size(), geom.data(), mat_type.data(), mat_param.data()};
  }

  // hittable_interface::hit (yk/hittable.hpp:32-34) for one ray, on the GPU, FP64-resolved
  std::optional<hit_record<T>> hit(const ray<T>& r, T t_min, T t_max) const {
    const float o[3] = {(float)r.origin.x, (float)r.origin.y, (float)r.origin.z};
    const float d[3] = {(float)r.direction.x, (float)r.direction.y, (float)r.direction.z};
    double t = 0, n[3] = {0, 0, 0};
    std::int32_t id = -1;
    std::uint8_t ff = 0;
    const uecrt_scene sc = flatten();
    if (uecrt_primary_hits(&sc, 1, o, d, (double)t_min, (double)t_max, 1, &t, n, &id, &ff) != UECRT_OK)
      throw std::runtime_error(std::string("uecrt_primary_hits: ") + uecrt_last_error());
    if (id < 0) return std::nullopt;
    hit_record<T> rec;
    rec.t = (T)t;
    rec.p = r.at(rec.t);
    rec.normal = vec3<T>((T)n[0], (T)n[1], (T)n[2]);
    rec.id = (std::size_t)id;
    rec.front_face = ff != 0;
    return rec;
  }
};

}  // namespace yk
