// yk/hittable_list.hpp -- host scene container with the reference's call shape.
//
// The reference's hittable_list<T, Hs...> is a compile-time std::tuple whose add() returns a
// NEW type (yk/hittable_list.hpp:18-30); 488 elements take 11 minutes to compile and 10k is
// infeasible.  This mirror keeps the call shape -- hittable_list<T>{}.add(sphere(
// pos3<T,world_tag>(..), radius, material)).add(..) -- but is runtime-sized: add() appends the
// flattened record {centre, radius, material kind, albedo, param} in insertion order, which IS
// hit_record::id (hittable_list.hpp:41) and fixes the tie rule.
//   hit(r, t_min, t_max)        one ray  } closest hit on the GPU (uecrt_ctx_primary_hits, FP64-resolved):
//   hit(rays, t_min, t_max)     a batch  } the scene is uploaded ONCE into a context the list caches
//   scatter<U>(r, rec, gen)     dispatch on rec.id to that sphere's material (hittable_list.hpp:60-73), whose
//                               scatter draws from gen like the reference and computes on the GPU (material.hpp)
// There is no host intersection or scatter arithmetic in the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/uecrt.h"
#include "hittable.hpp"
#include "material.hpp"
#include "sphere.hpp"
#include "vec3.hpp"

namespace yk {

template <concepts::arithmetic T>
struct hittable_list : public hittable_interface<T, hittable_list<T>> {
  std::vector<uecrt_float4> geom;
  std::vector<std::uint32_t> mat_type;
  std::vector<uecrt_float4> mat_param;

  template <class M, class... More>
  [[nodiscard]] hittable_list add(const sphere<T, M>& s, const More&... more) const {
    hittable_list out;
    out.geom = geom; out.mat_type = mat_type; out.mat_param = mat_param;      // a new list: no device context is shared
    out.push(s);
    (out.push(more), ...);
    return out;
  }
  template <class M>
  void push(const sphere<T, M>& s) {
    geom.push_back({(float)s.center.x, (float)s.center.y, (float)s.center.z, (float)s.radius});
    mat_type.push_back(M::kind);
    mat_param.push_back({(float)s.material.albedo.r, (float)s.material.albedo.g, (float)s.material.albedo.b, (float)s.material.param()});
    ctx_.reset();
  }
  std::size_t size() const { return geom.size(); }
  uecrt_scene flatten() const {
    return uecrt_scene{(std::uint32_t)geom.size(), geom.data(), mat_type.data(), mat_param.data()};
  }

  // hittable_interface::hit (yk/hittable.hpp:32-34) for a batch of rays: ONE kernel launch
  std::vector<std::optional<hit_record<T>>> hit(const std::vector<ray<T>>& rays, T t_min, T t_max) const {
    const std::size_t n = rays.size();
    std::vector<std::optional<hit_record<T>>> out(n);
    if (n == 0) return out;
    std::vector<float> o(3 * n), d(3 * n);
    for (std::size_t i = 0; i < n; ++i) {
      o[3 * i] = (float)rays[i].origin.x; o[3 * i + 1] = (float)rays[i].origin.y; o[3 * i + 2] = (float)rays[i].origin.z;
      d[3 * i] = (float)rays[i].direction.x; d[3 * i + 1] = (float)rays[i].direction.y; d[3 * i + 2] = (float)rays[i].direction.z;
    }
    std::vector<double> t(n), nr(3 * n);
    std::vector<std::int32_t> id(n);
    std::vector<std::uint8_t> ff(n);
    if (uecrt_ctx_primary_hits(context(), n, o.data(), d.data(), (double)t_min, (double)t_max, 1, t.data(), nr.data(), id.data(), ff.data()) != UECRT_OK)
      throw std::runtime_error(std::string("uecrt_ctx_primary_hits: ") + uecrt_last_error());
    for (std::size_t i = 0; i < n; ++i) {
      if (id[i] < 0) continue;
      hit_record<T> rec;
      rec.t = (T)t[i];
      rec.p = rays[i].at(rec.t);
      rec.normal = vec3<T>((T)nr[3 * i], (T)nr[3 * i + 1], (T)nr[3 * i + 2]);
      rec.id = (std::size_t)id[i];
      rec.front_face = ff[i] != 0;
      out[i] = rec;
    }
    return out;
  }
  std::optional<hit_record<T>> hit_impl(const ray<T>& r, T t_min, T t_max) const { return hit(std::vector<ray<T>>{r}, t_min, t_max)[0]; }
  using hittable_interface<T, hittable_list<T>>::hit;

  // hittable_list::scatter (yk/hittable_list.hpp:60-73): the material of sphere rec.id scatters
  template <concepts::arithmetic U, std::uniform_random_bit_generator Gen>
  std::optional<std::pair<color3<U>, ray<T>>> scatter(const ray<T>& r, const hit_record<T>& rec, Gen& gen) const {
    if (rec.id >= geom.size()) return std::nullopt;
    const uecrt_float4 mp = mat_param[rec.id];
    const color3<U> albedo((U)mp.x, (U)mp.y, (U)mp.z);
    switch (mat_type[rec.id]) {
      case UECRT_MAT_LAMBERTIAN: return lambertian<U>(albedo).template scatter<T>(r, rec, gen);
      case UECRT_MAT_METAL: return metal<U>(albedo, (U)mp.w).template scatter<T>(r, rec, gen);
      default: return dielectric<U>((U)mp.w).template scatter<T>(r, rec, gen);
    }
  }
  template <concepts::arithmetic U, std::uniform_random_bit_generator Gen>
  std::optional<std::pair<color3<U>, ray<T>>> scatter_impl(const ray<T>& r, const hit_record<T>& rec, Gen& gen) const {
    return scatter<U>(r, rec, gen);
  }

 private:
  struct ctx_deleter { void operator()(uecrt_ctx* c) const { uecrt_ctx_destroy(c); } };
  mutable std::shared_ptr<uecrt_ctx> ctx_;     // the scene on the device, created by the first hit()
  uecrt_ctx* context() const {
    if (!ctx_) {
      uecrt_ctx* c = nullptr;
      const uecrt_scene sc = flatten();
      if (uecrt_ctx_create(&sc, nullptr, &c) != UECRT_OK) throw std::runtime_error(std::string("uecrt_ctx_create: ") + uecrt_last_error());
      ctx_ = std::shared_ptr<uecrt_ctx>(c, ctx_deleter{});
    }
    return ctx_.get();
  }
};

}  // namespace yk
