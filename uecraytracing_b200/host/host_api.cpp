// host_api.cpp -- include/uecrt_host.h implemented with the yk:: host headers.
#include <cstring>
#include <string>

#include "../../include/uecrt_host.h"
#include "../csrc/philox.cuh"
#include "png_writer.hpp"
#include "yk/render.hpp"
#include "yk/scenes.hpp"

namespace {
bool build(const char* name, uint32_t n_hint, uint32_t seed, yk::hittable_list<double>& out) {
  const std::string s = name ? name : "";
  if (s == "ref4") out = yk::scenes::reference_scene();
  else if (s == "ref4d") out = yk::scenes::reference_scene_dielectric();
  else if (s == "final") out = yk::scenes::final_scene(seed);
  else if (s == "dense") out = yk::scenes::dense_scene(n_hint, seed);
  else return false;
  return true;
}
}  // namespace

extern "C" {
uint32_t uecrt_host_scene_count(const char* name, uint32_t n_hint, uint32_t seed) {
  yk::hittable_list<double> w;
  return build(name, n_hint, seed, w) ? (uint32_t)w.size() : 0u;
}
int uecrt_host_scene_fill(const char* name, uint32_t n_hint, uint32_t seed, uint32_t cap, uecrt_float4* geom,
                          uint32_t* mat_type, uecrt_float4* mat_param) {
  yk::hittable_list<double> w;
  if (!build(name, n_hint, seed, w)) return UECRT_ERR_INVALID;
  if (w.size() > cap || !geom || !mat_type || !mat_param) return UECRT_ERR_INVALID;
  std::memcpy(geom, w.geom.data(), w.size() * sizeof(uecrt_float4));
  std::memcpy(mat_type, w.mat_type.data(), w.size() * sizeof(uint32_t));
  std::memcpy(mat_param, w.mat_param.data(), w.size() * sizeof(uecrt_float4));
  return UECRT_OK;
}
int uecrt_host_camera(const char* name, uecrt_camera* out) {
  const std::string s = name ? name : "";
  if (!out) return UECRT_ERR_INVALID;
  if (s == "ref") *out = yk::scenes::reference_camera().flatten();
  else if (s == "final") *out = yk::scenes::final_camera().flatten();
  else if (s == "dense") *out = yk::scenes::dense_camera().flatten();
  else return UECRT_ERR_INVALID;
  return UECRT_OK;
}
int uecrt_host_camera_lookat(const double from[3], const double at[3], const double vup[3], double vfov_deg, double aspect,
                             double aperture, double focus_dist, uecrt_camera* out) {
  if (!from || !at || !vup || !out) return UECRT_ERR_INVALID;
  using P = yk::pos3<double, yk::world_tag>;
  *out = yk::camera<double>(P(from[0], from[1], from[2]), P(at[0], at[1], at[2]), yk::vec3<double>(vup[0], vup[1], vup[2]),
                            vfov_deg, aspect, aperture, focus_dist).flatten();
  return UECRT_OK;
}
void uecrt_host_philox(uint64_t seed, const uint32_t c[4], uint32_t out[4]) {
  const philox_out r = philox4x32_10(c[0], c[1], c[2], c[3], (uint32_t)(seed & 0xffffffffull), (uint32_t)(seed >> 32));
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
uint32_t uecrt_host_image_height(uint32_t width) { return yk::constants::height_for(width); }
int uecrt_host_write_png(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb) {
  return uecrt_host::write_png(path, (int)width, (int)height, 3, rgb, (int)(3 * width)) ? UECRT_OK : UECRT_ERR_INVALID;
}
}
