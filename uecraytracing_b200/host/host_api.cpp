// host_api.cpp -- include/uecrt_host.h implemented with the yk:: host headers.
#include <cstring>
#include <string>

#include "../../include/uecrt_host.h"
#include "../csrc/philox.cuh"
#include "png_writer.hpp"
#include "yk/render.hpp"
#include "yk/scene_io.hpp"
#include "yk/scenes.hpp"

namespace {
thread_local std::string g_host_err;
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
const char* uecrt_host_last_error(void) { return g_host_err.c_str(); }
int32_t uecrt_host_scene_file_count(const char* path) {
  try {
    return (int32_t)yk::scene_io::load(path ? path : "").world.size();
  } catch (const std::exception& e) {
    g_host_err = e.what();
    return -1;
  }
}
int uecrt_host_scene_file_fill(const char* path, uint32_t cap, uecrt_float4* geom, uint32_t* mat_type, uecrt_float4* mat_param,
                               uecrt_camera* cam_out, int* has_camera_out) {
  try {
    const yk::scene_io::scene_file f = yk::scene_io::load(path ? path : "");
    const auto& w = f.world;
    if (w.size() > cap) { g_host_err = "buffer too small"; return UECRT_ERR_INVALID; }
    if (w.size()) {
      std::memcpy(geom, w.geom.data(), w.size() * sizeof(uecrt_float4));
      std::memcpy(mat_type, w.mat_type.data(), w.size() * sizeof(uint32_t));
      std::memcpy(mat_param, w.mat_param.data(), w.size() * sizeof(uecrt_float4));
    }
    if (cam_out) *cam_out = f.cam.flatten();
    if (has_camera_out) *has_camera_out = f.has_camera ? 1 : 0;
    return UECRT_OK;
  } catch (const std::exception& e) {
    g_host_err = e.what();
    return UECRT_ERR_INVALID;
  }
}
int uecrt_host_scene_file_write(const char* path, uint32_t n, const uecrt_float4* geom, const uint32_t* mat_type,
                                const uecrt_float4* mat_param) {
  yk::hittable_list<double> w;
  w.geom.assign(geom, geom + n); w.mat_type.assign(mat_type, mat_type + n); w.mat_param.assign(mat_param, mat_param + n);
  std::ofstream f(path ? path : "");
  if (!f) { g_host_err = "cannot open for writing"; return UECRT_ERR_INVALID; }
  yk::scene_io::write(f, w);
  return f.good() ? UECRT_OK : UECRT_ERR_INVALID;
}
int uecrt_host_write_pfm(const char* path, uint32_t width, uint32_t height, const float* accum4, uint32_t spp) {
  return uecrt_host::write_pfm(path, (int)width, (int)height, accum4, spp) ? UECRT_OK : UECRT_ERR_INVALID;
}
uint32_t uecrt_host_image_height(uint32_t width) { return yk::constants::height_for(width); }
int uecrt_host_write_png(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb) {
  return uecrt_host::write_png(path, (int)width, (int)height, 3, rgb, (int)(3 * width)) ? UECRT_OK : UECRT_ERR_INVALID;
}
}
