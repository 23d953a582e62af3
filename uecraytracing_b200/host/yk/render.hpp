// yk/render.hpp -- re-hosted yk::render() (source.cpp:98-178).  Same role and the same four
// progress lines; the pixel x spp loop, ray_color, hit and scatter run in libuecrt.so (K1-K3).
// Width / spp / depth default to the reference's compile-time macros YK_IMAGE_WIDTH / YK_SPP /
// YK_MAX_DEPTH (source.cpp:43-53) and become runtime values, so image_t is a heap vector
// instead of a stack std::array (source.cpp:70-71 overflows the stack at 3840x2160).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <functional>
#include <iomanip>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/uecrt.h"
#include "camera.hpp"
#include "color.hpp"
#include "config.hpp"
#include "hittable_list.hpp"

#ifndef YK_IMAGE_WIDTH
#define YK_IMAGE_WIDTH 400
#endif
#ifndef YK_SPP
#define YK_SPP 100
#endif
#ifndef YK_MAX_DEPTH
#define YK_MAX_DEPTH 50
#endif

namespace yk {

namespace constants {
constexpr double aspect_ratio = 16.0 / 9.0;                      // source.cpp:57
constexpr std::uint32_t image_width = YK_IMAGE_WIDTH;
constexpr std::uint32_t samples_per_pixel = YK_SPP;
constexpr std::uint32_t max_depth = YK_MAX_DEPTH;
constexpr std::uint32_t height_for(std::uint32_t width) { return static_cast<std::uint32_t>(width / aspect_ratio); }  // source.cpp:59-62
}  // namespace constants

struct render_config {
  std::uint32_t width = constants::image_width;
  std::uint32_t spp = constants::samples_per_pixel;
  std::uint32_t max_depth = constants::max_depth;
  std::uint64_t seed = 0;
  int gpus = 1;
  bool exact_hits = false;
  int scan = UECRT_SCAN_AUTO;           // candidate scan of K1: tensor cores from 96 spheres up, else the FP32 pipe
};

struct image_t {
  std::uint32_t width = 0, height = 0;
  std::vector<color3b> pixels;          // row-major, top row first (source.cpp:170-171)
  std::vector<float> accum;             // optional W*H*4 linear sums
  uecrt_stats stats{};
  const std::uint8_t* data() const { return reinterpret_cast<const std::uint8_t*>(pixels.data()); }
};

// The reference's verbose levels 1 and 2 (source.cpp:128-152): one "(row,col) : (y,x)" line per pixel and, at level 2,
// one "(row,col,sam) : (y,x,s)" line per sample, in loop order, with its field widths.  They carry no result, so the
// re-hosted render() prints them up front (the pixels are no longer visited one at a time).  Level 3's per-ray lines
// (raytracer.hpp:21-25) exist on the host mirror's ray_color only: K1's rays live on the device.
inline void print_progress_lines(std::uint32_t width, std::uint32_t height, std::uint32_t spp) {
  if (!verbose) return;
  const int wy = (int)std::ceil(std::log10((double)height)) - 1, wx = (int)std::ceil(std::log10((double)width)) - 1,
            ws = (int)std::ceil(std::log10((double)spp)) - 1;
  for (std::uint32_t y = 0; y < height; ++y)
    for (std::uint32_t x = 0; x < width; ++x) {
      std::cout << "(row,col) : (" << std::setw(wy) << y << ',' << std::setw(wx) << x << ")\n";
      if (verbose > 1)
        for (std::uint32_t k = 0; k < spp; ++k)
          std::cout << "(row,col,sam) : (" << std::setw(wy) << y << ',' << std::setw(wx) << x << ',' << std::setw(ws) << k << ")\n";
    }
  std::cout.flush();
}

inline image_t render(const hittable_list<double>& world, const camera<double>& cam, const render_config& cfg, bool keep_accum = false) {
  image_t img;
  img.width = cfg.width;
  img.height = constants::height_for(cfg.width);
  img.pixels.resize((std::size_t)img.width * img.height);
  if (keep_accum) img.accum.resize((std::size_t)img.width * img.height * 4);
  std::cout << "rendering..." << std::endl;                      // source.cpp:114
  print_progress_lines(img.width, img.height, cfg.spp);
  const uecrt_scene sc = world.flatten();
  const uecrt_camera c = cam.flatten();
  uecrt_params p{};
  p.width = img.width; p.height = img.height; p.spp = cfg.spp; p.max_depth = cfg.max_depth;
  p.seed = cfg.seed; p.num_gpus = cfg.gpus; p.flags = (cfg.exact_hits ? UECRT_FLAG_EXACT_HITS : 0u) |
      (cfg.scan == UECRT_SCAN_FP32 ? UECRT_FLAG_SCAN_FP32 : cfg.scan == UECRT_SCAN_TENSOR ? UECRT_FLAG_SCAN_TENSOR
       : cfg.scan == UECRT_SCAN_GRID ? UECRT_FLAG_SCAN_GRID : 0u);
  static_assert(sizeof(color3b) == 3, "color3b must be 3 packed bytes");
  if (uecrt_render(&sc, &c, &p, reinterpret_cast<std::uint8_t*>(img.pixels.data()), keep_accum ? img.accum.data() : nullptr, &img.stats) != UECRT_OK)
    throw std::runtime_error(std::string("uecrt_render: ") + uecrt_last_error());
  if (verbose || stats) {
    const double samples = (double)img.width * img.height * cfg.spp;
    std::cout << "K1 render_paths (" << (img.stats.scan_mode == UECRT_SCAN_TENSOR ? "tensor-core scan" : img.stats.scan_mode == UECRT_SCAN_GRID ? "uniform grid" : "FP32-pipe scan") << ") "
              << img.stats.kernel_ms << " ms, reduce " << img.stats.reduce_ms << " ms, K3 resolve "
              << img.stats.resolve_ms << " ms; " << samples / (img.stats.kernel_ms * 1e-3) << " samples/s; rays "
              << img.stats.rays_traced << ", sphere tests " << img.stats.sphere_tests << std::endl;
  }
  std::cout << "rendering finished" << std::endl;                 // source.cpp:175
  return img;
}

// The camera turned by `degrees` about the world's y axis (every frame of an orbit sees the same scene from the
// same height and distance): all of its points and vectors are rotated, so any camera<T> works, positioned or not.
template <class T>
inline camera<T> orbit_y(const camera<T>& c, T degrees) {
  const T a = degrees * T(3.14159265358979323846) / T(180), cs = std::cos(a), sn = std::sin(a);
  auto rot = [&](auto& p) { const T x = p.x, z = p.z; p.x = cs * x + sn * z; p.z = -sn * x + cs * z; };
  camera<T> r = c;
  rot(r.origin); rot(r.lower_left_corner); rot(r.horizontal); rot(r.vertical); rot(r.u); rot(r.v);
  return r;
}

// A sequence of frames of one scene through the frame queue (include/uecrt.h uecrt_queue_*): K1 of frame k + 1 runs while
// the RGB image of frame k is copied into pinned host memory and handed to on_frame (the caller's PNG write), the overlap
// SURVEY 8(f)2 asks for after source.cpp:221-230.  on_frame(k, rgb, k1_ms) is called in frame order on the calling thread
// with the queue's pinned buffer, valid until it returns.  Returns the wall-clock seconds of the whole sequence.
inline double render_sequence(const hittable_list<double>& world, const std::vector<camera<double>>& cams, const render_config& cfg,
                              const std::function<void(std::size_t, const std::uint8_t*, float)>& on_frame) {
  if (cams.empty()) return 0.0;
  const std::uint32_t width = cfg.width, height = constants::height_for(cfg.width);
  const uecrt_scene sc = world.flatten();
  const uecrt_camera c0 = cams[0].flatten();
  uecrt_queue* q = nullptr;
  if (uecrt_queue_create(&sc, &c0, width, height, &q) != UECRT_OK) throw std::runtime_error(std::string("uecrt_queue_create: ") + uecrt_last_error());
  uecrt_params p{};
  p.width = width; p.height = height; p.spp = cfg.spp; p.max_depth = cfg.max_depth; p.num_gpus = 1;
  p.flags = (cfg.exact_hits ? UECRT_FLAG_EXACT_HITS : 0u) |
      (cfg.scan == UECRT_SCAN_FP32 ? UECRT_FLAG_SCAN_FP32 : cfg.scan == UECRT_SCAN_TENSOR ? UECRT_FLAG_SCAN_TENSOR
       : cfg.scan == UECRT_SCAN_GRID ? UECRT_FLAG_SCAN_GRID : 0u);
  auto submit = [&](std::size_t k) {
    const uecrt_camera c = cams[k].flatten();
    p.seed = cfg.seed + k;                          // independent Philox streams per frame
    if (uecrt_queue_submit(q, &c, &p, k) != UECRT_OK) { uecrt_queue_destroy(q); throw std::runtime_error(std::string("uecrt_queue_submit: ") + uecrt_last_error()); }
  };
  const auto t0 = std::chrono::steady_clock::now();
  submit(0);
  for (std::size_t k = 0; k < cams.size(); ++k) {
    if (k + 1 < cams.size()) submit(k + 1);
    const std::uint8_t* rgb = nullptr;
    std::uint64_t id = 0;
    float k1_ms = 0.f;
    if (uecrt_queue_next(q, &rgb, &id, &k1_ms) != UECRT_OK) { uecrt_queue_destroy(q); throw std::runtime_error(std::string("uecrt_queue_next: ") + uecrt_last_error()); }
    on_frame((std::size_t)id, rgb, k1_ms);
  }
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uecrt_queue_destroy(q);
  return s;
}

}  // namespace yk
