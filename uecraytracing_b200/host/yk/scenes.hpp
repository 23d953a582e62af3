// yk/scenes.hpp -- scene descriptions that feed hittable_list (the step BEFORE the hot path).
//   reference_scene() : the 4 spheres hard-coded in render() (source.cpp:103-112), verbatim
//   final_scene(seed) : RTIOW "final" random scene (absent upstream; book listing)
//   dense_scene(n,seed): BASELINE config 4, n random small spheres on a r=1000 ground
// Random numbers: std::mt19937 + std::uniform_real_distribution<double>, which the reference's
// yk::mt19937 / yk::uniform_real_distribution reproduce bit for bit (random.hpp:43-287).
#pragma once
#include <cmath>
#include <random>

#include "camera.hpp"
#include "hittable_list.hpp"

namespace yk::scenes {

using color = color3d;
using point = pos3<double, world_tag>;

inline hittable_list<double> reference_scene() {
  return hittable_list<double>{}
      .add(sphere(point(0, 0, -1), 0.5, lambertian<double>({0.7, 0.3, 0.3})))
      .add(sphere(point(0, -100.5, -1), 100.0, lambertian<double>({0.8, 0.8, 0.0})))
      .add(sphere(point(-1.0, 0.0, -1.0), 0.5, metal<double>({0.8, 0.8, 0.8})))
      .add(sphere(point(1.0, 0.0, -1.0), 0.5, metal<double>({0.8, 0.6, 0.2})));
}
// BASELINE config[0] as worded there ("lambertian/metal/dielectric"): id2 becomes glass
inline hittable_list<double> reference_scene_dielectric() {
  return hittable_list<double>{}
      .add(sphere(point(0, 0, -1), 0.5, lambertian<double>({0.7, 0.3, 0.3})))
      .add(sphere(point(0, -100.5, -1), 100.0, lambertian<double>({0.8, 0.8, 0.0})))
      .add(sphere(point(-1.0, 0.0, -1.0), 0.5, dielectric<double>(1.5)))
      .add(sphere(point(1.0, 0.0, -1.0), 0.5, metal<double>({0.8, 0.6, 0.2})));
}

inline hittable_list<double> final_scene(std::uint32_t seed) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<double> d01(0.0, 1.0);
  auto rnd = [&]() { return d01(gen); };
  auto rnd_in = [&](double lo, double hi) { return lo + (hi - lo) * d01(gen); };
  hittable_list<double> world;
  world.push(sphere(point(0, -1000, 0), 1000.0, lambertian<double>({0.5, 0.5, 0.5})));
  for (int a = -11; a < 11; ++a)
    for (int b = -11; b < 11; ++b) {
      const double choose_mat = rnd();
      const double cx = a + 0.9 * rnd();
      const double cz = b + 0.9 * rnd();
      const point center(cx, 0.2, cz);
      if (vec3<double>(center - point(4, 0.2, 0)).length() > 0.9) {
        if (choose_mat < 0.8) {
          const double r1 = rnd(), g1 = rnd(), b1 = rnd(), r2 = rnd(), g2 = rnd(), b2 = rnd();
          world.push(sphere(center, 0.2, lambertian<double>({r1 * r2, g1 * g2, b1 * b2})));
        } else if (choose_mat < 0.95) {
          const double r = rnd_in(0.5, 1), g = rnd_in(0.5, 1), bb = rnd_in(0.5, 1);
          const double fuzz = rnd_in(0, 0.5);
          world.push(sphere(center, 0.2, metal<double>({r, g, bb}, fuzz)));
        } else {
          world.push(sphere(center, 0.2, dielectric<double>(1.5)));
        }
      }
    }
  world.push(sphere(point(0, 1, 0), 1.0, dielectric<double>(1.5)));
  world.push(sphere(point(-4, 1, 0), 1.0, lambertian<double>({0.4, 0.2, 0.1})));
  world.push(sphere(point(4, 1, 0), 1.0, metal<double>({0.7, 0.6, 0.5}, 0.0)));
  return world;
}

inline hittable_list<double> dense_scene(std::uint32_t n, std::uint32_t seed) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<double> d01(0.0, 1.0);
  auto rnd = [&]() { return d01(gen); };
  hittable_list<double> world;
  if (n == 0) return world;
  world.push(sphere(point(0, -1000, 0), 1000.0, lambertian<double>({0.5, 0.5, 0.5})));
  for (std::uint32_t i = 1; i < n; ++i) {
    const double r = 0.1 + 0.3 * rnd();
    const point c(-50.0 + 100.0 * rnd(), r, -50.0 + 100.0 * rnd());
    const double m = rnd();
    if (m < 0.80) {
      const double cr = rnd() * rnd(), cg = rnd() * rnd(), cb = rnd() * rnd();
      world.push(sphere(c, r, lambertian<double>({cr, cg, cb})));
    } else if (m < 0.95) {
      const double cr = 0.5 + 0.5 * rnd(), cg = 0.5 + 0.5 * rnd(), cb = 0.5 + 0.5 * rnd();
      world.push(sphere(c, r, metal<double>({cr, cg, cb}, 0.5 * rnd())));
    } else {
      world.push(sphere(c, r, dielectric<double>(1.5)));
    }
  }
  return world;
}

inline camera<double> reference_camera() { return camera<double>{}; }
// RTIOW final-scene camera at the reference's fixed 16:9 aspect
inline camera<double> final_camera() {
  return camera<double>(point(13, 2, 3), point(0, 0, 0), vec3<double>(0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0);
}
inline camera<double> dense_camera() {
  return camera<double>(point(40, 8, 12), point(0, 0, 0), vec3<double>(0, 1, 0), 20.0, 16.0 / 9.0, 0.1, std::sqrt(40.0 * 40 + 8 * 8 + 12 * 12));
}

}  // namespace yk::scenes
