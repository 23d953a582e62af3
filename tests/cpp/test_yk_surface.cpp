// test_yk_surface.cpp -- the re-hosted yk:: C++ surface used the way the reference uses its own
// (source.cpp:100-112, 221): same spellings, same call shapes.  `--cpu-only` stops before the first
// call that needs the GPU (hit / render go through libuecrt.so; there is no host intersection code).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>

#include "yk/camera.hpp"
#include "yk/color.hpp"
#include "yk/config.hpp"
#include "yk/hittable_list.hpp"
#include "yk/material.hpp"
#include "yk/render.hpp"
#include "yk/scenes.hpp"

#define CHECK(c)                                                              \
  do {                                                                        \
    if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  using namespace yk;
  using T = double;
  using color = color3d;
  const bool cpu_only = argc > 1 && std::strcmp(argv[1], "--cpu-only") == 0;

  // ---- scene construction, verbatim from source.cpp:100-112
  const camera<T> cam = {};
  const auto world =
      hittable_list<T>{}
          .add(sphere(pos3<T, world_tag>(0, 0, -1), 0.5, lambertian<color::value_type>({0.7, 0.3, 0.3})))
          .add(sphere(pos3<T, world_tag>(0, -100.5, -1), 100.0, lambertian<color::value_type>({0.8, 0.8, 0.0})))
          .add(sphere(pos3<T, world_tag>(-1.0, 0.0, -1.0), 0.5, metal<color::value_type>({0.8, 0.8, 0.8})))
          .add(sphere(pos3<T, world_tag>(1.0, 0.0, -1.0), 0.5, metal<color::value_type>({0.8, 0.6, 0.2})));
  CHECK(world.size() == 4);
  const uecrt_scene sc = world.flatten();
  CHECK(sc.n == 4 && sc.geom[1].y == -100.5f && sc.geom[1].w == 100.0f && sc.mat_type[2] == UECRT_MAT_METAL);
  CHECK(sc.mat_param[3].y == 0.6f && sc.mat_param[0].w == 0.0f);
  // add() is [[nodiscard]] and returns a new list (hittable_list.hpp:26-30): the original is untouched
  const auto bigger = world.add(sphere(pos3<T, world_tag>(0, 1, -1), 0.25, dielectric<color::value_type>(1.5)),
                                sphere(pos3<T, world_tag>(0, 2, -1), 0.25, metal<color::value_type>({1, 1, 1}, 0.3)));
  CHECK(world.size() == 4 && bigger.size() == 6 && bigger.mat_type[4] == UECRT_MAT_DIELECTRIC && bigger.mat_param[5].w == 0.3f);
  CHECK(scenes::reference_scene().geom.size() == 4 && std::memcmp(scenes::reference_scene().geom.data(), world.geom.data(), 64) == 0);

  // ---- camera (camera.hpp:16-32): lower-left corner and the App. B centre ray
  CHECK(cam.lower_left_corner.x == -1.7777777777777777 && cam.lower_left_corner.y == -1 && cam.lower_left_corner.z == -1);
  const ray<T> centre = cam.get_ray(0.5, 0.5);
  CHECK(centre.origin.x == 0 && centre.direction.x == 0 && centre.direction.y == 0 && centre.direction.z == -1);
  CHECK(constants::height_for(400) == 225 && constants::height_for(3840) == 2160);
  const vec3<T> refl = reflect(vec3<T>(1, -1, 0), vec3<T>(0, 1, 0));
  CHECK(refl.x == 1 && refl.y == 1 && refl.z == 0);
  CHECK(vec3<T>(1e-9, 1e-9, 5).near_zero());                       // vec3.hpp:76-80: z is never tested
  std::printf("cpu checks ok\n");
  if (cpu_only) return 0;

  // ---- hittable_interface::hit on the GPU (SURVEY App. B known answers)
  auto rec = world.hit(centre, 0.001, std::numeric_limits<T>::infinity());
  CHECK(rec && rec->id == 0 && rec->t == 0.5 && rec->front_face);
  CHECK(rec->normal.x == 0 && rec->normal.y == 0 && rec->normal.z == 1 && rec->p.z == -0.5);
  rec = world.hit(cam.get_ray(0.5, 0.1), 0.001, std::numeric_limits<T>::infinity());
  CHECK(rec && rec->id == 1 && std::fabs(rec->t - 0.62587481340646434) < 1e-7);   // rays pass through float32
  CHECK(!world.hit(cam.get_ray(0.5, 0.9), 0.001, std::numeric_limits<T>::infinity()));
  CHECK(!world.hit(centre, 0.001, 0.4));                                          // closed [t_min, t_max]
  CHECK(world.hit(centre, 0.5, 0.5)->t == 0.5);
  const auto three = hittable_list<T>{}.add(sphere(pos3<T, world_tag>(0, 0, -1), 0.5, lambertian<T>({.5, .5, .5})),
                                            sphere(pos3<T, world_tag>(0, 0, -1), 0.5, lambertian<T>({.5, .5, .5})),
                                            sphere(pos3<T, world_tag>(0, 0, -1), 0.5, lambertian<T>({.5, .5, .5})));
  CHECK(three.hit(centre, 0.001, std::numeric_limits<T>::infinity())->id == 2);    // later index wins ties

  // ---- render(): progress lines and image geometry (source.cpp:114, 175, 59-62)
  std::ostringstream captured;
  auto* old = std::cout.rdbuf(captured.rdbuf());
  render_config cfg;
  cfg.width = 32; cfg.spp = 16; cfg.seed = 3;
  const image_t img = render(world, cam, cfg, true);
  std::cout.rdbuf(old);
  CHECK(captured.str() == "rendering...\nrendering finished\n");
  CHECK(img.width == 32 && img.height == 18 && img.pixels.size() == 32 * 18 && img.accum.size() == 32 * 18 * 4);
  CHECK(img.stats.paths == 32ull * 18 * 16 && img.accum[3] == 16.0f);
  CHECK(img.pixels[0].b > 200 && img.pixels[16 + 9 * 32].r > img.pixels[16 + 9 * 32].b);   // sky on top, red sphere in the centre
  std::printf("gpu checks ok\n");
  return 0;
}
