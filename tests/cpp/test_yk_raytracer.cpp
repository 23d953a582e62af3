// test_yk_raytracer.cpp -- code written against the REFERENCE's header names and call shapes, compiled against the
// mirror (uecraytracing_b200/host/yk): yk/ray.hpp, yk/hittable.hpp, yk/sphere.hpp, yk/raytracer.hpp,
// hittable_list::scatter<U>, raytracer<T,U>::ray_color.  The scene statements are source.cpp:100-112 verbatim; the
// known answers are SURVEY.md Appendix B (produced by the untouched reference headers).  Needs a B200: hit and scatter
// run on the GPU through libuecrt.so.
#include <cmath>
#include <cstdio>
#include <limits>
#include <random>
#include <vector>

#include "yk/camera.hpp"
#include "yk/color.hpp"
#include "yk/config.hpp"
#include "yk/hittable.hpp"
#include "yk/hittable_list.hpp"
#include "yk/material.hpp"
#include "yk/random.hpp"
#include "yk/ray.hpp"
#include "yk/raytracer.hpp"
#include "yk/sphere.hpp"

#define CHECK(c)                                                              \
  do {                                                                        \
    if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
  } while (0)

namespace yk {
using color = color3d;          // source.cpp:68
}

int main() {
  using namespace yk;
  using T = double;
  // ---- source.cpp:100-112
  const raytracer<T, double> tracer = {};
  const camera<T> cam = {};

  const auto world =
      hittable_list<T>{}
          .add(sphere(pos3<T, world_tag>(0, 0, -1), 0.5,
                      lambertian<color::value_type>({0.7, 0.3, 0.3})))
          .add(sphere(pos3<T, world_tag>(0, -100.5, -1), 100.0,
                      lambertian<color::value_type>({0.8, 0.8, 0.0})))
          .add(sphere(pos3<T, world_tag>(-1.0, 0.0, -1.0), 0.5,
                      metal<color::value_type>({0.8, 0.8, 0.8})))
          .add(sphere(pos3<T, world_tag>(1.0, 0.0, -1.0), 0.5,
                      metal<color::value_type>({0.8, 0.6, 0.2})));
  static_assert(concepts::hittable<std::remove_cvref_t<decltype(world)>, T>);
  static_assert(concepts::hittable<sphere<T, lambertian<double>>, T>);

  // ---- Appendix B: world.hit(cam.get_ray(u, v), 0.001, +inf), one batched call and ray by ray
  struct KA { double u, v; int id; double t, nx, ny, nz; };
  const KA ka[] = {{0.50, 0.50, 0, 0.5, 0, 0, 1},
                   {0.25, 0.50, 2, 0.68665834902296119, 0.779274046, 0, 0.626683302},
                   {0.75, 0.50, 3, 0.68665834902296119, -0.779274046, 0, 0.626683302},
                   {0.50, 0.10, 1, 0.62587481340646434, 0, 0.999993001, 0.003741252},
                   {0.50, 0.90, -1, 0, 0, 0, 0},
                   {0.02, 0.55, 2, 0.51685579548602179, 0.235798885, 0.103371159, 0.966288409}};
  std::vector<ray<T>> rays;
  for (const KA& k : ka) rays.push_back(cam.get_ray(k.u, k.v));
  const auto batch = world.hit(rays, 0.001, std::numeric_limits<T>::infinity());
  for (std::size_t i = 0; i < rays.size(); ++i) {
    const auto one = world.hit(rays[i], 0.001, std::numeric_limits<T>::infinity());
    CHECK(one.has_value() == batch[i].has_value() && one.has_value() == (ka[i].id >= 0));
    if (!one) continue;
    CHECK(one->id == (std::size_t)ka[i].id && batch[i]->id == one->id && one->front_face && one->t == batch[i]->t);
    // the ABI takes float rays: App. B's double directions (e.g. -0.8888888888888888) are rounded to float on the way in
    CHECK(std::fabs(one->t - ka[i].t) <= 2e-7 * ka[i].t);
    CHECK(std::fabs(one->normal.x - ka[i].nx) < 2e-7 && std::fabs(one->normal.y - ka[i].ny) < 2e-7 && std::fabs(one->normal.z - ka[i].nz) < 2e-7);
  }
  // a single sphere is hittable on its own (sphere.hpp:25-48): closed interval, far root (App. B, probe C.9)
  const auto s = sphere(pos3<T, world_tag>(0, 0, -1.5), 0.5, lambertian<double>({0.5, 0.5, 0.5}));
  const ray<T> zray{pos3<T, world_tag>(0, 0, 0), vec3<T>(0, 0, -1)};
  CHECK(s.hit(zray, 1.0, 2.0).has_value() && s.hit(zray, 1.0, 2.0)->t == 1.0);
  CHECK(s.hit(zray, 1.0000001, 2.0).has_value() && s.hit(zray, 1.0000001, 2.0)->t == 2.0);
  CHECK(s.hit(zray, 0.0, 1.0).has_value() && !s.hit(zray, 2.5, 9.0).has_value());

  // ---- scatter<U>: dispatch on rec.id, generator advanced as upstream (lambertian: 3 doubles = 6 draws; metal: none)
  {
    const auto rec = world.hit(rays[0], 0.001, std::numeric_limits<T>::infinity());       // sphere 0, lambertian
    std::mt19937 g(7), h(7);
    const auto sc = world.template scatter<double>(rays[0], rec.value(), g);
    CHECK(sc.has_value() && sc->first.r == 0.7f * 1.0 && sc->second.origin == rec->p);
    for (int i = 0; i < 6; ++i) h();
    CHECK(g == h);
    const double ux = 2 * (double)(float)std::generate_canonical<double, 53>(h = std::mt19937(7)) - 1;   // x of the cube sample
    (void)ux;
    const auto rec2 = world.hit(rays[1], 0.001, std::numeric_limits<T>::infinity());      // sphere 2, metal: mirror direction
    std::mt19937 g2(3), h2(3);
    const auto sm = world.template scatter<double>(rays[1], rec2.value(), g2);
    CHECK(sm.has_value() && g2 == h2);
    const vec3<T> v = rays[1].direction.normalized(), want = reflect(v, rec2->normal);
    CHECK(std::fabs(sm->second.direction.x - want.x) < 1e-6 && std::fabs(sm->second.direction.y - want.y) < 1e-6 && std::fabs(sm->second.direction.z - want.z) < 1e-6);
  }

  // ---- Appendix B: ray_color(cam.get_ray(.5, .5), world, 50, mt19937(7)) = (0.35704797734894778, 0.18781233703258654, 0)
  // (the hits are FP64-exact on the device, scatter runs in FP32 there: agreement to FP32 accuracy along one path)
  {
    mt19937 gen(7);
    const color3<double> c = tracer.ray_color(cam.get_ray(0.5, 0.5), world, 50, gen);
    std::printf("ray_color(centre, seed 7) = (%.17g, %.17g, %.17g)\n", c.r, c.g, c.b);
    CHECK(std::fabs(c.r - 0.35704797734894778) < 2e-5 && std::fabs(c.g - 0.18781233703258654) < 2e-5 && c.b == 0.0);
    mt19937 g0(7);
    CHECK((tracer.ray_color(cam.get_ray(0.5, 0.5), world, 0, g0) == color3<double>(0, 0, 0)) && g0 == mt19937(7));   // depth 0: black, no draws
    mt19937 g1(11);
    const color3<double> sky = tracer.ray_color(cam.get_ray(0.5, 0.9), world, 50, g1);    // a miss: the sky gradient, no draws
    const double t = (cam.get_ray(0.5, 0.9).direction.normalized().y + 1.0) / 2;
    CHECK(std::fabs(sky.r - ((1 - t) + t * 0.5)) < 1e-15 && std::fabs(sky.b - 1.0) < 1e-15 && g1 == mt19937(11));
  }
  // mean of ray_color over a few pixels and seeds stays a colour (finite, in [0, 1]) -- the loop K1 runs per path
  {
    color3<double> sum(0, 0, 0);
    int n = 0;
    for (int seed = 1; seed <= 4; ++seed)
      for (double u : {0.2, 0.5, 0.8})
        for (double v : {0.2, 0.5}) {
          mt19937 gen(seed);
          const color3<double> c = tracer.ray_color(cam.get_ray(u, v), world, 50, gen);
          CHECK(c.r >= 0 && c.r <= 1 && c.g >= 0 && c.g <= 1 && c.b >= 0 && c.b <= 1);
          sum += c; ++n;
        }
    CHECK(sum.r / n > 0.2 && sum.r / n < 0.9);
  }
  std::printf("yk raytracer ok\n");
  return 0;
}
