// dropin_demo.cpp -- TEST INFRASTRUCTURE ONLY: proof that the binding shown in INTEGRATION.md compiles
// against the UNMODIFIED reference.  The reference's translation unit is included where it lies
// (-I$(REF)); its own tuple-typed hittable_list, camera, constants, image_t and vendored stbi_write_png
// are used as they are, and only the body of render() -- the pixel x spp loop -- is replaced by ONE
// call through the C ABI.  Built into oracle/_ref/raytrace_dropin where /root/reference exists.
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <random>
#include <tuple>
#include <type_traits>
#include <vector>

#define main yk_reference_main
#include "source.cpp"                    // found through -I$(REF)
#undef main

#include "uecrt.h"

namespace yk {

template <class T, class... Hs>
static void flatten(const hittable_list<T, Hs...>& world, std::vector<uecrt_float4>& geom, std::vector<std::uint32_t>& type,
                    std::vector<uecrt_float4>& param) {
  std::apply(
      [&](const auto&... s) {            // tuple order == insertion order == hit_record::id
        (([&] {
           geom.push_back({(float)s.center.x, (float)s.center.y, (float)s.center.z, (float)s.radius});
           using M = std::decay_t<decltype(s.material)>;
           type.push_back(std::is_same_v<M, metal<double>> ? UECRT_MAT_METAL : UECRT_MAT_LAMBERTIAN);
           param.push_back({(float)s.material.albedo.r, (float)s.material.albedo.g, (float)s.material.albedo.b, 0.f});
         }()),
         ...);
      },
      world.objects);
}

// render() of source.cpp:98-178 with lines 116-172 replaced
image_t* render_b200(std::uint64_t seed) {
  using T = double;
  const camera<T> cam = {};
  const auto world =                     // source.cpp:103-112, verbatim
      hittable_list<T>{}
          .add(sphere(pos3<T, world_tag>(0, 0, -1), 0.5, lambertian<color::value_type>({0.7, 0.3, 0.3})))
          .add(sphere(pos3<T, world_tag>(0, -100.5, -1), 100.0, lambertian<color::value_type>({0.8, 0.8, 0.0})))
          .add(sphere(pos3<T, world_tag>(-1.0, 0.0, -1.0), 0.5, metal<color::value_type>({0.8, 0.8, 0.8})))
          .add(sphere(pos3<T, world_tag>(1.0, 0.0, -1.0), 0.5, metal<color::value_type>({0.8, 0.6, 0.2})));
  std::cout << "rendering..." << std::endl;
  auto* image = new image_t{};           // heap: image_t is a std::array (source.cpp:70-71)

  std::vector<uecrt_float4> geom, param;
  std::vector<std::uint32_t> type;
  flatten(world, geom, type, param);
  const uecrt_scene sc{(std::uint32_t)geom.size(), geom.data(), type.data(), param.data()};
  uecrt_camera c{};                      // camera.hpp:16-27, field for field
  c.origin[0] = (float)cam.origin.x; c.origin[1] = (float)cam.origin.y; c.origin[2] = (float)cam.origin.z;
  c.lower_left[0] = (float)cam.lower_left_corner.x; c.lower_left[1] = (float)cam.lower_left_corner.y; c.lower_left[2] = (float)cam.lower_left_corner.z;
  c.horizontal[0] = (float)cam.horizontal.x; c.horizontal[1] = (float)cam.horizontal.y; c.horizontal[2] = (float)cam.horizontal.z;
  c.vertical[0] = (float)cam.vertical.x; c.vertical[1] = (float)cam.vertical.y; c.vertical[2] = (float)cam.vertical.z;
  c.u[0] = 1.f; c.v[1] = 1.f;
  uecrt_params p{};
  p.width = constants::image_width; p.height = constants::image_height;
  p.spp = constants::samples_per_pixel; p.max_depth = constants::max_depth;
  p.seed = seed; p.num_gpus = 1;
  static_assert(sizeof(color3b) == 3, "image_t must be packed RGB");
  if (uecrt_render(&sc, &c, &p, reinterpret_cast<std::uint8_t*>(image->data()), nullptr, nullptr) != UECRT_OK) {
    std::cout << "error" << std::endl;
    std::cerr << uecrt_last_error() << std::endl;
    std::exit(EXIT_FAILURE);
  }
  std::cout << "rendering finished" << std::endl;
  return image;
}

}  // namespace yk

int main(int argc, char* argv[]) {
  const std::string filename = argc > 1 ? argv[1] : "image.png";
  const std::uint64_t seed = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : std::random_device{}();
  yk::image_t* image = yk::render_b200(seed);
  std::cout << "write to file : " << filename << std::endl;
  if (!stbi_write_png(filename.c_str(), yk::constants::image_width, yk::constants::image_height, 3, image->data(),
                      3 * yk::constants::image_width)) {   // the reference's vendored stb, source.cpp:224-226
    std::cout << "error" << std::endl;
    return EXIT_FAILURE;
  }
  std::cout << "success" << std::endl;
  delete image;
}
