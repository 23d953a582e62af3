// yk/raytracer.hpp -- host mirror of raytracer<T, U>::ray_color (yk/raytracer.hpp:16-38) over the mirror's GPU-backed
// world: every bounce is one closest-hit call and one scatter call on the device, the generator is advanced exactly as
// upstream (the materials draw from it), and the result is attenuation_1 * (attenuation_2 * (... * terminal)) in the
// reference's association.  A parity hook: K1 runs the same loop for millions of paths at once.
#pragma once
#include <iostream>
#include <limits>
#include <random>
#include <vector>

#include "color.hpp"
#include "concepts.hpp"
#include "config.hpp"
#include "hittable.hpp"
#include "ray.hpp"
#include "vec3.hpp"

namespace yk {

template <concepts::arithmetic T, concepts::arithmetic U>
struct raytracer {
  template <concepts::hittable<T> H, std::uniform_random_bit_generator Gen>
  color3<U> ray_color(const ray<T>& r, const H& world, unsigned int depth, Gen& gen) const {
    std::vector<color3<U>> attenuation;        // of the bounces taken so far
    color3<U> terminal(0, 0, 0);               // depth exhausted or absorbed: black (raytracer.hpp:26, 32-33)
    ray<T> cur = r;
    for (; depth != 0; --depth) {
      if (verbose > 2)
        std::cout << "ray { origin : (" << cur.origin.x << ", " << cur.origin.y << ", " << cur.origin.z << "), direction : (" << cur.direction.x
                  << ", " << cur.direction.y << ", " << cur.direction.z << ") }" << '\n';
      const auto rec = world.hit(cur, T(0.001), std::numeric_limits<T>::infinity());
      if (!rec) {                              // sky gradient (raytracer.hpp:35-36)
        const auto t = (cur.direction.normalized().y + 1.0) / 2;
        terminal = (1.0 - t) * color3<U>(1.0, 1.0, 1.0) + t * color3<U>(0.5, 0.7, 1.0);
        break;
      }
      const auto sc = world.template scatter<U>(cur, rec.value(), gen);
      if (!sc) break;
      attenuation.push_back(sc->first);
      cur = sc->second;
    }
    for (std::size_t k = attenuation.size(); k-- > 0;) terminal = attenuation[k] * terminal;
    return terminal;
  }
};

}  // namespace yk
