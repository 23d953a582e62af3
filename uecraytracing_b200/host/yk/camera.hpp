// yk/camera.hpp -- host mirror of camera<T> (yk/camera.hpp:14-38).  The default constructor
// reproduces the reference exactly (origin 0, viewport 2 x 32/9, focal length 1); the second
// constructor is the RTIOW positionable / defocus camera north_star's final scene needs
// (absent upstream).  Defaults of the second form equal the first.
#pragma once
#include <cmath>

#include "../../../include/uecrt.h"
#include "ray.hpp"
#include "vec3.hpp"

namespace yk {

struct camera_tag : world_tag {};

template <class T>
struct camera {
  pos3<T, world_tag> origin;
  pos3<T, camera_tag> lower_left_corner;
  vec3<T> horizontal, vertical, u, v;
  T lens_radius{};

  camera() {                                     // camera.hpp:16-27
    const T aspect_ratio = T(16.0) / T(9.0);
    const T viewport_height = 2.0;
    const T viewport_width = aspect_ratio * viewport_height;
    const T focal_length = 1.0;
    origin = pos3<T, world_tag>(0, 0, 0);
    horizontal = vec3<T>(viewport_width, 0, 0);
    vertical = vec3<T>(0, viewport_height, 0);
    lower_left_corner = pos3<T, camera_tag>(0, 0, 0) - horizontal / T(2) - vertical / T(2) - vec3<T>(0, 0, focal_length);
    u = vec3<T>(1, 0, 0);
    v = vec3<T>(0, 1, 0);
  }
  camera(pos3<T, world_tag> lookfrom, pos3<T, world_tag> lookat, vec3<T> vup, T vfov_deg, T aspect_ratio, T aperture, T focus_dist) {
    const T theta = vfov_deg * T(3.14159265358979323846) / T(180);
    const T h = std::tan(theta / 2);
    const T viewport_height = T(2) * h;
    const T viewport_width = aspect_ratio * viewport_height;
    const vec3<T> w = vec3<T>(lookfrom - lookat).normalized();
    u = cross(vup, w).normalized();
    v = cross(w, u);
    origin = lookfrom;
    horizontal = u * (focus_dist * viewport_width);
    vertical = v * (focus_dist * viewport_height);
    lower_left_corner = origin - horizontal / T(2) - vertical / T(2) - w * focus_dist;
    lens_radius = aperture / 2;
  }
  // camera.hpp:29-32 (pinhole; the defocus offset is sampled on the device)
  ray<T> get_ray(T s, T t) const {
    vec3<T> target = lower_left_corner + s * horizontal + t * vertical;
    return ray<T>{origin, target - vec3<T>(origin)};
  }
  uecrt_camera flatten() const {
    uecrt_camera c{};
    auto put = [](float* d, const auto& s) { d[0] = (float)s.x; d[1] = (float)s.y; d[2] = (float)s.z; };
    put(c.origin, origin); put(c.lower_left, lower_left_corner); put(c.horizontal, horizontal);
    put(c.vertical, vertical); put(c.u, u); put(c.v, v);
    c.lens_radius = (float)lens_radius;
    return c;
  }
};

}  // namespace yk
