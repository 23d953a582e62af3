// materials.cuh -- device scatter functions (FP32).
//   lambertian : yk/material.hpp:50-59 with random_unit_vector = normalize(U[-1,1)^3)
//                (material.hpp:33-36, vec3.hpp:134-142: cube-normalised, NOT rejection sampled)
//   metal      : yk/material.hpp:67-75; fuzz > 0 is the RTIOW extension using the reference's
//                random_in_unit_sphere helper distribution (material.hpp:27-31)
//   dielectric : RTIOW book (absent from the reference)
// hittable_list::scatter's O(N) dispatch-by-id chain (hittable_list.hpp:60-73) becomes a
// switch on the per-sphere material type.
#pragma once
#include "../../include/uecrt.h"
#include "closest_hit.cuh"
#include "philox.cuh"

namespace uecrt {

struct ScatterOut {
  bool ok;                 // false = absorbed (metal scattered below the surface, material.hpp:72-74)
  float ar, ag, ab;        // attenuation
  float dx, dy, dz;        // scattered direction, NOT normalised (as in the reference)
};

// d = unit incoming direction; n = unit normal against the ray; xi = 4 uniforms in [0,1)
__device__ __forceinline__ ScatterOut scatter_material(uint32_t type, float4 param, float dx, float dy, float dz,
                                                       float nx, float ny, float nz, bool front, float u0,
                                                       float u1, float u2, float u3) {
  ScatterOut s;
  s.ok = true;
  s.ar = param.x; s.ag = param.y; s.ab = param.z;
  if (type == UECRT_MAT_LAMBERTIAN) {
    const float vx = fmaf(2.0f, u0, -1.0f), vy = fmaf(2.0f, u1, -1.0f), vz = fmaf(2.0f, u2, -1.0f);
    const float inv = rsqrtf(fmaf(vz, vz, fmaf(vy, vy, vx * vx)));
    s.dx = fmaf(vx, inv, nx); s.dy = fmaf(vy, inv, ny); s.dz = fmaf(vz, inv, nz);
    // vec3.hpp:76-80 tests x, y, x (z never): kept
    if (fabsf(s.dx) < 1e-8f && fabsf(s.dy) < 1e-8f) { s.dx = nx; s.dy = ny; s.dz = nz; }
  } else if (type == UECRT_MAT_METAL) {
    const float dn = 2.0f * fmaf(dz, nz, fmaf(dy, ny, dx * nx));
    s.dx = fmaf(-dn, nx, dx); s.dy = fmaf(-dn, ny, dy); s.dz = fmaf(-dn, nz, dz);    // vec3.hpp:199-202
    if (param.w > 0.0f) {
      const float vx = fmaf(2.0f, u0, -1.0f), vy = fmaf(2.0f, u1, -1.0f), vz = fmaf(2.0f, u2, -1.0f);
      const float scale = rsqrtf(fmaf(vz, vz, fmaf(vy, vy, vx * vx))) * fmaf(0.98f, u3, 0.01f) * param.w;
      s.dx = fmaf(vx, scale, s.dx); s.dy = fmaf(vy, scale, s.dy); s.dz = fmaf(vz, scale, s.dz);
    }
    s.ok = fmaf(s.dz, nz, fmaf(s.dy, ny, s.dx * nx)) > 0.0f;                          // material.hpp:72
  } else {
    s.ar = s.ag = s.ab = 1.0f;
    const float eta = front ? __fdividef(1.0f, param.w) : param.w;
    const float ddn = fmaf(dz, nz, fmaf(dy, ny, dx * nx));
    const float cos_theta = fminf(-ddn, 1.0f);
    const float sin2 = fmaxf(fmaf(-cos_theta, cos_theta, 1.0f), 0.0f);
    const bool cannot = eta * eta * sin2 > 1.0f;
    float r0 = __fdividef(1.0f - eta, 1.0f + eta);
    r0 *= r0;
    const float m = 1.0f - cos_theta, m2 = m * m;
    const float refl = fmaf(1.0f - r0, m2 * m2 * m, r0);
    if (cannot || refl > u0) {
      const float dn = 2.0f * ddn;
      s.dx = fmaf(-dn, nx, dx); s.dy = fmaf(-dn, ny, dy); s.dz = fmaf(-dn, nz, dz);
    } else {
      const float px = eta * fmaf(cos_theta, nx, dx), py = eta * fmaf(cos_theta, ny, dy), pz = eta * fmaf(cos_theta, nz, dz);
      const float k = -sqrt_fast(fabsf(1.0f - fmaf(pz, pz, fmaf(py, py, px * px))));
      s.dx = fmaf(k, nx, px); s.dy = fmaf(k, ny, py); s.dz = fmaf(k, nz, pz);
    }
  }
  return s;
}

}  // namespace uecrt
