// oracle_port.cpp -- TEST INFRASTRUCTURE ONLY (see oracle_api.h).
//
// CPU restatement, in IEEE double, of the reference's path-tracing hot path.  Every function
// cites the reference file:line it follows and keeps the reference's operation ORDER, so that
// (compiled with -ffp-contract=off) it is bit-identical to the reference headers; tests pin it
// against oracle/_ref/liboracle_ref.so and against tests/golden/*.json generated from that.
//
// Parity status: PINNED for sphere/hittable_list hit, lambertian, metal(no fuzz), camera,
// ray_color, mt19937/uniform_real_distribution, math::sqrt, to_color3b (all exist in the
// reference).  UNPINNED BY THE REFERENCE for dielectric/Schlick/refract, metal fuzz and the
// defocus camera: those are absent upstream and restate the RTIOW book
// (App. D); liboracle_ref.so implements them with the reference's own vec3/ray/random types
// so the two libraries still cross-check each other.
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "oracle_api.h"

namespace {

struct V3 { double x, y, z; };
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }          // vec3.hpp:146-148
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }          // vec3.hpp:151-153
inline V3 mul(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }            // vec3.hpp:170-172
inline V3 divs(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }           // vec3.hpp:180-182
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }                                // vec3.hpp:84
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // vec3.hpp:114-117
inline double len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }           // vec3.hpp:129

// math.hpp:10-19 -- Newton iteration to a fixed point (loops forever on NaN, like upstream)
inline double nsqrt(double s) {
  double x = s / 2.0;
  double prev = 0.0;
  while (x != prev) {
    prev = x;
    x = (x + s / x) / 2.0;
  }
  return x;
}
inline double length(V3 a) { return nsqrt(a.x * a.x + a.y * a.y + a.z * a.z); }  // vec3.hpp:127
inline V3 normalized(V3 a) { return divs(a, length(a)); }                        // vec3.hpp:132
inline V3 reflect(V3 v, V3 n) { return sub(v, mul(n, 2 * dot(v, n))); }          // vec3.hpp:199-202
// vec3.hpp:76-80 -- NOTE upstream bug kept: z is never tested (x is tested twice)
inline bool near_zero(V3 a) {
  auto f = [](double x) { return x > 0 ? x : -x; };
  const double s = 1e-8;
  return (f(a.x) < s) && (f(a.y) < s) && (f(a.x) < s);
}

// random.hpp:43-151 (mersenne_twister_engine<uint_fast32_t,32,624,397,31,...>)
struct MT {
  uint32_t x[624];
  size_t p;
  explicit MT(uint32_t sd) {
    x[0] = sd;
    for (size_t i = 1; i < 624; ++i) {
      uint32_t v = x[i - 1];
      v ^= v >> 30;
      v *= 1812433253u;
      v += (uint32_t)(i % 624);
      x[i] = v;
    }
    p = 624;
  }
  void gen() {  // random.hpp:111-128
    const uint32_t upper = 0x80000000u, lower = 0x7fffffffu, a = 0x9908b0dfu;
    for (size_t k = 0; k < 624 - 397; ++k) {
      uint32_t y = (x[k] & upper) | (x[k + 1] & lower);
      x[k] = x[k + 397] ^ (y >> 1) ^ ((y & 1) ? a : 0);
    }
    for (size_t k = 624 - 397; k < 623; ++k) {
      uint32_t y = (x[k] & upper) | (x[k + 1] & lower);
      x[k] = x[k + 397 - 624] ^ (y >> 1) ^ ((y & 1) ? a : 0);
    }
    uint32_t y = (x[623] & upper) | (x[0] & lower);
    x[623] = x[396] ^ (y >> 1) ^ ((y & 1) ? a : 0);
    p = 0;
  }
  uint32_t next() {  // random.hpp:92-102
    if (p >= 624) gen();
    uint32_t z = x[p++];
    z ^= (z >> 11) & 0xffffffffu;
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= (z >> 18);
    return z;
  }
};

// The random source seen by the path code: either the engine or a replay list of canonicals.
struct Rng {
  MT* mt = nullptr;
  const double* replay = nullptr;
  uint32_t n_replay = 0, i_replay = 0;
  // random.hpp:161-183 generate_canonical<double,53>: m = 2 engine outputs, (lo + hi*2^32)/2^64
  double canonical() {
    if (replay) return i_replay < n_replay ? replay[i_replay++] : 0.0;
    const long double r = 4294967296.0L;
    double sum = 0, tmp = 1;
    for (int k = 2; k != 0; --k) {
      sum += double(mt->next()) * tmp;
      tmp *= r;
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = 1.0 - std::numeric_limits<double>::epsilon() / 2.0;
    return ret;
  }
  // random.hpp:267-272 uniform_real_distribution::operator()
  double uniform(double a, double b) { return (canonical() * (b - a)) + a; }
};

struct Sphere { V3 c; double r; uint32_t mat; double albedo[3]; double param; };
struct Hit { V3 p, normal; double t; int64_t id; bool front; };

// sphere.hpp:25-48 + hittable.hpp:23-27
inline bool sphere_hit(const Sphere& s, V3 o, V3 d, double t_min, double t_max, Hit& rec) {
  V3 oc = sub(o, s.c);
  double a = len2(d);
  double half_b = dot(oc, d);
  double c = len2(oc) - s.r * s.r;
  double disc = half_b * half_b - a * c;
  if (disc < 0) return false;
  double sqrtd = nsqrt(disc);
  double root = (-half_b - sqrtd) / a;
  if (root < t_min || t_max < root) {
    root = (-half_b + sqrtd) / a;
    if (root < t_min || t_max < root) return false;
  }
  rec.t = root;
  rec.p = add(o, mul(d, rec.t));                 // ray.hpp:16-19
  V3 outward = divs(sub(rec.p, s.c), s.r);
  rec.front = dot(d, outward) < 0;               // hittable.hpp:25
  rec.normal = rec.front ? outward : neg(outward);
  return true;
}

// hittable_list.hpp:32-58: left fold in insertion order with a shrinking closed t_max;
// an equal t is accepted, so the LATER index wins ties.
inline bool world_hit(const std::vector<Sphere>& w, V3 o, V3 d, double t_min, double t_max, Hit& best) {
  bool any = false;
  double closest = t_max;
  Hit rec;
  for (size_t i = 0; i < w.size(); ++i) {
    if (sphere_hit(w[i], o, d, t_min, closest, rec)) {
      any = true;
      closest = rec.t;
      rec.id = (int64_t)i;
      best = rec;
    }
  }
  return any;
}

// material.hpp:33-36 + vec3.hpp:134-142: normalize(U[-1,1)^3), x then y then z, no rejection
inline V3 random_unit_vector(Rng& g) {
  V3 v;
  v.x = g.uniform(-1, 1);
  v.y = g.uniform(-1, 1);
  v.z = g.uniform(-1, 1);
  return divs(v, length(v));                     // vec3.hpp:131 normalize(): *this /= length()
}
// material.hpp:27-31 (never called upstream; the fuzz extension uses its distribution).  Upstream
// writes normalize(random) * dist(gen): both operands draw, in an order C++ leaves unspecified;
// fixed here (and in ref_harness.cpp) as direction x,y,z first, then the radius.
inline V3 random_in_unit_sphere(Rng& g) {
  V3 v = random_unit_vector(g);
  double radius = g.uniform(0.01, 0.99);
  return mul(v, radius);
}

// [EXTERNAL, RTIOW book; SURVEY.md App. D] -- not in the reference
inline V3 refract(V3 uv, V3 n, double eta) {
  double c = dot(neg(uv), n);
  double cos_theta = c < 1.0 ? c : 1.0;
  V3 perp = mul(add(uv, mul(n, cos_theta)), eta);
  double k = 1.0 - len2(perp);
  double ak = k > 0 ? k : -k;
  V3 par = mul(n, -nsqrt(ak));
  return add(perp, par);
}
inline double schlick(double cosine, double eta) {
  double r0 = (1 - eta) / (1 + eta);
  r0 = r0 * r0;
  double m = 1 - cosine;
  return r0 + (1 - r0) * (m * m * m * m * m);
}

inline bool scatter(const Sphere& s, V3 d_in, const Hit& rec, Rng& g, double att[3], V3& o_out, V3& d_out) {
  att[0] = s.albedo[0]; att[1] = s.albedo[1]; att[2] = s.albedo[2];
  o_out = rec.p;
  switch (s.mat) {
    case UECRT_MAT_LAMBERTIAN: {                 // material.hpp:50-59
      V3 dir = add(rec.normal, random_unit_vector(g));
      if (near_zero(dir)) dir = rec.normal;
      d_out = dir;
      return true;
    }
    case UECRT_MAT_METAL: {                      // material.hpp:67-75 (+ fuzz, App. D)
      V3 reflected = reflect(normalized(d_in), rec.normal);
      if (s.param > 0) reflected = add(reflected, mul(random_in_unit_sphere(g), s.param));
      d_out = reflected;
      return dot(d_out, rec.normal) > 0;
    }
    default: {                                   // dielectric, App. D
      att[0] = att[1] = att[2] = 1.0;
      double eta = rec.front ? (1.0 / s.param) : s.param;
      V3 unit = normalized(d_in);
      double c = dot(neg(unit), rec.normal);
      double cos_theta = c < 1.0 ? c : 1.0;
      double sin_theta = nsqrt(1.0 - cos_theta * cos_theta);
      bool cannot = eta * sin_theta > 1.0;
      if (cannot || schlick(cos_theta, eta) > g.uniform(0, 1))
        d_out = reflect(unit, rec.normal);
      else
        d_out = refract(unit, rec.normal, eta);
      return true;
    }
  }
}

// raytracer.hpp:19-37 (recursive, same multiplication order)
void ray_color(const std::vector<Sphere>& w, V3 o, V3 d, unsigned depth, Rng& g, double out[3], uint64_t& rays) {
  if (depth == 0) { out[0] = out[1] = out[2] = 0; return; }
  ++rays;
  Hit rec;
  if (world_hit(w, o, d, 0.001, std::numeric_limits<double>::infinity(), rec)) {
    double att[3]; V3 so, sd;
    if (scatter(w[(size_t)rec.id], d, rec, g, att, so, sd)) {
      double sub_[3];
      ray_color(w, so, sd, depth - 1, g, sub_, rays);
      out[0] = att[0] * sub_[0]; out[1] = att[1] * sub_[1]; out[2] = att[2] * sub_[2];
    } else {
      out[0] = out[1] = out[2] = 0;
    }
    return;
  }
  double t = (normalized(d).y + 1.0) / 2;
  out[0] = (1.0 - t) * 1.0 + t * 0.5;
  out[1] = (1.0 - t) * 1.0 + t * 0.7;
  out[2] = (1.0 - t) * 1.0 + t * 1.0;
}

std::vector<Sphere> widen(const uecrt_scene* sc) {
  std::vector<Sphere> w(sc->n);
  for (uint32_t i = 0; i < sc->n; ++i) {
    w[i].c = {(double)sc->geom[i].x, (double)sc->geom[i].y, (double)sc->geom[i].z};
    w[i].r = (double)sc->geom[i].w;
    w[i].mat = sc->mat_type[i];
    w[i].albedo[0] = sc->mat_param[i].x; w[i].albedo[1] = sc->mat_param[i].y; w[i].albedo[2] = sc->mat_param[i].z;
    w[i].param = sc->mat_param[i].w;
  }
  return w;
}

struct Cam { V3 origin, llc, hor, ver, u, v; double lens; };
Cam widen(const uecrt_camera* c) {
  auto v3 = [](const float* f) { return V3{(double)f[0], (double)f[1], (double)f[2]}; };
  return {v3(c->origin), v3(c->lower_left), v3(c->horizontal), v3(c->vertical), v3(c->u), v3(c->v), (double)c->lens_radius};
}
// camera.hpp:29-32; with lens>0 the RTIOW defocus form (App. D)
inline void camera_ray(const Cam& c, double s, double t, double rdx, double rdy, V3& o, V3& d) {
  V3 target = add(add(c.llc, mul(c.hor, s)), mul(c.ver, t));   // lower_left + u*horizontal + v*vertical
  if (c.lens > 0) {
    V3 off = add(mul(c.u, c.lens * rdx), mul(c.v, c.lens * rdy));
    o = add(c.origin, off);
    d = sub(sub(target, c.origin), off);
  } else {
    // the reference never subtracts its (zero) origin (camera.hpp:30-31); a positioned pinhole must
    o = c.origin;
    bool zero = c.origin.x == 0 && c.origin.y == 0 && c.origin.z == 0;
    d = zero ? target : sub(target, c.origin);
  }
}
}  // namespace

extern "C" {

int orc_kind(void) { return 0; }

uint32_t orc_mt19937_nth(uint32_t seed, uint32_t n) {
  MT m(seed);
  uint32_t v = 0;
  for (uint32_t i = 0; i <= n; ++i) v = m.next();
  return v;
}
double orc_uniform_nth(uint32_t seed, double a, double b, uint32_t n) {
  MT m(seed);
  Rng g; g.mt = &m;
  double v = 0;
  for (uint32_t i = 0; i <= n; ++i) v = g.uniform(a, b);
  return v;
}
double orc_sqrt(double s) { return nsqrt(s); }

int orc_hits(const uecrt_scene* scene, uint64_t n_rays, const double* o, const double* d, double t_min,
             double t_max, double* t, double* nrm, int32_t* id, uint8_t* ff) {
  std::vector<Sphere> w = widen(scene);
  for (uint64_t i = 0; i < n_rays; ++i) {
    Hit rec;
    V3 ro{o[3 * i], o[3 * i + 1], o[3 * i + 2]}, rd{d[3 * i], d[3 * i + 1], d[3 * i + 2]};
    if (world_hit(w, ro, rd, t_min, t_max, rec)) {
      t[i] = rec.t; nrm[3 * i] = rec.normal.x; nrm[3 * i + 1] = rec.normal.y; nrm[3 * i + 2] = rec.normal.z;
      id[i] = (int32_t)rec.id; ff[i] = rec.front ? 1 : 0;
    } else {
      t[i] = -1; nrm[3 * i] = nrm[3 * i + 1] = nrm[3 * i + 2] = 0; id[i] = -1; ff[i] = 0;
    }
  }
  return 0;
}

int orc_camera_ray(const uecrt_camera* cam, double s, double t, double rdx, double rdy, double* o, double* d) {
  Cam c = widen(cam);
  V3 ro, rd;
  camera_ray(c, s, t, rdx, rdy, ro, rd);
  o[0] = ro.x; o[1] = ro.y; o[2] = ro.z; d[0] = rd.x; d[1] = rd.y; d[2] = rd.z;
  return 0;
}

int orc_scatter(uint32_t mat_type, const uecrt_float4* mp, const double* d_in, const double* p, const double* n,
                int front_face, const double* uniforms, uint32_t n_uniforms, int* scattered, double* att,
                double* o_out, double* d_out) {
  Sphere s{};
  s.mat = mat_type; s.albedo[0] = mp->x; s.albedo[1] = mp->y; s.albedo[2] = mp->z; s.param = mp->w;
  Hit rec{};
  rec.p = {p[0], p[1], p[2]}; rec.normal = {n[0], n[1], n[2]}; rec.front = front_face != 0;
  Rng g; g.replay = uniforms; g.n_replay = n_uniforms;
  V3 so, sd;
  bool ok = scatter(s, V3{d_in[0], d_in[1], d_in[2]}, rec, g, att, so, sd);
  *scattered = ok ? 1 : 0;
  o_out[0] = so.x; o_out[1] = so.y; o_out[2] = so.z; d_out[0] = sd.x; d_out[1] = sd.y; d_out[2] = sd.z;
  return 0;
}

int orc_ray_color(const uecrt_scene* scene, const double* o, const double* d, uint32_t depth, uint32_t seed,
                  double* rgb, uint64_t* rays_traced) {
  std::vector<Sphere> w = widen(scene);
  MT m(seed);
  Rng g; g.mt = &m;
  uint64_t rays = 0;
  ray_color(w, V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]}, depth, g, rgb, rays);
  if (rays_traced) *rays_traced = rays;
  return 0;
}

// source.cpp:122-172
int orc_render_rows(const uecrt_scene* scene, const uecrt_camera* cam, uint32_t W, uint32_t H, uint32_t spp,
                    uint32_t max_depth, uint32_t seed, uint32_t y0, uint32_t y1, double* sum_rgb,
                    uint64_t* rays_traced) {
  std::vector<Sphere> w = widen(scene);
  Cam c = widen(cam);
  uint64_t rays = 0;
  for (uint32_t y = y0; y < y1; ++y)
    for (uint32_t x = 0; x < W; ++x) {
      double px[3] = {0, 0, 0};
      for (uint32_t s = 0; s < spp; ++s) {
        MT m(seed + (y * W + x) * spp + s);                  // source.cpp:154-158 (uint32 wrap kept)
        Rng g; g.mt = &m;
        double u = (x + g.uniform(0, 1)) / W;                  // source.cpp:162
        double v = (H - y - 1 + g.uniform(0, 1)) / H;          // source.cpp:163-164
        double rdx = 0, rdy = 0;
        if (c.lens > 0) {                                      // RTIOW random_in_unit_disk (App. D)
          for (;;) {
            rdx = g.uniform(-1, 1); rdy = g.uniform(-1, 1);
            if (rdx * rdx + rdy * rdy >= 1) continue;
            break;
          }
        }
        V3 ro, rd;
        camera_ray(c, u, v, rdx, rdy, ro, rd);
        double col[3];
        ray_color(w, ro, rd, max_depth, g, col, rays);
        px[0] = px[0] + col[0]; px[1] = px[1] + col[1]; px[2] = px[2] + col[2];  // std::plus, in order
      }
      double* dst = sum_rgb + ((size_t)(y - y0) * W + x) * 3;
      dst[0] = px[0]; dst[1] = px[1]; dst[2] = px[2];
    }
  if (rays_traced) *rays_traced = rays;
  return 0;
}

// source.cpp:73-83: /spp, math::sqrt per channel, clamp [0,0.999], *256, truncate
void orc_to_color3b(const double* sum_rgb, uint64_t n_pixels, uint32_t spp, uint8_t* rgb) {
  for (uint64_t i = 0; i < n_pixels * 3; ++i) {
    double c = sum_rgb[i] / spp;
    double g = (c == c && c > 0) ? nsqrt(c) : 0.0;   // upstream would hang on NaN (math.hpp:14); 0 -> 0
    double cl = g < 0.0 ? 0.0 : (g > 0.999 ? 0.999 : g);
    rgb[i] = (uint8_t)(cl * 256);
  }
}

int orc_ref4_tuple_hits(uint64_t, const double*, const double*, double, double, double*, double*, int32_t*, uint8_t*) { return -1; }
int orc_ref4_tuple_ray_color(const double*, const double*, uint32_t, uint32_t, double*) { return -1; }
}
