// yk/scene_io.hpp -- runtime scene description ("the step before the path").
// The reference hard-codes its scene in render() (source.cpp:103-112); this reads the same
// information from a text file so the drop-in binary can render arbitrary sphere scenes.
//
//   # comment
//   sphere  cx cy cz radius  lambertian r g b
//   sphere  cx cy cz radius  metal      r g b [fuzz]
//   sphere  cx cy cz radius  dielectric ior
//   camera  fromx fromy fromz  atx aty atz  vfov_deg [aperture [focus_dist]]     (optional, once)
//
// Lines appear in insertion order, which is hit_record::id (hittable_list.hpp:41).
#pragma once
#include <cmath>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>

#include "camera.hpp"
#include "hittable_list.hpp"

namespace yk::scene_io {

struct scene_file {
  hittable_list<double> world;
  camera<double> cam;          // the reference camera unless the file has a `camera` line
  bool has_camera = false;
};

inline scene_file parse(std::istream& in, const std::string& name = "<scene>") {
  scene_file out;
  std::string line;
  int lineno = 0;
  auto fail = [&](const std::string& why) { throw std::runtime_error(name + ":" + std::to_string(lineno) + ": " + why); };
  while (std::getline(in, line)) {
    ++lineno;
    const auto hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    std::istringstream ls(line);
    std::string kw;
    if (!(ls >> kw)) continue;
    if (kw == "sphere") {
      double cx, cy, cz, r;
      std::string mat;
      if (!(ls >> cx >> cy >> cz >> r >> mat)) fail("expected: sphere cx cy cz radius material ...");
      const pos3<double, world_tag> c(cx, cy, cz);
      if (mat == "lambertian") {
        double cr, cg, cb;
        if (!(ls >> cr >> cg >> cb)) fail("lambertian needs r g b");
        out.world.push(sphere(c, r, lambertian<double>({cr, cg, cb})));
      } else if (mat == "metal") {
        double cr, cg, cb, fuzz = 0.0;
        if (!(ls >> cr >> cg >> cb)) fail("metal needs r g b [fuzz]");
        ls >> fuzz;
        out.world.push(sphere(c, r, metal<double>({cr, cg, cb}, fuzz)));
      } else if (mat == "dielectric") {
        double ior;
        if (!(ls >> ior)) fail("dielectric needs an index of refraction");
        out.world.push(sphere(c, r, dielectric<double>(ior)));
      } else {
        fail("unknown material '" + mat + "'");
      }
    } else if (kw == "camera") {
      double f[3], a[3], vfov, aperture = 0.0, focus = -1.0;
      if (!(ls >> f[0] >> f[1] >> f[2] >> a[0] >> a[1] >> a[2] >> vfov)) fail("expected: camera from(3) at(3) vfov [aperture [focus]]");
      ls >> aperture >> focus;
      if (focus <= 0) focus = std::sqrt((f[0] - a[0]) * (f[0] - a[0]) + (f[1] - a[1]) * (f[1] - a[1]) + (f[2] - a[2]) * (f[2] - a[2]));
      out.cam = camera<double>(pos3<double, world_tag>(f[0], f[1], f[2]), pos3<double, world_tag>(a[0], a[1], a[2]),
                               vec3<double>(0, 1, 0), vfov, 16.0 / 9.0, aperture, focus);
      out.has_camera = true;
    } else {
      fail("unknown keyword '" + kw + "'");
    }
  }
  return out;
}

inline scene_file load(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open scene file " + path);
  return parse(f, path);
}

// the inverse, from the flattened arrays (float values print exactly with 9 significant digits)
inline void write(std::ostream& os, const hittable_list<double>& w) {
  os << std::setprecision(9);
  for (std::size_t i = 0; i < w.size(); ++i) {
    const auto& g = w.geom[i];
    const auto& m = w.mat_param[i];
    os << "sphere " << g.x << ' ' << g.y << ' ' << g.z << ' ' << g.w << ' ';
    if (w.mat_type[i] == UECRT_MAT_LAMBERTIAN) os << "lambertian " << m.x << ' ' << m.y << ' ' << m.z;
    else if (w.mat_type[i] == UECRT_MAT_METAL) os << "metal " << m.x << ' ' << m.y << ' ' << m.z << ' ' << m.w;
    else os << "dielectric " << m.w;
    os << '\n';
  }
}

}  // namespace yk::scene_io
