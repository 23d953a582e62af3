/* oracle_api.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C API shared by the two CPU checkers of the path-tracing hot path:
 *   oracle/liboracle_port.so      oracle_port.cpp : a plain restatement of the reference's
 *                                 algorithm in double precision (travels to the GPU box);
 *   oracle/_ref/liboracle_ref.so  ref_harness.cpp : the same API implemented by calling the
 *                                 UNMODIFIED reference headers under /root/reference
 *                                 (compiled where they lie, never copied).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load these libraries.  The product (libuecrt.so, the raytrace binary, the python
 * package) never links, loads or calls anything declared here.
 *
 * Scene and camera use the product's flattened float structs (include/uecrt.h) widened to
 * double inside the oracle, so both sides see bit-identical inputs.
 */
#ifndef UECRT_ORACLE_API_H
#define UECRT_ORACLE_API_H
#include <stdint.h>
#include "../include/uecrt.h"
#ifdef __cplusplus
extern "C" {
#endif

/* 0 = restatement ("port"), 1 = reference headers ("reference") */
int orc_kind(void);

/* yk::mt19937 (random.hpp:43-151): n-th (0-based) output for a seed */
uint32_t orc_mt19937_nth(uint32_t seed, uint32_t n);
/* yk::uniform_real_distribution<double>(a,b) (random.hpp:209-287): n-th draw */
double orc_uniform_nth(uint32_t seed, double a, double b, uint32_t n);
/* yk::math::sqrt (math.hpp:10-19) */
double orc_sqrt(double s);

/* hittable_list::hit_impl (hittable_list.hpp:32-58) over sphere::hit_impl (sphere.hpp:25-48)
 * on n_rays double rays.  Miss: id=-1, t=-1.  */
int orc_hits(const uecrt_scene* scene, uint64_t n_rays, const double* o_xyz, const double* d_xyz,
             double t_min, double t_max, double* t, double* n_xyz, int32_t* id,
             uint8_t* front_face);

/* camera::get_ray (camera.hpp:29-32) + RTIOW defocus offset (rdx,rdy = unit-disk sample) */
int orc_camera_ray(const uecrt_camera* cam, double s, double t, double rdx, double rdy,
                   double* o_xyz, double* d_xyz);

/* material scatter with a REPLAY generator: the uniforms (each in [0,1), <= 32 significant
 * bits) are what generate_canonical returns, in draw order.  material.hpp:50-59 (lambertian),
 * 67-75 (metal; fuzz>0 and dielectric are the RTIOW extensions).  *scattered = 0 on absorb. */
int orc_scatter(uint32_t mat_type, const uecrt_float4* mat_param, const double* d_in_xyz,
                const double* p_xyz, const double* normal_xyz, int front_face,
                const double* uniforms, uint32_t n_uniforms, int* scattered,
                double* attenuation_rgb, double* o_out_xyz, double* d_out_xyz);

/* raytracer::ray_color (raytracer.hpp:19-37) with yk::mt19937(seed) */
int orc_ray_color(const uecrt_scene* scene, const double* o_xyz, const double* d_xyz,
                  uint32_t depth, uint32_t seed, double* rgb, uint64_t* rays_traced);

/* render() rows [y0,y1) (source.cpp:122-172) with the constexpr-mode per-sample seed
 * seed + (y*W+x)*spp + s (source.cpp:155-158).  sum_rgb: (y1-y0)*W*3 doubles, the
 * un-normalised sample sum (pixel_color before to_color3b). Thread-safe. */
int orc_render_rows(const uecrt_scene* scene, const uecrt_camera* cam, uint32_t width,
                    uint32_t height, uint32_t spp, uint32_t max_depth, uint32_t seed,
                    uint32_t y0, uint32_t y1, double* sum_rgb, uint64_t* rays_traced);

/* to_color3b (source.cpp:73-83) on n pixels */
void orc_to_color3b(const double* sum_rgb, uint64_t n_pixels, uint32_t spp, uint8_t* rgb);

/* reference-only extras (return -1 in the port): the TRUE tuple-typed hittable_list of the
 * reference scene (source.cpp:103-112) through world.hit / ray_color, to pin the restated fold */
int orc_ref4_tuple_hits(uint64_t n_rays, const double* o_xyz, const double* d_xyz, double t_min,
                        double t_max, double* t, double* n_xyz, int32_t* id, uint8_t* front_face);
int orc_ref4_tuple_ray_color(const double* o_xyz, const double* d_xyz, uint32_t depth,
                             uint32_t seed, double* rgb);

#ifdef __cplusplus
}
#endif
#endif
