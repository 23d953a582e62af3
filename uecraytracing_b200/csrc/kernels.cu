// kernels.cu -- hand-written sm_100a kernels of the path-tracing hot path.
//
//   K1 k_render_paths   render<T>() pixel x spp loop (source.cpp:122-172) + ray_color
//                       (raytracer.hpp:19-37) + hittable_list hit/scatter (hittable_list.hpp:32-73)
//                       + sphere::hit_impl (sphere.hpp:25-48) + materials (material.hpp:44-76)
//   K2 k_primary_hits   hittable_list::hit_impl on caller rays (parity kernel)
//   K3 k_resolve        to_color3b (source.cpp:73-83)
//   probes              scatter_material / philox on caller inputs (unit parity)
//
// K1 design (B200: 148 SMs, 4 sub-partitions each, FP32 128 lanes/clk/SM):
//  * persistent CTAs; every thread owns one live path; the reference's recursion is an
//    iterative loop carrying the throughput product (exact up to FP reassociation);
//  * the packed scene (32 B per sphere pair) is staged into shared memory with cp.async.bulk (TMA,
//    SASS UBLKCP) completing on an mbarrier -- ONCE per CTA when it fits (RESIDENT), else TILED:
//    the CTA sweeps the scene through a two-buffer ring, every live ray of the block testing tile
//    k while the TMA fills the other buffer with tile k+1; the scan reads shared memory with
//    warp-broadcast LDS.128, so there is no bank conflict and no global traffic in the loop;
//  * a work item is ONE path (pixel, sample), numbered sample-major so that consecutive items are
//    consecutive pixels of the same sample; lanes whose path ended refill together with ONE
//    atomicAdd per warp (__ballot_sync/__popc/__shfl_sync) and so pick up neighbouring pixels, which
//    keeps the warp's rays coherent (measured: 69% of FP32 peak vs 66% with 16-sample items) and
//    all 32 lanes scanning on every iteration until the queue is empty;
//  * a finished path publishes its radiance with one red.global.add.v4.f32 (16 B vector atomic
//    {r,g,b,1} into the W*H float4 accumulation buffer, which lives in L2);
//  * RNG: Philox4x32-10 keyed by seed, counter (pixel, sample, bounce, sub) -- see philox.cuh.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

#include "closest_hit.cuh"
#include "kernels.cuh"
#include "materials.cuh"
#include "philox.cuh"
#include "grid_walk.cuh"
#include "tc_collect.cuh"
#include "tc_self.cuh"

namespace uecrt {

// ---------------------------------------------------------------------------------------
// TMA bulk staging of the packed scene: global -> shared, completion on an mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_scene_tma(uint64_t* bar, void* dst, const void* src, uint32_t bytes) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    const uint32_t kChunk = 32768;   // bytes per bulk copy (multiple of 16)
    for (uint32_t off = 0; off < bytes; off += kChunk) {
      const uint32_t sz = bytes - off < kChunk ? bytes - off : kChunk;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32((char*)dst + off)),
                   "l"((const char*)src + off), "r"(sz), "r"(smem_u32(bar))
                   : "memory");
    }
  }
  // every thread waits for phase 0 of the barrier (HW sleep, not a spin on memory)
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar))
        : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Tile ring for scenes that do not fit in one CTA's shared memory.  Tiles are consumed in a fixed
// global order s = 0, 1, 2, ... (tile s % ntiles into buffer s & 1); the k-th fill of a buffer
// completes phase k of its full-barrier.  A buffer is refilled only after a __syncthreads() that
// follows every thread's scan of it.  `seq` is identical in all threads of the CTA.
struct TileRing {
  uint64_t* full;          // 2 mbarriers
  unsigned char* buf;      // 2 buffers of tile_bytes
  uint32_t tile_bytes;
  uint32_t seq;            // tiles consumed so far
};

__device__ __forceinline__ void ring_issue(const TileRing& r, const DeviceScene& sc, uint32_t s) {   // one thread
  const uint32_t b = s & 1u, tile = s % (uint32_t)sc.ntiles;
  const char* src = reinterpret_cast<const char*>(sc.pairs) + (size_t)tile * r.tile_bytes;
  const uint32_t bar = smem_u32(r.full + b);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(r.tile_bytes) : "memory");
  const uint32_t kChunk = 32768;
  for (uint32_t off = 0; off < r.tile_bytes; off += kChunk) {
    const uint32_t sz = r.tile_bytes - off < kChunk ? r.tile_bytes - off : kChunk;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(r.buf + (size_t)b * r.tile_bytes + off)),
                 "l"(src + off), "r"(sz), "r"(bar)
                 : "memory");
  }
}
__device__ __forceinline__ void ring_wait(const TileRing& r, uint32_t s) {   // every thread
  const uint32_t bar = smem_u32(r.full + (s & 1u)), parity = (s >> 1) & 1u;
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void ring_init(TileRing& r, const DeviceScene& sc, unsigned char* smem_raw) {
  r.full = reinterpret_cast<uint64_t*>(smem_raw);
  r.buf = smem_raw + 32;
  r.tile_bytes = (uint32_t)(sc.npairs + 2) * 32u;
  r.seq = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(r.full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(r.full + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) { ring_issue(r, sc, 0); ring_issue(r, sc, 1); }
}
// before the CTA exits: the two fills that are always in flight ahead of `seq` must land
__device__ __forceinline__ void ring_drain(const TileRing& r) { ring_wait(r, r.seq); ring_wait(r, r.seq + 1); }

__device__ __forceinline__ void red_add_v4(float4* addr, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// rare path of the scan: the candidate list is full -> fold it into the running best
template <int STRIDE>
__device__ __noinline__ void fold_full_fast(const float4* geom, uint16_t* base, int cnt, float ox, float oy,
                                            float oz, float dx, float dy, float dz, float tmin, int self_id,
                                            FastBest* best) {
  CandList<STRIDE> L{base};
  resolve_fast(geom, L, cnt, ox, oy, oz, dx, dy, dz, tmin, self_id, *best);
}
template <int STRIDE>
__device__ __noinline__ void fold_full_exact(const float4* geom, uint16_t* base, int cnt, double ox,
                                             double oy, double oz, double dx, double dy, double dz, double a,
                                             double tmin, ExactBest* best, int self_id = -1) {
  CandList<STRIDE> L{base};
  resolve_exact(geom, L, cnt, ox, oy, oz, dx, dy, dz, a, tmin, *best, self_id);
}

// Candidate collection over the whole scene for the R rays of a thread.  RESIDENT: one call of the
// scan on the resident tile, no synchronisation.  TILED: block-synchronous sweep; EVERY thread of
// the CTA must call this (threads without a live ray only take part in the barriers).
template <bool TILED, int R, int STRIDE, class OnFull>
__device__ __forceinline__ void collect_candidates(const DeviceScene& sc, const ulonglong2* s_pairs, TileRing& ring,
                                                   const CandList<STRIDE> (&L)[R], bool any_ray, const ScanRay (&sr)[R],
                                                   int (&cnt)[R], OnFull&& on_full) {
#pragma unroll
  for (int r = 0; r < R; ++r) cnt[r] = 0;
  if (!TILED) {
    scan_pairs<R>(s_pairs, sc.npairs, 0, sr, L, cnt, on_full);
    return;
  }
  for (int t = 0; t < sc.ntiles; ++t) {
    const uint32_t sq = ring.seq;
    ring_wait(ring, sq);
    if (any_ray)
      scan_pairs<R>(reinterpret_cast<const ulonglong2*>(ring.buf + (size_t)(sq & 1u) * ring.tile_bytes), sc.npairs,
                    t * sc.npairs * 2, sr, L, cnt, on_full);
    __syncthreads();                                      // every thread is done with this buffer
    if (threadIdx.x == 0) ring_issue(ring, sc, sq + 2);   // refill it with the tile after next
    ring.seq = sq + 1;
  }
}

// What one closest-hit query needs: unit direction, t in world units.
struct RayQuery {
  float ox, oy, oz, dx, dy, dz, tmin;
  int self_id;
  bool live;
};

// R closest-hit queries of one thread (the renderer's form).
template <bool EXACT, bool TILED, int R, int STRIDE>
__device__ __forceinline__ void closest_hits(const DeviceScene& sc, const ulonglong2* s_pairs, TileRing& ring,
                                             const CandList<STRIDE> (&L)[R], const RayQuery (&q)[R], Hit (&h)[R]) {
  ScanRay sr[R];
  int cnt[R];
  bool any = false;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    sr[r] = make_scan_ray(q[r].ox, q[r].oy, q[r].oz, q[r].dx, q[r].dy, q[r].dz);
    if (!q[r].live) sr[r].thr = 0x7f800000;     // +inf: a dead slot never produces a candidate
    h[r].id = -1;
    any = any || q[r].live;
  }
  if (!EXACT) {
    FastBest best[R];
#pragma unroll
    for (int r = 0; r < R; ++r) best[r] = FastBest{CUDART_INF_F, -1};
    collect_candidates<TILED, R>(sc, s_pairs, ring, L, any, sr, cnt, [&](int r, int c) {
      // r is a compile-time constant after unrolling; the switch keeps best[] in registers
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (k == r)
          fold_full_fast<STRIDE>(sc.geom, L[k].base, c, q[k].ox, q[k].oy, q[k].oz, q[k].dx, q[k].dy, q[k].dz, q[k].tmin,
                                 q[k].self_id, &best[k]);
    });
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!q[r].live) continue;
      resolve_fast(sc.geom, L[r], cnt[r], q[r].ox, q[r].oy, q[r].oz, q[r].dx, q[r].dy, q[r].dz, q[r].tmin, q[r].self_id, best[r]);
      if (best[r].id >= 0) {
        h[r].id = best[r].id;
        h[r].t = best[r].t;
        finish_hit_fast(sc.geom, q[r].ox, q[r].oy, q[r].oz, q[r].dx, q[r].dy, q[r].dz, best[r].far_root, h[r]);
      }
    }
  } else {
    ExactBest best[R];
    double a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      best[r] = ExactBest{(double)CUDART_INF, -1};
      const double Dx = q[r].dx, Dy = q[r].dy, Dz = q[r].dz;
      a[r] = __dadd_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(Dy, Dy)), __dmul_rn(Dz, Dz));
    }
    collect_candidates<TILED, R>(sc, s_pairs, ring, L, any, sr, cnt, [&](int r, int c) {
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (k == r)
          fold_full_exact<STRIDE>(sc.geom, L[k].base, c, (double)q[k].ox, (double)q[k].oy, (double)q[k].oz, (double)q[k].dx,
                                  (double)q[k].dy, (double)q[k].dz, a[k], (double)q[k].tmin, &best[k], q[k].self_id);
    });
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!q[r].live) continue;
      const double Ox = q[r].ox, Oy = q[r].oy, Oz = q[r].oz, Dx = q[r].dx, Dy = q[r].dy, Dz = q[r].dz;
      resolve_exact(sc.geom, L[r], cnt[r], Ox, Oy, Oz, Dx, Dy, Dz, a[r], (double)q[r].tmin, best[r], q[r].self_id);
      if (best[r].id >= 0) {
        const float4 g = __ldg(&sc.geom[best[r].id]);
        const double px = Ox + Dx * best[r].t, py = Oy + Dy * best[r].t, pz = Oz + Dz * best[r].t;   // ray.hpp:16-19
        const double rad = g.w;
        double nx = (px - (double)g.x) / rad, ny = (py - (double)g.y) / rad, nz = (pz - (double)g.z) / rad;   // sphere.hpp:44
        const bool front = (Dx * nx + Dy * ny + Dz * nz) < 0;                                     // hittable.hpp:25
        if (!front) { nx = -nx; ny = -ny; nz = -nz; }
        h[r].id = best[r].id; h[r].t = (float)best[r].t; h[r].front = front;
        h[r].px = (float)px; h[r].py = (float)py; h[r].pz = (float)pz;
        h[r].nx = (float)nx; h[r].ny = (float)ny; h[r].nz = (float)nz;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// K1
// Per-slot path state: a thread owns RAYS independent paths ("slots").
struct PathState {
  bool active, done, valid;
  uint32_t pixel, sample, bounce, depth_left;
  int self_id;
  float ox, oy, oz, dx, dy, dz, dlen;
  float thr_r, thr_g, thr_b;
};

// work item w -> (pixel, sample) of the path this slot runs next, or done when the queue is exhausted.
// w = sample * n_slots + slot, slot = (tile << 5) | within: tiles of 8x4 pixels in row-major tile
// order, so the 32 lanes of a warp start on a compact 8x4 pixel block (coherent rays).
// FP64 / FP32 reciprocal estimates + fix-up instead of integer divides.
__device__ __forceinline__ void slot_assign(PathState& p, const RenderArgs& a, unsigned long long w,
                                            unsigned long long n_pixels, double inv_np) {
  if (w < a.total_items) {
    w += a.item_begin;
    unsigned long long q = __double2ull_rz(__ull2double_rz(w) * inv_np);
    long long rem = (long long)(w - q * n_pixels);
    if (rem < 0) { --q; rem += (long long)n_pixels; }
    if (rem >= (long long)n_pixels) { ++q; rem -= (long long)n_pixels; }
    const uint32_t slot = (uint32_t)rem, tile = slot >> 5, within = slot & 31u;
    uint32_t ty = (uint32_t)__fmul_rz((float)tile, a.inv_tiles_x);
    uint32_t tx = tile - ty * a.tiles_x;
    if ((int32_t)tx < 0) { --ty; tx += a.tiles_x; }
    if (tx >= a.tiles_x) { ++ty; tx -= a.tiles_x; }
    const uint32_t x = tx * 8u + (within & 7u), y = ty * 4u + (within >> 3);
    p.pixel = y * a.width + x;
    p.sample = a.sample_begin + (uint32_t)q;
    p.valid = (x < a.width) && (y < a.height);      // ragged right / bottom edge of the tile grid
  } else {
    p.done = true;
  }
}

__device__ __forceinline__ void slot_start_path(PathState& p, const RenderArgs& a, bool need, float inv_w, float inv_h,
                                                unsigned& n_paths);

// refill + start of the next path for one slot; warp-convergent (uses __ballot_sync)
__device__ __forceinline__ void slot_refill_and_start(PathState& p, const RenderArgs& a, unsigned lane,
                                                      unsigned long long n_pixels, double inv_np, float inv_w,
                                                      float inv_h, unsigned& n_paths) {
  // lanes whose path ended pull the next paths together: one atomicAdd per warp
  const bool need = !p.active && !p.done;
  const unsigned need_mask = __ballot_sync(0xffffffffu, need);
  if (need_mask) {
    const int leader = __ffs(need_mask) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(a.counters, (unsigned long long)__popc(need_mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (need) slot_assign(p, a, base + __popc(need_mask & ((1u << lane) - 1u)), n_pixels, inv_np);
  }
  slot_start_path(p, a, need, inv_w, inv_h, n_paths);
}

// Work items of a warp drawn from a range claimed one round AHEAD (tensor-scan K1): the global
// atomicAdd that claims the next 32 items is issued when the previous claim is opened, and its
// ~700-cycle round trip to L2 overlaps a whole scan instead of stalling the warp every round.
// cur: [next, end) being handed out; nxt: start of the claimed-but-unopened range of kBatch items.
struct WorkPool {
  static constexpr unsigned kBatch = 32;
  unsigned long long next, end, nxt;
};
__device__ __forceinline__ void pool_init(WorkPool& wp, const RenderArgs& a, unsigned lane) {
  unsigned long long b0 = 0, b1 = 0;
  if (lane == 0) { b0 = atomicAdd(a.counters, (unsigned long long)WorkPool::kBatch); b1 = atomicAdd(a.counters, (unsigned long long)WorkPool::kBatch); }
  wp.next = __shfl_sync(0xffffffffu, b0, 0);
  wp.end = wp.next + WorkPool::kBatch;
  wp.nxt = __shfl_sync(0xffffffffu, b1, 0);
}
__device__ __forceinline__ void slot_refill_pooled(PathState& p, WorkPool& wp, const RenderArgs& a, unsigned lane,
                                                   unsigned long long n_pixels, double inv_np, float inv_w, float inv_h,
                                                   unsigned& n_paths) {
  const bool need = !p.active && !p.done;
  const unsigned need_mask = __ballot_sync(0xffffffffu, need);
  if (need_mask) {
    const unsigned n = __popc(need_mask), k = __popc(need_mask & ((1u << lane) - 1u));
    const unsigned long long avail = wp.end - wp.next;
    const unsigned long long w = k < avail ? wp.next + k : wp.nxt + (k - avail);
    if (n <= avail) {
      wp.next += n;
    } else {                                   // open the range claimed earlier, claim the one after it
      wp.next = wp.nxt + (n - avail);
      wp.end = wp.nxt + WorkPool::kBatch;
      unsigned long long b = 0;
      if (lane == 0) b = atomicAdd(a.counters, (unsigned long long)WorkPool::kBatch);
      wp.nxt = __shfl_sync(0xffffffffu, b, 0);   // consumed at the earliest one round from now
    }
    if (need) slot_assign(p, a, w, n_pixels, inv_np);
  }
  slot_start_path(p, a, need, inv_w, inv_h, n_paths);
}

// the camera ray of a fresh path from its Philox block (pixel, sample, 0, 0) (source.cpp:154-165, camera.hpp:29-32)
__device__ __forceinline__ void slot_camera_ray(PathState& p, const RenderArgs& a, const philox_out r, float inv_w, float inv_h, unsigned& n_paths) {
  // pixel -> (row, column) without the ~20-instruction runtime integer divide: float reciprocal
  // estimate (exact to +-1 for pixel < 2^24 * width) and one fix-up step each way
  uint32_t py = (uint32_t)__fmul_rz((float)p.pixel, inv_w);
  uint32_t px = p.pixel - py * a.width;
  if ((int32_t)px < 0) { --py; px += a.width; }
  if (px >= a.width) { ++py; px -= a.width; }
  const float u = ((float)px + uecrt_u01(r.x)) * inv_w;
  const float v = ((float)(a.height - py - 1u) + uecrt_u01(r.y)) * inv_h;
  float tx = fmaf(v, a.cam.ver[0], fmaf(u, a.cam.hor[0], a.cam.ll[0]));
  float ty = fmaf(v, a.cam.ver[1], fmaf(u, a.cam.hor[1], a.cam.ll[1]));
  float tz = fmaf(v, a.cam.ver[2], fmaf(u, a.cam.hor[2], a.cam.ll[2]));
  p.ox = a.cam.origin[0]; p.oy = a.cam.origin[1]; p.oz = a.cam.origin[2];
  if (a.cam.lens > 0.f) {
    // uniform point in the unit disk from the block's last two uniforms (polar map: r = sqrt(u),
    // phi = 2*pi*v).  Same distribution as the RTIOW random_in_unit_disk rejection loop of the
    // book (and of the CPU checker), but exactly two uniforms and no divergent retry loop.
    const float rr = sqrt_fast(uecrt_u01(r.z)), phi = 6.28318530717958647692f * uecrt_u01(r.w);
    const float rx = rr * __cosf(phi), ry = rr * __sinf(phi);
    const float lx = a.cam.lens * rx, ly = a.cam.lens * ry;
    const float offx = fmaf(ly, a.cam.v[0], lx * a.cam.u[0]);
    const float offy = fmaf(ly, a.cam.v[1], lx * a.cam.u[1]);
    const float offz = fmaf(ly, a.cam.v[2], lx * a.cam.u[2]);
    tx = tx - p.ox - offx; ty = ty - p.oy - offy; tz = tz - p.oz - offz;
    p.ox += offx; p.oy += offy; p.oz += offz;
  } else if (!a.cam.origin_is_zero) {
    tx -= p.ox; ty -= p.oy; tz -= p.oz;
  }
  const float len2 = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
  const float inv = rsqrt_nr(len2);
  p.dlen = len2 * inv;
  p.dx = tx * inv; p.dy = ty * inv; p.dz = tz * inv;
  p.thr_r = p.thr_g = p.thr_b = 1.f;
  p.depth_left = a.max_depth;
  p.self_id = -1;
  p.bounce = 0;
  p.active = p.depth_left > 0;
  if (!p.active) {                               // depth 0: black (raytracer.hpp:26)
    red_add_v4(a.accum + p.pixel, 0.f, 0.f, 0.f, 1.f);
    ++n_paths;
  }
}

__device__ __forceinline__ void slot_start_path(PathState& p, const RenderArgs& a, bool need, float inv_w, float inv_h,
                                                unsigned& n_paths) {
  if (need && !p.done && p.valid) slot_camera_ray(p, a, philox4x32_10(p.pixel, p.sample, 0u, 0u, a.key0, a.key1), inv_w, inv_h, n_paths);
}

// one path segment after its closest hit is known: scatter or terminate (raytracer.hpp:27-36)
__device__ __forceinline__ void slot_shade(PathState& p, const Hit& h, const RenderArgs& a, unsigned& n_paths) {
  float rad_r = 0.f, rad_g = 0.f, rad_b = 0.f;
  bool ended = true;
  if (h.id < 0) {
    const float t = 0.5f * (p.dy + 1.0f);                       // raytracer.hpp:35
    const float w0 = 1.0f - t;
    rad_r = p.thr_r * fmaf(t, 0.5f, w0); rad_g = p.thr_g * fmaf(t, 0.7f, w0); rad_b = p.thr_b * (w0 + t);
  } else {
    ++p.bounce;
    const uint32_t mt = __ldg(&a.sc.mat_type[h.id]);
    const float4 mp = __ldg(&a.sc.mat_param[h.id]);
    float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
    if (mt != UECRT_MAT_METAL || mp.w > 0.f) {
      const philox_out r = philox4x32_10(p.pixel, p.sample, p.bounce, 0u, a.key0, a.key1);
      u0 = uecrt_u01(r.x); u1 = uecrt_u01(r.y); u2 = uecrt_u01(r.z); u3 = uecrt_u01(r.w);
    }
    const ScatterOut s = scatter_material(mt, mp, p.dx, p.dy, p.dz, h.nx, h.ny, h.nz, h.front, u0, u1, u2, u3);
    if (s.ok && p.depth_left > 1u) {                             // depth-1 == 0 returns black (raytracer.hpp:26)
      p.thr_r *= s.ar; p.thr_g *= s.ag; p.thr_b *= s.ab;
      p.ox = h.px; p.oy = h.py; p.oz = h.pz;
      const float len2 = fmaf(s.dz, s.dz, fmaf(s.dy, s.dy, s.dx * s.dx));
      const float inv = rsqrt_nr(len2);
      p.dlen = len2 * inv;
      p.dx = s.dx * inv; p.dy = s.dy * inv; p.dz = s.dz * inv;
      p.self_id = h.id;
      --p.depth_left;
      ended = false;
    }
  }
  if (ended) {
    // a NaN radiance would hang the reference's math::sqrt (math.hpp:14); it contributes 0 here
    red_add_v4(a.accum + p.pixel, (rad_r == rad_r) ? rad_r : 0.f, (rad_g == rad_g) ? rad_g : 0.f,
               (rad_b == rad_b) ? rad_b : 0.f, 1.f);
    ++n_paths;
    p.active = false;
  }
}

// Shade + refill of one slot with ONE Philox block per round (self-scan K1).  slot_shade draws the scatter's uniforms and
// slot_refill_pooled / slot_start_path the camera ray's in two separately diverged regions (each ran with ~14 of 32 lanes
// and 70 instructions of Philox4x32-10).  Path ends that need no random number -- a miss, or a hit with no depth left --
// are known before the draw, so those lanes take their next work item first and the block is computed once for both kinds
// of lanes: counter (pixel, sample, bounce) for a scattering lane, (pixel', sample', 0) for a fresh one.  Same counters,
// same paths; a lane whose scatter is absorbed (metal below the surface) is refilled one round later.  WARP-CONVERGENT.
__device__ __forceinline__ void slot_shade_refill_fused(PathState& p, const Hit& h, const RenderArgs& a, WorkPool& wp, unsigned lane,
                                                        unsigned long long n_pixels, double inv_np, float inv_w, float inv_h,
                                                        unsigned& n_rays, unsigned& n_paths) {
  bool scatter_lane = false;
  if (p.active) {
    ++n_rays;
    if (h.id < 0) {                                                // raytracer.hpp:35: the sky
      const float t = 0.5f * (p.dy + 1.0f), w0 = 1.0f - t;
      const float rr = p.thr_r * fmaf(t, 0.5f, w0), rg = p.thr_g * fmaf(t, 0.7f, w0), rb = p.thr_b * (w0 + t);
      red_add_v4(a.accum + p.pixel, (rr == rr) ? rr : 0.f, (rg == rg) ? rg : 0.f, (rb == rb) ? rb : 0.f, 1.f);
      ++n_paths;
      p.active = false;
    } else if (p.depth_left <= 1u) {                               // depth - 1 == 0 returns black whatever the scatter does (raytracer.hpp:26)
      red_add_v4(a.accum + p.pixel, 0.f, 0.f, 0.f, 1.f);
      ++n_paths;
      p.active = false;
    } else {
      scatter_lane = true;
    }
  }
  // lanes without a path take the next work items together (slot_refill_pooled)
  const bool need = !p.active && !p.done;
  const unsigned need_mask = __ballot_sync(0xffffffffu, need);
  if (need_mask) {
    const unsigned n = __popc(need_mask), k = __popc(need_mask & ((1u << lane) - 1u));
    const unsigned long long avail = wp.end - wp.next;
    const unsigned long long w = k < avail ? wp.next + k : wp.nxt + (k - avail);
    if (n <= avail) {
      wp.next += n;
    } else {
      wp.next = wp.nxt + (n - avail);
      wp.end = wp.nxt + WorkPool::kBatch;
      unsigned long long b = 0;
      if (lane == 0) b = atomicAdd(a.counters, (unsigned long long)WorkPool::kBatch);
      wp.nxt = __shfl_sync(0xffffffffu, b, 0);
    }
    if (need) slot_assign(p, a, w, n_pixels, inv_np);
  }
  const bool fresh = need && !p.done && p.valid;
  uint32_t mt = 0;
  float4 mp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scatter_lane) {
    ++p.bounce;
    mt = __ldg(&a.sc.mat_type[h.id]);
    mp = __ldg(&a.sc.mat_param[h.id]);
  }
  philox_out r{0u, 0u, 0u, 0u};
  if (fresh || (scatter_lane && (mt != UECRT_MAT_METAL || mp.w > 0.f)))
    r = philox4x32_10(p.pixel, p.sample, fresh ? 0u : p.bounce, 0u, a.key0, a.key1);
  if (fresh) {
    slot_camera_ray(p, a, r, inv_w, inv_h, n_paths);
  } else if (scatter_lane) {
    const ScatterOut s = scatter_material(mt, mp, p.dx, p.dy, p.dz, h.nx, h.ny, h.nz, h.front, uecrt_u01(r.x), uecrt_u01(r.y), uecrt_u01(r.z), uecrt_u01(r.w));
    if (s.ok) {
      p.thr_r *= s.ar; p.thr_g *= s.ag; p.thr_b *= s.ab;
      p.ox = h.px; p.oy = h.py; p.oz = h.pz;
      const float len2 = fmaf(s.dz, s.dz, fmaf(s.dy, s.dy, s.dx * s.dx));
      const float inv = rsqrt_nr(len2);
      p.dlen = len2 * inv;
      p.dx = s.dx * inv; p.dy = s.dy * inv; p.dz = s.dz * inv;
      p.self_id = h.id;
      --p.depth_left;
    } else {                                                       // absorbed: black; the lane is refilled next round
      red_add_v4(a.accum + p.pixel, 0.f, 0.f, 0.f, 1.f);
      ++n_paths;
      p.active = false;
    }
  }
}

#ifndef UECRT_K1_MINB
#define UECRT_K1_MINB 3          // CTAs per SM the register allocator must allow for the 1-ray fast kernel
#endif
// shared memory: RESIDENT [16 B mbarrier][(npairs+2)*32 B pairs][lists]; TILED [32 B: 2 mbarriers][2 tile buffers][lists]
// lists: RAYS x kListCap x 256 uint16
template <bool EXACT, bool TILED, int RAYS>
__global__ void __launch_bounds__(kRenderThreads, (EXACT || RAYS > 1) ? 2 : UECRT_K1_MINB) k_render_paths(const RenderArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t pair_bytes = (uint32_t)(a.sc.npairs + 2) * 32u;   // one tile + its prefetch pad group
  ulonglong2* s_pairs = reinterpret_cast<ulonglong2*>(smem_raw + 16);
  uint16_t* s_lists = reinterpret_cast<uint16_t*>(smem_raw + (TILED ? 32u + 2u * pair_bytes : 16u + pair_bytes));
  TileRing ring{};
  if (TILED) ring_init(ring, a.sc, smem_raw);
  else stage_scene_tma(reinterpret_cast<uint64_t*>(smem_raw), s_pairs, a.sc.pairs, pair_bytes);

  CandList<kRenderThreads> L[RAYS];
#pragma unroll
  for (int s = 0; s < RAYS; ++s) L[s].base = s_lists + (size_t)s * kListCap * kRenderThreads + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long n_pixels = a.n_slots;     // pixel slots of the 8x4 tile grid (>= width*height)
  const double inv_np = 1.0 / (double)n_pixels;
  const float inv_w = 1.0f / (float)a.width, inv_h = 1.0f / (float)a.height;

  PathState st[RAYS];
#pragma unroll
  for (int s = 0; s < RAYS; ++s) {
    st[s] = PathState{};
    st[s].dz = 1.f; st[s].dlen = 1.f; st[s].self_id = -1;
  }
  unsigned n_rays = 0, n_paths = 0;

  for (;;) {
    bool all_done = true;
#pragma unroll
    for (int s = 0; s < RAYS; ++s) {
      slot_refill_and_start(st[s], a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths);
      all_done = all_done && st[s].done;
    }
    if (TILED ? __syncthreads_and(all_done) : __all_sync(0xffffffffu, all_done)) break;   // TILED: the whole CTA sweeps together

    // ---- closest hits of this thread's live slots: ONE scan over the scene for all of them
    RayQuery q[RAYS];
    Hit h[RAYS];
    bool any_active = false;
#pragma unroll
    for (int s = 0; s < RAYS; ++s) {
      q[s] = RayQuery{st[s].ox, st[s].oy, st[s].oz, st[s].dx, st[s].dy, st[s].dz, 0.001f * st[s].dlen, st[s].self_id, st[s].active};
      h[s].id = -1;
      any_active = any_active || st[s].active;
    }
    if (TILED || any_active) closest_hits<EXACT, TILED, RAYS>(a.sc, s_pairs, ring, L, q, h);
#pragma unroll
    for (int s = 0; s < RAYS; ++s)
      if (st[s].active) {
        ++n_rays;
        slot_shade(st[s], h[s], a, n_paths);
      }
  }

  if (TILED) ring_drain(ring);
  if (a.count_stats) {
    unsigned r = n_rays, p = n_paths;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); p += __shfl_xor_sync(0xffffffffu, p, o); }
    if (lane == 0) {
      atomicAdd(a.counters + 1, (unsigned long long)r);
      atomicAdd(a.counters + 2, (unsigned long long)p);
    }
  }
}

// ---------------------------------------------------------------------------------------
// K2: closest hit on caller rays.  One thread per ray, same scan + resolve as K1.  In TILED mode the
// CTA sweeps the tile ring once per batch of blockDim rays, so every thread runs the same number of
// batches (threads past the end of the ray array carry has_ray = false).
template <bool EXACT, bool TILED>
__global__ void __launch_bounds__(kHitsThreads) k_primary_hits(const DeviceScene sc, unsigned long long n_rays,
                                                               const float* __restrict__ o, const float* __restrict__ d,
                                                               double tmin, double tmax, double* __restrict__ t_out,
                                                               double* __restrict__ n_out, int32_t* __restrict__ id_out,
                                                               uint8_t* __restrict__ ff_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t pair_bytes = (uint32_t)(sc.npairs + 2) * 32u;     // one tile + its prefetch pad group
  ulonglong2* s_pairs = reinterpret_cast<ulonglong2*>(smem_raw + 16);
  uint16_t* s_lists = reinterpret_cast<uint16_t*>(smem_raw + (TILED ? 32u + 2u * pair_bytes : 16u + pair_bytes));
  TileRing ring{};
  if (TILED) ring_init(ring, sc, smem_raw);
  else stage_scene_tma(reinterpret_cast<uint64_t*>(smem_raw), s_pairs, sc.pairs, pair_bytes);
  CandList<kHitsThreads> L[1];
  L[0].base = s_lists + threadIdx.x;

  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long batches = (n_rays + stride - 1) / stride;
  for (unsigned long long b = 0; b < batches; ++b) {
    const unsigned long long i = b * stride + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool has_ray = i < n_rays;
    if (!TILED && !has_ray) break;
    const unsigned long long j = has_ray ? i : 0;
    const float ox = o[3 * j], oy = o[3 * j + 1], oz = o[3 * j + 2];
    const float rx = d[3 * j], ry = d[3 * j + 1], rz = d[3 * j + 2];
    const float dlen = sqrtf(fmaf(rz, rz, fmaf(ry, ry, rx * rx)));
    const float inv = 1.0f / dlen;
    const float dx = rx * inv, dy = ry * inv, dz = rz * inv;
    ScanRay sr[1] = {make_scan_ray(ox, oy, oz, dx, dy, dz)};
    if (!has_ray) sr[0].thr = 0x7f800000;
    int cnt[1];
    double t = -1.0, nx = 0.0, ny = 0.0, nz = 0.0;
    int id = -1;
    bool front = false;
    if (!EXACT) {
      const float tmin_w = (float)tmin * dlen;
      FastBest best{(float)tmax * dlen, -1};
      collect_candidates<TILED, 1>(sc, s_pairs, ring, L, has_ray, sr, cnt, [&](int, int c) {
        fold_full_fast<kHitsThreads>(sc.geom, L[0].base, c, ox, oy, oz, dx, dy, dz, tmin_w, -1, &best);
      });
      if (has_ray) resolve_fast(sc.geom, L[0], cnt[0], ox, oy, oz, dx, dy, dz, tmin_w, -1, best);
      if (has_ray && best.id >= 0) {
        // the winner's root and normal in double for the CALLER's direction (closest_hit.cuh refine_winner)
        const float4 gw = __ldg(&sc.geom[best.id]);
        const HitD w = refine_winner<false>(gw, ox, oy, oz, rx, ry, rz, !best.far_root);
        id = best.id; t = w.t; nx = w.nx; ny = w.ny; nz = w.nz; front = w.front;
      }
    } else {
      const double Ox = ox, Oy = oy, Oz = oz, Dx = rx, Dy = ry, Dz = rz;   // ORIGINAL direction: t == reference root
      const double a = __dadd_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(Dy, Dy)), __dmul_rn(Dz, Dz));  // sphere.hpp:28
      ExactBest best{tmax, -1};
      collect_candidates<TILED, 1>(sc, s_pairs, ring, L, has_ray, sr, cnt, [&](int, int c) {
        fold_full_exact<kHitsThreads>(sc.geom, L[0].base, c, Ox, Oy, Oz, Dx, Dy, Dz, a, tmin, &best);
      });
      if (has_ray) resolve_exact(sc.geom, L[0], cnt[0], Ox, Oy, Oz, Dx, Dy, Dz, a, tmin, best);
      if (has_ray && best.id >= 0) {
        const float4 g = __ldg(&sc.geom[best.id]);
        const double px = __dadd_rn(Ox, __dmul_rn(Dx, best.t)), py = __dadd_rn(Oy, __dmul_rn(Dy, best.t)),
                     pz = __dadd_rn(Oz, __dmul_rn(Dz, best.t));                                   // ray.hpp:16-19
        const double r = g.w;
        nx = __ddiv_rn(__dsub_rn(px, (double)g.x), r);                                             // sphere.hpp:44
        ny = __ddiv_rn(__dsub_rn(py, (double)g.y), r);
        nz = __ddiv_rn(__dsub_rn(pz, (double)g.z), r);
        front = __dadd_rn(__dadd_rn(__dmul_rn(Dx, nx), __dmul_rn(Dy, ny)), __dmul_rn(Dz, nz)) < 0;  // hittable.hpp:25
        if (!front) { nx = -nx; ny = -ny; nz = -nz; }
        id = best.id; t = best.t;
      }
    }
    if (has_ray) {
      t_out[i] = t; n_out[3 * i] = nx; n_out[3 * i + 1] = ny; n_out[3 * i + 2] = nz;
      id_out[i] = id; ff_out[i] = front ? 1 : 0;
    }
  }
  if (TILED) ring_drain(ring);
}

// ---------------------------------------------------------------------------------------
// K1 / K2 on the TENSOR-CORE scan (tc_scan.cuh, tc_collect.cuh).  Same path loop, same resolve and
// shading as above; only the all-pairs discriminant evaluation moves from the FP32 pipe to
// tcgen05.mma.kind::tf32 (three-way TF32 split, FP32 accumulation in TMEM), and the CTA is
// warp-specialised: 4 shade groups of 128 threads, one scan warpgroup, one issuer warp (1 CTA/SM).
#ifndef UECRT_TC_ISSUER_WARP
#define UECRT_TC_ISSUER_WARP 0     // which warp of the last warpgroup issues the MMAs (= the SM sub-partition it shares with a scan warp)
#endif
#ifdef UECRT_TC_PROFILE
struct TcProf { long long t[8]; long long last; };
#define UECRT_TCP(p, i) { __syncwarp(); const long long now_ = clock64(); (p).t[i] += now_ - (p).last; (p).last = now_; }
#else
struct TcProf {};
#define UECRT_TCP(p, i)
#endif

// closest hit of one shade thread's ray through the pipeline; every shade thread calls it once per round.
// `m` is the pipeline's pass counter (tc_scan.cuh PassIter): one pass per round when the sphere operand is
// resident; SWEEP: npasses passes of pass_rows spheres each, folded into one running closest hit in
// ascending sphere index, so the later-index-wins tie rule holds across passes as it does inside one.
// candidates of one flagged quad (tc_collect.cuh): the FP32 scan's packed discriminant test on its four spheres, then
// the resolve test on the survivors in ascending index
template <class Test>
__device__ __forceinline__ void tc_quad(const DeviceScene& sc, int quad, const tc::QuadRay& qr, Test&& test) {
  uint32_t cm = tc::quad_candidates(sc.tc_quads, quad, qr);
  while (cm) {
    const int j = __ffs((int)cm) - 1;
    cm &= cm - 1u;
    test(4 * quad + j);
  }
}

template <bool EXACT, bool SWEEP, class InWindow>
__device__ __forceinline__ void tc_closest_hit(tc::Shade& sh, uint32_t& m, uint32_t npasses, int pass_rows, const DeviceScene& sc,
                                               const RayQuery& q, Hit& h, TcProf& prof, InWindow&& in_window) {
  const ScanRay sr = make_scan_ray(q.ox, q.oy, q.oz, q.dx, q.dy, q.dz);
  h.id = -1;
  tc::shade_submit(sh, sr, q.live);
  UECRT_TCP(prof, 2)
  in_window();                                    // work that does not depend on this round's hit
  UECRT_TCP(prof, 1)
  const tc::QuadRay qr = tc::make_quad_ray(sr);
  const int pass_quads = pass_rows / tc::kQuad;   // bitmap bits per pass
  if (!EXACT) {
    FastBest best{CUDART_INF_F, -1};
    for (uint32_t p = 0, base = 0;; base += (uint32_t)pass_quads) {
      tc::shade_wait(sh, m++);
      UECRT_TCP(prof, 3)
      tc::for_each_flag(sh, SWEEP ? 0u : ((m - 1u) & 1u), q.live, [&](int k) {
        auto test = [&](int j) { test_fast(sc.geom, j, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz, q.tmin, q.self_id, best); };
        if (tc::kQuad == 4) tc_quad(sc, (int)base + k, qr, test); else test((int)base + k);
      });
      if (!SWEEP || ++p == npasses) break;
      tc::shade_next_pass(sh);
    }
    if (best.id >= 0) {
      h.id = best.id;
      h.t = best.t;
      finish_hit_fast(sc.geom, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz, best.far_root, h);
    }
  } else {
    ExactBest best{(double)CUDART_INF, -1};
    const double Ox = q.ox, Oy = q.oy, Oz = q.oz, Dx = q.dx, Dy = q.dy, Dz = q.dz;
    const double a = __dadd_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(Dy, Dy)), __dmul_rn(Dz, Dz));
    for (uint32_t p = 0, base = 0;; base += (uint32_t)pass_quads) {
      tc::shade_wait(sh, m++);
      UECRT_TCP(prof, 3)
      tc::for_each_flag(sh, SWEEP ? 0u : ((m - 1u) & 1u), q.live, [&](int k) {
        auto test = [&](int j) { test_exact(sc.geom, j, Ox, Oy, Oz, Dx, Dy, Dz, a, (double)q.tmin, best, q.self_id); };
        if (tc::kQuad == 4) tc_quad(sc, (int)base + k, qr, test); else test((int)base + k);
      });
      if (!SWEEP || ++p == npasses) break;
      tc::shade_next_pass(sh);
    }
    if (best.id >= 0) {
      const float4 g = __ldg(&sc.geom[best.id]);
      const double px = Ox + Dx * best.t, py = Oy + Dy * best.t, pz = Oz + Dz * best.t;   // ray.hpp:16-19
      const double rad = g.w;
      double nx = (px - (double)g.x) / rad, ny = (py - (double)g.y) / rad, nz = (pz - (double)g.z) / rad;   // sphere.hpp:44
      const bool front = (Dx * nx + Dy * ny + Dz * nz) < 0;                                     // hittable.hpp:25
      if (!front) { nx = -nx; ny = -ny; nz = -nz; }
      h.id = best.id; h.t = (float)best.t; h.front = front;
      h.px = (float)px; h.py = (float)py; h.pz = (float)pz;
      h.nx = (float)nx; h.ny = (float)ny; h.nz = (float)nz;
    }
  }
}

// The collect half alone (resident operand): the ray was submitted one tile earlier, its quad flags are in bitmap buffer `buf`.
// s_quads / s_geom: shared addresses of the CTA's copies of the quad table and the sphere array (tc_scan.cuh Layout).
// plane = 0: the table is the global one's copy (64 B per quad); plane > 0: planar copy (tc_collect.cuh quad_candidates_sp)
template <class Test>
__device__ __forceinline__ void tc_quad_s(uint32_t s_quads, uint32_t plane, int quad, const tc::QuadRay& qr, Test&& test) {
  uint32_t cm = plane ? tc::quad_candidates_sp(s_quads, plane, quad, qr) : tc::quad_candidates_s(s_quads, quad, qr);
  while (cm) {
    const int j = __ffs((int)cm) - 1;
    cm &= cm - 1u;
    test(4 * quad + j);
  }
}
// fe(f): calls f(flag index) for every set flag of the ray, ascending, warp-convergently (tc::for_each_flag / ts::for_each_flag)
template <bool EXACT, class ForEach>
__device__ __forceinline__ void tc_collect_hit_with(ForEach&& fe, uint32_t s_quads, uint32_t plane, uint32_t s_geom, const RayQuery& q, Hit& h) {
  h.id = -1;
  const tc::QuadRay qr = tc::make_quad_ray(make_scan_ray(q.ox, q.oy, q.oz, q.dx, q.dy, q.dz));
  if (!EXACT) {
    FastBest best{CUDART_INF_F, -1};
    fe([&](int k) {
      auto test = [&](int j) { test_fast_g(tc::lds_f4(s_geom + 16u * (uint32_t)j), j, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz, q.tmin, q.self_id, best); };
      if (tc::kQuad == 4) tc_quad_s(s_quads, plane, k, qr, test); else test(k);
    });
    if (best.id >= 0) {
      h.id = best.id;
      h.t = best.t;
      finish_hit_fast_g(tc::lds_f4(s_geom + 16u * (uint32_t)best.id), q.ox, q.oy, q.oz, q.dx, q.dy, q.dz, best.far_root, h);
    }
  } else {
    ExactBest best{(double)CUDART_INF, -1};
    const double Ox = q.ox, Oy = q.oy, Oz = q.oz, Dx = q.dx, Dy = q.dy, Dz = q.dz;
    const double a = __dadd_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(Dy, Dy)), __dmul_rn(Dz, Dz));
    fe([&](int k) {
      auto test = [&](int j) { test_exact_g(tc::lds_f4(s_geom + 16u * (uint32_t)j), j, Ox, Oy, Oz, Dx, Dy, Dz, a, (double)q.tmin, best, q.self_id); };
      if (tc::kQuad == 4) tc_quad_s(s_quads, plane, k, qr, test); else test(k);
    });
    if (best.id >= 0) {
      const float4 g = tc::lds_f4(s_geom + 16u * (uint32_t)best.id);
      const double px = Ox + Dx * best.t, py = Oy + Dy * best.t, pz = Oz + Dz * best.t;   // ray.hpp:16-19
      const double rad = g.w;
      double nx = (px - (double)g.x) / rad, ny = (py - (double)g.y) / rad, nz = (pz - (double)g.z) / rad;   // sphere.hpp:44
      const bool front = (Dx * nx + Dy * ny + Dz * nz) < 0;                                     // hittable.hpp:25
      if (!front) { nx = -nx; ny = -ny; nz = -nz; }
      h.id = best.id; h.t = (float)best.t; h.front = front;
      h.px = (float)px; h.py = (float)py; h.pz = (float)pz;
      h.nx = (float)nx; h.ny = (float)ny; h.nz = (float)nz;
    }
  }
}
template <bool EXACT>
__device__ __forceinline__ void tc_collect_hit(tc::Shade& sh, uint32_t buf, uint32_t s_quads, uint32_t s_geom, const RayQuery& q, Hit& h) {
  tc_collect_hit_with<EXACT>([&](auto&& f) { tc::for_each_flag(sh, buf, q.live, f); }, s_quads, 0u, s_geom, q, h);
}

// SHADE role of K1, resident operand: TWO independent paths per thread.  While the tile that carries path Q goes
// through the MMA + scan pipeline, the thread resolves and shades path P from the bitmap of ITS last tile, so the
// closed loop of four groups around one scanner always has a tile to scan (round 1: one path per thread, each
// group waited 7.3 k of a 13.5 k-cycle round for its own tile).  Tile m of a group carries the group's P paths
// for even m and its Q paths for odd m; its quad flags go to bitmap buffer m & 1.  Per tile m the thread
//   waits for tile m's flags (so the tile's MMAs have completed and the group's A rows in TMEM are free),
//   submits the OTHER path's ray as tile m + 1,
//   then resolves and shades tile m's path, which leaves that path's next ray (or the next path's camera ray).
template <bool EXACT>
__device__ __forceinline__ void shade_role_resident(const RenderArgs& a, unsigned char* smem_raw, uint32_t tmem_base, int nchunks) {
  tc::Header* hdr = reinterpret_cast<tc::Header*>(smem_raw);
  tc::Shade sh = tc::make_shade(smem_raw, tmem_base, nchunks, 1u);
  const tc::Layout lay = tc::make_layout((uint32_t)nchunks * tc::kNC, 1u);
  const uint32_t s_quads = tc::smem_addr(smem_raw + lay.quads), s_geom = tc::smem_addr(smem_raw + lay.geom);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long n_pixels = a.n_slots;
  const double inv_np = 1.0 / (double)n_pixels;
  const float inv_w = 1.0f / (float)a.width, inv_h = 1.0f / (float)a.height;
  unsigned n_rays = 0, n_paths = 0;
  TcProf prof{};
#ifdef UECRT_TC_PROFILE
  unsigned long long pc_rounds = 0;
  prof.last = clock64();
#endif
  WorkPool pool;
  pool_init(pool, a, lane);
  PathState P{}, Q{};                           // Q: in the pipeline; P: its next ray is ready
  P.dz = 1.f; P.dlen = 1.f; P.self_id = -1;
  Q = P;
  slot_refill_pooled(Q, pool, a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths);
  slot_refill_pooled(P, pool, a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths);
  auto vote = [&](uint32_t tile) {              // termination vote (tc_scan.cuh): stored before this warp's a_ready arrive of `tile`
    const bool warp_done = __all_sync(0xffffffffu, P.done && Q.done);
    if (lane == 0) hdr->done_flags[tile & 7u][threadIdx.x >> 5] = warp_done ? 1 : 0;
  };
  auto submit = [&](const PathState& s) {
    tc::shade_submit(sh, make_scan_ray(s.ox, s.oy, s.oz, s.dx, s.dy, s.dz), s.active);
  };
  vote(0u);
  submit(Q);                                    // tile 0
  for (uint32_t m = 0; tc::round_continues(hdr, m, -1); ++m) {
#ifdef UECRT_TC_PROFILE
    const long long pc_start = prof.last;
#endif
    tc::shade_wait(sh, m);                      // tile m (path Q) is scanned
    UECRT_TCP(prof, 3)
    if (tc::round_continues(hdr, m + 1u, -1)) { // the flags of tile m - 1 are final: every role decides alike
      vote(m + 1u);
      submit(P);                                // tile m + 1
    }
    UECRT_TCP(prof, 2)
    { const PathState t = P; P = Q; Q = t; }    // P: tile m's path, results pending; Q: in the pipeline
    const RayQuery q{P.ox, P.oy, P.oz, P.dx, P.dy, P.dz, 0.001f * P.dlen, P.self_id, P.active};
    Hit h;
    tc_collect_hit<EXACT>(sh, m & 1u, s_quads, s_geom, q, h);
    UECRT_TCP(prof, 4)
    if (P.active) {
      ++n_rays;
      slot_shade(P, h, a, n_paths);
    }
    UECRT_TCP(prof, 5)
    slot_refill_pooled(P, pool, a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths);   // a path that ended: the next one's camera ray
    UECRT_TCP(prof, 1)
#ifdef UECRT_TC_PROFILE
    prof.t[6] += prof.last - pc_start;
    ++pc_rounds;
#endif
  }
#ifdef UECRT_TC_PROFILE
  if (lane == 0) {
    atomicAdd(a.counters + 4, pc_rounds);
    for (int i = 0; i < 7; ++i) atomicAdd(a.counters + 5 + i, (unsigned long long)prof.t[i]);
    atomicAdd(a.counters + 16, sh.prof_trips);
  }
  atomicAdd(a.counters + 18, sh.prof_cands);
#endif
  if (a.count_stats) {
    unsigned r = n_rays, p = n_paths;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); p += __shfl_xor_sync(0xffffffffu, p, o); }
    if (lane == 0) {
      atomicAdd(a.counters + 1, (unsigned long long)r);
      atomicAdd(a.counters + 2, (unsigned long long)p);
    }
  }
}

template <bool EXACT, bool SWEEP>
__global__ void __launch_bounds__(tc::kThreads, 1) k_render_paths_tc(const RenderArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t npasses = SWEEP ? (uint32_t)a.sc.tc_npasses : 1u;
  const int pass_rows = a.sc.tc_nrows / (int)npasses;
  const int nchunks = pass_rows / tc::kNC;
  const uint32_t tmem_base = tc::cta_setup(smem_raw, a.sc.tc_rows, SWEEP ? 0u : (uint32_t)a.sc.tc_nrows * 128u, (uint32_t)nchunks);
  if (!SWEEP) {   // the shade side's tables: quad table and spheres (rows beyond n: never-hit quads, zero spheres)
    const tc::Layout lay = tc::make_layout((uint32_t)a.sc.tc_nrows, 1u);
    float4* sq = reinterpret_cast<float4*>(smem_raw + lay.quads);
    float4* sg = reinterpret_cast<float4*>(smem_raw + lay.geom);
    for (int i = threadIdx.x; i < a.sc.tc_nrows; i += tc::kThreads) {
      sq[i] = __ldg(&a.sc.tc_quads[i]);
      sg[i] = i < a.sc.n ? __ldg(&a.sc.geom[i]) : make_float4(0.f, 0.f, 0.f, 1.f);
    }
    __syncthreads();
  }
  const unsigned lane = threadIdx.x & 31u;
  if (tc::is_issuer_thread()) {
    tc::reg_dealloc<tc::kIssuerRegs>();
    const uint32_t iw = (uint32_t)__shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) - tc::kIssuerFirstThread / 32;   // warp-uniform
    if (iw == UECRT_TC_ISSUER_WARP) {
      if (UECRT_TC_ISSUER_LANES != 1 || tc::elect_one()) tc::issuer_role<SWEEP>(smem_raw, tmem_base, nchunks, npasses, a.sc.tc_rows, -1, a.counters);
    }
  } else if (tc::is_scan_thread()) {
    tc::reg_alloc<tc::kScanRegs>();
    tc::scanner_role<SWEEP>(smem_raw, tmem_base, nchunks, npasses, -1, a.counters);
  } else if (!SWEEP) {
    if (tc::kShadeRegs < tc::kLaunchRegs) tc::reg_dealloc<tc::kShadeRegs>();
    else if (tc::kShadeRegs > tc::kLaunchRegs) tc::reg_alloc<tc::kShadeRegs>();
    shade_role_resident<EXACT>(a, smem_raw, tmem_base, nchunks);
  } else {
    // swept operand: one path per thread; the thread serves every pass of its tile (tc_closest_hit)
    if (tc::kShadeRegs < tc::kLaunchRegs) tc::reg_dealloc<tc::kShadeRegs>();
    else if (tc::kShadeRegs > tc::kLaunchRegs) tc::reg_alloc<tc::kShadeRegs>();
    tc::Header* hdr = reinterpret_cast<tc::Header*>(smem_raw);
    tc::Shade sh = tc::make_shade(smem_raw, tmem_base, nchunks, SWEEP ? 2u : 1u);
#ifdef UECRT_TC_TRACE
    sh.trace = a.counters;
#endif
    uint32_t pass_m = 0;
    const unsigned long long n_pixels = a.n_slots;
    const double inv_np = 1.0 / (double)n_pixels;
    const float inv_w = 1.0f / (float)a.width, inv_h = 1.0f / (float)a.height;
    PathState st{};
    st.dz = 1.f; st.dlen = 1.f; st.self_id = -1;
    unsigned n_rays = 0, n_paths = 0;
    TcProf prof{};
#ifdef UECRT_TC_PROFILE
    unsigned long long pc_rounds = 0;
#endif
    WorkPool pool;
    pool_init(pool, a, lane);
    // The NEXT path of every lane is started ahead of time (work item claimed, Philox block, camera ray) in the
    // window where the warp waits for its tile's scan, so a lane whose path ends only copies it in: the
    // shade warps' own work per round is what starves the scan pipeline (a closed loop of 4 groups
    // around one scanner), and this moves a quarter of it off that loop.
    PathState nx{};
    nx.dz = 1.f; nx.dlen = 1.f; nx.self_id = -1;
    slot_refill_pooled(nx, pool, a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths);
    for (uint32_t round = 0; tc::round_continues(hdr, round, -1); ++round) {
#ifdef UECRT_TC_PROFILE
      prof.last = clock64();
      const long long pc_start = prof.last;
#endif
      if (!st.active && !st.done) {
        if (nx.active) { st = nx; nx.active = false; }
        else if (nx.done) st.done = true;
      }
      // termination vote (tc_scan.cuh): this warp's flag of this round, stored before its a_ready arrive
      const bool warp_done = __all_sync(0xffffffffu, st.done);
      if (lane == 0) hdr->done_flags[round & 7u][threadIdx.x >> 5] = warp_done ? 1 : 0;
      UECRT_TCP(prof, 0)
      const RayQuery q{st.ox, st.oy, st.oz, st.dx, st.dy, st.dz, 0.001f * st.dlen, st.self_id, st.active};
      Hit h;
      tc_closest_hit<EXACT, SWEEP>(sh, pass_m, npasses, pass_rows, a.sc, q, h, prof,
                                   [&]() { slot_refill_pooled(nx, pool, a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths); });
      UECRT_TCP(prof, 4)
      if (st.active) {
        ++n_rays;
        slot_shade(st, h, a, n_paths);
      }
#ifdef UECRT_TC_PROFILE
      UECRT_TCP(prof, 5)
      prof.t[6] += prof.last - pc_start;
      ++pc_rounds;
#endif
    }
#ifdef UECRT_TC_PROFILE
    if (lane == 0) {
      atomicAdd(a.counters + 4, pc_rounds);
      for (int i = 0; i < 7; ++i) atomicAdd(a.counters + 5 + i, (unsigned long long)prof.t[i]);
      atomicAdd(a.counters + 16, sh.prof_trips);
    }
#endif
#ifdef UECRT_TC_PROFILE
    atomicAdd(a.counters + 18, sh.prof_cands);
#endif
    if (a.count_stats) {
      unsigned r = n_rays, p = n_paths;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); p += __shfl_xor_sync(0xffffffffu, p, o); }
      if (lane == 0) {
        atomicAdd(a.counters + 1, (unsigned long long)r);
        atomicAdd(a.counters + 2, (unsigned long long)p);
      }
    }
  }
  tc::cta_teardown(tmem_base);
}

#ifndef UECRT_TS_PIN
#define UECRT_TS_PIN 1      // loop-invariant shared addresses pinned in registers (17.33 -> 17.03 ms at 100 spp)
#endif
#ifndef UECRT_TS_FUSED_SHADE
#define UECRT_TS_FUSED_SHADE 0  // 1: one Philox block per round for the scatter and the next path's camera ray (slot_shade_refill_fused); same paths, measured SLOWER (17.55 against 17.35 ms at 100 spp)
#endif
#ifndef UECRT_TS_PLANAR
#define UECRT_TS_PLANAR 0      // 1: planar copy of the quad table in shared memory (fewer bank conflicts, measured SLOWER: 17.05 against 16.65 ms at 100 spp)
#endif
// K1, tensor-core scan, SELF-SCAN pipeline (tc_self.cuh): resident operand, one path per shade thread, every shade warp
// reads the accumulators of its own rays.  Same path loop, same work queue, same tests as k_render_paths_tc.
template <bool EXACT>
__device__ __forceinline__ void shade_role_self(const RenderArgs& a, unsigned char* smem_raw, uint32_t tmem_base, int nchunks) {
  ts::Header* hdr = reinterpret_cast<ts::Header*>(smem_raw);
  ts::Shade sh = ts::make_shade(smem_raw, tmem_base, nchunks);
  const ts::Layout lay = ts::make_layout((uint32_t)nchunks * ts::kNC);
  uint32_t s_quads = tc::smem_addr(smem_raw + lay.quads), s_geom = tc::smem_addr(smem_raw + lay.geom);
#if UECRT_TS_PIN
  // keep the loop-invariant shared addresses in registers: ptxas otherwise re-derives them from S2UR SR_CgaCtaId inside the
  // resolve loop (3 % of the executed instructions, ncu line attribution of __cvta_generic_to_shared)
  tc::pin(s_quads); tc::pin(s_geom); tc::pin(sh.words); tc::pin(sh.full); tc::pin(sh.empty0); tc::pin(sh.lane_taddr);
  tc::pin(sh.a_ready0); tc::pin(sh.a_free0); tc::pin(sh.is_lane0);
#endif
  uint32_t flags0 = tc::smem_addr(&hdr->done_flags[0][0]), my_flag0 = flags0 + (threadIdx.x >> 5), my_lane = sh.g % (uint32_t)ts::kLanes;
  unsigned lane = threadIdx.x & 31u;
#if UECRT_TS_PIN
  tc::pin(flags0); tc::pin(my_flag0); tc::pin(my_lane); tc::pin(lane);
#endif
  const unsigned long long n_pixels = a.n_slots;
  const double inv_np = 1.0 / (double)n_pixels;
  const float inv_w = 1.0f / (float)a.width, inv_h = 1.0f / (float)a.height;
  unsigned n_rays = 0, n_paths = 0;
  TcProf prof{};
#ifdef UECRT_TC_PROFILE
  unsigned long long pc_rounds = 0;
  prof.last = clock64();
#endif
  WorkPool pool;
  pool_init(pool, a, lane);
  PathState st{};
  st.dz = 1.f; st.dlen = 1.f; st.self_id = -1;
  slot_refill_pooled(st, pool, a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths);
  uint32_t T = sh.g;                              // this group's tile of the round: T = round * kGroups + group
  for (uint32_t r = 0; ts::round_continues_s(flags0, my_lane, r); ++r, T += (uint32_t)ts::kGroups) {
#ifdef UECRT_TC_PROFILE
    const long long pc_start = prof.last;
#endif
    // termination vote (tc_self.cuh): this warp's flag of this round, stored before its a_ready arrive
    const bool warp_done = __all_sync(0xffffffffu, st.done);
    if (lane == 0) ts::vote_s(my_flag0, r, warp_done);
    ts::submit(sh, T, make_scan_ray(st.ox, st.oy, st.oz, st.dx, st.dy, st.dz), st.active);
    UECRT_TCP(prof, 2)
#ifdef UECRT_TC_PROFILE
    ts::scan_tile(sh, T, UECRT_TS_IDLE_SKIP && !__any_sync(0xffffffffu, st.active), &prof.t[0]);
#else
    ts::scan_tile(sh, T, UECRT_TS_IDLE_SKIP && !__any_sync(0xffffffffu, st.active));
#endif
    UECRT_TCP(prof, 3)
    const RayQuery q{st.ox, st.oy, st.oz, st.dx, st.dy, st.dz, 0.001f * st.dlen, st.self_id, st.active};
    Hit h;
    tc_collect_hit_with<EXACT>([&](auto&& f) { ts::for_each_flag(sh, q.live, f); }, s_quads, UECRT_TS_PLANAR ? (uint32_t)nchunks * (ts::kNC * 4u) : 0u, s_geom, q, h);
    UECRT_TCP(prof, 4)
#if UECRT_TS_FUSED_SHADE
    slot_shade_refill_fused(st, h, a, pool, lane, n_pixels, inv_np, inv_w, inv_h, n_rays, n_paths);   // shade, and the next path's camera ray, on one Philox block
    UECRT_TCP(prof, 5)
#else
    if (st.active) {
      ++n_rays;
      slot_shade(st, h, a, n_paths);
    }
    UECRT_TCP(prof, 5)
    slot_refill_pooled(st, pool, a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths);   // a path that ended: the next one's camera ray
    UECRT_TCP(prof, 1)
#endif
#ifdef UECRT_TC_PROFILE
    prof.t[6] += prof.last - pc_start;
    ++pc_rounds;
#endif
  }
#ifdef UECRT_TC_PROFILE
  if (lane == 0) {
    atomicAdd(a.counters + 4, pc_rounds);
    for (int i = 0; i < 7; ++i) atomicAdd(a.counters + 5 + i, (unsigned long long)prof.t[i]);
  }
#endif
  if (a.count_stats) {
    unsigned r = n_rays, p = n_paths;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); p += __shfl_xor_sync(0xffffffffu, p, o); }
    if (lane == 0) {
      atomicAdd(a.counters + 1, (unsigned long long)r);
      atomicAdd(a.counters + 2, (unsigned long long)p);
    }
  }
}

template <bool EXACT>
__global__ void __launch_bounds__(ts::kThreads, 1) k_render_paths_ts(const RenderArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nchunks = a.sc.tc_nrows / ts::kNC;
  const uint32_t tmem_base = ts::cta_setup(smem_raw, a.sc.tc_rows, (uint32_t)a.sc.tc_nrows * 128u);
  {   // the shade side's tables: quad table and spheres (rows beyond n: never-hit quads, zero spheres)
    const ts::Layout lay = ts::make_layout((uint32_t)a.sc.tc_nrows);
    float4* sq = reinterpret_cast<float4*>(smem_raw + lay.quads);
    float4* sg = reinterpret_cast<float4*>(smem_raw + lay.geom);
    const int nquads = a.sc.tc_nrows / 4;
    for (int i = threadIdx.x; i < a.sc.tc_nrows; i += ts::kThreads) {
      // global table: float4 4q + k = plane k of quad q; planar copy: plane k at k * nquads (tc_collect.cuh quad_candidates_sp)
      sq[UECRT_TS_PLANAR ? (i & 3) * nquads + (i >> 2) : i] = __ldg(&a.sc.tc_quads[i]);
      sg[i] = i < a.sc.n ? __ldg(&a.sc.geom[i]) : make_float4(0.f, 0.f, 0.f, 1.f);
    }
    __syncthreads();
  }
  if (ts::is_issuer_thread()) {
    tc::reg_dealloc<ts::kIssuerRegs>();
    const uint32_t iw = (uint32_t)__shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) - ts::kShadeThreads / 32;   // warp-uniform
    if (iw < (uint32_t)ts::kLanes) ts::issuer_role(smem_raw, tmem_base, nchunks, iw, a.counters);
  } else {
    if (ts::kShadeRegs < ts::kLaunchRegs) tc::reg_dealloc<ts::kShadeRegs>();
    else if (ts::kShadeRegs > ts::kLaunchRegs) tc::reg_alloc<ts::kShadeRegs>();
    shade_role_self<EXACT>(a, smem_raw, tmem_base, nchunks);
  }
  ts::cta_teardown(tmem_base);
}

// K2 on the tensor-core scan: every shade thread takes one caller ray per round; all CTAs run the
// same number of rounds (threads past the end of the ray array carry has_ray = false).
template <bool EXACT, bool SWEEP>
__global__ void __launch_bounds__(tc::kThreads, 1) k_primary_hits_tc(const DeviceScene sc, unsigned long long n_rays,
                                                                     const float* __restrict__ o, const float* __restrict__ d,
                                                                     double tmin, double tmax, double* __restrict__ t_out,
                                                                     double* __restrict__ n_out, int32_t* __restrict__ id_out,
                                                                     uint8_t* __restrict__ ff_out, uint32_t* __restrict__ dbg_words) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t npasses = SWEEP ? (uint32_t)sc.tc_npasses : 1u;
  const int pass_rows = sc.tc_nrows / (int)npasses;
  const int nchunks = pass_rows / tc::kNC;
  const uint32_t tmem_base = tc::cta_setup(smem_raw, sc.tc_rows, SWEEP ? 0u : (uint32_t)sc.tc_nrows * 128u, (uint32_t)nchunks);
  const unsigned long long stride = (unsigned long long)gridDim.x * tc::kShadeThreads;
  const int batches = (int)((n_rays + stride - 1) / stride);
  if (tc::is_issuer_thread()) {
    tc::reg_dealloc<tc::kIssuerRegs>();
    const uint32_t iw = (uint32_t)__shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) - tc::kIssuerFirstThread / 32;   // warp-uniform
    if (iw == UECRT_TC_ISSUER_WARP) {
      if (UECRT_TC_ISSUER_LANES != 1 || tc::elect_one()) tc::issuer_role<SWEEP>(smem_raw, tmem_base, nchunks, npasses, sc.tc_rows, batches);
    }
  } else if (tc::is_scan_thread()) {
    tc::reg_alloc<tc::kScanRegs>();
    tc::scanner_role<SWEEP>(smem_raw, tmem_base, nchunks, npasses, batches);
  } else {
    if (tc::kShadeRegs < tc::kLaunchRegs) tc::reg_dealloc<tc::kShadeRegs>();
    else if (tc::kShadeRegs > tc::kLaunchRegs) tc::reg_alloc<tc::kShadeRegs>();
    tc::Shade sh = tc::make_shade(smem_raw, tmem_base, nchunks, SWEEP ? 2u : 1u);
    uint32_t pass_m = 0;
    for (int b = 0; b < batches; ++b) {
      const unsigned long long i = (unsigned long long)b * stride + (unsigned long long)blockIdx.x * tc::kShadeThreads + threadIdx.x;
      const bool has_ray = i < n_rays;
      const unsigned long long j = has_ray ? i : 0;
      const float ox = o[3 * j], oy = o[3 * j + 1], oz = o[3 * j + 2];
      const float rx = d[3 * j], ry = d[3 * j + 1], rz = d[3 * j + 2];
      const float dlen = sqrtf(fmaf(rz, rz, fmaf(ry, ry, rx * rx)));
      const float inv = 1.0f / dlen;
      const float dx = rx * inv, dy = ry * inv, dz = rz * inv;
      const ScanRay sr = make_scan_ray(ox, oy, oz, dx, dy, dz);
      tc::shade_submit(sh, sr, has_ray);
      const tc::QuadRay qr = tc::make_quad_ray(sr);
      const int pass_quads = pass_rows / tc::kQuad;   // bitmap bits per pass
      double t = -1.0, nx = 0.0, ny = 0.0, nz = 0.0;
      int id = -1;
      bool front = false;
      if (!EXACT) {
        const float tmin_w = (float)tmin * dlen;
        FastBest best{(float)tmax * dlen, -1};
        for (uint32_t p = 0, base = 0;; base += (uint32_t)pass_quads) {
          tc::shade_wait(sh, pass_m++);
          if (dbg_words && has_ray) tc::dump_flags(sh, SWEEP ? 0u : ((pass_m - 1u) & 1u), dbg_words + (i * npasses + p) * (unsigned long long)sh.nwords);
          tc::for_each_flag(sh, SWEEP ? 0u : ((pass_m - 1u) & 1u), has_ray, [&](int k) {
            auto test = [&](int j) { test_fast(sc.geom, j, ox, oy, oz, dx, dy, dz, tmin_w, -1, best); };
            if (tc::kQuad == 4) tc_quad(sc, (int)base + k, qr, test); else test((int)base + k);
          });
          if (!SWEEP || ++p == npasses) break;
          tc::shade_next_pass(sh);
        }
        if (best.id >= 0) {
          // the winner's root and normal in double for the CALLER's direction (closest_hit.cuh refine_winner)
          const float4 gw = __ldg(&sc.geom[best.id]);
        const HitD w = refine_winner<false>(gw, ox, oy, oz, rx, ry, rz, !best.far_root);
          id = best.id; t = w.t; nx = w.nx; ny = w.ny; nz = w.nz; front = w.front;
        }
      } else {
        const double Ox = ox, Oy = oy, Oz = oz, Dx = rx, Dy = ry, Dz = rz;   // ORIGINAL direction: t == reference root
        const double a = __dadd_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(Dy, Dy)), __dmul_rn(Dz, Dz));  // sphere.hpp:28
        ExactBest best{tmax, -1};
        for (uint32_t p = 0, base = 0;; base += (uint32_t)pass_quads) {
          tc::shade_wait(sh, pass_m++);
          if (dbg_words && has_ray) tc::dump_flags(sh, SWEEP ? 0u : ((pass_m - 1u) & 1u), dbg_words + (i * npasses + p) * (unsigned long long)sh.nwords);
          tc::for_each_flag(sh, SWEEP ? 0u : ((pass_m - 1u) & 1u), has_ray, [&](int k) {
            auto test = [&](int j) { test_exact(sc.geom, j, Ox, Oy, Oz, Dx, Dy, Dz, a, tmin, best); };
            if (tc::kQuad == 4) tc_quad(sc, (int)base + k, qr, test); else test((int)base + k);
          });
          if (!SWEEP || ++p == npasses) break;
          tc::shade_next_pass(sh);
        }
        if (best.id >= 0) {
          const float4 g = __ldg(&sc.geom[best.id]);
          const double px = __dadd_rn(Ox, __dmul_rn(Dx, best.t)), py = __dadd_rn(Oy, __dmul_rn(Dy, best.t)),
                       pz = __dadd_rn(Oz, __dmul_rn(Dz, best.t));                                   // ray.hpp:16-19
          const double r = g.w;
          nx = __ddiv_rn(__dsub_rn(px, (double)g.x), r);                                             // sphere.hpp:44
          ny = __ddiv_rn(__dsub_rn(py, (double)g.y), r);
          nz = __ddiv_rn(__dsub_rn(pz, (double)g.z), r);
          front = __dadd_rn(__dadd_rn(__dmul_rn(Dx, nx), __dmul_rn(Dy, ny)), __dmul_rn(Dz, nz)) < 0;  // hittable.hpp:25
          if (!front) { nx = -nx; ny = -ny; nz = -nz; }
          id = best.id; t = best.t;
        }
      }
      if (has_ray) {
        t_out[i] = t; n_out[3 * i] = nx; n_out[3 * i + 1] = ny; n_out[3 * i + 2] = nz;
        id_out[i] = id; ff_out[i] = front ? 1 : 0;
      }
    }
  }
  tc::cta_teardown(tmem_base);
}

// ---------------------------------------------------------------------------------------
// K1 / K2 behind the uniform grid (grid_build.hpp, grid_walk.cuh).  Same path loop,
// same per-sphere test, same shading; a ray tests only the spheres listed in the cells it passes (plus the
// always-list).  Opt-in: the algorithmic work is no longer 18 FLOP x N x rays, so it is reported separately.
template <bool EXACT>
__device__ __forceinline__ void grid_closest_hit(const DeviceScene& sc, const RayQuery& q, Hit& h, unsigned& n_tests) {
  h.id = -1;
  if (!q.live) return;
  if (!EXACT) {
    FastBest best{CUDART_INF_F, -1};
    grid_walk(sc.grid, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz,
              [&](int j) { test_fast<false>(sc.geom, j, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz, q.tmin, q.self_id, best); },
              [&]() { return best.t; }, n_tests);
    if (best.id >= 0) {
      h.id = best.id;
      h.t = best.t;
      finish_hit_fast(sc.geom, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz, best.far_root, h);
    }
  } else {
    ExactBest best{(double)CUDART_INF, -1};
    const double Ox = q.ox, Oy = q.oy, Oz = q.oz, Dx = q.dx, Dy = q.dy, Dz = q.dz;
    const double a = __dadd_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(Dy, Dy)), __dmul_rn(Dz, Dz));
    grid_walk(sc.grid, q.ox, q.oy, q.oz, q.dx, q.dy, q.dz,
              [&](int j) { test_exact<false>(sc.geom, j, Ox, Oy, Oz, Dx, Dy, Dz, a, (double)q.tmin, best, q.self_id); },
              [&]() { return (float)best.t; }, n_tests);
    if (best.id >= 0) {
      const float4 g = __ldg(&sc.geom[best.id]);
      const double px = Ox + Dx * best.t, py = Oy + Dy * best.t, pz = Oz + Dz * best.t;   // ray.hpp:16-19
      const double rad = g.w;
      double nx = (px - (double)g.x) / rad, ny = (py - (double)g.y) / rad, nz = (pz - (double)g.z) / rad;   // sphere.hpp:44
      const bool front = (Dx * nx + Dy * ny + Dz * nz) < 0;                                     // hittable.hpp:25
      if (!front) { nx = -nx; ny = -ny; nz = -nz; }
      h.id = best.id; h.t = (float)best.t; h.front = front;
      h.px = (float)px; h.py = (float)py; h.pz = (float)pz;
      h.nx = (float)nx; h.ny = (float)ny; h.nz = (float)nz;
    }
  }
}

template <bool EXACT>
__global__ void __launch_bounds__(kRenderThreads, 2) k_render_paths_grid(const RenderArgs a) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long n_pixels = a.n_slots;
  const double inv_np = 1.0 / (double)n_pixels;
  const float inv_w = 1.0f / (float)a.width, inv_h = 1.0f / (float)a.height;
  PathState st{};
  st.dz = 1.f; st.dlen = 1.f; st.self_id = -1;
  unsigned n_rays = 0, n_paths = 0, n_tests = 0;
  WorkPool pool;
  pool_init(pool, a, lane);
  for (;;) {
    slot_refill_pooled(st, pool, a, lane, n_pixels, inv_np, inv_w, inv_h, n_paths);
    if (__all_sync(0xffffffffu, st.done)) break;
    const RayQuery q{st.ox, st.oy, st.oz, st.dx, st.dy, st.dz, 0.001f * st.dlen, st.self_id, st.active};
    Hit h;
    grid_closest_hit<EXACT>(a.sc, q, h, n_tests);
    if (st.active) {
      ++n_rays;
      slot_shade(st, h, a, n_paths);
    }
  }
  if (a.count_stats) {
    unsigned r = n_rays, p = n_paths;
    unsigned long long t = n_tests;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r += __shfl_xor_sync(0xffffffffu, r, o); p += __shfl_xor_sync(0xffffffffu, p, o); t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    if (lane == 0) {
      atomicAdd(a.counters + 1, (unsigned long long)r);
      atomicAdd(a.counters + 2, (unsigned long long)p);
      atomicAdd(a.counters + 3, t);
    }
  }
}

template <bool EXACT>
__global__ void __launch_bounds__(kHitsThreads) k_primary_hits_grid(const DeviceScene sc, unsigned long long n_rays,
                                                                    const float* __restrict__ o, const float* __restrict__ d,
                                                                    double tmin, double tmax, double* __restrict__ t_out,
                                                                    double* __restrict__ n_out, int32_t* __restrict__ id_out,
                                                                    uint8_t* __restrict__ ff_out) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += stride) {
    const float ox = o[3 * i], oy = o[3 * i + 1], oz = o[3 * i + 2];
    const float rx = d[3 * i], ry = d[3 * i + 1], rz = d[3 * i + 2];
    const float dlen = sqrtf(fmaf(rz, rz, fmaf(ry, ry, rx * rx)));
    const float inv = 1.0f / dlen;
    const float dx = rx * inv, dy = ry * inv, dz = rz * inv;
    double t = -1.0, nx = 0.0, ny = 0.0, nz = 0.0;
    int id = -1;
    bool front = false;
    unsigned n_tests = 0;
    if (!EXACT) {
      const float tmin_w = (float)tmin * dlen;
      FastBest best{(float)tmax * dlen, -1};
      grid_walk(sc.grid, ox, oy, oz, dx, dy, dz, [&](int j) { test_fast<false>(sc.geom, j, ox, oy, oz, dx, dy, dz, tmin_w, -1, best); },
                [&]() { return best.t; }, n_tests);
      if (best.id >= 0) {
        // the winner's root and normal in double for the CALLER's direction (closest_hit.cuh refine_winner)
        const float4 gw = __ldg(&sc.geom[best.id]);
        const HitD w = refine_winner<false>(gw, ox, oy, oz, rx, ry, rz, !best.far_root);
        id = best.id; t = w.t; nx = w.nx; ny = w.ny; nz = w.nz; front = w.front;
      }
    } else {
      const double Ox = ox, Oy = oy, Oz = oz, Dx = rx, Dy = ry, Dz = rz;   // ORIGINAL direction: t == reference root
      const double a = __dadd_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(Dy, Dy)), __dmul_rn(Dz, Dz));  // sphere.hpp:28
      ExactBest best{tmax, -1};
      // the walk measures distance along the unit direction: root * |D|
      grid_walk(sc.grid, ox, oy, oz, dx, dy, dz, [&](int j) { test_exact<false>(sc.geom, j, Ox, Oy, Oz, Dx, Dy, Dz, a, tmin, best); },
                [&]() { return (float)(best.t * (double)dlen); }, n_tests);
      if (best.id >= 0) {
        const float4 g = __ldg(&sc.geom[best.id]);
        const double px = __dadd_rn(Ox, __dmul_rn(Dx, best.t)), py = __dadd_rn(Oy, __dmul_rn(Dy, best.t)),
                     pz = __dadd_rn(Oz, __dmul_rn(Dz, best.t));                                   // ray.hpp:16-19
        const double r = g.w;
        nx = __ddiv_rn(__dsub_rn(px, (double)g.x), r);                                             // sphere.hpp:44
        ny = __ddiv_rn(__dsub_rn(py, (double)g.y), r);
        nz = __ddiv_rn(__dsub_rn(pz, (double)g.z), r);
        front = __dadd_rn(__dadd_rn(__dmul_rn(Dx, nx), __dmul_rn(Dy, ny)), __dmul_rn(Dz, nz)) < 0;  // hittable.hpp:25
        if (!front) { nx = -nx; ny = -ny; nz = -nz; }
        id = best.id; t = best.t;
      }
    }
    t_out[i] = t; n_out[3 * i] = nx; n_out[3 * i + 1] = ny; n_out[3 * i + 2] = nz;
    id_out[i] = id; ff_out[i] = front ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------
// K3: to_color3b (source.cpp:73-83): sum / spp -> sqrt -> clamp [0, 0.999] -> * 256 -> truncate.
// Done in double like the reference (the input is an FP32 sum); NaN / negative -> 0.
__global__ void k_resolve(const float4* __restrict__ accum, uint32_t n_pixels, uint32_t spp, uint8_t* __restrict__ rgb) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pixels) return;
  const float4 a = accum[i];
  const double inv = 1.0 / (double)spp;
  auto q = [&](float c) -> uint8_t {
    double v = (double)c * inv;
    v = (v == v && v > 0.0) ? sqrt(v) : 0.0;
    v = v > 0.999 ? 0.999 : v;
    return (uint8_t)(v * 256.0);
  };
  rgb[3 * i] = q(a.x); rgb[3 * i + 1] = q(a.y); rgb[3 * i + 2] = q(a.z);
}

// ---------------------------------------------------------------------------------------
// probes
__global__ void k_scatter(unsigned long long n, const uint32_t* __restrict__ mat_type, const float4* __restrict__ mat_param,
                          const float* __restrict__ d_in, const float* __restrict__ p, const float* __restrict__ nrm,
                          const uint8_t* __restrict__ ff, const float* __restrict__ un, uint8_t* __restrict__ scattered,
                          float* __restrict__ att, float* __restrict__ o_out, float* __restrict__ d_out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float dx = d_in[3 * i], dy = d_in[3 * i + 1], dz = d_in[3 * i + 2];
  const float inv = 1.0f / sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));   // material.hpp:70 normalises d_in
  dx *= inv; dy *= inv; dz *= inv;
  const ScatterOut s = scatter_material(mat_type[i], mat_param[i], dx, dy, dz, nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2],
                                        ff[i] != 0, un[4 * i], un[4 * i + 1], un[4 * i + 2], un[4 * i + 3]);
  scattered[i] = s.ok ? 1 : 0;
  att[3 * i] = s.ar; att[3 * i + 1] = s.ag; att[3 * i + 2] = s.ab;
  o_out[3 * i] = p[3 * i]; o_out[3 * i + 1] = p[3 * i + 1]; o_out[3 * i + 2] = p[3 * i + 2];
  d_out[3 * i] = s.dx; d_out[3 * i + 1] = s.dy; d_out[3 * i + 2] = s.dz;
}

__global__ void k_philox(unsigned long long n, uint32_t k0, uint32_t k1, const uint32_t* __restrict__ c, uint32_t* __restrict__ out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const philox_out r = philox4x32_10(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3], k0, k1);
  out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

// ---------------------------------------------------------------------------------------
// host launchers
size_t render_smem_bytes(int npairs, int ntiles, int threads, int rays) {
  const size_t tile = (size_t)(npairs + 2) * 32, lists = (size_t)rays * kListCap * threads * sizeof(uint16_t);
  return ntiles > 1 ? 32 + 2 * tile + lists : 16 + tile + lists;
}

// Kernel table: [exact][tiled][rays-1].  Two paths per thread (one shared scene sweep for both) is
// correct and tested but MEASURED SLOWER on B200 (54% vs 65% of FP32 peak on the final scene: the
// duplicated per-slot code grows the kernel to ~60 KB of SASS, past the instruction cache, and the
// 128 registers leave 2 CTAs/SM).  It is compiled only with -DUECRT_ENABLE_RAYS2.
typedef void (*RenderKernel)(const RenderArgs);
#ifdef UECRT_ENABLE_RAYS2
#define UECRT_R2(e, t) k_render_paths<e, t, 2>
#else
#define UECRT_R2(e, t) k_render_paths<e, t, 1>
#endif
static RenderKernel render_kernel(bool exact, bool tiled, int rays) {
  static const RenderKernel table[2][2][2] = {
      {{k_render_paths<false, false, 1>, UECRT_R2(false, false)}, {k_render_paths<false, true, 1>, UECRT_R2(false, true)}},
      {{k_render_paths<true, false, 1>, UECRT_R2(true, false)}, {k_render_paths<true, true, 1>, UECRT_R2(true, true)}}};
  return table[exact ? 1 : 0][tiled ? 1 : 0][rays == 2 ? 1 : 0];
}
int max_rays_per_thread() {
#ifdef UECRT_ENABLE_RAYS2
  return 2;
#else
  return 1;
#endif
}

// max_blocks_per_sm[0] = fast variant, [1] = exact variant (for `rays` rays per thread)
cudaError_t configure_kernels(int* regs_render, int* max_blocks_per_sm, size_t smem_render, bool tiled, int rays) {
  cudaError_t e;
  const int optin = 232448;
  static bool attrs_set[16] = {};
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if (dev < 16 && !attrs_set[dev]) {
    for (int ex = 0; ex < 2; ++ex)
      for (int ti = 0; ti < 2; ++ti)
        for (int ry = 1; ry <= 2; ++ry)
          if ((e = cudaFuncSetAttribute(render_kernel(ex, ti, ry), cudaFuncAttributeMaxDynamicSharedMemorySize, optin)) != cudaSuccess) return e;
#define UECRT_SET(k) if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin)) != cudaSuccess) return e;
    UECRT_SET((k_primary_hits<false, false>)) UECRT_SET((k_primary_hits<true, false>))
    UECRT_SET((k_primary_hits<false, true>)) UECRT_SET((k_primary_hits<true, true>))
#undef UECRT_SET
    attrs_set[dev] = true;
  }
  cudaFuncAttributes fa;
  if ((e = cudaFuncGetAttributes(&fa, render_kernel(false, tiled, rays))) != cudaSuccess) return e;
  if (regs_render) *regs_render = fa.numRegs;
  for (int exact = 0; exact < 2; ++exact) {
    int n1 = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, render_kernel(exact, tiled, rays), kRenderThreads, smem_render)) != cudaSuccess) return e;
    if (max_blocks_per_sm) max_blocks_per_sm[exact] = n1;
  }
  return cudaSuccess;
}

void launch_render_paths(const RenderArgs& a, bool exact, int rays, int blocks, int threads, size_t smem, cudaStream_t st) {
  render_kernel(exact, a.sc.ntiles > 1, rays)<<<blocks, threads, smem, st>>>(a);
}

void launch_primary_hits(const DeviceScene& sc, unsigned long long n_rays, const float* o, const float* d, double tmin,
                         double tmax, bool exact, double* t, double* nrm, int32_t* id, uint8_t* ff, cudaStream_t st) {
  const size_t smem = render_smem_bytes(sc.npairs, sc.ntiles, kHitsThreads, 1);
  unsigned long long want = (n_rays + kHitsThreads - 1) / kHitsThreads;
  const int blocks = (int)(want < 1 ? 1 : (want > 148ull * 16 ? 148ull * 16 : want));
  const bool tiled = sc.ntiles > 1;
  if (exact && tiled) k_primary_hits<true, true><<<blocks, kHitsThreads, smem, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff);
  else if (exact) k_primary_hits<true, false><<<blocks, kHitsThreads, smem, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff);
  else if (tiled) k_primary_hits<false, true><<<blocks, kHitsThreads, smem, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff);
  else k_primary_hits<false, false><<<blocks, kHitsThreads, smem, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff);
}

// ---- uniform-grid variants
cudaError_t configure_kernels_grid(int* regs_render, int* max_blocks_per_sm) {
  cudaError_t e;
  cudaFuncAttributes fa;
  if ((e = cudaFuncGetAttributes(&fa, k_render_paths_grid<false>)) != cudaSuccess) return e;
  if (regs_render) *regs_render = fa.numRegs;
  for (int exact = 0; exact < 2; ++exact) {
    int n1 = 0;
    e = exact ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k_render_paths_grid<true>, kRenderThreads, 0)
              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k_render_paths_grid<false>, kRenderThreads, 0);
    if (e != cudaSuccess) return e;
    if (max_blocks_per_sm) max_blocks_per_sm[exact] = n1;
  }
  return cudaSuccess;
}
void launch_render_paths_grid(const RenderArgs& a, bool exact, int blocks, cudaStream_t st) {
  if (exact) k_render_paths_grid<true><<<blocks, kRenderThreads, 0, st>>>(a);
  else k_render_paths_grid<false><<<blocks, kRenderThreads, 0, st>>>(a);
}
void launch_primary_hits_grid(const DeviceScene& sc, unsigned long long n_rays, const float* o, const float* d, double tmin,
                              double tmax, bool exact, double* t, double* nrm, int32_t* id, uint8_t* ff, cudaStream_t st) {
  unsigned long long want = (n_rays + kHitsThreads - 1) / kHitsThreads;
  const int blocks = (int)(want < 1 ? 1 : (want > 148ull * 16 ? 148ull * 16 : want));
  if (exact) k_primary_hits_grid<true><<<blocks, kHitsThreads, 0, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff);
  else k_primary_hits_grid<false><<<blocks, kHitsThreads, 0, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff);
}

// ---- tensor-core scan variants
int tc_resident_max_spheres() {
  const size_t fixed = tc::make_layout(0).total;
  // per sphere row: 128 B operand, two bitmap buffers of 512 rays x 1 bit per kQuad spheres, 16 B quad table, 16 B sphere
  const int fit = (int)((232448 - fixed) / (128 + 2 * 64 / tc::kQuad + 32) / tc::kNC * tc::kNC);
  return fit < tc::kMaxRows ? fit : tc::kMaxRows;
}
#ifndef UECRT_TC_SWEEP_ROWS
#define UECRT_TC_SWEEP_ROWS 512      // two operand buffers of 64 KB + 32 KB of bitmaps
#endif
// slice size for a swept scene of n spheres: at most UECRT_TC_SWEEP_ROWS rows, and one 128-row chunk less per
// slice when that leaves fewer never-hit pad rows in total (1025 spheres: 3 slices of 384 rows, not 3 of 512)
int tc_sweep_rows(int n) {
  const int chunks = (n + tc::kNC - 1) / tc::kNC, max_per = UECRT_TC_SWEEP_ROWS / tc::kNC;
  int best_per = max_per, best_total = (chunks + max_per - 1) / max_per * max_per;
  if (max_per > 1) {
    const int per = max_per - 1, total = (chunks + per - 1) / per * per;
    if (total < best_total) { best_per = per; best_total = total; }
  }
  return best_per * tc::kNC;
}
int tc_threads() { return tc::kThreads; }
int tc_quad() { return tc::kQuad; }
int tc_chunk() { return tc::kNC; }
size_t tc_smem_bytes(int tc_nrows, int tc_npasses) {
  return tc_npasses > 1 ? tc::make_layout((uint32_t)(tc_nrows / tc_npasses), 2u).total : tc::make_layout((uint32_t)tc_nrows).total;
}

// Which pipeline K1 runs a tensor scan with: 1 = self-scan (tc_self.cuh: resident operand up to ts::kMaxRows rows),
// 0 = scan warps (tc_scan.cuh / tc_collect.cuh: every scene; always for swept operands and for K2).
// UECRT_TC_PIPELINE=scanwarps|self in the environment overrides the built-in choice (tuning and tests).
#ifndef UECRT_TC_SELF_DEFAULT
#define UECRT_TC_SELF_DEFAULT 1
#endif
static std::atomic<int> g_tc_pipeline{-1};      // -1: not set (environment, else the built-in default)
void tc_set_pipeline(int p) { g_tc_pipeline.store(p == 0 ? 0 : 1); }
static bool tc_self_enabled() {
  static const int env = [] {
    const char* e = getenv("UECRT_TC_PIPELINE");
    if (e && !strcmp(e, "scanwarps")) return 0;
    if (e && !strcmp(e, "self")) return 1;
    return UECRT_TC_SELF_DEFAULT;
  }();
  const int v = g_tc_pipeline.load();
  return (v < 0 ? env : v) != 0;
}
static bool tc_use_self(const DeviceScene& sc) { return tc_self_enabled() && sc.tc_npasses == 1 && sc.tc_nrows <= ts::kMaxRows; }
void tc_render_config(const DeviceScene& sc, int* threads, size_t* smem, int* pipeline) {
  const bool self = tc_use_self(sc);
  if (threads) *threads = self ? ts::kThreads : tc::kThreads;
  if (smem) *smem = self ? ts::make_layout((uint32_t)sc.tc_nrows).total : tc_smem_bytes(sc.tc_nrows, sc.tc_npasses);
  if (pipeline) *pipeline = self ? 1 : 0;
}

cudaError_t configure_kernels_tc(int* regs_render, size_t smem) {
  cudaError_t e;
  static bool attrs_set[16] = {};
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if (dev < 16 && !attrs_set[dev]) {
    const int optin = 232448;
#define UECRT_SET(k) if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin)) != cudaSuccess) return e;
    UECRT_SET((k_render_paths_tc<false, false>)) UECRT_SET((k_render_paths_tc<true, false>))
    UECRT_SET((k_primary_hits_tc<false, false>)) UECRT_SET((k_primary_hits_tc<true, false>))
    UECRT_SET((k_render_paths_tc<false, true>)) UECRT_SET((k_render_paths_tc<true, true>))
    UECRT_SET((k_primary_hits_tc<false, true>)) UECRT_SET((k_primary_hits_tc<true, true>))
    UECRT_SET((k_render_paths_ts<false>)) UECRT_SET((k_render_paths_ts<true>))
#undef UECRT_SET
    attrs_set[dev] = true;
  }
  cudaFuncAttributes fa;
  if ((e = cudaFuncGetAttributes(&fa, k_render_paths_tc<false, false>)) != cudaSuccess) return e;
  if (regs_render) *regs_render = fa.numRegs;
  int n1 = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k_render_paths_tc<false, false>, tc::kThreads, smem)) != cudaSuccess) return e;
  return n1 >= 1 ? cudaSuccess : cudaErrorInvalidConfiguration;
}

void launch_render_paths_tc(const RenderArgs& a, bool exact, int blocks, size_t smem, cudaStream_t st) {
  const bool sweep = a.sc.tc_npasses > 1;
  if (tc_use_self(a.sc)) {
    const size_t smem_self = ts::make_layout((uint32_t)a.sc.tc_nrows).total;
    if (exact) k_render_paths_ts<true><<<blocks, ts::kThreads, smem_self, st>>>(a);
    else k_render_paths_ts<false><<<blocks, ts::kThreads, smem_self, st>>>(a);
  } else if (sweep) {
    if (exact) k_render_paths_tc<true, true><<<blocks, tc::kThreads, smem, st>>>(a);
    else k_render_paths_tc<false, true><<<blocks, tc::kThreads, smem, st>>>(a);
  } else {
    if (exact) k_render_paths_tc<true, false><<<blocks, tc::kThreads, smem, st>>>(a);
    else k_render_paths_tc<false, false><<<blocks, tc::kThreads, smem, st>>>(a);
  }
}

void launch_primary_hits_tc(const DeviceScene& sc, unsigned long long n_rays, const float* o, const float* d, double tmin,
                            double tmax, bool exact, double* t, double* nrm, int32_t* id, uint8_t* ff, int sm_count,
                            cudaStream_t st, uint32_t* dbg_words) {
  const size_t smem = tc_smem_bytes(sc.tc_nrows, sc.tc_npasses);
  unsigned long long want = (n_rays + tc::kShadeThreads - 1) / tc::kShadeThreads;
  const int blocks = (int)(want < 1 ? 1 : (want > (unsigned long long)sm_count ? (unsigned long long)sm_count : want));
  if (sc.tc_npasses > 1) {
    if (exact) k_primary_hits_tc<true, true><<<blocks, tc::kThreads, smem, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff, dbg_words);
    else k_primary_hits_tc<false, true><<<blocks, tc::kThreads, smem, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff, dbg_words);
  } else {
    if (exact) k_primary_hits_tc<true, false><<<blocks, tc::kThreads, smem, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff, dbg_words);
    else k_primary_hits_tc<false, false><<<blocks, tc::kThreads, smem, st>>>(sc, n_rays, o, d, tmin, tmax, t, nrm, id, ff, dbg_words);
  }
}

void launch_resolve(const float4* accum, uint32_t n_pixels, uint32_t spp, uint8_t* rgb, cudaStream_t st) {
  k_resolve<<<(n_pixels + 255) / 256, 256, 0, st>>>(accum, n_pixels, spp, rgb);
}

void launch_scatter(unsigned long long n, const uint32_t* mat_type, const float4* mat_param, const float* d_in,
                    const float* p, const float* nrm, const uint8_t* ff, const float* uniforms4, uint8_t* scattered,
                    float* att, float* o_out, float* d_out, cudaStream_t st) {
  k_scatter<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, mat_type, mat_param, d_in, p, nrm, ff, uniforms4, scattered,
                                                        att, o_out, d_out);
}

// Probe: the camera ray K1 starts path (pixel, sample) with -- slot_start_path itself on caller (pixel, sample) pairs
__global__ void k_camera_rays(const RenderArgs a, unsigned long long n, const uint32_t* __restrict__ pixel, const uint32_t* __restrict__ sample,
                              float* __restrict__ o_out, float* __restrict__ d_out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PathState p{};
  p.pixel = pixel[i]; p.sample = sample[i]; p.valid = true;
  unsigned n_paths = 0;
  slot_start_path(p, a, true, 1.0f / (float)a.width, 1.0f / (float)a.height, n_paths);
  o_out[3 * i] = p.ox; o_out[3 * i + 1] = p.oy; o_out[3 * i + 2] = p.oz;
  d_out[3 * i] = p.dx * p.dlen; d_out[3 * i + 1] = p.dy * p.dlen; d_out[3 * i + 2] = p.dz * p.dlen;   // the unnormalised direction
}
void launch_camera_rays(const RenderArgs& a, unsigned long long n, const uint32_t* pixel, const uint32_t* sample, float* o, float* d, cudaStream_t st) {
  k_camera_rays<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, n, pixel, sample, o, d);
}

void launch_philox(unsigned long long n, uint32_t k0, uint32_t k1, const uint32_t* ctr4, uint32_t* out4, cudaStream_t st) {
  k_philox<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, k0, k1, ctr4, out4);
}

}  // namespace uecrt
