// tor_kernels.cuh — the sm_100a render kernel (exact mode).
//
// One persistent thread per pixel *stream*: a lane owns one pixel at a time, runs all of its
// samples sequentially (the reference shares one RNG stream across a pixel's samples,
// render.nim:59-67, so samples of a pixel cannot be split), and pulls the next pixel from a global
// atomic queue when it is done.  The kernel body is a state machine over bounce *segments* rather
// than nested sample / depth loops, so every lane of a warp always has a live ray when the warp
// enters the sphere loop — the only part that matters for throughput (>= 95 % of the work).
//
// Per segment:
//   A. conservative FMA filter over every sphere (filter records in shared memory, staged once per
//      CTA with a TMA bulk copy; all lanes read the same record -> shared-memory broadcast);
//      survivors are appended to a small per-lane candidate list in shared memory;
//   B. the reference's exact arithmetic (spheres.nim:28-49, moving_spheres.nim:39-67; no FMA
//      contraction, IEEE div/sqrt) on the candidates -> closest hit, ties to the lowest index like
//      the in-order scan of hittables_lists.nim:48-55;
//   C. scatter (materials.nim) or sky (render.nim:40-45), then next segment / sample / pixel.
//
// Compiled with -fmad=false: every fused multiply-add in this file is an explicit fma().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tor_b200.h"
#include "tor_detmath.h"
#include "tor_scene_pack.hpp"

namespace tor {

struct RenderParams {
  SceneView sv;
  const uint8_t* blob;  // device copy of PackedScene::blob
  tor_camera cam;
  double* pixels;  // compact selected rows: row i of the selection at pixels + i*ncols*3
  int32_t nrows, ncols, spp;
  int32_t max_depth;
  double inv_spp;    // 1.0 / float64(spp)            canvas.nim:49
  double inv_gamma;  // 1.0 / float64(float32 gamma)  canvas.nim:50
  int32_t row_begin, row_step, nsel_rows;
  uint32_t count_segments;
  unsigned long long* work_counter;  // pixel queue head
  unsigned long long* counters;      // [0] primary rays, [1] segments
};

static constexpr int kMaxCand = 24;  // per-lane candidate slots (u16) before an early resolve

// ------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------- RNG
// support/rng.nim:18-74 — xoshiro256+ seeded by the reference's (sic) SplitMix64 variant.
struct Rng {
  uint64_t s0, s1, s2, s3;
};
__device__ __forceinline__ uint64_t splitmix64(uint64_t& state) {  // rng.nim:31-36
  state += 0x9e3779b97f4a7c15ull;
  uint64_t r = state;
  r = (r ^ (r >> 30)) * 0xbf58476d1ce4e5b9ull;
  r = (r ^ (r >> 27)) * 0xbf58476d1ce4e5b9ull;  // same multiplier twice, as in the reference
  return r ^ (r >> 31);
}
// sub = 0 is the reference's seed(row, col).  Split-stream mode (TOR_MODE_FAST) seeds sample range `sub` of the pixel
// with outputs 4*sub .. 4*sub+3 of the same SplitMix64 sequence (its state advances by a constant per output).
__device__ __forceinline__ void rng_seed_pixel(Rng& g, int32_t row, int32_t col, uint32_t sub = 0) {  // rng.nim:21-29,46-53
  uint64_t sm = (((uint64_t)(int64_t)row << 32) ^ (uint64_t)(int64_t)col) + (uint64_t)(4u * sub) * 0x9e3779b97f4a7c15ull;
  g.s0 = splitmix64(sm);
  g.s1 = splitmix64(sm);
  g.s2 = splitmix64(sm);
  g.s3 = splitmix64(sm);
}
__device__ __forceinline__ uint64_t rng_next(Rng& g) {  // rng.nim:58-74
  uint64_t result = g.s0 + g.s3;
  uint64_t t = g.s1 << 17;
  g.s2 ^= g.s0;
  g.s3 ^= g.s1;
  g.s1 ^= g.s2;
  g.s0 ^= g.s3;
  g.s2 ^= t;
  g.s3 = (g.s3 << 45) | (g.s3 >> 19);
  return result;
}
__device__ __forceinline__ double rng_u01(Rng& g) {  // rng.nim:129-133
  return __longlong_as_double((long long)((rng_next(g) >> 12) | 0x3ff0000000000000ull)) - 1.0;
}
__device__ __forceinline__ double rng_umax(Rng& g, double mx) { return rng_u01(g) * mx; }  // rng.nim:135-143
__device__ __forceinline__ double rng_urange(Rng& g, double lo, double hi) {               // rng.nim:116-127
  double v = rng_u01(g) * (hi - lo) + lo;
  return (v <= lo) ? lo : v;
}

// ---------------------------------------------------------------------------------- vectors
struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 v3(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }  // vec3s.nim:86-91
__device__ __forceinline__ V3 operator*(double s, V3 a) { return a * s; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }     // :96-98
__device__ __forceinline__ double len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }          // :23-24
__device__ __forceinline__ V3 unit_vector(V3 a) { return a * (1.0 / sqrt(len2(a))); }               // :93-94,106-107
__device__ __forceinline__ V3 reflect(V3 u, V3 n) { return u - (2 * dot(u, n)) * n; }               // rays.nim:27-28

// ------------------------------------------------------------------------ exact sphere test
// The reference's arithmetic on one object (spheres.nim:28-49 / moving_spheres.nim:39-67) up to the
// accepted root; returns +inf when no root lies in (t_min, +inf).  The caller takes the
// lexicographic minimum of (t, index), which equals the in-order scan with a shrinking t_max
// (hittables_lists.nim:48-55): a sphere's first root above t_min is the only one that can win.
__device__ __forceinline__ V3 obj_center(const tor_hittable* __restrict__ h, double time) {
  V3 c0 = v3(h->center0[0], h->center0[1], h->center0[2]);
  if (h->kind == TOR_MOVING_SPHERE) {  // moving_spheres.nim:39-44
    V3 c1 = v3(h->center1[0], h->center1[1], h->center1[2]);
    double q = (time - h->time0) / (h->time1 - h->time0);
    return c0 + (q * (c1 - c0));
  }
  return c0;
}

__device__ __forceinline__ double exact_first_root(const tor_hittable* __restrict__ h, V3 o, V3 d, double time,
                                                   double a, double t_min) {
  V3 oc = o - obj_center(h, time);
  double half_b = dot(oc, d);
  double c = len2(oc) - h->radius * h->radius;
  double disc = half_b * half_b - a * c;
  double t = __longlong_as_double(0x7ff0000000000000ll);
  if (disc > 0) {
    double root = sqrt(disc);
    double sol = (-half_b - root) / a;
    if (t_min < sol) {
      t = sol;
    } else {
      sol = (-half_b + root) / a;
      if (t_min < sol) t = sol;
    }
  }
  return t;
}


// ------------------------------------------------------------------ per-lane path state + helpers
// Shared by the brute-force kernel below and the BVH kernel (tor_kernels_bvh.cuh): ONE restatement of
// render.nim:62-67 (sample start), cameras.nim:47-57 and materials.nim:24-86 on the device.
struct Lane {
  Rng rng;
  V3 pix, att, o, d;
  double time;
  int32_t row, col, sample, depth;
};

// render.nim:64-66 + cameras.nim:47-57: jitter, lens sample, shutter time -> primary ray
__device__ __forceinline__ void start_sample(Lane& L, const tor_camera& cam, int32_t nrows, int32_t ncols) {
  double u = ((double)L.col + rng_u01(L.rng)) / (double)(ncols - 1);
  double v = ((double)L.row + rng_u01(L.rng)) / (double)(nrows - 1);
  double rx, ry;
  for (;;) {  // sampling.nim:64-68
    rx = rng_urange(L.rng, -1.0, 1.0);
    ry = rng_urange(L.rng, -1.0, 1.0);
    if (rx * rx + ry * ry + 0.0 * 0.0 < 1) break;
  }
  double rdx = rx * cam.lens_radius, rdy = ry * cam.lens_radius;  // Vec3 * scalar
  V3 cu = v3(cam.u[0], cam.u[1], cam.u[2]), cv = v3(cam.v[0], cam.v[1], cam.v[2]);
  V3 offset = cu * rdx + cv * rdy;
  V3 org = v3(cam.origin[0], cam.origin[1], cam.origin[2]);
  V3 llc = v3(cam.lower_left_corner[0], cam.lower_left_corner[1], cam.lower_left_corner[2]);
  V3 hor = v3(cam.horizontal[0], cam.horizontal[1], cam.horizontal[2]);
  V3 ver = v3(cam.vertical[0], cam.vertical[1], cam.vertical[2]);
  L.o = org + offset;
  L.d = (((llc + u * hor) + v * ver) - org) - offset;
  L.time = rng_urange(L.rng, cam.shutter_open, cam.shutter_close);
  L.att = v3(1, 1, 1);
  L.depth = 0;
}

// The surface the closest hit landed on, as the materials need it.
struct Surface {
  V3 center;  // at the ray's time (moving_spheres.nim:61 recomputes it on the accepted root)
  double inv_r;
  V3 albedo;
  double fuzz_or_ior;
  uint32_t mat_kind;
};

// Record fill (spheres.nim:41-46, core.nim:47-49) + scatter (materials.nim:24-86) + render.nim:35-38.
// Returns true when the sample ends here (absorbed, or depth exhausted -> black).
// ud_pre: unit_vector(L.d) when the caller has it already (Metal, Dielectric and the sky all need it: the BVH kernel
// computes it once per segment for all of them together), else NULL.
__device__ __forceinline__ bool shade_hit(Lane& L, double best_t, const Surface& S, int32_t max_depth,
                                          const V3* ud_pre = nullptr) {
  const V3 o = L.o, d = L.d;
  V3 p = o + best_t * d;
  V3 outward = (p - S.center) * S.inv_r;
  bool front_face = dot(d, outward) < 0;  // core.nim:47-49
  V3 n = front_face ? outward : -outward;
  bool scattered = true;
  V3 nd;
  double ntime = 0.0;  // rays.nim:19 default time (Metal / Dielectric)
  V3 matt = v3(1, 1, 1);
  if (S.mat_kind == TOR_LAMBERTIAN) {  // materials.nim:24-30 + sampling.nim:51-55
    double ang = rng_umax(L.rng, 6.283185307179586);
    double z = rng_urange(L.rng, -1.0, 1.0);
    double r = sqrt(1.0 - z * z);
    double sn, cs;
    detmath::sincos(ang, &sn, &cs);
    nd = n + v3(r * cs, r * sn, z);
    ntime = L.time;
    matt = S.albedo;
  } else if (S.mat_kind == TOR_METAL) {  // materials.nim:39-47
    V3 refl = reflect(ud_pre ? *ud_pre : unit_vector(d), n);
    V3 s;
    for (;;) {  // sampling.nim:45-49
      s.x = rng_urange(L.rng, -1, 1);
      s.y = rng_urange(L.rng, -1, 1);
      s.z = rng_urange(L.rng, -1, 1);
      if (len2(s) < 1.0) break;
    }
    nd = refl + S.fuzz_or_ior * s;
    scattered = dot(nd, n) > 0;
    matt = S.albedo;
  } else {  // materials.nim:62-86
    double ior = S.fuzz_or_ior;
    double eta = front_face ? 1.0 / ior : ior;
    V3 ud = ud_pre ? *ud_pre : unit_vector(d);
    double dt = dot(-ud, n);
    double cos_theta = (dt <= 1.0) ? dt : 1.0;
    double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
    bool do_reflect = eta * sin_theta > 1.0;
    if (!do_reflect) {
      double r0 = (1 - eta) / (1 + eta);  // schlick, materials.nim:55-60
      r0 *= r0;
      double reflect_prob = r0 + (1 - r0) * detmath::pow(1 - cos_theta, 5.0);
      do_reflect = rng_u01(L.rng) < reflect_prob;
    }
    if (do_reflect) {
      nd = reflect(ud, n);
    } else {  // rays.nim:30-37
      double ct = dot(-ud, n);
      V3 r_par = eta * (ud + ct * n);
      V3 r_perp = (-sqrt(1.0 - len2(r_par))) * n;
      nd = r_par + r_perp;
    }
  }
  if (!scattered) return true;  // render.nim:38 -> black
  L.att.x *= matt.x;            // render.nim:35-37
  L.att.y *= matt.y;
  L.att.z *= matt.z;
  L.o = p;
  L.d = nd;
  L.time = ntime;
  ++L.depth;
  return L.depth >= max_depth;  // render.nim:47 -> black
}

// render.nim:40-45 — the ray left the scene: sky colour times the attenuation so far
__device__ __forceinline__ V3 shade_miss(const Lane& L, const V3* ud_pre = nullptr) {
  V3 ud = ud_pre ? *ud_pre : unit_vector(L.d);
  double t = 0.5 * ud.y + 1.0;
  V3 color = v3((1.0 - t) + t * 0.5, (1.0 - t) + t * 0.7, (1.0 - t) + t);
  color.x *= L.att.x;
  color.y *= L.att.y;
  color.z *= L.att.z;
  return color;
}

// canvas.nim:47-54 `draw`
__device__ __forceinline__ void draw_pixel(double* out, V3 pix, double inv_spp, double inv_gamma) {
  out[0] = detmath::pow(inv_spp * pix.x, inv_gamma);
  out[1] = detmath::pow(inv_spp * pix.y, inv_gamma);
  out[2] = detmath::pow(inv_spp * pix.z, inv_gamma);
}

// ------------------------------------------------------------------------------ the kernel
// STAGE: 2 = whole blob staged in shared memory; 1 = filter records only; 0 = nothing (scene too large)
template <int BLOCK, int STAGE>
__global__ void __launch_bounds__(BLOCK, 512 / BLOCK) render_exact_kernel(const __grid_constant__ RenderParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t stage_bar;

  const int tid = threadIdx.x;
  const SceneView& sv = P.sv;
  const uint32_t staged_bytes = STAGE == 2 ? sv.total_bytes : (STAGE == 1 ? sv.hot_bytes : 0u);
  uint16_t* cand = reinterpret_cast<uint16_t*>(smem + staged_bytes);  // [kMaxCand][BLOCK]

  // ---- stage the scene blob: one elected thread issues TMA bulk copies, everyone waits on the mbarrier
  if (tid == 0) {
    mbar_init(&stage_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (staged_bytes) {
    if (tid == 0) {
      mbar_expect_tx(&stage_bar, staged_bytes);
      const uint32_t kChunk = 32768;
      for (uint32_t off = 0; off < staged_bytes; off += kChunk) {
        uint32_t n = staged_bytes - off < kChunk ? staged_bytes - off : kChunk;
        tma_bulk_g2s(smem + off, P.blob + off, n, &stage_bar);
      }
    }
    mbar_wait(&stage_bar, 0);
  }

  // generic pointers: shared memory when staged, the (L1/L2-resident) global blob otherwise
  const uint8_t* hot = STAGE >= 1 ? smem : P.blob;
  const double2* __restrict__ rec_static = reinterpret_cast<const double2*>(hot + sv.off_static);
  const double2* __restrict__ rec_ymov = reinterpret_cast<const double2*>(hot + sv.off_ymov);
  const double2* __restrict__ rec_gmov = reinterpret_cast<const double2*>(hot + sv.off_gmov);
  const TimeClass* __restrict__ classes = reinterpret_cast<const TimeClass*>(hot + sv.off_classes);
  const uint16_t* __restrict__ idx_static = reinterpret_cast<const uint16_t*>(hot + sv.off_idx_static);
  const uint16_t* __restrict__ idx_ymov = reinterpret_cast<const uint16_t*>(hot + sv.off_idx_ymov);
  const uint16_t* __restrict__ idx_gmov = reinterpret_cast<const uint16_t*>(hot + sv.off_idx_gmov);
  const tor_hittable* __restrict__ exact =
      reinterpret_cast<const tor_hittable*>((STAGE == 2 ? smem : P.blob) + sv.off_exact);

  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  const double t_min = 0.001;  // render.nim:28
  const unsigned long long total_px = (unsigned long long)P.nsel_rows * (unsigned long long)P.ncols;

  // ---- per-lane persistent state
  Lane L;
  L.pix = v3(0, 0, 0);
  L.att = v3(1, 1, 1);
  L.o = v3(0, 0, 0);
  L.d = v3(0, 0, 1);
  L.time = 0.0;
  L.row = L.col = L.sample = L.depth = 0;
  unsigned long long px = 0;  // index into the selected-pixel queue
  bool active = false;
  bool need_pixel = true;
  bool need_sample = false;
  unsigned long long seg_count = 0, ray_count = 0;

  for (;;) {
    // ------------------------------------------------------------ fetch work / start a sample
    if (need_pixel) {
      need_pixel = false;
      active = false;
      for (;;) {
        px = atomicAdd(P.work_counter, 1ull);
        if (px >= total_px) break;
        if (P.spp > 0) {
          int32_t ri = (int32_t)(px / (unsigned long long)P.ncols);
          L.col = (int32_t)(px - (unsigned long long)ri * (unsigned long long)P.ncols);
          L.row = P.row_begin + ri * P.row_step;
          rng_seed_pixel(L.rng, L.row, L.col);  // render.nim:59-60
          L.pix = v3(0, 0, 0);
          L.sample = 0;
          active = true;
          need_sample = true;
          break;
        }
        // no samples: draw() of the zero colour (0 * inf = NaN, as in canvas.nim:49-54)
        double* out = P.pixels + 3ull * px;
        out[0] = out[1] = out[2] = detmath::pow(P.inv_spp * 0.0, P.inv_gamma);
      }
    }
    if (!__any_sync(0xffffffffu, active)) break;

    if (active && need_sample) {
      start_sample(L, P.cam, P.nrows, P.ncols);
      need_sample = false;
      ++ray_count;
    }
    const V3 o = L.o, d = L.d;
    const double time = L.time;

    // ------------------------------------------------------------ A. conservative filter
    const double a = len2(d);  // spheres.nim:30 — also needed by the exact stage
    double best_t = INF;
    int best_i = 0x7fffffff;
    int ncand = 0;
    {
      const double K = kFilterSlack;
      const double inv_len = rsqrt(a);
      const double ndx = -d.x * inv_len, ndy = -d.y * inv_len, ndz = -d.z * inv_len;
      const double p2ox = 2.0 * o.x, p2oy = 2.0 * o.y, p2oz = 2.0 * o.z;
      const double oo = fma(o.x, o.x, fma(o.y, o.y, o.z * o.z));
      const double od = fma(o.x, d.x, fma(o.y, d.y, o.z * d.z)) * inv_len;
      // pass  <=>  hb^2 + 2 o.c - W  >  oo (1 - K) - K   (the trailing K is an absolute floor)
      const double thr = fma(-K, oo, oo) - K;

      // early resolve when the list is nearly full (rare): evaluate exactly and restart the list
      auto resolve = [&](int n) {
        for (int k = 0; k < n; ++k) {
          int i = cand[k * BLOCK + tid];
          double t = exact_first_root(exact + i, o, d, time, a, t_min);
          if (t < best_t || (t == best_t && i < best_i && t < INF)) {
            best_t = t;
            best_i = i;
          }
        }
      };
      auto push_mask = [&](uint32_t mask, const uint16_t* __restrict__ idx, int base) {
        while (mask) {
          int j = __ffs(mask) - 1;
          mask &= mask - 1;
          if (ncand == kMaxCand) {
            resolve(ncand);
            ncand = 0;
          }
          cand[ncand * BLOCK + tid] = idx[base + j];
          ++ncand;
        }
      };

      // static spheres: {cx, cy, cz, nW}
      for (int g = 0; g < sv.n_static_pad; g += 32) {
        uint32_t mask = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          double2 r0 = rec_static[2 * (g + j)], r1 = rec_static[2 * (g + j) + 1];
          double s = fma(p2ox, r0.x, fma(p2oy, r0.y, fma(p2oz, r1.x, r1.y)));
          double hb = fma(ndx, r0.x, fma(ndy, r0.y, fma(ndz, r1.x, od)));
          double s2 = fma(hb, hb, s);
          if (s2 > thr) mask |= (1u << j);
        }
        if (mask && active) push_mask(mask, idx_static, g);
      }
      // movers, one time class at a time
      for (int ci = 0; ci < sv.n_classes; ++ci) {
        const TimeClass tc = classes[ci];
        const double q = (time - tc.t0) / (tc.t1 - tc.t0);  // moving_spheres.nim:41-42
        // y-only movers: {cx, cz, nWxz, c0y, dcy, -}
        for (int g = tc.y_begin; g < tc.y_end; g += 32) {
          uint32_t mask = 0;
#pragma unroll 8
          for (int j = 0; j < 32; ++j) {
            double2 r0 = rec_ymov[3 * (g + j)], r1 = rec_ymov[3 * (g + j) + 1], r2 = rec_ymov[3 * (g + j) + 2];
            double cy = fma(q, r2.x, r1.y);
            double nW = fma(-cy, cy, r1.x);
            double s = fma(p2ox, r0.x, fma(p2oy, cy, fma(p2oz, r0.y, nW)));
            double hb = fma(ndx, r0.x, fma(ndy, cy, fma(ndz, r0.y, od)));
            double s2 = fma(hb, hb, s);
            if (s2 > thr) mask |= (1u << j);
          }
          if (mask && active) push_mask(mask, idx_ymov, g);
        }
        // general movers: {c0x, c0y, c0z, r2m, dcx, dcy, dcz, -}
        for (int g = tc.g_begin; g < tc.g_end; g += 32) {
          uint32_t mask = 0;
#pragma unroll 8
          for (int j = 0; j < 32; ++j) {
            double2 r0 = rec_gmov[4 * (g + j)], r1 = rec_gmov[4 * (g + j) + 1];
            double2 r2 = rec_gmov[4 * (g + j) + 2], r3 = rec_gmov[4 * (g + j) + 3];
            double cx = fma(q, r2.x, r0.x), cy = fma(q, r2.y, r0.y), cz = fma(q, r3.x, r1.x);
            double nW = fma(-cx, cx, fma(-cy, cy, fma(-cz, cz, r1.y)));
            double s = fma(p2ox, cx, fma(p2oy, cy, fma(p2oz, cz, nW)));
            double hb = fma(ndx, cx, fma(ndy, cy, fma(ndz, cz, od)));
            double s2 = fma(hb, hb, s);
            if (s2 > thr) mask |= (1u << j);
          }
          if (mask && active) push_mask(mask, idx_gmov, g);
        }
      }
      // ---------------------------------------------------------- B. exact stage
      resolve(ncand);
    }

    // ------------------------------------------------------------ C. shade
    if (active) {
      bool sample_done = false;
      V3 color = v3(0, 0, 0);
      if (P.max_depth <= 0) {  // render.nim:25 — the bounce loop body never runs; the sample is black
        sample_done = true;
      } else if (++seg_count, best_t < INF) {
        const tor_hittable* __restrict__ h = exact + best_i;
        Surface S;
        S.center = obj_center(h, time);
        S.inv_r = 1.0 / h->radius;
        S.albedo = v3(h->albedo[0], h->albedo[1], h->albedo[2]);
        S.fuzz_or_ior = h->fuzz_or_ior;
        S.mat_kind = h->mat_kind;
        sample_done = shade_hit(L, best_t, S, P.max_depth);
      } else {
        color = shade_miss(L);
        sample_done = true;
      }
      if (sample_done) {
        L.pix.x += color.x;  // render.nim:67
        L.pix.y += color.y;
        L.pix.z += color.z;
        ++L.sample;
        need_sample = true;
        if (L.sample >= P.spp) {
          draw_pixel(P.pixels + 3ull * px, L.pix, P.inv_spp, P.inv_gamma);
          need_pixel = true;
          need_sample = false;
          active = false;
        }
      }
    }
  }

  if (P.count_segments) {
    // warp-shuffle reduction, one atomic per warp
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
      seg_count += __shfl_down_sync(0xffffffffu, seg_count, ofs);
      ray_count += __shfl_down_sync(0xffffffffu, ray_count, ofs);
    }
    if ((tid & 31) == 0) {
      atomicAdd(P.counters + 0, ray_count);
      atomicAdd(P.counters + 1, seg_count);
    }
  }
}

// ------------------------------------------------------------------- FP64 peak microbenchmark
// Register-only DFMA chains: the denominator of the ALU roofline (MEASURED_PEAKS.json has no FP64 entry).
static constexpr int kPeakChains = 16;
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double m, double c) {
  double acc[kPeakChains];
#pragma unroll
  for (int k = 0; k < kPeakChains; ++k) acc[k] = (double)(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) acc[k] = fma(acc[k], m, c);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < kPeakChains; ++k) s += acc[k];
  if (s == 12345.678) *sink = s;  // never true; keeps the chains alive
}

}  // namespace tor
