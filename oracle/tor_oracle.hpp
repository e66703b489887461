// tor_oracle.hpp — CPU oracle: a float64 restatement of trace-of-radiance's render path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under trace_of_radiance_b200/ links, imports
// or executes this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs do, and only as the checker / the timed CPU baseline.
//
// The reference (mratsim/trace-of-radiance @ c174493) is Nim + the un-vendored Weave runtime;
// neither `nim` nor Weave exists in this environment, so the reference cannot be compiled
// (oracle/_ref is therefore absent — see DESIGN.md).  This file restates the algorithm
// operation by operation, in the reference's evaluation order, so that float64 results are
// bit-identical to what the Nim program computes given the same libm.  Every function cites
// the reference file:line it follows (paths relative to /root/reference/trace_of_radiance/).
//
// Parity pins (tests/test_oracle_*.py): RNG known-answer vectors, random_scene object counts
// and first objects, the camera, and the reference's own media/book2_motion_blur.png.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math (the reference build is scalar SSE2 with no
// contraction, README.md:78-82).  Templated on a Math policy: LibmMath (glibc sin/cos/pow —
// what the reference links) or DetMath (tor_detmath.h — the portable routines the GPU kernels
// use, giving a bit-comparable CPU image).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../trace_of_radiance_b200/csrc/tor_detmath.h"

namespace oracle {

// ------------------------------------------------------------------------------- math policy
struct LibmMath {
  static double sin_(double a) { return ::sin(a); }
  static double cos_(double a) { return ::cos(a); }
  static double pow_(double x, double y) { return ::pow(x, y); }
};
struct DetMath {
  static double sin_(double a) {
    double s, c;
    tor::detmath::sincos(a, &s, &c);
    return s;
  }
  static double cos_(double a) {
    double s, c;
    tor::detmath::sincos(a, &s, &c);
    return c;
  }
  static double pow_(double x, double y) { return tor::detmath::pow(x, y); }
};

// ------------------------------------------------------------------------ support/rng.nim
struct Rng {  // rng.nim:18-19
  uint64_t s0, s1, s2, s3;
};

// rng.nim:21-29
static inline uint64_t pair(int64_t x, int64_t y) { return ((uint64_t)x << 32) ^ (uint64_t)y; }

// rng.nim:31-36 — NB the first multiplier is used twice (sic); parity depends on it.
static inline uint64_t splitMix64(uint64_t& state) {
  state += 0x9e3779b97f4a7c15ull;
  uint64_t r = state;
  r = (r ^ (r >> 30)) * 0xbf58476d1ce4e5b9ull;
  r = (r ^ (r >> 27)) * 0xbf58476d1ce4e5b9ull;
  r = r ^ (r >> 31);
  return r;
}

// rng.nim:38-44
static inline void seed(Rng& rng, uint64_t x) {
  uint64_t sm = x;
  rng.s0 = splitMix64(sm);
  rng.s1 = splitMix64(sm);
  rng.s2 = splitMix64(sm);
  rng.s3 = splitMix64(sm);
}
// rng.nim:46-53
static inline void seed(Rng& rng, int64_t x, int64_t y) { seed(rng, pair(x, y)); }

// rng.nim:58-74 — xoshiro256+
static inline uint64_t next(Rng& rng) {
  uint64_t result = rng.s0 + rng.s3;
  uint64_t t = rng.s1 << 17;
  rng.s2 ^= rng.s0;
  rng.s3 ^= rng.s1;
  rng.s1 ^= rng.s2;
  rng.s0 ^= rng.s3;
  rng.s2 ^= t;
  rng.s3 = (rng.s3 << 45) | (rng.s3 >> 19);
  return result;
}

static inline double mantissa_to_unit(uint64_t bits) {  // rng.nim:131-133
  uint64_t fl = (bits >> 12) | 0x3ff0000000000000ull;
  double d;
  memcpy(&d, &fl, 8);
  return d - 1.0;
}
// rng.nim:129-133
static inline double uniform01(Rng& rng) { return mantissa_to_unit(next(rng)); }
// rng.nim:135-143
static inline double uniform_max(Rng& rng, double maxExcl) { return mantissa_to_unit(next(rng)) * maxExcl; }
// rng.nim:116-127; Nim's max(x, y) is `if y <= x: x else: y`
static inline double uniform_range(Rng& rng, double minIncl, double maxExcl) {
  double debiased = mantissa_to_unit(next(rng));
  double v = debiased * (maxExcl - minIncl) + minIncl;
  return (v <= minIncl) ? minIncl : v;
}

// ------------------------------------------------------------------- primitives/vec3s.nim
struct Vec3 {
  double x, y, z;
};
static inline Vec3 vec3(double x, double y, double z) { return Vec3{x, y, z}; }
static inline double length_squared(Vec3 u) { return u.x * u.x + u.y * u.y + u.z * u.z; }  // :23-24
static inline double length(Vec3 u) { return sqrt(length_squared(u)); }                    // :26-27
static inline Vec3 operator+(Vec3 u, Vec3 v) { return Vec3{u.x + v.x, u.y + v.y, u.z + v.z}; }
static inline Vec3 operator-(Vec3 u, Vec3 v) { return Vec3{u.x - v.x, u.y - v.y, u.z - v.z}; }
static inline Vec3 operator-(Vec3 u) { return Vec3{-u.x, -u.y, -u.z}; }
static inline Vec3 operator*(Vec3 u, double s) { return Vec3{u.x * s, u.y * s, u.z * s}; }  // :86-88
static inline Vec3 operator*(double s, Vec3 u) { return u * s; }                             // :90-91
static inline Vec3 operator/(Vec3 u, double s) { return u * (1.0 / s); }                     // :93-94 (sic)
static inline double dot(Vec3 u, Vec3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }       // :96-98
static inline Vec3 cross(Vec3 u, Vec3 v) {                                                   // :100-104
  return Vec3{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
}
static inline Vec3 unit_vector(Vec3 u) { return u / length(u); }  // :106-107

// ------------------------------------------------------------------- primitives/rays.nim
struct Ray {  // :14-17
  Vec3 origin, direction;
  double time;
};
static inline Vec3 ray_at(const Ray& r, double t) { return r.origin + t * r.direction; }  // :24-25
static inline Vec3 reflect(Vec3 u, Vec3 n) { return u - 2 * dot(u, n) * n; }              // :27-28
static inline Vec3 refract(Vec3 uv, Vec3 n, double eta) {                                 // :30-37
  double cos_theta = dot(-uv, n);
  Vec3 r_out_parallel = eta * (uv + cos_theta * n);
  Vec3 r_out_perp = -sqrt(1.0 - length_squared(r_out_parallel)) * n;
  return r_out_parallel + r_out_perp;
}

// --------------------------------------------------------------------- physics/core.nim
enum MaterialKind : uint32_t { kLambertian = 0, kMetal = 1, kDielectric = 2 };  // core.nim:24-28
enum HittableKind : uint32_t { kSphere = 0, kMovingSphere = 1 };  // hittables_variants.nim:50-52

// One flat record per HittableVariant (spheres.nim:15-18, moving_spheres.nim:15-20, core.nim:16-22).
// Same field order as include/tor_b200.h's tor_hittable so tests can hand one array to both.
struct Hittable {
  uint32_t kind;      // HittableKind
  uint32_t mat_kind;  // MaterialKind
  double center0[3];
  double center1[3];  // MovingSphere only
  double time0, time1;
  double radius;
  double albedo[3];     // Lambertian / Metal
  double fuzz_or_ior;   // Metal: fuzz (already clamped, materials.nim:35-37); Dielectric: ior
};
static_assert(sizeof(Hittable) == 112, "flat hittable layout");

struct HitRecord {  // core.nim:31-36 (material referenced by object index instead of copied)
  Vec3 p, normal;
  const Hittable* obj;
  double t;
  bool front_face;
};

// core.nim:47-49
static inline void set_face_normal(HitRecord& rec, const Ray& r, Vec3 outward_normal) {
  rec.front_face = dot(r.direction, outward_normal) < 0;
  rec.normal = rec.front_face ? outward_normal : -outward_normal;
}

// ------------------------------------------------------------------ physics/hittables/*.nim
static inline Vec3 c0(const Hittable& h) { return Vec3{h.center0[0], h.center0[1], h.center0[2]}; }
static inline Vec3 c1(const Hittable& h) { return Vec3{h.center1[0], h.center1[1], h.center1[2]}; }

// moving_spheres.nim:39-44
static inline Vec3 moving_center(const Hittable& h, double time) {
  return c0(h) + (((time - h.time0) / (h.time1 - h.time0)) * (c1(h) - c0(h)));
}

// spheres.nim:28-49 and moving_spheres.nim:46-67 (identical but for center)
static inline bool hit_one(const Hittable& h, const Ray& r, double t_min, double t_max, HitRecord& rec) {
  const bool moving = h.kind == kMovingSphere;
  Vec3 center = moving ? moving_center(h, r.time) : c0(h);
  Vec3 oc = r.origin - center;
  double a = length_squared(r.direction);
  double half_b = dot(oc, r.direction);
  double c = length_squared(oc) - h.radius * h.radius;
  double discriminant = half_b * half_b - a * c;
  if (discriminant > 0) {
    double root = sqrt(discriminant);
    double sol = (-half_b - root) / a;
    for (int k = 0; k < 2; ++k) {
      if (t_min < sol && sol < t_max) {
        rec.t = sol;
        rec.p = ray_at(r, rec.t);
        Vec3 ctr = moving ? moving_center(h, r.time) : c0(h);  // recomputed, moving_spheres.nim:61
        Vec3 outward_normal = (rec.p - ctr) / h.radius;
        set_face_normal(rec, r, outward_normal);
        rec.obj = &h;
        return true;
      }
      sol = (-half_b + root) / a;
    }
  }
  return false;
}

// hittables_lists.nim:48-55
static inline bool hit_list(const Hittable* objs, int64_t n, const Ray& r, double t_min, double t_max,
                            HitRecord& rec) {
  double closest_so_far = t_max;
  bool result = false;
  for (int64_t i = 0; i < n; ++i) {
    if (hit_one(objs[i], r, t_min, closest_so_far, rec)) {
      closest_so_far = rec.t;
      result = true;
    }
  }
  return result;
}

// ------------------------------------------------------------------------- sampling.nim
static inline Vec3 random_in_unit_sphere(Rng& rng) {  // :45-49 (x, y, z drawn in field order :35-38)
  for (;;) {
    Vec3 p;
    p.x = uniform_range(rng, -1, 1);
    p.y = uniform_range(rng, -1, 1);
    p.z = uniform_range(rng, -1, 1);
    if (length_squared(p) < 1.0) return p;
  }
}
template <class M>
static inline Vec3 random_unit_vector(Rng& rng) {  // :51-55
  const double TWO_PI = 2 * 3.141592653589793;    // Nim folds 2*PI at compile time
  double a = uniform_max(rng, TWO_PI);
  double z = uniform_range(rng, -1.0, 1.0);
  double r = sqrt(1.0 - z * z);
  return vec3(r * M::cos_(a), r * M::sin_(a), z);
}
static inline Vec3 random_in_unit_disk(Rng& rng) {  // :64-68
  for (;;) {
    Vec3 p;
    p.x = uniform_range(rng, -1.0, 1.0);
    p.y = uniform_range(rng, -1.0, 1.0);
    p.z = 0;
    if (length_squared(p) < 1) return p;
  }
}
static inline Vec3 random_attenuation(Rng& rng) {  // :71-74
  Vec3 r;
  r.x = uniform01(rng);
  r.y = uniform01(rng);
  r.z = uniform01(rng);
  return r;
}
static inline Vec3 random_attenuation(Rng& rng, double mn, double mx) {  // :81-84
  Vec3 r;
  r.x = uniform_range(rng, mn, mx);
  r.y = uniform_range(rng, mn, mx);
  r.z = uniform_range(rng, mn, mx);
  return r;
}

// --------------------------------------------------------------------- physics/cameras.nim
struct Camera {  // :15-22 — 24 contiguous float64 in declaration order
  Vec3 origin, lower_left_corner, horizontal, vertical, u, v, w;
  double lens_radius, shutterOpen, shutterClose;
};
static_assert(sizeof(Camera) == 192, "camera layout");

// :24-45
static inline Camera make_camera(Vec3 lookFrom, Vec3 lookAt, Vec3 view_up, double vfov_degrees,
                                 double aspect_ratio, double aperture, double focus_distance,
                                 double shutterOpen, double shutterClose) {
  Camera cam;
  const double RadPerDeg = 3.141592653589793 / 180.0;  // std/math degToRad
  double theta = vfov_degrees * RadPerDeg;
  double h = tan(theta / 2.0);
  double viewport_height = 2.0 * h;
  double viewport_width = aspect_ratio * viewport_height;
  cam.w = unit_vector(lookFrom - lookAt);
  cam.u = unit_vector(cross(view_up, cam.w));
  cam.v = cross(cam.w, cam.u);
  cam.origin = lookFrom;
  cam.horizontal = focus_distance * viewport_width * cam.u;
  cam.vertical = focus_distance * viewport_height * cam.v;
  cam.lower_left_corner = cam.origin - cam.horizontal / 2 - cam.vertical / 2 - focus_distance * cam.w;
  cam.lens_radius = aperture / 2;
  cam.shutterOpen = shutterOpen;
  cam.shutterClose = shutterClose;
  return cam;
}

// :47-57
static inline Ray camera_ray(const Camera& self, double s, double t, Rng& rng) {
  Vec3 rd = self.lens_radius * random_in_unit_disk(rng);
  Vec3 offset = self.u * rd.x + self.v * rd.y;
  Ray r;
  r.origin = self.origin + offset;
  r.direction = self.lower_left_corner + s * self.horizontal + t * self.vertical - self.origin - offset;
  r.time = uniform_range(rng, self.shutterOpen, self.shutterClose);
  return r;
}

// ------------------------------------------------------------------- physics/materials.nim
template <class M>
static inline double schlick(double cosine, double ri) {  // :55-60
  double r0 = (1 - ri) / (1 + ri);
  r0 *= r0;
  return r0 + (1 - r0) * M::pow_(1 - cosine, 5);
}

// dispatcher :91-96 over Lambertian :24-30, Metal :39-47, Dielectric :62-86
template <class M>
static inline bool scatter(const Hittable& obj, const Ray& r_in, const HitRecord& rec, Rng& rng,
                           Vec3& attenuation, Ray& scattered) {
  switch (obj.mat_kind) {
    case kLambertian: {
      Vec3 scatter_direction = rec.normal + random_unit_vector<M>(rng);
      scattered = Ray{rec.p, scatter_direction, r_in.time};
      attenuation = Vec3{obj.albedo[0], obj.albedo[1], obj.albedo[2]};
      return true;
    }
    case kMetal: {
      Vec3 reflected = reflect(unit_vector(r_in.direction), rec.normal);
      scattered = Ray{rec.p, reflected + obj.fuzz_or_ior * random_in_unit_sphere(rng), 0.0};
      if (dot(scattered.direction, rec.normal) > 0) {
        attenuation = Vec3{obj.albedo[0], obj.albedo[1], obj.albedo[2]};
        return true;
      }
      return false;
    }
    default: {  // kDielectric
      attenuation = Vec3{1, 1, 1};
      double eta = rec.front_face ? 1.0 / obj.fuzz_or_ior : obj.fuzz_or_ior;
      Vec3 unit_direction = unit_vector(r_in.direction);
      double d = dot(-unit_direction, rec.normal);
      double cos_theta = (d <= 1.0) ? d : 1.0;  // Nim min(x, y): if x <= y: x else: y
      double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
      if (eta * sin_theta > 1.0) {
        scattered = Ray{rec.p, reflect(unit_direction, rec.normal), 0.0};
        return true;
      }
      double reflect_prob = schlick<M>(cos_theta, eta);
      if (uniform01(rng) < reflect_prob) {
        scattered = Ray{rec.p, reflect(unit_direction, rec.normal), 0.0};
        return true;
      }
      scattered = Ray{rec.p, refract(unit_direction, rec.normal, eta), 0.0};
      return true;
    }
  }
}

// ------------------------------------------------------------------------------ render.nim
struct Counters {
  uint64_t primary_rays = 0;
  uint64_t segments = 0;      // world.hit calls
  uint64_t sphere_tests = 0;  // segments * N
};

// :21-47
template <class M>
static inline Vec3 radiance(Ray ray, const Hittable* world, int64_t n, int64_t max_depth, Rng& rng,
                            uint64_t& segments) {
  Vec3 attenuation{1.0, 1.0, 1.0};
  for (int64_t depth = 0; depth < max_depth; ++depth) {
    HitRecord rec;
    ++segments;
    if (hit_list(world, n, ray, 0.001, INFINITY, rec)) {
      Vec3 materialAttenuation;
      Ray scattered;
      if (scatter<M>(*rec.obj, ray, rec, rng, materialAttenuation, scattered)) {
        attenuation.x *= materialAttenuation.x;  // colors.nim:44-48
        attenuation.y *= materialAttenuation.y;
        attenuation.z *= materialAttenuation.z;
        ray = scattered;
        continue;
      }
      return Vec3{0, 0, 0};
    }
    Vec3 unit_direction = unit_vector(ray.direction);
    double t = 0.5 * unit_direction.y + 1.0;  // (sic) render.nim:42
    Vec3 result = (1.0 - t) * Vec3{1, 1, 1} + t * Vec3{0.5, 0.7, 1};
    result.x *= attenuation.x;  // colors.nim:62-66
    result.y *= attenuation.y;
    result.z *= attenuation.z;
    return result;
  }
  return Vec3{0, 0, 0};
}

// One pixel of render.nim:59-68 (+ canvas.nim:47-54 `draw`).
template <class M>
static inline void render_pixel(double* out3, int64_t row, int64_t col, int32_t nrows, int32_t ncols,
                                int32_t spp, float gamma_correction, const Camera& cam,
                                const Hittable* world, int64_t n, int64_t max_depth, uint64_t& segments) {
  Rng rng;
  seed(rng, row, col);
  Vec3 pixel{0, 0, 0};
  for (int32_t s = 0; s < spp; ++s) {
    double u = ((double)col + uniform01(rng)) / (double)(ncols - 1);
    double v = ((double)row + uniform01(rng)) / (double)(nrows - 1);
    Ray r = camera_ray(cam, u, v, rng);
    Vec3 c = radiance<M>(r, world, n, max_depth, rng, segments);
    pixel.x += c.x;
    pixel.y += c.y;
    pixel.z += c.z;
  }
  double scale = 1.0 / (double)spp;                 // canvas.nim:49
  double gamma = 1.0 / (double)gamma_correction;    // canvas.nim:50 (float32 widened)
  out3[0] = M::pow_(scale * pixel.x, gamma);
  out3[1] = M::pow_(scale * pixel.y, gamma);
  out3[2] = M::pow_(scale * pixel.z, gamma);
}

// ---- split-stream ("fast") mode: NOT in the reference.  It is the restatement of what TOR_MODE_FAST of
// include/tor_b200.h defines, so that the CUDA path can be checked bit for bit in that mode too:
//   * the pixel's sample loop (render.nim:62-67) is cut into nsub = 2^k consecutive ranges
//     [floor(j*spp/nsub), floor((j+1)*spp/nsub));
//   * range j has its own xoshiro256+ state: outputs 4j .. 4j+3 of the SplitMix64 sequence that rng.nim:46-53
//     starts at pair(row, col) — range 0 is the reference's own stream, and nsub = 1 is the reference's render;
//   * the per-range sums are added pairwise, lower index on the left: ((s0+s1)+(s2+s3))+...
static inline void seed_substream(Rng& rng, int64_t row, int64_t col, uint32_t sub) {
  uint64_t sm = pair(row, col) + (uint64_t)(4u * sub) * 0x9e3779b97f4a7c15ull;  // SplitMix64 advanced 4*sub steps
  rng.s0 = splitMix64(sm);
  rng.s1 = splitMix64(sm);
  rng.s2 = splitMix64(sm);
  rng.s3 = splitMix64(sm);
}

// Linear (pre-draw) sum of one pixel in split-stream mode; sq3 (optional) receives the per-channel sum of squares
// of the sample colours (for the Monte-Carlo standard error the statistical parity test needs).
template <class M>
static inline Vec3 pixel_sum_split(int64_t row, int64_t col, int32_t nrows, int32_t ncols, int32_t spp,
                                   uint32_t nsub, const Camera& cam, const Hittable* world, int64_t n,
                                   int64_t max_depth, uint64_t& segments, double* sq3) {
  Vec3 part[64];
  for (uint32_t j = 0; j < nsub; ++j) {
    Rng rng;
    seed_substream(rng, row, col, j);
    const int32_t s_begin = (int32_t)(((uint64_t)j * (uint64_t)spp) / nsub);
    const int32_t s_end = (int32_t)(((uint64_t)(j + 1) * (uint64_t)spp) / nsub);
    Vec3 pixel{0, 0, 0};
    for (int32_t s = s_begin; s < s_end; ++s) {
      double u = ((double)col + uniform01(rng)) / (double)(ncols - 1);
      double v = ((double)row + uniform01(rng)) / (double)(nrows - 1);
      Ray r = camera_ray(cam, u, v, rng);
      Vec3 c = radiance<M>(r, world, n, max_depth, rng, segments);
      pixel.x += c.x;
      pixel.y += c.y;
      pixel.z += c.z;
      if (sq3) {
        sq3[0] += c.x * c.x;
        sq3[1] += c.y * c.y;
        sq3[2] += c.z * c.z;
      }
    }
    part[j] = pixel;
  }
  for (uint32_t w = 1; w < nsub; w <<= 1)
    for (uint32_t j = 0; j + w < nsub; j += 2 * w) part[j] = part[j] + part[j + w];
  return part[0];
}

// The split-stream render over the selected rows.  linear_sum (optional, canvas-sized) receives the pre-draw sums,
// sum_sq (optional) the per-channel sums of squares.
template <class M>
static void render_split(double* pixels, int32_t nrows, int32_t ncols, int32_t spp, float gamma_correction,
                         const Camera& cam, const Hittable* world, int64_t n, int64_t max_depth, int32_t row_begin,
                         int32_t row_end, int32_t row_step, uint32_t nsub, double* linear_sum, double* sum_sq,
                         Counters* counters) {
  uint64_t segments = 0;
  int64_t nsel = (row_end > row_begin) ? (row_end - row_begin + row_step - 1) / row_step : 0;
  const double scale = 1.0 / (double)spp;               // canvas.nim:49
  const double gamma = 1.0 / (double)gamma_correction;  // canvas.nim:50
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : segments) collapse(2)
  for (int64_t ri = 0; ri < nsel; ++ri) {
    for (int64_t cb = 0; cb < (ncols + 31) / 32; ++cb) {
      int64_t row = row_begin + ri * row_step;
      int64_t cend = (cb + 1) * 32 < ncols ? (cb + 1) * 32 : ncols;
      for (int64_t col = cb * 32; col < cend; ++col) {
        const int64_t at = 3 * (row * ncols + col);
        double sq[3] = {0, 0, 0};
        Vec3 sum = pixel_sum_split<M>(row, col, nrows, ncols, spp, nsub, cam, world, n, max_depth, segments,
                                      sum_sq ? sq : nullptr);
        if (linear_sum) {
          linear_sum[at] = sum.x;
          linear_sum[at + 1] = sum.y;
          linear_sum[at + 2] = sum.z;
        }
        if (sum_sq) {
          sum_sq[at] = sq[0];
          sum_sq[at + 1] = sq[1];
          sum_sq[at + 2] = sq[2];
        }
        if (pixels) {
          pixels[at] = M::pow_(scale * sum.x, gamma);
          pixels[at + 1] = M::pow_(scale * sum.y, gamma);
          pixels[at + 2] = M::pow_(scale * sum.z, gamma);
        }
      }
    }
  }
  if (counters) {
    counters->primary_rays += (uint64_t)nsel * ncols * spp;
    counters->segments += segments;
    counters->sphere_tests += segments * (uint64_t)n;
  }
}

// render.nim:49-68 over rows row_begin, row_begin+row_step, ... < row_end.  `pixels` is the
// full canvas (nrows*ncols*3 doubles, row 0 = bottom); only the selected rows are written.
template <class M>
static void render(double* pixels, int32_t nrows, int32_t ncols, int32_t spp, float gamma_correction,
                   const Camera& cam, const Hittable* world, int64_t n, int64_t max_depth,
                   int32_t row_begin, int32_t row_end, int32_t row_step, Counters* counters) {
  uint64_t segments = 0;
  int64_t nsel = (row_end > row_begin) ? (row_end - row_begin + row_step - 1) / row_step : 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : segments) collapse(2)
  for (int64_t ri = 0; ri < nsel; ++ri) {
    for (int64_t cb = 0; cb < (ncols + 31) / 32; ++cb) {
      int64_t row = row_begin + ri * row_step;
      int64_t cend = (cb + 1) * 32 < ncols ? (cb + 1) * 32 : ncols;
      for (int64_t col = cb * 32; col < cend; ++col) {
        render_pixel<M>(pixels + 3 * (row * ncols + col), row, col, nrows, ncols, spp, gamma_correction,
                        cam, world, n, max_depth, segments);
      }
    }
  }
  if (counters) {
    counters->primary_rays += (uint64_t)nsel * ncols * spp;
    counters->segments += segments;
    counters->sphere_tests += segments * (uint64_t)n;
  }
}

// ------------------------------------------------------------------------------- io/ppm.nim
// :15-16 `int(256 * clamp(c, 0.0, 0.999))`; a NaN channel (UB in the reference) is defined as 0.
static inline int ppm_conv(double c) {
  if (c != c) return 0;
  double cl = c < 0.0 ? 0.0 : (c > 0.999 ? 0.999 : c);  // safe_math.nim:10-14
  return (int)(256 * cl);
}
// :14-27 — rows from nrows-1 down to 0
static inline std::string export_ppm(const double* pixels, int32_t nrows, int32_t ncols) {
  std::string out = "P3\n" + std::to_string(ncols) + " " + std::to_string(nrows) + "\n255\n";
  for (int32_t i = nrows - 1; i >= 0; --i)
    for (int32_t j = 0; j < ncols; ++j) {
      const double* p = pixels + 3 * ((int64_t)i * ncols + j);
      out += std::to_string(ppm_conv(p[0])) + " " + std::to_string(ppm_conv(p[1])) + " " +
             std::to_string(ppm_conv(p[2])) + "\n";
    }
  return out;
}

// ------------------------------------------------------------------------------- scenes.nim
static inline Hittable make_sphere(Vec3 center, double radius, uint32_t mat, Vec3 albedo, double fz) {
  Hittable h;
  memset(&h, 0, sizeof(h));
  h.kind = kSphere;
  h.mat_kind = mat;
  h.center0[0] = center.x; h.center0[1] = center.y; h.center0[2] = center.z;
  h.radius = radius;
  h.albedo[0] = albedo.x; h.albedo[1] = albedo.y; h.albedo[2] = albedo.z;
  h.fuzz_or_ior = fz;
  return h;
}
static inline Hittable make_moving_sphere(Vec3 center0, double time0, Vec3 center1, double time1,
                                          double radius, uint32_t mat, Vec3 albedo, double fz) {
  Hittable h = make_sphere(center0, radius, mat, albedo, fz);
  h.kind = kMovingSphere;
  h.center1[0] = center1.x; h.center1[1] = center1.y; h.center1[2] = center1.z;
  h.time0 = time0;
  h.time1 = time1;
  return h;
}
static inline double metal_fuzz(double fuzz) { return fuzz <= 1.0 ? fuzz : 1.0; }  // materials.nim:35-37

// scenes.nim:13-50, generalised to a (2*half)^2 grid: half = 11 is the reference scene; other
// values give SURVEY §8(d)'s random_scene_grid used for the stress / roofline-sweep configs.
static inline std::vector<Hittable> random_scene(Rng& rng, int half = 11) {
  std::vector<Hittable> w;
  w.push_back(make_sphere(vec3(0, -1000, 0), 1000, kLambertian, vec3(0.5, 0.5, 0.5), 0));
  for (int a = -half; a < half; ++a)
    for (int b = -half; b < half; ++b) {
      double cx = (double)a + 0.9 * uniform01(rng);
      double cz = (double)b + 0.9 * uniform01(rng);
      Vec3 center = vec3(cx, 0.2, cz);
      if (length(center - vec3(4, 0.2, 0)) > 0.9) {
        double choose_mat = uniform01(rng);
        if (choose_mat < 0.8) {
          Vec3 a1 = random_attenuation(rng);
          Vec3 a2 = random_attenuation(rng);
          Vec3 albedo{a1.x * a2.x, a1.y * a2.y, a1.z * a2.z};  // colors.nim:50-54
          Vec3 center2 = center + vec3(0, uniform_max(rng, 0.5), 0);
          w.push_back(make_moving_sphere(center, 0.0, center2, 1.0, 0.2, kLambertian, albedo, 0));
        } else if (choose_mat < 0.95) {
          Vec3 albedo = random_attenuation(rng, 0.5, 1);
          double fuzz = uniform_max(rng, 0.5);
          w.push_back(make_sphere(center, 0.2, kMetal, albedo, metal_fuzz(fuzz)));
        } else {
          w.push_back(make_sphere(center, 0.2, kDielectric, vec3(0, 0, 0), 1.5));
        }
      }
    }
  w.push_back(make_sphere(vec3(0, 1, 0), 1.0, kDielectric, vec3(0, 0, 0), 1.5));
  w.push_back(make_sphere(vec3(-4, 1, 0), 1.0, kLambertian, vec3(0.4, 0.2, 0.1), 0));
  w.push_back(make_sphere(vec3(4, 1, 0), 1.0, kMetal, vec3(0.7, 0.6, 0.5), metal_fuzz(0.0)));
  return w;
}

// ---------------------------------------------------------------------- scenes_animated.nim
struct AnimSphere {  // :42-55
  double velocity, pos_y, coef_restitution, x, z, radius;
  uint32_t mat_kind;
  Vec3 albedo;
  double fuzz_or_ior;
};
struct Animation {  // :57-69
  int32_t nrows, ncols;
  float dt, t_min, t_max, t;
  double lookFromAngle;
  std::vector<AnimSphere> spheres;
};

// :90-154
static inline Animation random_moving_spheres(Rng& rng, int32_t height, int32_t width, float dt, float t_min,
                                              float t_max) {
  Animation an;
  an.nrows = height;
  an.ncols = width;
  an.dt = dt;
  an.t_min = t_min;
  an.t_max = t_max;
  an.t = 0.0f;
  an.lookFromAngle = 2 * 3.141592653589793;
  const double SmallRadius = 0.2;
  for (int a = -20; a < 20; ++a)
    for (int b = -20; b < 20; ++b) {
      double cx = (double)a + 0.9 * uniform01(rng);
      double cz = (double)b + 0.9 * uniform01(rng);
      Vec3 center = vec3(cx, SmallRadius, cz);
      if (length(center - vec3(4, SmallRadius, 0)) > 0.9) {
        double choose_mat = uniform01(rng);
        AnimSphere s;
        s.x = center.x;
        s.pos_y = center.y;
        s.z = center.z;
        s.radius = SmallRadius;
        if (choose_mat < 0.65) {
          Vec3 a1 = random_attenuation(rng);
          Vec3 a2 = random_attenuation(rng);
          s.albedo = Vec3{a1.x * a2.x, a1.y * a2.y, a1.z * a2.z};
          s.coef_restitution = 0.6;
          s.velocity = 10.0 + (4 * uniform01(rng) - 2.0);  // random(float32) resolves to the f64 draw
          s.mat_kind = kLambertian;
          s.fuzz_or_ior = 0;
        } else if (choose_mat < 0.95) {
          s.albedo = random_attenuation(rng, 0.5, 1);
          double fuzz = uniform_max(rng, 0.5);
          s.coef_restitution = 0.5;
          s.velocity = 10.0 + (4 * uniform01(rng) - 2.0);
          s.mat_kind = kMetal;
          s.fuzz_or_ior = metal_fuzz(fuzz);
        } else {
          s.albedo = vec3(0, 0, 0);
          s.coef_restitution = 0.5;
          s.velocity = 10.0 + (4 * uniform01(rng) - 2.0);
          s.mat_kind = kDielectric;
          s.fuzz_or_ior = 1.5;
        }
        an.spheres.push_back(s);
      }
    }
  return an;
}

// :156-174
static inline void anim_step(Animation& an) {
  an.lookFromAngle -= 2.0 * 3.141592653589793 / 1200.0;
  an.t += an.dt;  // float32 accumulate
  const double G = 9.80665, SmallRadius = 0.2;
  for (AnimSphere& s : an.spheres) {
    if (s.velocity < 0.0 && s.pos_y < SmallRadius)
      s.velocity = -s.coef_restitution * s.velocity;
    else
      s.velocity -= G * (double)an.dt;
    s.pos_y += s.velocity * (double)an.dt;
  }
}

// One iteration of `iterator scenes` :176-225.  Returns false when t >= t_max (no frame).
static inline bool anim_next_frame(Animation& an, int skip, bool first, Camera& cam, std::vector<Hittable>& scene) {
  if (first)
    while (an.t < an.t_min) anim_step(an);
  else
    for (int i = 0; i < skip; ++i) anim_step(an);
  if (!(an.t < an.t_max)) return false;
  double aspect_ratio = (double)an.ncols / (double)an.nrows;
  const double r = sqrt(200.0);
  Vec3 lookFrom = vec3(r * cos(an.lookFromAngle), 2.0, r * sin(an.lookFromAngle));
  cam = make_camera(lookFrom, vec3(4, 1, 0), vec3(0, 1, 0), 20.0, aspect_ratio, 0.1, 10.0, 0.0, 0.0);
  scene.clear();
  scene.push_back(make_sphere(vec3(0, -1000, 0), 1000, kLambertian, vec3(0.5, 0.5, 0.5), 0));
  for (const AnimSphere& s : an.spheres)
    scene.push_back(make_sphere(vec3(s.x, s.pos_y, s.z), s.radius, s.mat_kind, s.albedo, s.fuzz_or_ior));
  scene.push_back(make_sphere(vec3(0, 1, 0), 1.0, kDielectric, vec3(0, 0, 0), 1.5));
  scene.push_back(make_sphere(vec3(-4, 1, 0), 1.0, kLambertian, vec3(0.4, 0.2, 0.1), 0));
  scene.push_back(make_sphere(vec3(4, 1, 0), 1.0, kMetal, vec3(0.7, 0.6, 0.5), metal_fuzz(0.0)));
  return true;
}

}  // namespace oracle
