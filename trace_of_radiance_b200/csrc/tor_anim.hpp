// tor_anim.hpp — host-side restatement of the reference's scene generators' building blocks, shared by tor_host.cc
// (tor_random_scene, tor_animation_*: the host iterator) and tor_api.cu (tor_animation_dev_*: the device-resident
// animation, which keeps only time, camera angle and frame logic on the host).  Test against the oracle, not the
// other way round.  Must be compiled without floating-point contraction (-ffp-contract=off).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/tor_b200.h"

namespace tor_anim {

// ---- support/rng.nim:18-74,116-143 (host copy for the scene generators) ----
struct HostRng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& st) {  // rng.nim:31-36 — first multiplier used twice (sic)
    st += 0x9e3779b97f4a7c15ull;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0xbf58476d1ce4e5b9ull;
    return z ^ (z >> 31);
  }
  explicit HostRng(uint64_t seed) {  // rng.nim:38-44
    for (int i = 0; i < 4; ++i) s[i] = splitmix(seed);
  }
  uint64_t next() {  // rng.nim:58-74
    uint64_t r = s[0] + s[3], t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = (s[3] << 45) | (s[3] >> 19);
    return r;
  }
  double u01() {  // rng.nim:129-133
    uint64_t b = (next() >> 12) | 0x3ff0000000000000ull;
    double d;
    memcpy(&d, &b, 8);
    return d - 1.0;
  }
  double umax(double mx) { return u01() * mx; }  // rng.nim:135-143
  double urange(double lo, double hi) {          // rng.nim:116-127
    double v = u01() * (hi - lo) + lo;
    return v <= lo ? lo : v;
  }
};

struct V {
  double x, y, z;
};
inline V sub(V a, V b) { return V{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V mul(V a, double s) { return V{a.x * s, a.y * s, a.z * s}; }
inline V cross(V a, V b) { return V{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double len(V a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V unit(V a) { return mul(a, 1.0 / len(a)); }  // vec3s.nim:93-94,106-107: `/` multiplies by the reciprocal
inline void put(double* d, V a) {
  d[0] = a.x;
  d[1] = a.y;
  d[2] = a.z;
}

inline tor_hittable mk_sphere(V c, double r, uint32_t mat, V albedo, double fz) {
  tor_hittable h;
  memset(&h, 0, sizeof(h));
  h.kind = TOR_SPHERE;
  h.mat_kind = mat;
  put(h.center0, c);
  h.radius = r;
  put(h.albedo, albedo);
  h.fuzz_or_ior = fz;
  return h;
}

// ---- scenes_animated.nim:42-69 ----
struct AnimSphere {
  double velocity, pos_y, coef_restitution, x, z, radius;
  uint32_t mat_kind;
  V albedo;
  double fuzz_or_ior;
};
struct Animation {
  int32_t nrows, ncols;
  float dt, t_min, t_max, t;  // ATime is float32 (scenes_animated.nim:36): the frame count depends on it
  double look_from_angle;
  bool started;
  std::vector<AnimSphere> spheres;
};

inline void anim_step(Animation& an) {  // scenes_animated.nim:156-174: camera first, then physics
  an.look_from_angle -= 2.0 * 3.141592653589793 / 1200.0;
  an.t += an.dt;
  const double G = 9.80665, small_radius = 0.2;
  for (AnimSphere& s : an.spheres) {
    if (s.velocity < 0.0 && s.pos_y < small_radius)
      s.velocity = -s.coef_restitution * s.velocity;
    else
      s.velocity -= G * (double)an.dt;
    s.pos_y += s.velocity * (double)an.dt;
  }
}


// physics/cameras.nim:24-45
inline void camera_make(tor_camera* out, const double look_from[3], const double look_at[3], const double view_up[3],
                        double vfov_degrees, double aspect_ratio, double aperture, double focus_distance,
                        double shutter_open, double shutter_close) {
  V from{look_from[0], look_from[1], look_from[2]}, at{look_at[0], look_at[1], look_at[2]};
  V vup{view_up[0], view_up[1], view_up[2]};
  double theta = vfov_degrees * (3.141592653589793 / 180.0);  // std/math degToRad
  double h = tan(theta / 2.0);
  double viewport_height = 2.0 * h;
  double viewport_width = aspect_ratio * viewport_height;
  V w = unit(sub(from, at));
  V u = unit(cross(vup, w));
  V v = cross(w, u);
  V horizontal = mul(u, focus_distance * viewport_width);
  V vertical = mul(v, focus_distance * viewport_height);
  V llc = sub(sub(sub(from, mul(horizontal, 1.0 / 2)), mul(vertical, 1.0 / 2)), mul(w, focus_distance));
  put(out->origin, from);
  put(out->lower_left_corner, llc);
  put(out->horizontal, horizontal);
  put(out->vertical, vertical);
  put(out->u, u);
  put(out->v, v);
  put(out->w, w);
  out->lens_radius = aperture / 2;
  out->shutter_open = shutter_open;
  out->shutter_close = shutter_close;
}

// scenes_animated.nim:90-154 `random_moving_spheres` with rng.seed(seed) (trace_of_radiance_animation.nim:61-63)
inline Animation* animation_create(uint64_t seed, int32_t height, int32_t width, float dt, float t_min, float t_max) {
  HostRng rng(seed);
  Animation* an = new Animation();
  an->nrows = height;
  an->ncols = width;
  an->dt = dt;
  an->t_min = t_min;
  an->t_max = t_max;
  an->t = 0.0f;
  an->look_from_angle = 2 * 3.141592653589793;
  an->started = false;
  const double small_radius = 0.2;
  for (int a = -20; a < 20; ++a)
    for (int b = -20; b < 20; ++b) {
      double cx = (double)a + 0.9 * rng.u01();
      double cz = (double)b + 0.9 * rng.u01();
      V center{cx, small_radius, cz};
      if (len(sub(center, V{4, small_radius, 0})) > 0.9) {
        double choose = rng.u01();
        AnimSphere s;
        s.x = center.x;
        s.pos_y = center.y;
        s.z = center.z;
        s.radius = small_radius;
        if (choose < 0.65) {
          V a1{0, 0, 0}, a2{0, 0, 0};
          a1.x = rng.u01(); a1.y = rng.u01(); a1.z = rng.u01();
          a2.x = rng.u01(); a2.y = rng.u01(); a2.z = rng.u01();
          s.albedo = V{a1.x * a2.x, a1.y * a2.y, a1.z * a2.z};
          s.coef_restitution = 0.6;
          s.velocity = 10.0 + (4 * rng.u01() - 2.0);  // rng.random(float32) is the single float64 draw
          s.mat_kind = TOR_LAMBERTIAN;
          s.fuzz_or_ior = 0;
        } else if (choose < 0.95) {
          V al{0, 0, 0};
          al.x = rng.urange(0.5, 1); al.y = rng.urange(0.5, 1); al.z = rng.urange(0.5, 1);
          s.albedo = al;
          double fuzz = rng.umax(0.5);
          s.coef_restitution = 0.5;
          s.velocity = 10.0 + (4 * rng.u01() - 2.0);
          s.mat_kind = TOR_METAL;
          s.fuzz_or_ior = fuzz <= 1.0 ? fuzz : 1.0;
        } else {
          s.albedo = V{0, 0, 0};
          s.coef_restitution = 0.5;
          s.velocity = 10.0 + (4 * rng.u01() - 2.0);
          s.mat_kind = TOR_DIELECTRIC;
          s.fuzz_or_ior = 1.5;
        }
        an->spheres.push_back(s);
      }
    }
  return an;
}

// The camera of the frame (scenes_animated.nim:187-202): orbit radius sqrt(200), height 2, looking at (4, 1, 0).
inline void animation_camera(const Animation& an, tor_camera* cam) {
  const double aspect_ratio = (double)an.ncols / (double)an.nrows;
  const double r = sqrt(200.0);
  const double from[3] = {r * cos(an.look_from_angle), 2.0, r * sin(an.look_from_angle)};
  const double at[3] = {4, 1, 0}, up[3] = {0, 1, 0};
  camera_make(cam, from, at, up, 20.0, aspect_ratio, 0.1, 10.0, 0.0, 0.0);
}

// The frame's HittableList (scenes_animated.nim:204-220): ground, the moving spheres at their current height as
// static spheres, the three big spheres.
inline void animation_scene(const Animation& an, std::vector<tor_hittable>* w) {
  w->clear();
  w->push_back(mk_sphere(V{0, -1000, 0}, 1000, TOR_LAMBERTIAN, V{0.5, 0.5, 0.5}, 0));
  for (const AnimSphere& s : an.spheres)
    w->push_back(mk_sphere(V{s.x, s.pos_y, s.z}, s.radius, s.mat_kind, s.albedo, s.fuzz_or_ior));
  w->push_back(mk_sphere(V{0, 1, 0}, 1.0, TOR_DIELECTRIC, V{0, 0, 0}, 1.5));
  w->push_back(mk_sphere(V{-4, 1, 0}, 1.0, TOR_LAMBERTIAN, V{0.4, 0.2, 0.1}, 0));
  w->push_back(mk_sphere(V{4, 1, 0}, 1.0, TOR_METAL, V{0.7, 0.6, 0.5}, 0.0));
}

// Camera and clock part of one step (scenes_animated.nim:156-160); the spheres are stepped by the caller (host loop
// in anim_step, or the device kernel of tor_animation_dev_*).
inline void anim_step_clock(Animation& an) {
  an.look_from_angle -= 2.0 * 3.141592653589793 / 1200.0;
  an.t += an.dt;
}

}  // namespace tor_anim
