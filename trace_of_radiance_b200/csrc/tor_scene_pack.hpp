// tor_scene_pack.hpp — host-side re-layout of the HittableList for the sm_100a kernels.
//
// The reference scans an array of 120-byte tagged unions (physics/hittables/hittables_lists.nim:
// 48-55) and evaluates the full quadratic for every object.  The kernels split that into
//   (1) a conservative float64 FMA *filter* over compact structure-of-arrays records staged in
//       shared memory — it can only say "certainly no root" or "maybe", and
//   (2) the reference's exact, non-fused arithmetic on the few "maybe" objects,
// so the hit that comes out is bit-identical to the reference's (DESIGN.md §filter has the error
// bound).  This file builds the "scene blob" both stages read:
//
//   [static filter records   : n_static_pad  x 4 f64 {cx, cy, cz, nW}]
//   [y-mover filter records  : n_ymov_pad    x 6 f64 {cx, cz, nWxz, c0y, dcy, 0}]
//   [general mover records   : n_gmov_pad    x 8 f64 {c0x, c0y, c0z, r2m, dcx, dcy, dcz, 0}]
//   [time classes            : n_classes     x 32 B  {t0, t1, y_begin, y_end, g_begin, g_end}]
//   [index maps              : (n_static_pad + n_ymov_pad + n_gmov_pad) x u16 -> original index]
//   [exact records           : n_objects     x 112 B tor_hittable]
//
// Lists are padded to a multiple of 32 with never-pass sentinels so the unrolled filter loop has
// no bounds checks.  Movers are grouped by (time0, time1) "time class": the lerp parameter of
// moving_spheres.nim:39-44 depends only on (ray.time, time0, time1), so it is computed once per
// class per segment instead of once per test.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tor_b200.h"

namespace tor {

// Relative slack of the conservative filter: 2^-40 is ~8000 ulp of float64; the filter's own
// rounding error is < 64 ulp of (|o|^2 + |c|^2 + r^2).
static constexpr double kFilterSlack = 0x1p-40;

struct TimeClass {  // 32 bytes
  double t0, t1;
  int32_t y_begin, y_end;  // record ranges (multiples of 32) inside the y-mover / general lists
  int32_t g_begin, g_end;
};
static_assert(sizeof(TimeClass) == 32, "TimeClass layout");

struct SceneView {  // kernel parameter: where things are inside the blob (byte offsets)
  int32_t n_objects;
  int32_t n_static_pad;
  int32_t n_ymov_pad;
  int32_t n_gmov_pad;
  int32_t n_classes;
  uint32_t off_static, off_ymov, off_gmov, off_classes;
  uint32_t off_idx_static, off_idx_ymov, off_idx_gmov;
  uint32_t off_exact;
  uint32_t hot_bytes;    // blob prefix every CTA stages into shared memory (filter + classes + idx)
  uint32_t total_bytes;  // hot + exact records
};

struct PackedScene {
  SceneView view;
  std::vector<uint8_t> blob;
  int n_static = 0, n_ymov = 0, n_gmov = 0;
};

static inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// Decode one object of either encoding into the flat record.  Returns false on a bad tag.
static inline bool decode_object(const uint8_t* p, int64_t stride, tor_hittable* out) {
  if (stride == TOR_STRIDE_FLAT) {
    memcpy(out, p, sizeof(tor_hittable));
  } else {  // TOR_STRIDE_NIM_VARIANT: see include/tor_b200.h
    memset(out, 0, sizeof(*out));
    uint8_t kind = p[0];
    const uint8_t* u = p + 8;
    const uint8_t* mat;
    if (kind == TOR_SPHERE) {
      memcpy(out->center0, u, 24);
      memcpy(&out->radius, u + 24, 8);
      mat = u + 32;
    } else if (kind == TOR_MOVING_SPHERE) {
      memcpy(out->center0, u, 24);
      memcpy(out->center1, u + 24, 24);
      memcpy(&out->time0, u + 48, 8);
      memcpy(&out->time1, u + 56, 8);
      memcpy(&out->radius, u + 64, 8);
      mat = u + 72;
    } else {
      return false;
    }
    out->kind = kind;
    uint8_t mk = mat[0];
    out->mat_kind = mk;
    if (mk == TOR_LAMBERTIAN) {
      memcpy(out->albedo, mat + 8, 24);
    } else if (mk == TOR_METAL) {
      memcpy(out->albedo, mat + 8, 24);
      memcpy(&out->fuzz_or_ior, mat + 32, 8);
    } else if (mk == TOR_DIELECTRIC) {
      memcpy(&out->fuzz_or_ior, mat + 8, 8);
    }
  }
  return out->kind <= TOR_MOVING_SPHERE && out->mat_kind <= TOR_DIELECTRIC;
}

// |c0 + q (c1 - c0)|^2
static inline double center_norm2_at(const tor_hittable& h, double q) {
  double s = 0;
  for (int k = 0; k < 3; ++k) {
    double c = h.center0[k] + q * (h.center1[k] - h.center0[k]);
    s += c * c;
  }
  return s;
}

// Upper bound of |center(time)|^2 over every ray time the render can produce: the camera draws
// time in [shutter_open, shutter_close] (cameras.nim:56) and Metal / Dielectric scattering resets it
// to 0.0 (rays.nim:19).  |c(q)|^2 is convex in q, so the maximum sits at an end of the q range.
static inline double mover_cc_bound(const tor_hittable& h, double time_lo, double time_hi) {
  double q_lo = (time_lo - h.time0) / (h.time1 - h.time0);
  double q_hi = (time_hi - h.time0) / (h.time1 - h.time0);
  double m = center_norm2_at(h, q_lo);
  double m2 = center_norm2_at(h, q_hi);
  if (m2 > m) m = m2;
  m2 = center_norm2_at(h, 0.0);
  if (m2 > m) m = m2;
  m2 = center_norm2_at(h, 1.0);
  if (m2 > m) m = m2;
  return 2.0 * m + 1.0;  // NaN / inf propagate: the record then always passes the filter
}

static inline bool pack_scene(const void* objects, int64_t len, int64_t stride, double shutter_open,
                              double shutter_close, PackedScene* ps, std::string* err) {
  if (!objects || len <= 0) {
    *err = "empty HittableList (hittables_lists.nim:42 asserts len > 0)";
    return false;
  }
  if (stride != TOR_STRIDE_FLAT && stride != TOR_STRIDE_NIM_VARIANT) {
    *err = "stride must be 112 (tor_hittable) or 120 (Nim HittableVariant)";
    return false;
  }
  if (len > 65535) {
    *err = "more than 65535 objects";
    return false;
  }
  std::vector<tor_hittable> objs((size_t)len);
  for (int64_t i = 0; i < len; ++i)
    if (!decode_object((const uint8_t*)objects + i * stride, stride, &objs[(size_t)i])) {
      *err = "object " + std::to_string(i) + ": unknown kind / material tag";
      return false;
    }

  const double K = kFilterSlack;
  double time_lo = 0.0, time_hi = 0.0;
  if (shutter_open < time_lo) time_lo = shutter_open;
  if (shutter_close < time_lo) time_lo = shutter_close;
  if (shutter_open > time_hi) time_hi = shutter_open;
  if (shutter_close > time_hi) time_hi = shutter_close;
  struct Rec4 { double v[4]; };
  struct Rec6 { double v[6]; };
  struct Rec8 { double v[8]; };
  std::vector<Rec4> st;
  std::vector<uint16_t> st_idx;
  // movers grouped by time class, keeping array order inside a class.  The key is the bit pattern of (time0, time1):
  // a NaN time would break the strict weak ordering a map of doubles needs.
  std::map<std::pair<uint64_t, uint64_t>, int> class_of;
  std::vector<std::pair<double, double>> class_key;
  std::vector<std::vector<int>> class_y, class_g;

  for (int64_t i = 0; i < len; ++i) {
    const tor_hittable& h = objs[(size_t)i];
    const double r2 = h.radius * h.radius;
    if (h.kind == TOR_SPHERE) {
      const double* c = h.center0;
      double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
      double W = (cc - r2) - K * (cc + r2);
      st.push_back(Rec4{{c[0], c[1], c[2], -W}});
      st_idx.push_back((uint16_t)i);
    } else {
      uint64_t b0, b1;
      memcpy(&b0, &h.time0, 8);
      memcpy(&b1, &h.time1, 8);
      auto key = std::make_pair(b0, b1);
      auto it = class_of.find(key);
      int ci;
      if (it == class_of.end()) {
        ci = (int)class_key.size();
        class_of[key] = ci;
        class_key.push_back(std::make_pair(h.time0, h.time1));
        class_y.emplace_back();
        class_g.emplace_back();
      } else {
        ci = it->second;
      }
      bool y_only = (h.center1[0] == h.center0[0]) && (h.center1[2] == h.center0[2]);
      (y_only ? class_y : class_g)[(size_t)ci].push_back((int)i);
    }
  }

  std::vector<Rec6> ym;
  std::vector<uint16_t> ym_idx;
  std::vector<Rec8> gm;
  std::vector<uint16_t> gm_idx;
  std::vector<TimeClass> classes;
  const double NEG_INF = -INFINITY;
  auto pad32 = [](size_t n) { return (n + 31) / 32 * 32; };
  for (size_t ci = 0; ci < class_key.size(); ++ci) {
    TimeClass tc;
    tc.t0 = class_key[ci].first;
    tc.t1 = class_key[ci].second;
    tc.y_begin = (int32_t)ym.size();
    for (int i : class_y[ci]) {
      const tor_hittable& h = objs[(size_t)i];
      const double r2 = h.radius * h.radius;
      const double* c0 = h.center0;
      const double* c1 = h.center1;
      // the slack must cover |c(time)|^2 for every time a ray can carry
      double ccmax = mover_cc_bound(h, time_lo, time_hi);
      double r2m = r2 + K * (ccmax + r2);
      if (!(r2m < INFINITY)) r2m = INFINITY;  // also catches NaN
      double nWxz = r2m - (c0[0] * c0[0] + c0[2] * c0[2]);
      ym.push_back(Rec6{{c0[0], c0[2], nWxz, c0[1], c1[1] - c0[1], 0.0}});
      ym_idx.push_back((uint16_t)i);
    }
    while (ym.size() != pad32(ym.size())) {
      ym.push_back(Rec6{{0, 0, NEG_INF, 0, 0, 0}});
      ym_idx.push_back(0);
    }
    tc.y_end = (int32_t)ym.size();
    tc.g_begin = (int32_t)gm.size();
    for (int i : class_g[ci]) {
      const tor_hittable& h = objs[(size_t)i];
      const double r2 = h.radius * h.radius;
      const double* c0 = h.center0;
      const double* c1 = h.center1;
      double ccmax = mover_cc_bound(h, time_lo, time_hi);
      double r2m = r2 + K * (ccmax + r2);
      if (!(r2m < INFINITY)) r2m = INFINITY;
      gm.push_back(Rec8{{c0[0], c0[1], c0[2], r2m, c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2], 0.0}});
      gm_idx.push_back((uint16_t)i);
    }
    while (gm.size() != pad32(gm.size())) {
      gm.push_back(Rec8{{0, 0, 0, NEG_INF, 0, 0, 0, 0}});
      gm_idx.push_back(0);
    }
    tc.g_end = (int32_t)gm.size();
    classes.push_back(tc);
  }
  ps->n_static = (int)st.size();
  ps->n_ymov = 0;
  ps->n_gmov = 0;
  for (size_t ci = 0; ci < class_key.size(); ++ci) {
    ps->n_ymov += (int)class_y[ci].size();
    ps->n_gmov += (int)class_g[ci].size();
  }
  while (st.size() != pad32(st.size())) {
    st.push_back(Rec4{{0, 0, 0, NEG_INF}});
    st_idx.push_back(0);
  }

  SceneView& v = ps->view;
  v.n_objects = (int32_t)len;
  v.n_static_pad = (int32_t)st.size();
  v.n_ymov_pad = (int32_t)ym.size();
  v.n_gmov_pad = (int32_t)gm.size();
  v.n_classes = (int32_t)classes.size();
  uint32_t off = 0;
  v.off_static = off;
  off += (uint32_t)(st.size() * sizeof(Rec4));
  v.off_ymov = off;
  off += (uint32_t)(ym.size() * sizeof(Rec6));
  v.off_gmov = off;
  off += (uint32_t)(gm.size() * sizeof(Rec8));
  v.off_classes = off;
  off += (uint32_t)(classes.size() * sizeof(TimeClass));
  v.off_idx_static = off;
  off += (uint32_t)(st_idx.size() * 2);
  v.off_idx_ymov = off;
  off += (uint32_t)(ym_idx.size() * 2);
  v.off_idx_gmov = off;
  off += (uint32_t)(gm_idx.size() * 2);
  off = align_up(off, 16);
  v.hot_bytes = off;
  v.off_exact = off;
  off += (uint32_t)(objs.size() * sizeof(tor_hittable));
  off = align_up(off, 16);
  v.total_bytes = off;

  ps->blob.assign(off, 0);
  uint8_t* b = ps->blob.data();
  if (!st.empty()) memcpy(b + v.off_static, st.data(), st.size() * sizeof(Rec4));
  if (!ym.empty()) memcpy(b + v.off_ymov, ym.data(), ym.size() * sizeof(Rec6));
  if (!gm.empty()) memcpy(b + v.off_gmov, gm.data(), gm.size() * sizeof(Rec8));
  if (!classes.empty()) memcpy(b + v.off_classes, classes.data(), classes.size() * sizeof(TimeClass));
  if (!st_idx.empty()) memcpy(b + v.off_idx_static, st_idx.data(), st_idx.size() * 2);
  if (!ym_idx.empty()) memcpy(b + v.off_idx_ymov, ym_idx.data(), ym_idx.size() * 2);
  if (!gm_idx.empty()) memcpy(b + v.off_idx_gmov, gm_idx.data(), gm_idx.size() * 2);
  memcpy(b + v.off_exact, objs.data(), objs.size() * sizeof(tor_hittable));
  return true;
}

}  // namespace tor
