// tor_bvh.hpp — host-side build of the bounding-volume hierarchy the render kernel traverses.
//
// The reference scans every object for every bounce segment (physics/hittables/hittables_lists.nim:48-55;
// README.md:36-37 notes the missing acceleration structure).  The closest hit of that scan is
//     argmin over objects of (t_i, i)   with t_i = the object's first root in (t_min, +inf)
// (see tor_kernels.cuh), so ANY superset of the objects that can have a root may be tested, in any order,
// and the result is bit-identical.  The hierarchy only has to be conservative:
//
//   * every object's box covers the sphere at every time a ray can carry (camera shutter interval and
//     the 0.0 that Metal / Dielectric scattering resets to, rays.nim:19), and is padded by 2^-19 of the
//     scene's largest coordinate — this absorbs (a) the float32 rounding of the slab test
//     (<= 2^-20.4 * S in position, derivation in DESIGN.md §bvh) and (b) the reference's own float64
//     rounding (~2^-45 * S);
//   * boxes are stored in float32 rounded OUTWARD;
//   * objects whose box is not finite (NaN / inf centres, zero-length time interval) are not put in the
//     tree at all: they sit in an "always" list that every segment tests exactly.
//
// Blob layout:  [BvhNode x n_nodes][ObjRec x n_objects (tree leaves first, then the "always" list)][cluster boxes]
//               [object boxes]
// (the box tables come last: kernels without cooperative warps stage only the prefix in shared memory, which leaves
// more of the SM's L1 to the traversal stacks in local memory)
//
// The two box tables serve the warp-cooperative search of the render kernel (one warp traces one expensive
// pixel, tor_kernels_bvh.cuh): tree records are in leaf order, i.e. spatially sorted, so 32 consecutive records
// form a "cluster".  The warp tests 32 cluster boxes at once, then the 32 object boxes of every cluster the ray
// enters, and evaluates the surviving objects in parallel lanes.  Both tables hold the same padded, outward
// rounded float32 boxes as the tree (the same conservativeness argument applies), as structure-of-arrays so that
// the 32 lanes read consecutive words: cluster boxes float[6][ncl_pad] (lo.x, lo.y, lo.z, hi.x, hi.y, hi.z),
// object boxes float[6][32] per cluster.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/tor_b200.h"
#include "tor_scene_pack.hpp"

namespace tor {

// Two children per node, each with its own box.  child >= 0: inner node index; child < 0: leaf,
// v = ~child, first record = v >> 4, count = v & 15 (0 = empty child).
// The strides (80 and 144 bytes) are deliberately not powers of two: lanes of a warp read *different* nodes /
// records with 16-byte shared-memory loads, and a stride of 20 (36) words spreads consecutive indices over all
// eight 4-bank groups, where 64 (128) bytes would put every record on the same two (one) groups.
struct alignas(16) BvhNode {  // 80 bytes, 56 used
  float lo0[3], hi0[3];
  float lo1[3], hi1[3];
  int32_t child0, child1;
  int32_t pad[6];
};
static_assert(sizeof(BvhNode) == 80, "BvhNode layout");
static constexpr int kNodeStride16 = sizeof(BvhNode) / 16;

// Everything the kernel needs about one object, in leaf order.  Derived fields are the SAME IEEE operations
// the reference performs on the same inputs (so precomputing them changes no bit):
//   dc = center1 - center0 (moving_spheres.nim:43), r2 = radius*radius (spheres.nim:32),
//   inv_r = 1.0/radius (spheres.nim:43 through vec3s.nim:93-94).
struct alignas(16) ObjRec {  // 144 bytes; the hit test of a static sphere reads only the first 48
  double c0[3];
  double r2;
  uint32_t kind_mat;  // kind | mat_kind << 8
  uint32_t orig;      // index in the caller's HittableList (ties go to the lowest, hittables_lists.nim:48-55)
  double inv_r;
  double dc[3];
  double pad;
  double t0, t1;  // time0, time1 (movers only)
  double albedo[3];
  double fuzz_or_ior;
  double pad2[2];
};
static_assert(sizeof(ObjRec) == 144, "ObjRec layout");
static constexpr int kRecStride16 = sizeof(ObjRec) / 16;
// kind_mat flag: a mover with time0 == +0.0 and time1 == 1.0, whose lerp parameter
// (time - time0) / (time1 - time0) (moving_spheres.nim:41-42) is exactly `time` in IEEE arithmetic
static constexpr uint32_t kObjUnitInterval = 1u << 16;

struct BvhView {  // kernel parameter
  int32_t n_nodes;
  int32_t n_tree_objs;    // records [0, n_tree_objs) are reachable through the tree
  int32_t n_objects;      // records [n_tree_objs, n_objects) are the "always" list
  int32_t has_movers;
  uint32_t off_nodes, off_objs;
  uint32_t nodes_bytes;   // blob prefix holding the nodes
  uint32_t lane_bytes;    // blob prefix holding nodes + records: all that kernels without cooperative warps read
  uint32_t boxes_bytes;   // the two box tables (cluster boxes, then object boxes), contiguous at the end of the blob
  uint32_t total_bytes;
  uint32_t off_cboxes, off_oboxes;  // cluster boxes float[6][ncl_pad]; object boxes float[n_clusters][6][32]
  int32_t n_clusters, ncl_pad;      // clusters of 32 consecutive tree records; ncl_pad = n_clusters rounded up to 32
  float s_limit;          // ray origins with a larger |coordinate| take the "test everything" route
  int32_t max_depth;
};

struct PackedBvh {
  BvhView view;
  std::vector<uint8_t> blob;
  int max_depth = 0;
  int n_leaves = 0;
};

// objects per leaf: at most 15 fit the leaf reference; developer knob TOR_BVH_LEAF
static inline int bvh_max_leaf() {
  static const int v = [] {
    const char* e = getenv("TOR_BVH_LEAF");
    int n = e ? atoi(e) : 4;
    return n < 1 ? 1 : (n > 15 ? 15 : n);
  }();
  return v;
}
#define kBvhMaxLeaf (::tor::bvh_max_leaf())
static constexpr int kBvhStackDepth = 40;  // >= max tree depth (forced median splits below bound it)
static constexpr int kBvhMaxObjects = 1 << 24;  // leaf references hold first_record << 4 in 31 bits; depth <= 16 + 22

namespace bvh_detail {

struct Box {
  double lo[3], hi[3];
};
struct Prim {
  Box b;
  double c[3];
  int obj;
};

inline void box_reset(Box& b) {
  for (int k = 0; k < 3; ++k) {
    b.lo[k] = INFINITY;
    b.hi[k] = -INFINITY;
  }
}
inline void box_grow(Box& b, const Box& o) {
  for (int k = 0; k < 3; ++k) {
    if (o.lo[k] < b.lo[k]) b.lo[k] = o.lo[k];
    if (o.hi[k] > b.hi[k]) b.hi[k] = o.hi[k];
  }
}
inline double box_area(const Box& b) {
  double dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  if (!(dx >= 0 && dy >= 0 && dz >= 0)) return 0.0;
  return 2.0 * (dx * dy + dy * dz + dz * dx);
}
// nextafterf(f, -inf) / nextafterf(f, +inf) for finite f (the boxes are finite by construction), on the bits
inline float f32_pred(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) == 0u) u = 0x80000001u;  // +-0 -> smallest negative subnormal
  else if (u & 0x80000000u) ++u;                   // negative: larger magnitude
  else --u;
  memcpy(&f, &u, 4);
  return f;
}
inline float f32_succ(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) == 0u) u = 0x00000001u;
  else if (u & 0x80000000u) --u;
  else ++u;
  memcpy(&f, &u, 4);
  return f;
}
inline float f32_down(double v) {
  float f = (float)v;
  if (!(f == f) || f - f != 0.0f) {  // NaN / inf: the library call defines the result
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return nextafterf(f, -INFINITY);
  }
  if ((double)f > v) f = f32_pred(f);
  return f32_pred(f);
}
inline float f32_up(double v) {
  float f = (float)v;
  if (!(f == f) || f - f != 0.0f) {
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return nextafterf(f, INFINITY);
  }
  if ((double)f < v) f = f32_succ(f);
  return f32_succ(f);
}

struct Builder {
  std::vector<Prim>& prims;
  std::vector<BvhNode>& nodes;
  std::vector<int>& order;  // object indices in leaf order
  double pad;
  int max_depth = 0;
  int n_leaves = 0;

  void store_box(float* lo, float* hi, const Box& b) {
    for (int k = 0; k < 3; ++k) {
      lo[k] = f32_down(b.lo[k] - pad);
      hi[k] = f32_up(b.hi[k] + pad);
    }
  }
  static void store_empty(float* lo, float* hi) {
    for (int k = 0; k < 3; ++k) {
      lo[k] = 3.0e38f;
      hi[k] = -3.0e38f;
    }
  }
  int32_t make_leaf(int begin, int end) {
    int first = (int)order.size();
    // ascending original index inside a leaf (not required for correctness; keeps runs deterministic)
    std::sort(prims.begin() + begin, prims.begin() + end, [](const Prim& a, const Prim& b) { return a.obj < b.obj; });
    for (int i = begin; i < end; ++i) order.push_back(prims[(size_t)i].obj);
    ++n_leaves;
    return ~(int32_t)((first << 4) | (end - begin));
  }

  // Chooses a split of prims[begin, end) (count > 1); returns mid, or -1 to make a leaf.
  // `all` = the union of the objects' boxes (the caller has it already).
  int split(int begin, int end, int depth, const Box& all) {
    const int n = end - begin;
    Box cb;
    box_reset(cb);
    for (int i = begin; i < end; ++i) {
      const Prim& p = prims[(size_t)i];
      for (int k = 0; k < 3; ++k) {
        if (p.c[k] < cb.lo[k]) cb.lo[k] = p.c[k];
        if (p.c[k] > cb.hi[k]) cb.hi[k] = p.c[k];
      }
    }
    int axis = 0;
    double ext = cb.hi[0] - cb.lo[0];
    for (int k = 1; k < 3; ++k)
      if (cb.hi[k] - cb.lo[k] > ext) {
        ext = cb.hi[k] - cb.lo[k];
        axis = k;
      }
    auto median = [&](int ax) {
      int mid = begin + n / 2;
      std::nth_element(prims.begin() + begin, prims.begin() + mid, prims.begin() + end,
                       [ax](const Prim& a, const Prim& b) { return a.c[ax] < b.c[ax] || (a.c[ax] == b.c[ax] && a.obj < b.obj); });
      return mid;
    };
    if (!(ext > 0.0)) return n <= kBvhMaxLeaf ? -1 : median(axis);  // coincident centroids
    if (depth >= 16) return n <= kBvhMaxLeaf ? -1 : median(axis);   // bounds the depth: <= 16 + log2(2^24/leaf)

    // binned surface-area heuristic on every axis
    constexpr int kBins = 16;
    constexpr int kSmall = 24;  // object count up to which the sorted-object sweep below replaces the bin boxes
    // developer knob TOR_BVH_ISECT: cost of one sphere test relative to one node visit
    static const double c_isect_env = [] {
      const char* e = getenv("TOR_BVH_ISECT");
      return e ? atof(e) : 0.7;
    }();
    const double c_trav = 1.0, c_isect = c_isect_env;
    double best_cost = INFINITY;
    int best_axis = -1, best_bin = -1;
    Box bb3[3][kBins];  // bins of the three axes (nodes with more than kSmall objects)
    int cnt3[3][kBins];
    bool binned = false;
    for (int ax = 0; ax < 3; ++ax) {
      double e = cb.hi[ax] - cb.lo[ax];
      if (!(e > 0.0)) continue;
      if (n <= kSmall) {
        // Few objects (most calls): the same candidates — a split after every occupied bin — evaluated on the
        // objects sorted by bin instead of on 16 mostly empty bin boxes.  Unions and areas are exact min / max
        // arithmetic on the same boxes and the candidates are visited in the same order, so the choice is identical.
        const double scale = kBins / e;
        int bin[kSmall], idx[kSmall];
        for (int i = 0; i < n; ++i) {
          const Prim& p = prims[(size_t)(begin + i)];
          int b = (int)((p.c[ax] - cb.lo[ax]) * scale);
          if (b < 0) b = 0;
          if (b >= kBins) b = kBins - 1;
          int j = i;  // insertion sort by bin
          while (j > 0 && bin[j - 1] > b) {
            bin[j] = bin[j - 1];
            idx[j] = idx[j - 1];
            --j;
          }
          bin[j] = b;
          idx[j] = i;
        }
        double suffix_area[kSmall];
        Box acc;
        box_reset(acc);
        for (int k = n - 1; k > 0; --k) {
          box_grow(acc, prims[(size_t)(begin + idx[k])].b);
          suffix_area[k] = box_area(acc);
        }
        box_reset(acc);
        for (int k = 0; k < n - 1; ++k) {
          box_grow(acc, prims[(size_t)(begin + idx[k])].b);
          if (bin[k + 1] == bin[k]) continue;  // not a bin boundary
          double cost = box_area(acc) * (k + 1) + suffix_area[k + 1] * (n - k - 1);
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = ax;
            best_bin = bin[k];
          }
        }
        continue;
      }
      if (!binned) {  // one pass over the objects fills the bins of all three axes
        binned = true;
        double scale3[3];
        for (int k = 0; k < 3; ++k) {
          const double ek = cb.hi[k] - cb.lo[k];
          scale3[k] = ek > 0.0 ? kBins / ek : 0.0;  // an axis without extent is skipped above; its bins stay unused
          for (int b = 0; b < kBins; ++b) {
            box_reset(bb3[k][b]);
            cnt3[k][b] = 0;
          }
        }
        for (int i = begin; i < end; ++i) {
          const Prim& p = prims[(size_t)i];
          for (int k = 0; k < 3; ++k) {
            if (!(scale3[k] > 0.0)) continue;
            int b = (int)((p.c[k] - cb.lo[k]) * scale3[k]);
            if (b < 0) b = 0;
            if (b >= kBins) b = kBins - 1;
            box_grow(bb3[k][b], p.b);
            ++cnt3[k][b];
          }
        }
      }
      const Box* bb = bb3[ax];
      const int* cnt = cnt3[ax];
      // Empty bins add nothing to a sweep: the right-hand area is carried over, and a left-hand candidate after an
      // empty bin is the partition of the candidate before it (same cost, so never strictly better).
      double right_area[kBins];
      int right_cnt[kBins];
      Box acc;
      box_reset(acc);
      int c = 0;
      double area = 0.0;
      for (int b = kBins - 1; b > 0; --b) {
        if (cnt[b]) {
          box_grow(acc, bb[b]);
          c += cnt[b];
          area = box_area(acc);
        }
        right_area[b] = area;
        right_cnt[b] = c;
      }
      box_reset(acc);
      c = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        if (!cnt[b]) continue;
        box_grow(acc, bb[b]);
        c += cnt[b];
        if (c == 0 || right_cnt[b + 1] == 0) continue;
        double cost = box_area(acc) * c + right_area[b + 1] * right_cnt[b + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = ax;
          best_bin = b;
        }
      }
    }
    const double parent_area = box_area(all);
    if (best_axis < 0) return n <= kBvhMaxLeaf ? -1 : median(axis);
    if (n <= kBvhMaxLeaf && parent_area > 0.0) {
      double split_cost = c_trav + c_isect * best_cost / parent_area;
      if (split_cost >= c_isect * n) return -1;
    }
    const double lo = cb.lo[best_axis], scale = kBins / (cb.hi[best_axis] - cb.lo[best_axis]);
    const int ax = best_axis, bin = best_bin;
    auto it = std::partition(prims.begin() + begin, prims.begin() + end, [=](const Prim& p) {
      int b = (int)((p.c[ax] - lo) * scale);
      if (b < 0) b = 0;
      if (b >= kBins) b = kBins - 1;
      return b <= bin;
    });
    int mid = (int)(it - prims.begin());
    if (mid == begin || mid == end) return median(axis);
    return mid;
  }

  // Builds the subtree over prims[begin, end) and returns its child reference + its box.
  int32_t build(int begin, int end, int depth, Box* out_box) {
    box_reset(*out_box);
    for (int i = begin; i < end; ++i) box_grow(*out_box, prims[(size_t)i].b);
    if (depth > max_depth) max_depth = depth;
    const int n = end - begin;
    int mid = n == 1 ? -1 : split(begin, end, depth, *out_box);
    if (mid < 0) return make_leaf(begin, end);
    int me = (int)nodes.size();
    nodes.emplace_back();
    Box b0, b1;
    int32_t c0 = build(begin, mid, depth + 1, &b0);
    int32_t c1 = build(mid, end, depth + 1, &b1);
    BvhNode& nd = nodes[(size_t)me];
    memset(&nd, 0, sizeof(nd));
    store_box(nd.lo0, nd.hi0, b0);
    store_box(nd.lo1, nd.hi1, b1);
    nd.child0 = c0;
    nd.child1 = c1;
    return me;
  }
};

}  // namespace bvh_detail

// objects: decoded flat records.  shutter_*: the camera's interval.  Returns false with *err set on bad input.
static inline bool pack_bvh(const std::vector<tor_hittable>& objs, const tor_camera& cam, PackedBvh* out,
                            std::string* err) {
  using namespace bvh_detail;
  const int n = (int)objs.size();
  if (n <= 0 || n > kBvhMaxObjects) {
    *err = "object count must be in 1..16777216";
    return false;
  }
  double time_lo = 0.0, time_hi = 0.0;  // rays.nim:19 — scattered Metal / Dielectric rays carry time 0.0
  for (double t : {cam.shutter_open, cam.shutter_close}) {
    if (t < time_lo) time_lo = t;
    if (t > time_hi) time_hi = t;
  }
  const bool time_ok = isfinite(cam.shutter_open) && isfinite(cam.shutter_close);

  std::vector<Prim> prims;
  std::vector<int> always;
  std::vector<Box> obj_box((size_t)n);  // by original index (tree objects only)
  double S = 0.0;  // largest |coordinate| of any box or of the camera (ray origins live on the objects or the lens)
  for (int k = 0; k < 3; ++k) {
    double lens = fabs(cam.lens_radius) * (fabs(cam.u[k]) + fabs(cam.v[k]));
    double v = fabs(cam.origin[k]) + lens;
    if (isfinite(v) && v > S) S = v;
  }
  bool has_movers = false;
  for (int i = 0; i < n; ++i) {
    const tor_hittable& h = objs[(size_t)i];
    Prim p;
    p.obj = i;
    box_reset(p.b);
    const double r = fabs(h.radius);
    auto grow_at = [&](double q) {
      for (int k = 0; k < 3; ++k) {
        double c = h.kind == TOR_MOVING_SPHERE ? h.center0[k] + q * (h.center1[k] - h.center0[k]) : h.center0[k];
        // 1e-9 relative head-room: the reference rounds c0 + q*dc (and q itself) in float64
        double e = r + 1e-9 * (fabs(c) + r);
        if (!(c - e >= p.b.lo[k])) p.b.lo[k] = c - e;  // written so that NaN poisons the box
        if (!(c + e <= p.b.hi[k])) p.b.hi[k] = c + e;
      }
    };
    if (h.kind == TOR_MOVING_SPHERE) {
      has_movers = true;
      double q0 = (time_lo - h.time0) / (h.time1 - h.time0);
      double q1 = (time_hi - h.time0) / (h.time1 - h.time0);
      grow_at(q0);
      grow_at(q1);
      if (!time_ok) p.b.lo[0] = NAN;
    } else {
      grow_at(0.0);
    }
    bool finite = true;
    for (int k = 0; k < 3; ++k) {
      finite = finite && isfinite(p.b.lo[k]) && isfinite(p.b.hi[k]) && p.b.lo[k] <= p.b.hi[k];
      finite = finite && fabs(p.b.lo[k]) < 1e30 && fabs(p.b.hi[k]) < 1e30;  // stays far from float32 overflow
    }
    if (!finite) {
      always.push_back(i);
      continue;
    }
    for (int k = 0; k < 3; ++k) {
      p.c[k] = 0.5 * (p.b.lo[k] + p.b.hi[k]);
      S = std::max(S, std::max(fabs(p.b.lo[k]), fabs(p.b.hi[k])));
    }
    obj_box[(size_t)i] = p.b;
    prims.push_back(p);
  }

  // Objects whose box spans most of the scene (the r = 1000 ground sphere of scenes.nim:15) prune nothing from
  // inside the tree.  They join the "always" list instead: tested first, on every segment, by all lanes of the
  // warp together, and their root bounds the traversal before it starts.
  if (prims.size() > 4) {
    Box all;
    box_reset(all);
    for (const Prim& p : prims) box_grow(all, p.b);
    const double limit = 0.5 * box_area(all);
    std::vector<Prim> kept;
    for (const Prim& p : prims) {
      if (always.size() < 8 && box_area(p.b) > limit)
        always.push_back(p.obj);
      else
        kept.push_back(p);
    }
    prims.swap(kept);
  }

  std::vector<BvhNode> nodes;
  std::vector<int> order;
  Builder B{prims, nodes, order, /*pad=*/ldexp(S, -19)};
  if (prims.empty()) {
    BvhNode nd;
    memset(&nd, 0, sizeof(nd));
    Builder::store_empty(nd.lo0, nd.hi0);
    Builder::store_empty(nd.lo1, nd.hi1);
    nd.child0 = ~0;
    nd.child1 = ~0;
    nodes.push_back(nd);
  } else {
    Box rb;
    int32_t root = B.build(0, (int)prims.size(), 0, &rb);
    if (root < 0) {  // everything fits one leaf: wrap it in a root node
      BvhNode nd;
      memset(&nd, 0, sizeof(nd));
      B.store_box(nd.lo0, nd.hi0, rb);
      Builder::store_empty(nd.lo1, nd.hi1);
      nd.child0 = root;
      nd.child1 = ~0;
      nodes.push_back(nd);
    }
  }
  if (B.max_depth + 1 > kBvhStackDepth) {
    *err = "internal: BVH deeper than the traversal stack";
    return false;
  }
  const int n_tree = (int)order.size();
  for (int i : always) order.push_back(i);

  std::vector<ObjRec> recs((size_t)n);
  for (int j = 0; j < n; ++j) {
    const tor_hittable& h = objs[(size_t)order[(size_t)j]];
    ObjRec& r = recs[(size_t)j];
    memset(&r, 0, sizeof(r));
    const bool mover = h.kind == TOR_MOVING_SPHERE;
    for (int k = 0; k < 3; ++k) {
      r.c0[k] = h.center0[k];
      r.dc[k] = mover ? h.center1[k] - h.center0[k] : 0.0;  // moving_spheres.nim:43
      r.albedo[k] = h.albedo[k];
    }
    r.r2 = h.radius * h.radius;  // spheres.nim:32
    r.inv_r = 1.0 / h.radius;    // vec3s.nim:93-94
    r.t0 = h.time0;
    r.t1 = h.time1;
    r.fuzz_or_ior = h.fuzz_or_ior;
    r.kind_mat = h.kind | (h.mat_kind << 8);
    if (mover && h.time0 == 0.0 && !signbit(h.time0) && h.time1 == 1.0) r.kind_mat |= kObjUnitInterval;
    r.orig = (uint32_t)order[(size_t)j];
  }

  // box tables of the warp-cooperative search: clusters of 32 consecutive tree records
  const int n_clusters = (n_tree + 31) / 32;
  const int ncl_pad = std::max(32, (n_clusters + 31) / 32 * 32);
  std::vector<float> cboxes((size_t)6 * ncl_pad), oboxes((size_t)n_clusters * 6 * 32);
  for (int c = 0; c < ncl_pad; ++c) {
    float lo[3], hi[3];
    Builder::store_empty(lo, hi);
    if (c < n_clusters) {
      Box cb;
      box_reset(cb);
      for (int j = 32 * c; j < std::min(n_tree, 32 * c + 32); ++j) box_grow(cb, obj_box[(size_t)order[(size_t)j]]);
      B.store_box(lo, hi, cb);
      for (int i = 0; i < 32; ++i) {
        float olo[3], ohi[3];
        Builder::store_empty(olo, ohi);
        const int j = 32 * c + i;
        if (j < n_tree) B.store_box(olo, ohi, obj_box[(size_t)order[(size_t)j]]);
        for (int k = 0; k < 3; ++k) {
          oboxes[((size_t)c * 6 + k) * 32 + i] = olo[k];
          oboxes[((size_t)c * 6 + 3 + k) * 32 + i] = ohi[k];
        }
      }
    }
    for (int k = 0; k < 3; ++k) {
      cboxes[(size_t)k * ncl_pad + c] = lo[k];
      cboxes[(size_t)(3 + k) * ncl_pad + c] = hi[k];
    }
  }

  BvhView& v = out->view;
  v.n_nodes = (int32_t)nodes.size();
  v.n_tree_objs = n_tree;
  v.n_objects = n;
  v.has_movers = has_movers ? 1 : 0;
  v.off_nodes = 0;
  v.nodes_bytes = (uint32_t)(nodes.size() * sizeof(BvhNode));
  v.off_objs = v.nodes_bytes;
  v.lane_bytes = v.off_objs + (uint32_t)(recs.size() * sizeof(ObjRec));
  v.off_cboxes = v.lane_bytes;
  v.off_oboxes = v.off_cboxes + (uint32_t)(cboxes.size() * sizeof(float));
  v.boxes_bytes = (uint32_t)((cboxes.size() + oboxes.size()) * sizeof(float));
  v.total_bytes = v.off_cboxes + v.boxes_bytes;
  v.n_clusters = n_clusters;
  v.ncl_pad = ncl_pad;
  v.s_limit = bvh_detail::f32_down(1.001 * S);  // the padding was sized for origins within S (see header)
  v.max_depth = B.max_depth;
  out->blob.resize(v.total_bytes);  // nodes + box tables + records cover every byte
  memcpy(out->blob.data() + v.off_nodes, nodes.data(), nodes.size() * sizeof(BvhNode));
  memcpy(out->blob.data() + v.off_cboxes, cboxes.data(), cboxes.size() * sizeof(float));
  if (!oboxes.empty()) memcpy(out->blob.data() + v.off_oboxes, oboxes.data(), oboxes.size() * sizeof(float));
  memcpy(out->blob.data() + v.off_objs, recs.data(), recs.size() * sizeof(ObjRec));
  out->max_depth = B.max_depth;
  out->n_leaves = B.n_leaves;
  return true;
}

}  // namespace tor
