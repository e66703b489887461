// bvh_check.cc — CPU-side check of the host BVH builder (trace_of_radiance_b200/csrc/tor_bvh.hpp).
// Test infrastructure: compiled by tests/test_bvh_host.py with g++.  It restates the kernel's float32 slab
// traversal (tor_kernels_bvh.cuh) on the host and compares the set of objects it reaches with the objects
// that have a root under the reference's float64 test (spheres.nim:28-49, moving_spheres.nim:39-67).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../trace_of_radiance_b200/csrc/tor_bvh.hpp"

using namespace tor;

namespace {

struct Hit {
  double t;
  uint32_t orig;
};

// the reference's test on one record; returns +inf when there is no root above t_min
double first_root(const ObjRec& r, const double o[3], const double d[3], double time, double a) {
  double c[3];
  const bool mover = (r.kind_mat & 0xffu) == TOR_MOVING_SPHERE;
  for (int k = 0; k < 3; ++k) {
    c[k] = r.c0[k];
    if (mover) {
      double q = (time - r.t0) / (r.t1 - r.t0);
      c[k] = r.c0[k] + (q * r.dc[k]);
    }
  }
  double oc[3] = {o[0] - c[0], o[1] - c[1], o[2] - c[2]};
  double half_b = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
  double cc = (oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2]) - r.r2;
  double disc = half_b * half_b - a * cc;
  if (disc > 0) {
    double root = sqrt(disc);
    double sol = (-half_b - root) / a;
    if (0.001 < sol) return sol;
    sol = (-half_b + root) / a;
    if (0.001 < sol) return sol;
  }
  return INFINITY;
}

void better(Hit& best, double t, uint32_t orig) {
  if (t < best.t || (t == best.t && orig < best.orig && t < INFINITY)) {
    best.t = t;
    best.orig = orig;
  }
}

}  // namespace

extern "C" {

// Structure checks.  Returns 0 when every object appears exactly once and every tree object's swept sphere lies
// inside its leaf box and every ancestor's box; otherwise a negative code.
int bvh_check_structure(const tor_hittable* objs, int n, const tor_camera* cam, int64_t* info) {
  std::vector<tor_hittable> v(objs, objs + n);
  PackedBvh pb;
  std::string err;
  if (!pack_bvh(v, *cam, &pb, &err)) return -1;
  const BvhNode* nodes = (const BvhNode*)(pb.blob.data() + pb.view.off_nodes);
  const ObjRec* recs = (const ObjRec*)(pb.blob.data() + pb.view.off_objs);
  std::vector<int> seen((size_t)n, 0);
  for (int i = 0; i < pb.view.n_objects; ++i) {
    if (recs[i].orig >= (uint32_t)n) return -2;
    seen[recs[i].orig]++;
  }
  for (int i = 0; i < n; ++i)
    if (seen[(size_t)i] != 1) return -3;
  double t_lo = 0, t_hi = 0;
  for (double t : {cam->shutter_open, cam->shutter_close}) {
    if (t < t_lo) t_lo = t;
    if (t > t_hi) t_hi = t;
  }
  // walk the tree with an explicit stack carrying the chain of enclosing boxes
  struct Item {
    int32_t ref;
    std::vector<const float*> boxes;  // pairs lo, hi
  };
  std::vector<Item> st;
  st.push_back({0, {}});
  std::vector<int> reached((size_t)pb.view.n_objects, 0);
  int leaves = 0;
  while (!st.empty()) {
    Item it = st.back();
    st.pop_back();
    if (it.ref >= 0) {
      const BvhNode& nd = nodes[it.ref];
      Item a = it, b = it;
      a.ref = nd.child0;
      a.boxes.push_back(nd.lo0);
      a.boxes.push_back(nd.hi0);
      b.ref = nd.child1;
      b.boxes.push_back(nd.lo1);
      b.boxes.push_back(nd.hi1);
      st.push_back(a);
      st.push_back(b);
    } else {
      int32_t lv = ~it.ref, first = lv >> 4, cnt = lv & 15;
      if (cnt) ++leaves;
      for (int k = 0; k < cnt; ++k) {
        const ObjRec& r = recs[first + k];
        reached[(size_t)(first + k)]++;
        const bool mover = (r.kind_mat & 0xffu) == TOR_MOVING_SPHERE;
        for (double time : {t_lo, t_hi, 0.5 * (t_lo + t_hi)}) {
          for (int ax = 0; ax < 3; ++ax) {
            double c = r.c0[ax];
            if (mover) c = r.c0[ax] + ((time - r.t0) / (r.t1 - r.t0)) * r.dc[ax];
            double rad = sqrt(r.r2);
            for (size_t bi = 0; bi < it.boxes.size(); bi += 2) {
              if (!((double)it.boxes[bi][ax] < c - rad && (double)it.boxes[bi + 1][ax] > c + rad)) return -4;
            }
          }
        }
      }
    }
  }
  for (int i = 0; i < pb.view.n_tree_objs; ++i)
    if (reached[(size_t)i] != 1) return -5;
  if (info) {
    info[0] = pb.view.n_nodes;
    info[1] = leaves;
    info[2] = pb.max_depth;
    info[3] = pb.view.n_objects - pb.view.n_tree_objs;
  }
  return 0;
}

// Traces `nrays` rays (o, d, time: 7 doubles each) both ways; returns the number of rays whose closest hit
// (t, original index) differs between the float32-filtered traversal and the full scan (must be 0).
// stats[0] += node visits, stats[1] += sphere tests of the traversal.
int64_t bvh_check_rays(const tor_hittable* objs, int n, const tor_camera* cam, const double* rays, int64_t nrays,
                       int64_t* stats) {
  std::vector<tor_hittable> v(objs, objs + n);
  PackedBvh pb;
  std::string err;
  if (!pack_bvh(v, *cam, &pb, &err)) return -1;
  const BvhView& bv = pb.view;
  const BvhNode* nodes = (const BvhNode*)(pb.blob.data() + bv.off_nodes);
  const ObjRec* recs = (const ObjRec*)(pb.blob.data() + bv.off_objs);
  int64_t bad = 0;
  for (int64_t ri = 0; ri < nrays; ++ri) {
    const double* o = rays + 7 * ri;
    const double* d = o + 3;
    const double time = o[6];
    const double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    Hit scan{INFINITY, 0xffffffffu};
    for (int i = 0; i < bv.n_objects; ++i) better(scan, first_root(recs[i], o, d, time, a), recs[i].orig);

    // ---- the kernel's traversal, restated (setup + while-while)
    Hit best{INFINITY, 0xffffffffu};
    float best_f = INFINITY;
    float df[3], of[3], id[3], oi[3];
    for (int k = 0; k < 3; ++k) {
      df[k] = (float)d[k];
      of[k] = (float)o[k];
    }
    float dmax = fmaxf(fabsf(df[0]), fmaxf(fabsf(df[1]), fabsf(df[2])));
    float omax = fmaxf(fabsf(of[0]), fmaxf(fabsf(of[1]), fabsf(of[2])));
    bool ok = dmax >= 0x1p-40f && dmax <= 0x1p40f && omax <= bv.s_limit;
    for (int k = 0; k < 3; ++k) ok = ok && df[k] == df[k] && of[k] == of[k];
    for (int k = 0; k < 3; ++k) {
      if (ok) {
        float dmin = dmax * 0x1p-60f;
        if (fabsf(df[k]) < dmin) df[k] = copysignf(dmin, df[k]);
        id[k] = 1.0f / df[k];
        oi[k] = of[k] * id[k];
      } else {
        id[k] = oi[k] = 0.f;
      }
    }
    for (int i = bv.n_tree_objs; i < bv.n_objects; ++i) {
      double t = first_root(recs[i], o, d, time, a);
      Hit before = best;
      better(best, t, recs[i].orig);
      if (best.t != before.t) best_f = nextafterf((float)best.t, INFINITY);  // >= fl32_up(t)
      if (stats) stats[1]++;
    }
    int32_t stk[kBvhStackDepth];
    float stk_t[kBvhStackDepth];
    int sp = 0;
    int32_t cur = 0;
    bool done = false;
    while (!done) {
      while (cur >= 0) {
        const BvhNode& nd = nodes[cur];
        if (stats) stats[0]++;
        float nr[2], fr[2];
        const float* lo[2] = {nd.lo0, nd.lo1};
        const float* hi[2] = {nd.hi0, nd.hi1};
        for (int c = 0; c < 2; ++c) {
          float t0[3], t1[3];
          for (int k = 0; k < 3; ++k) {
            t0[k] = fmaf(lo[c][k], id[k], -oi[k]);
            t1[k] = fmaf(hi[c][k], id[k], -oi[k]);
          }
          nr[c] = fmaxf(fmaxf(fminf(t0[0], t1[0]), fminf(t0[1], t1[1])), fmaxf(fminf(t0[2], t1[2]), 0.f));
          fr[c] = fminf(fminf(fmaxf(t0[0], t1[0]), fmaxf(t0[1], t1[1])), fminf(fmaxf(t0[2], t1[2]), best_f));
        }
        bool h0 = nr[0] <= fr[0], h1 = nr[1] <= fr[1];
        if (h0 && h1) {
          bool first0 = nr[0] <= nr[1];
          stk[sp] = first0 ? nd.child1 : nd.child0;
          stk_t[sp] = first0 ? nr[1] : nr[0];
          ++sp;
          cur = first0 ? nd.child0 : nd.child1;
        } else if (h0 || h1) {
          cur = h0 ? nd.child0 : nd.child1;
        } else {
          done = true;
          while (sp > 0) {
            --sp;
            if (stk_t[sp] <= best_f) {
              cur = stk[sp];
              done = false;
              break;
            }
          }
          if (done) break;
        }
      }
      if (done) break;
      int32_t lv = ~cur, first = lv >> 4, cnt = lv & 15;
      for (int k = 0; k < cnt; ++k) {
        double t = first_root(recs[first + k], o, d, time, a);
        Hit before = best;
        better(best, t, recs[first + k].orig);
        if (best.t != before.t) best_f = nextafterf((float)best.t, INFINITY);
        if (stats) stats[1]++;
      }
      done = true;
      while (sp > 0) {
        --sp;
        if (stk_t[sp] <= best_f) {
          cur = stk[sp];
          done = false;
          break;
        }
      }
    }
    if (!(best.t == scan.t && (best.orig == scan.orig || scan.t == INFINITY))) ++bad;
  }
  return bad;
}

// The warp-cooperative search of the render kernel (tor_kernels_bvh.cuh, COOP kernels), restated: 32 cluster boxes per
// step, then the 32 object boxes of every cluster the ray enters, all with best_f = +inf (nothing is pruned by a
// closer hit), then the reference's test on the survivors and the "always" list.  Returns the number of rays whose
// closest hit differs from the full scan (must be 0).  stats[0] += box rounds, stats[1] += sphere tests.
int64_t bvh_check_rays_coop(const tor_hittable* objs, int n, const tor_camera* cam, const double* rays, int64_t nrays,
                            int64_t* stats) {
  std::vector<tor_hittable> v(objs, objs + n);
  PackedBvh pb;
  std::string err;
  if (!pack_bvh(v, *cam, &pb, &err)) return -1;
  const BvhView& bv = pb.view;
  const ObjRec* recs = (const ObjRec*)(pb.blob.data() + bv.off_objs);
  const float* cboxes = (const float*)(pb.blob.data() + bv.off_cboxes);
  const float* oboxes = (const float*)(pb.blob.data() + bv.off_oboxes);
  if (bv.n_clusters != (bv.n_tree_objs + 31) / 32 || bv.ncl_pad % 32 || bv.ncl_pad < bv.n_clusters) return -2;
  if (bv.off_oboxes != bv.off_cboxes + 24u * (uint32_t)bv.ncl_pad) return -3;
  if (bv.boxes_bytes != 24u * (uint32_t)bv.ncl_pad + 768u * (uint32_t)bv.n_clusters) return -4;
  if (bv.off_cboxes != bv.lane_bytes || bv.total_bytes != bv.lane_bytes + bv.boxes_bytes || bv.off_objs != bv.nodes_bytes)
    return -4;
  int64_t bad = 0;
  for (int64_t ri = 0; ri < nrays; ++ri) {
    const double* o = rays + 7 * ri;
    const double* d = o + 3;
    const double time = o[6];
    const double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    Hit scan{INFINITY, 0xffffffffu};
    for (int i = 0; i < bv.n_objects; ++i) better(scan, first_root(recs[i], o, d, time, a), recs[i].orig);

    float df[3], of[3], id[3], oi[3];
    for (int k = 0; k < 3; ++k) {
      df[k] = (float)d[k];
      of[k] = (float)o[k];
    }
    float dmax = fmaxf(fabsf(df[0]), fmaxf(fabsf(df[1]), fabsf(df[2])));
    float omax = fmaxf(fabsf(of[0]), fmaxf(fabsf(of[1]), fabsf(of[2])));
    bool ok = dmax >= 0x1p-40f && dmax <= 0x1p40f && omax <= bv.s_limit;
    for (int k = 0; k < 3; ++k) ok = ok && df[k] == df[k] && of[k] == of[k];
    for (int k = 0; k < 3; ++k) {
      if (ok) {
        float dmin = dmax * 0x1p-60f;
        if (fabsf(df[k]) < dmin) df[k] = copysignf(dmin, df[k]);
        id[k] = 1.0f / df[k];
        oi[k] = of[k] * id[k];
      } else {
        id[k] = oi[k] = 0.f;
      }
    }
    auto slab = [&](const float* base, int stride) {  // box components at base[k * stride], lo then hi
      float t0[3], t1[3];
      for (int k = 0; k < 3; ++k) {
        t0[k] = fmaf(base[k * stride], id[k], -oi[k]);
        t1[k] = fmaf(base[(3 + k) * stride], id[k], -oi[k]);
      }
      float nr = fmaxf(fmaxf(fminf(t0[0], t1[0]), fminf(t0[1], t1[1])), fmaxf(fminf(t0[2], t1[2]), 0.f));
      float fr = fminf(fminf(fmaxf(t0[0], t1[0]), fmaxf(t0[1], t1[1])), fminf(fmaxf(t0[2], t1[2]), INFINITY));
      return nr <= fr;
    };
    Hit best{INFINITY, 0xffffffffu};
    for (int i = bv.n_tree_objs; i < bv.n_objects; ++i) {
      better(best, first_root(recs[i], o, d, time, a), recs[i].orig);
      if (stats) stats[1]++;
    }
    for (int c = 0; c < bv.n_clusters; ++c) {
      if (stats && c % 32 == 0) stats[0]++;
      if (!slab(cboxes + c, bv.ncl_pad)) continue;
      if (stats) stats[0]++;
      for (int i = 0; i < 32; ++i) {
        const int obj = 32 * c + i;
        if (obj >= bv.n_tree_objs) break;
        if (!slab(oboxes + (size_t)c * 192 + i, 32)) continue;
        better(best, first_root(recs[obj], o, d, time, a), recs[obj].orig);
        if (stats) stats[1]++;
      }
    }
    if (!(best.t == scan.t && (best.orig == scan.orig || scan.t == INFINITY))) ++bad;
  }
  return bad;
}

// FNV-1a digest of the packed BVH blob's nodes and object records (the box tables of the cooperative search are
// derived from the same boxes and checked by bvh_check_rays_coop): freezes the builder's output in
// tests/test_bvh_host.py so that speed work on the builder cannot silently change the trees the kernel was tuned on.
uint64_t bvh_blob_hash(const tor_hittable* objs, int n, const tor_camera* cam, int64_t* n_nodes) {
  std::vector<tor_hittable> v(objs, objs + n);
  PackedBvh pb;
  std::string err;
  if (!pack_bvh(v, *cam, &pb, &err)) return 0;
  if (n_nodes) *n_nodes = pb.view.n_nodes;
  uint64_t h = 1469598103934665603ull;
  auto feed = [&](size_t begin, size_t end) {
    for (size_t i = begin; i < end; ++i) {
      h ^= pb.blob[i];
      h *= 1099511628211ull;
    }
  };
  feed(pb.view.off_nodes, pb.view.off_nodes + (size_t)pb.view.n_nodes * sizeof(BvhNode));
  feed(pb.view.off_objs, pb.view.lane_bytes);
  return h;
}

}  // extern "C"
