// tor_kernels_anim.cuh — the animation's per-frame work either side of render(), on the device
// (scenes_animated.nim:156-225, trace_of_radiance_animation.nim:173-199): the toy physics step and the rebuild of
// the frame's scene.  The reference rebuilds a Scene of 1 601 static spheres on the host every frame; only the
// height of the bouncing spheres ever changes, so here the packed scene stays in HBM and a frame costs
//   anim_step_kernel : `skip` physics steps per sphere (the reference's float64 operations, non-fused), and
//   anim_refit_kernel: the new heights written into the object records and every box that contains them re-fitted
//                      (object and cluster boxes of the cooperative search, leaf and inner boxes of the tree).
// The tree's topology is the one built for the first frame: x and z never change, so it stays a valid (and good)
// partition; boxes only have to be conservative for the image to be bit-identical (tor_bvh.hpp).
#pragma once
#include "tor_bvh.hpp"
#include "tor_kernels.cuh"

namespace tor {

// scenes_animated.nim:159-170 `stepPhysics`, nsteps times: G * dt is the same float64 product in every step
// (dt is float32 widened), so it is formed once on the host.
__global__ void __launch_bounds__(256) anim_step_kernel(double* __restrict__ vel, double* __restrict__ pos_y,
                                                        const double* __restrict__ restitution, int32_t n,
                                                        int32_t nsteps, double dt, double g_dt) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = vel[i], y = pos_y[i];
  const double r = restitution[i];
  for (int32_t s = 0; s < nsteps; ++s) {
    if (v < 0.0 && y < 0.2)  // descending and touching the ground (SmallRadius): bounce
      v = -r * v;
    else
      v -= g_dt;
    y += v * dt;
  }
  vel[i] = v;
  pos_y[i] = y;
}

// float32 strictly below / above a double (round towards the outside, then one more step: the host builder's
// f32_down / f32_up are at most this far out as well)
__device__ __forceinline__ float f32_below(double v) {
  float f = __double2float_rd(v);
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) == 0u) u = 0x80000001u;
  else if (u & 0x80000000u) ++u;
  else --u;
  return __uint_as_float(u);
}
__device__ __forceinline__ float f32_above(double v) {
  float f = __double2float_ru(v);
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) == 0u) u = 0x00000001u;
  else if (u & 0x80000000u) --u;
  else ++u;
  return __uint_as_float(u);
}

struct AnimRefitParams {
  uint8_t* blob;  // the frame's packed BVH (PackedBvh layout), rewritten in place
  BvhView bv;
  const double* pos_y;        // current heights of the moving spheres
  const double* radius;       // their radii
  const int32_t* rec_dyn;     // per tree record: index of its moving sphere, or -1 (static object)
  const int32_t* node_order;  // inner nodes sorted by height (nodes whose children are all leaves first)
  const int32_t* level_off;   // n_levels + 1 offsets into node_order
  int32_t n_levels;
  double pad;      // box padding of the build (2^-19 of the scene's largest coordinate, tor_bvh.hpp)
  double s_limit;  // largest coordinate the padding was derived for
  float* node_y;   // scratch [2 * n_nodes]: y extent of every node's subtree
  int32_t* flag;   // set to 1 when a sphere left the range the padding covers (the caller rebuilds on the host)
};

// One CTA: the tree has ~1 100 nodes and a dozen levels.
__global__ void __launch_bounds__(1024) anim_refit_kernel(const AnimRefitParams P) {
  const BvhView& bv = P.bv;
  double* const recs = reinterpret_cast<double*>(P.blob + bv.off_objs);
  float* const oboxes = reinterpret_cast<float*>(P.blob + bv.off_oboxes);
  float* const cboxes = reinterpret_cast<float*>(P.blob + bv.off_cboxes);
  float* const nodes = reinterpret_cast<float*>(P.blob + bv.off_nodes);
  const int32_t* const nodes_i = reinterpret_cast<const int32_t*>(P.blob + bv.off_nodes);
  constexpr int kRecDoubles = sizeof(ObjRec) / 8, kNodeFloats = sizeof(BvhNode) / 4;
  // A. heights into the records, y extent of every moving object's box
  for (int32_t j = threadIdx.x; j < bv.n_tree_objs; j += blockDim.x) {
    const int32_t dyn = P.rec_dyn[j];
    if (dyn < 0) continue;
    const double y = P.pos_y[dyn], r = fabs(P.radius[dyn]);
    recs[(size_t)j * kRecDoubles + 1] = y;                // ObjRec::c0[1]
    const double e = r + 1e-9 * (fabs(y) + r);            // the builder's head-room for the reference's own rounding
    if (!(fabs(y) + e + P.pad <= P.s_limit)) *P.flag = 1;  // also catches NaN
    const int32_t c = j >> 5, i = j & 31;
    oboxes[c * 192 + 32 + i] = f32_below(y - e - P.pad);
    oboxes[c * 192 + 128 + i] = f32_above(y + e + P.pad);
  }
  __syncthreads();
  // B. cluster boxes
  for (int32_t c = threadIdx.x; c < bv.n_clusters; c += blockDim.x) {
    float lo = 3.0e38f, hi = -3.0e38f;
    const int32_t cnt = min(32, bv.n_tree_objs - 32 * c);
    for (int32_t i = 0; i < cnt; ++i) {
      lo = fminf(lo, oboxes[c * 192 + 32 + i]);
      hi = fmaxf(hi, oboxes[c * 192 + 128 + i]);
    }
    cboxes[bv.ncl_pad + c] = lo;
    cboxes[4 * bv.ncl_pad + c] = hi;
  }
  // C. the tree, bottom up: one height class at a time
  for (int32_t lv = 0; lv < P.n_levels; ++lv) {
    for (int32_t k = P.level_off[lv] + threadIdx.x; k < P.level_off[lv + 1]; k += blockDim.x) {
      const int32_t n = P.node_order[k];
      float nlo = 3.0e38f, nhi = -3.0e38f;
      for (int ch = 0; ch < 2; ++ch) {
        const int32_t ref = nodes_i[n * kNodeFloats + 12 + ch];
        float lo = 3.0e38f, hi = -3.0e38f;
        if (ref >= 0) {
          lo = P.node_y[2 * ref];
          hi = P.node_y[2 * ref + 1];
        } else {
          const int32_t v = ~ref, first = v >> 4, cnt = v & 15;
          for (int32_t q = first; q < first + cnt; ++q) {
            lo = fminf(lo, oboxes[(q >> 5) * 192 + 32 + (q & 31)]);
            hi = fmaxf(hi, oboxes[(q >> 5) * 192 + 128 + (q & 31)]);
          }
          if (cnt == 0) continue;  // an empty child keeps its empty box
        }
        nodes[n * kNodeFloats + 6 * ch + 1] = lo;  // lo{ch}[1]
        nodes[n * kNodeFloats + 6 * ch + 4] = hi;  // hi{ch}[1]
        nlo = fminf(nlo, lo);
        nhi = fmaxf(nhi, hi);
      }
      P.node_y[2 * n] = nlo;
      P.node_y[2 * n + 1] = nhi;
    }
    __syncthreads();
  }
}

}  // namespace tor
