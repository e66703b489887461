// tor_kernels_bvh.cuh — the sm_100a render kernel with a bounding-volume hierarchy in front of the
// reference's sphere test (the image is bit-identical to the brute-force scan), plus the small kernels around it:
// the cost sort of the exact mode's pixel queue, the partial-sum reduction of the split-stream mode, draw (gamma),
// RGB8 quantisation and the BT.601 Y'CbCr 4:2:0 conversion of the video export.
//
// Work units.  Exact mode: a pixel with all its samples (one RNG stream, render.nim:59-67).  Split-stream mode
// (TOR_MODE_FAST, include/tor_b200.h): a (pixel, sample range) pair with its own counter-seeded RNG substream;
// template parameter CHUNKED selects the warp-level queue that hands neighbouring units to the lanes of a warp.
//
// Why the result cannot change: the in-order scan of hittables_lists.nim:48-55 returns
//     argmin over objects of (t_i, i),   t_i = the object's first root in (t_min, +inf)
// (see exact_first_root in tor_kernels.cuh).  The hierarchy (tor_bvh.hpp) is only a conservative
// *filter*: a subtree is skipped when the ray misses its padded float32 box or enters it beyond the
// closest root found so far; every object that survives is evaluated with the reference's own
// non-fused float64 arithmetic, and ties go to the lowest original index.
//
// Execution model: persistent lanes, one work unit per lane at a time.  The kernel alternates two phases:
//   S  "shade": lanes whose traversal has finished shade the hit / the sky, start the next segment,
//      sample or pixel (pixels come from a global atomic queue) and set up the next traversal;
//   T  "traverse": while-while traversal (inner nodes until a leaf, then the leaf's spheres), which
//      keeps running until at least `refill` lanes of the warp wait for phase S (or nobody traverses).
// A lane keeps its traversal state (node, stack, closest hit) across phase S of its neighbours, so the
// expensive float64 shading code (sincos, sqrt, divides) always runs with many lanes, and the
// traversal loop always with at least 32 - refill.  The warp re-converges explicitly (__syncwarp) between
// the steps: left to itself the compiler's schedule ran the sphere test at 2 lanes per instruction.
#pragma once
#include "tor_bvh.hpp"
#include "tor_kernels.cuh"

namespace tor {

// Cooperative pixels run in their OWN kernel (render_coop_kernel below): one CTA of 512 threads per SM that is set
// aside — with 128 registers per thread such a CTA takes the SM's whole register file, so no CTA of the lane kernel can
// share the SM: a cooperative warp is a serial dependency chain, and the hundreds of lanes of a lane CTA on the same
// schedulers would take most of its issue slots.  `cw` of the CTA's 16 warps work (all 16 by default: a warp alone on
// its SM needs 1.0 us per segment, sixteen need 2.3 us each but deliver three times the pixels per SM, and the SMs
// set aside are what the lanes lose).  The lane kernel is launched with its usual grid; its CTAs that find no SM free
// start when cooperative CTAs are done and go straight to the queue.  CoopLayout tells both kernels and the scatter
// kernels how the ranked pixels are laid out.
static constexpr uint32_t kCoopWarps = 16;   // working warps per cooperative CTA (default; CoopLayout::cw)
static constexpr uint32_t kCoopBlock = 512;  // threads per cooperative CTA (fills an SM's register file)
struct CoopLayout {
  uint32_t coop_grid;  // CTAs of the cooperative kernel (0: no cooperative pixels in this launch)
  uint32_t cw;         // working warps per cooperative CTA (default kCoopWarps; the rest only hold the registers)
  uint32_t n_fast, cw_fast;  // the first n_fast CTAs run only cw_fast warps and start on the cw_fast * n_fast most
                             // expensive pixels (pixel k -> CTA k mod n_fast): the longest chains of the launch get
                             // SMs with few neighbours; everything else comes from the shared queue
  uint32_t grid, wpc;  // lane kernel: CTAs and warps per CTA
  uint32_t per_sm;     // lane-kernel CTAs that one cooperative CTA keeps off its SM
  uint32_t lanes;      // lanes of each warp that take pixels (BvhRenderParams::lanes_per_warp): the dealt wave gives a
                       // warp `lanes` pixels; its slots in `order` stay 32 apart
  uint32_t sorted;     // dealt wave: 0 = every warp gets the same mix of costs (tiers), 1 = every warp gets `lanes`
                       // consecutive ranks and every CTA the same mix of warps (place_rank)
};
// cooperative CTAs that get pixels when K pixels are cooperative
__host__ __device__ __forceinline__ uint32_t coop_ctas_used(const CoopLayout& c, uint32_t K) {
  const uint32_t head = c.n_fast * c.cw_fast;  // pixels the fast CTAs start on
  uint32_t want;
  if (K <= head) {
    want = K < c.n_fast ? K : c.n_fast;
  } else {
    want = c.n_fast + (K - head + c.cw - 1u) / c.cw;
  }
  return want < c.coop_grid ? want : c.coop_grid;
}
// warps of the lane kernel that are dealt pixels: those of the CTAs that start at once
__host__ __device__ __forceinline__ uint32_t deal_warps(const CoopLayout& c, uint32_t K) {
  const uint32_t late = coop_ctas_used(c, K) * c.per_sm;
  return (late < c.grid ? c.grid - late : 0u) * c.wpc;
}

struct BvhRenderParams {
  BvhView bv;
  const uint8_t* blob;  // device copy of PackedBvh::blob
  tor_camera cam;
  double* pixels;
  int32_t nrows, ncols, spp;
  int32_t max_depth;
  double inv_spp, inv_gamma;
  int32_t row_begin, row_step, nsel_rows;
  uint32_t count_segments;
  unsigned long long* work_counter;
  unsigned long long* counters;  // [0] primary rays, [1] segments, [2] box-pair tests, [3] exact sphere tests
  int32_t refill;                // waiting lanes of a warp that trigger a shade phase
  // Pixel scheduling.  order == NULL: queue slot i is pixel i (row-major over the selected rows); otherwise pixel
  // order[i].  cost != NULL marks the cost pre-pass: nothing is written to `pixels`, cost[pixel] receives the
  // number of bounce segments the pixel's first `spp` samples took.
  const uint32_t* order;
  uint32_t* cost;
  // order == NULL and scramble != 0: queue slot i is pixel (i * scramble) mod total (scramble coprime to total): lanes
  // of a warp get pixels from all over the image, so the few expensive pixels of a latency-bound render end up in
  // different warps and each runs in a nearly idle warp once its cheap neighbours are done.
  uint32_t scramble;
  // order != NULL and first_wave != 0: the head of `order` is not queued but dealt, 32 entries per warp: lane l of
  // the warp with dealing rank w starts on order[32 * w + l] (0xffffffff = nothing), and the queue serves what
  // follows.  The scatter kernels deal the most expensive pixels: consecutive ranks per warp (CoopLayout::sorted)
  // or tiers (place_rank).  first_wave = 32 * (warps of the grid): the dealt region when no warp is
  // cooperative; with n_coop cooperative warps (below) the region is 32 * (warps - n_coop) entries and the queue
  // starts right behind it.  Both the scatter kernels and this kernel derive the layout from sched[0].
  uint32_t first_wave;
  // Warp-cooperative pixels (exact mode).  sched[0] = n_coop, computed on the device from the cost histogram
  // (cost_offsets_kernel): the n_coop most expensive pixels are not given to single lanes; render_coop_kernel traces
  // pixel coop_list[k] with a whole warp whose lanes share the closest-hit search of every bounce segment.
  // NULL = no cooperative pixels.  deal_ticket[0]: the lane kernel's CTAs take their dealing rank in arrival order;
  // [1]: cooperative CTAs resident (coop_gate_kernel); [2]: head of the cooperative pixel queue.
  const uint32_t* sched;
  const uint32_t* coop_list;
  CoopLayout coop;
  unsigned int* deal_ticket;
  // Developer aid (NULL in production): %globaltimer stamps — [3k..3k+2] = start, end, segments of cooperative warp k
  // (k = CTA * kCoopWarps + warp), from entry 3 * 4096 on, [2w], [2w+1] = start and end of lane-kernel warp w, and
  // from entry 3 * 4096 + 2 * 8192 on the warps of the hand-off launch like the cooperative ones.
  unsigned long long* dbg_times;
  // Lanes of each warp that take pixels (1..32; 32 unless the TOR_BVH_LANES tuning knob says otherwise).
  int32_t lanes_per_warp;
  // Split-stream mode (TOR_MODE_FAST, include/tor_b200.h): every pixel's sample loop is cut into 2^sub_log2
  // consecutive ranges with their own RNG substream, and the unit of work a lane pulls from the queue is one
  // (pixel, range) pair: unit = pixel << sub_log2 | range.  `pixels` then holds one partial sum per unit
  // (3 doubles at 3*unit) and substream_reduce_kernel adds them up.  sub_log2 == 0 is the exact mode: the unit is
  // the pixel and its one stream is the reference's (render.nim:59-67).
  uint32_t sub_log2;
  // Queue slots a warp takes at once when there is no cost-ranked order (multiple of 32), and the distance from the
  // end of the queue below which warps take 32 at a time so that the render does not end on one warp's long chunk.
  uint32_t chunk;
  unsigned long long chunk_guard;
  uint32_t endgame_min_chunk;  // cost-ranked queue: smallest share a warp takes near the end (0: always `chunk`)
  // Late hand-off (exact mode, cost-ranked launches; NULL = off).  Once `handoff_pct` % of the lane kernel's dealt
  // warps have run out of work the launch is in its tail: the pixels still in flight are single serial chains on
  // warps that are mostly idle.  From then on every lane — and every warp of the concurrent render_coop_kernel —
  // parks its pixel at its next sample boundary as a HandoffRec {pixel, samples done, generator state, colour sum}
  // and retires; a second launch of render_coop_kernel (handoff_mode = 1) on the whole GPU continues each parked
  // pixel with one warp from exactly that state.  A pixel's stream and the order of its additions are untouched, so
  // the image is the same bit for bit.  Records: [handoff_cap_a from cooperative warps — the longest chains, taken
  // first][one per lane].  deal_ticket[3] = dealt lane warps that are done, [4] = the tail flag, [5], [6] = records
  // in the two regions, [7] = queue head of the second launch.
  uint8_t* handoff;
  uint32_t handoff_cap_a;
  uint32_t handoff_pct;
  uint32_t handoff_min_left;  // a pixel with fewer samples left than this stays where it is
  uint32_t handoff_mode;
};

struct __align__(16) HandoffRec {
  uint32_t pid;
  int32_t sample;  // samples already in the sum (Lane::sample)
  uint64_t s0, s1, s2, s3;
  double px, py, pz;
};
static_assert(sizeof(HandoffRec) == 64, "HandoffRec is one 64-byte record");

// One object of a leaf (or of the "always" list) against the ray: the reference's arithmetic
// (spheres.nim:28-49 / moving_spheres.nim:39-67), operation for operation.  r2 = radius*radius and
// dc = center1 - center0 were computed on the host with the same IEEE operations.
// 16-byte loads of the read-only scene data.  For the shared-memory copies the 32-bit shared address is formed once
// per kernel (a generic pointer costs a window-base computation per access inside the traversal loop).
template <bool SHARED>
__device__ __forceinline__ float4 ld16f(const void* generic_base, uint32_t shared_base, int32_t idx16) {
  if (SHARED) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "r"(shared_base + 16u * (uint32_t)idx16));
    return v;
  }
  return reinterpret_cast<const float4*>(generic_base)[idx16];
}
template <bool SHARED>
__device__ __forceinline__ int4 ld16i(const void* generic_base, uint32_t shared_base, int32_t idx16) {
  if (SHARED) {
    int4 v;
    asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
        : "r"(shared_base + 16u * (uint32_t)idx16));
    return v;
  }
  return reinterpret_cast<const int4*>(generic_base)[idx16];
}

template <bool SHARED>
__device__ __forceinline__ float ld4f(const void* generic_base, uint32_t shared_base, int32_t idx4) {
  if (SHARED) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(shared_base + 4u * (uint32_t)idx4));
    return v;
  }
  return reinterpret_cast<const float*>(generic_base)[idx4];
}

struct QCache {  // lerp parameter of moving_spheres.nim:41-42, one divide per (time0, time1) per segment
  double t0, t1, q;
};

__device__ __forceinline__ V3 rec_center(const double2* __restrict__ r, uint32_t kind_mat, double time, QCache& qc) {
  double2 a0 = r[0], a1 = r[1];
  V3 c0 = v3(a0.x, a0.y, a1.x);
  if ((kind_mat & 0xffu) == TOR_MOVING_SPHERE) {
    double2 a3 = r[3], a4 = r[4];
    double q;
    if (kind_mat & kObjUnitInterval) {
      // time0 = +0.0, time1 = 1.0: (time - 0.0) / (1.0 - 0.0) is `time` itself, bit for bit
      q = time;
    } else {
      double2 a5 = r[5];
      if (!(a5.x == qc.t0 && a5.y == qc.t1)) {
        qc.t0 = a5.x;
        qc.t1 = a5.y;
        qc.q = (time - a5.x) / (a5.y - a5.x);
      }
      q = qc.q;
    }
    return c0 + (q * v3(a3.x, a3.y, a4.x));
  }
  return c0;
}

// What a kernel variant keeps in shared memory: [blob prefix of bytes0][box tables, bytes1].
struct StagePlan {
  uint32_t bytes0, bytes1;
};
template <int STAGE>
__host__ __device__ __forceinline__ StagePlan stage_plan(const BvhView& bv) {
  StagePlan s;
  s.bytes0 = STAGE == 2 ? bv.lane_bytes : (STAGE == 1 ? bv.nodes_bytes : 0u);
  // The box tables of the cooperative search are never staged (render_coop_kernel reads everything through L1).
  s.bytes1 = 0u;
  return s;
}

// Closest hit of the current bounce segment + the per-segment constants of the sphere test.
struct Closest {
  double a;       // d.d  (spheres.nim:30)
  double best_t;  // closest root so far
  float best_f;   // the same rounded up to float32
  uint32_t best_orig;
  int32_t best_rec;
  QCache qc;
};
// The float32 ray of the slab tests: 1/d and o/d.
struct SlabRay {
  float idx, idy, idz, oix, oiy, oiz;
};

// One object (record ri) against the ray: the reference's arithmetic (spheres.nim:28-49 / moving_spheres.nim:39-67),
// operation for operation.  r2 = radius*radius and dc = center1 - center0 were computed on the host with the same
// IEEE operations.
__device__ __forceinline__ void test_record(const double2* __restrict__ recs, int32_t ri, const V3 o, const V3 d,
                                            double time, Closest& C) {
  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  const double t_min = 0.001;  // render.nim:28
  const double a = C.a;
  const double2* __restrict__ r = recs + kRecStride16 * ri;
  const double2 a1 = r[1], a2 = r[2];
  const uint32_t kind_mat = (uint32_t)__double_as_longlong(a2.x);
  const uint32_t orig = (uint32_t)((unsigned long long)__double_as_longlong(a2.x) >> 32);
  V3 oc = o - rec_center(r, kind_mat, time, C.qc);
  double half_b = dot(oc, d);
  double c = len2(oc) - a1.y;
  double disc = half_b * half_b - a * c;
  if (disc > 0) {
    const double root = sqrt(disc);
    const double x1 = -half_b - root, x2 = -half_b + root;  // the reference's two numerators (spheres.nim:37-48)
    // Three shortcuts that skip IEEE divides without changing the outcome (division by a > 0 and rounding are
    // monotonic; the 2^-40 margins cover the rounding of the products below):
    //   x2 <= t_min*a*(1 - 2^-40)  =>  both roots round to <= t_min: no root in (t_min, inf)   [the sphere the
    //                                   ray starts on, spheres behind the origin]
    //   x1 >= best_t*a*(1 + 2^-40) =>  the first root is valid and strictly beyond the closest so far
    //   x1 <= t_min*a*(1 - 2^-40)  =>  the first root rounds to <= t_min: go straight to the second
    //                                   (a ray that starts on this sphere and crosses it: every glass-internal segment)
    const bool a_ok = a >= 1e-200 && a <= 1e200;
    const double ta_lo = (t_min * a) * (1.0 - 0x1p-40);
    const double ba = C.best_t * a;
    const bool none = a_ok && x2 <= ta_lo;
    const bool behind = a_ok && x1 >= ba * (1.0 + 0x1p-40);
    if (!none && !behind) {
      double t = INF;
      const bool skip1 = a_ok && x1 <= ta_lo;
      double sol = skip1 ? 0.0 : x1 / a;
      if (!skip1 && t_min < sol) {
        t = sol;
      } else {
        sol = x2 / a;
        if (t_min < sol) t = sol;
      }
      if (t < C.best_t || (t == C.best_t && orig < C.best_orig && t < INF)) {
        C.best_t = t;
        C.best_orig = orig;
        C.best_rec = ri;
        C.best_f = __double2float_ru(t);
      }
    }
  }
}

// A new bounce segment: resets the closest hit and derives the float32 ray of the slab tests.  Components of d far
// below the largest one are replaced by +-2^-60 * dmax (a direction change below 2^-60, negligible against the box
// padding) so that 1/d stays finite; rays outside the range the padding was derived for test everything instead.
__device__ __forceinline__ void setup_ray(const V3 o, const V3 d, float s_limit, Closest& C, SlabRay& R) {
  C.best_t = __longlong_as_double(0x7ff0000000000000ll);
  C.best_orig = 0xffffffffu;
  C.best_rec = -1;
  C.best_f = __int_as_float(0x7f800000);
  C.qc.t0 = C.qc.t1 = __longlong_as_double(0x7ff8000000000000ll);  // NaN: never equal, forces the first divide
  C.a = len2(d);
  float dxf = __double2float_rn(d.x), dyf = __double2float_rn(d.y), dzf = __double2float_rn(d.z);
  float oxf = __double2float_rn(o.x), oyf = __double2float_rn(o.y), ozf = __double2float_rn(o.z);
  float dmax = fmaxf(fabsf(dxf), fmaxf(fabsf(dyf), fabsf(dzf)));
  float omax = fmaxf(fabsf(oxf), fmaxf(fabsf(oyf), fabsf(ozf)));
  bool ok = dmax >= 0x1p-40f && dmax <= 0x1p40f && omax <= s_limit;
  ok = ok && dxf == dxf && dyf == dyf && dzf == dzf && oxf == oxf && oyf == oyf && ozf == ozf;
  if (ok) {
    float dmin = dmax * 0x1p-60f;
    if (fabsf(dxf) < dmin) dxf = copysignf(dmin, dxf);
    if (fabsf(dyf) < dmin) dyf = copysignf(dmin, dyf);
    if (fabsf(dzf) < dmin) dzf = copysignf(dmin, dzf);
    R.idx = __frcp_rn(dxf);
    R.idy = __frcp_rn(dyf);
    R.idz = __frcp_rn(dzf);
    R.oix = oxf * R.idx;
    R.oiy = oyf * R.idy;
    R.oiz = ozf * R.idz;
  } else {
    R.idx = R.idy = R.idz = 0.f;  // every slab interval becomes [0, 0]: all boxes pass
    R.oix = R.oiy = R.oiz = 0.f;
  }
}

// The slab test of the traversal loop on one box (the same expression, so the same conservativeness argument).
__device__ __forceinline__ bool slab_test(const SlabRay& R, float best_f, float lx, float ly, float lz, float hx,
                                          float hy, float hz) {
  const float ax0 = fmaf(lx, R.idx, -R.oix), ax1 = fmaf(hx, R.idx, -R.oix);
  const float ay0 = fmaf(ly, R.idy, -R.oiy), ay1 = fmaf(hy, R.idy, -R.oiy);
  const float az0 = fmaf(lz, R.idz, -R.oiz), az1 = fmaf(hz, R.idz, -R.oiz);
  const float nr = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.f));
  const float fr = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fminf(fmaxf(az0, az1), best_f));
  return nr <= fr;
}

// ================================================================= warp-cooperative pixels
// A pixel's samples are one serial chain (render.nim:59-67), so the render cannot end before its most expensive
// pixel does, and a single lane needs microseconds per bounce segment.  The n_coop most expensive pixels of the
// cost pre-pass are therefore traced by a whole warp each: every lane carries the same path state (same RNG
// stream, same arithmetic, hence the same bits — nothing is broadcast), and the lanes share the only part of a
// segment that has parallelism, the closest-hit search:
//   1. 32 cluster boxes per step (a cluster = 32 consecutive tree records), float32 slab test;
//   2. the 32 object boxes of every cluster the ray enters; a lane remembers the records whose box it saw entered;
//   3. the remembered candidates through the reference's float64 sphere test (test_rec), all lanes at once;
//   4. argmin over the lanes of (t, original index) — the order-free form of hittables_lists.nim:48-55 — with
//      three warp reductions on the bit pattern of t (monotonic: t_min < t <= +inf).
// Shading then runs on all lanes redundantly (converged, one pass).  The set of objects tested is a superset of
// the objects with a root (same padded boxes as the tree), so the hit is the lane mode's, bit for bit.
// Launched beside the lane kernel on its own high-priority stream (tor_api.cu); see CoopLayout for the placement.
__global__ void __launch_bounds__(kCoopBlock, 1) render_coop_kernel(const __grid_constant__ BvhRenderParams P) {
  const uint32_t warp = threadIdx.x >> 5;
  const bool resume = P.handoff_mode != 0;  // second launch: continue the parked pixels (BvhRenderParams::handoff)
  const bool fast_cta = !resume && blockIdx.x < P.coop.n_fast;
  if (warp >= (fast_cta ? P.coop.cw_fast : P.coop.cw)) return;  // the other warps only hold the SM's registers
  const uint32_t n_coop = resume ? 0u : P.sched[0];
  if (!resume && blockIdx.x >= coop_ctas_used(P.coop, n_coop)) return;
  // first pixel of a fast CTA's warp: fixed (the head of the ranking, spread over the fast CTAs)
  uint32_t first_px = fast_cta ? warp * P.coop.n_fast + blockIdx.x : 0xffffffffu;
  if (!resume && threadIdx.x == 0) atomicAdd(P.deal_ticket + 1, 1u);  // resident: coop_gate_kernel lets the lane kernel start
  // parked pixels: region A first; the first gridDim.x * cw records go out round-robin over the CTAs (the longest
  // chains end up on different SMs), the rest through deal_ticket[7]
  const uint32_t n_rec_a = resume ? P.deal_ticket[5] : 0u, n_rec = resume ? n_rec_a + P.deal_ticket[6] : 0u;
  uint32_t first_rec = resume ? warp * gridDim.x + blockIdx.x : 0xffffffffu;
  HandoffRec* const rec_a = reinterpret_cast<HandoffRec*>(P.handoff);
  HandoffRec* const rec_b = rec_a + P.handoff_cap_a;
  const bool may_park = !resume && P.handoff != nullptr;
  uint32_t tail_seen = 0;  // lane 0: the tail flag as of the previous sample (the load is consumed a sample later)
  unsigned long long dbg_t0 = 0;
  if (P.dbg_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  const BvhView& bv = P.bv;
  // everything is read from global memory: the SM is this CTA's alone, its whole 256 KB is L1
  const double2* __restrict__ recs = reinterpret_cast<const double2*>(P.blob + bv.off_objs);
  const float* __restrict__ cboxes = reinterpret_cast<const float*>(P.blob + bv.off_cboxes);
  const float* __restrict__ oboxes = reinterpret_cast<const float*>(P.blob + bv.off_oboxes);
  const uint32_t cboxes_sa = 0u, oboxes_sa = 0u;
  Lane L;
  L.pix = v3(0, 0, 0);
  L.att = v3(1, 1, 1);
  L.o = v3(0, 0, 0);
  L.d = v3(0, 0, 1);
  L.time = 0.0;
  L.row = L.col = L.sample = L.depth = 0;
  Closest C;
  C.a = 1.0;
  C.qc.t0 = C.qc.t1 = C.qc.q = 0.0;
  SlabRay R;
  unsigned long long seg_count = 0, ray_count = 0;
  uint32_t box_count = 0, test_count = 0;
  const int lane = threadIdx.x & 31;
  const int32_t n_always = bv.n_objects - bv.n_tree_objs;
  // The cooperative pixels are a queue, most expensive first: a warp that finishes one takes the next, so the SMs
  // that are set aside stay busy until the list is empty (there may be several times more pixels than warps).
  for (;;) {
    uint32_t pid;
    int32_t s_first = 0;
    if (resume) {
      uint32_t k = first_rec;
      first_rec = 0xffffffffu;
      if (k == 0xffffffffu) {
        if (lane == 0) k = gridDim.x * P.coop.cw + atomicAdd(P.deal_ticket + 7, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
      }
      if (k >= n_rec) break;
      const HandoffRec h = k < n_rec_a ? rec_a[k] : rec_b[k - n_rec_a];
      pid = h.pid;
      s_first = h.sample;
      L.rng.s0 = h.s0;
      L.rng.s1 = h.s1;
      L.rng.s2 = h.s2;
      L.rng.s3 = h.s3;
      L.pix = v3(h.px, h.py, h.pz);
    } else {
      uint32_t cr = first_px;
      first_px = 0xffffffffu;
      if (cr == 0xffffffffu) {  // the shared queue starts behind the fast CTAs' first pixels
        if (lane == 0) cr = P.coop.n_fast * P.coop.cw_fast + atomicAdd(P.deal_ticket + 2, 1u);
        cr = __shfl_sync(0xffffffffu, cr, 0);
      }
      if (cr >= n_coop) {
        if (cr < P.coop.n_fast * P.coop.cw_fast) continue;  // a fast warp without a first pixel: on to the queue
        break;
      }
      pid = P.coop_list[cr];
    }
    {
      const int32_t ri = (int32_t)(pid / (uint32_t)P.ncols);
      L.col = (int32_t)(pid - (uint32_t)ri * (uint32_t)P.ncols);
      L.row = P.row_begin + ri * P.row_step;
    }
    if (!resume) {
      rng_seed_pixel(L.rng, L.row, L.col, 0);  // render.nim:59-60
      L.pix = v3(0, 0, 0);
    }
    bool parked = false;
    for (int32_t s = s_first; s < P.spp; ++s) {  // render.nim:62
      if (may_park) {
        // the launch is in its tail: park the pixel (also one not yet started) for the second launch, which spreads
        // what is left over every SM
        const uint32_t tail = __shfl_sync(0xffffffffu, tail_seen, 0);
        if (tail && P.spp - s >= (int32_t)P.handoff_min_left) {
          if (lane == 0) {
            HandoffRec h;
            h.pid = pid;
            h.sample = s;
            h.s0 = L.rng.s0;
            h.s1 = L.rng.s1;
            h.s2 = L.rng.s2;
            h.s3 = L.rng.s3;
            h.px = L.pix.x;
            h.py = L.pix.y;
            h.pz = L.pix.z;
            rec_a[atomicAdd(P.deal_ticket + 5, 1u)] = h;
          }
          parked = true;
          break;
        }
        if (lane == 0) tail_seen = *(volatile const unsigned int*)(P.deal_ticket + 4);
      }
      start_sample(L, P.cam, P.nrows, P.ncols);
      if (lane == 0) ++ray_count;
      V3 color = v3(0, 0, 0);
      for (;;) {  // render.nim:25-47, one bounce segment per pass
        if (lane == 0) ++seg_count;
        setup_ray(L.o, L.d, bv.s_limit, C, R);
        // Candidate records of this lane, kept in registers: a lane whose object box is entered in a cluster
        // round remembers the record, and the exact tests run afterwards, all lanes together.  Objects without a
        // finite box ("always" list: the ground sphere) start out as the candidates of the top lanes.  More
        // than three candidates in one lane, or more than 32 always-objects, fall back to testing every record.
        int32_t p0 = -1, p1 = -1, p2 = -1;
        bool overflow = n_always > 32;
        if (lane >= 32 - n_always) p0 = bv.n_tree_objs + (31 - lane);
        // one object-box test of this lane in cluster ci (ci < 0: none); branch-free so that the four tests of a
        // pass overlap in the pipeline
        auto obj_round = [&](int32_t ci) {
          const int32_t cj = ci < 0 ? 0 : ci;
          const int32_t obj = cj * 32 + lane;
          const int32_t ob = cj * 192 + lane;
          const bool in =
              slab_test(R, C.best_f, ld4f<false>(oboxes, oboxes_sa, ob), ld4f<false>(oboxes, oboxes_sa, ob + 32),
                       ld4f<false>(oboxes, oboxes_sa, ob + 64), ld4f<false>(oboxes, oboxes_sa, ob + 96),
                       ld4f<false>(oboxes, oboxes_sa, ob + 128), ld4f<false>(oboxes, oboxes_sa, ob + 160));
          return (ci >= 0 && obj < bv.n_tree_objs && in) ? obj : -1;
        };
        auto remember = [&](int32_t obj) {
          const bool has = obj >= 0;
          overflow = overflow || (has && p2 >= 0);
          p2 = (has && p1 >= 0) ? obj : p2;
          p1 = (has && p0 >= 0 && p1 < 0) ? obj : p1;
          p0 = (has && p0 < 0) ? obj : p0;
        };
        for (int32_t cb = 0; cb < bv.n_clusters; cb += 32) {
          const int32_t c = cb + lane;
          const int32_t cc = c < bv.n_clusters ? c : 0;  // padding lanes read cluster 0 and are masked out
          const bool h =
              slab_test(R, C.best_f, ld4f<false>(cboxes, cboxes_sa, cc), ld4f<false>(cboxes, cboxes_sa, bv.ncl_pad + cc),
                       ld4f<false>(cboxes, cboxes_sa, 2 * bv.ncl_pad + cc),
                       ld4f<false>(cboxes, cboxes_sa, 3 * bv.ncl_pad + cc),
                       ld4f<false>(cboxes, cboxes_sa, 4 * bv.ncl_pad + cc),
                       ld4f<false>(cboxes, cboxes_sa, 5 * bv.ncl_pad + cc)) &&
              c < bv.n_clusters;
          unsigned hit_clusters = __ballot_sync(0xffffffffu, h);
          if (lane == 0) box_count += 1u + (uint32_t)__popc(hit_clusters);
          while (hit_clusters) {  // four entered clusters per pass
            int32_t ci[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ci[j] = hit_clusters ? cb + __ffs(hit_clusters) - 1 : -1;
              hit_clusters &= hit_clusters - 1u;
            }
            int32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = obj_round(ci[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) remember(o[j]);
          }
        }
        __syncwarp();
        // the reference's test on the candidates: usually one pass, every lane with a candidate at once
        if (p0 >= 0) {
          test_record(recs, p0, L.o, L.d, L.time, C);
          ++test_count;
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, p1 >= 0 || overflow)) {  // rare
          for (int k = 1; k < 3; ++k) {
            const int32_t cand = k == 1 ? p1 : p2;
            if (cand >= 0) {
              test_record(recs, cand, L.o, L.d, L.time, C);
              ++test_count;
            }
            __syncwarp();
          }
          if (__any_sync(0xffffffffu, overflow)) {
            for (int32_t ri = lane; ri < bv.n_objects; ri += 32) {
              test_record(recs, ri, L.o, L.d, L.time, C);
              ++test_count;
            }
            __syncwarp();
          }
        }
        // closest hit of the warp: lexicographic minimum of (t, original index), then the winner's record
        {
          const unsigned long long tb = (unsigned long long)__double_as_longlong(C.best_t);
          const uint32_t hi = (uint32_t)(tb >> 32), lo = (uint32_t)tb;
          const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
          const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
          const bool is_min = hi == mhi && lo == mlo;
          const uint32_t morig = __reduce_min_sync(0xffffffffu, is_min ? C.best_orig : 0xffffffffu);
          // several lanes can hold the winner only when they hold the same record (original indices are unique)
          C.best_rec = (int32_t)__reduce_min_sync(0xffffffffu, (is_min && C.best_orig == morig) ? (uint32_t)C.best_rec
                                                                                              : 0xffffffffu);
          C.best_t = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | (unsigned long long)mlo));
        }
        if (C.best_rec < 0) {
          color = shade_miss(L);
          break;
        }
        const double2* __restrict__ r = recs + kRecStride16 * C.best_rec;
        const double2 a2 = r[2], a6 = r[6], a7 = r[7];
        const uint32_t kind_mat = (uint32_t)__double_as_longlong(a2.x);
        Surface S;
        S.center = rec_center(r, kind_mat, L.time, C.qc);  // moving_spheres.nim:61
        S.inv_r = a2.y;
        S.albedo = v3(a6.x, a6.y, a7.x);
        S.fuzz_or_ior = a7.y;
        S.mat_kind = (kind_mat >> 8) & 0xffu;
        if (shade_hit(L, C.best_t, S, P.max_depth)) break;  // absorbed or depth exhausted: black
      }
      L.pix.x += color.x;  // render.nim:67
      L.pix.y += color.y;
      L.pix.z += color.z;
    }
    if (lane == 0 && !parked) {
      double* out = P.pixels + 3ull * pid;  // the sum; draw_kernel applies canvas.nim:47-54 afterwards
      out[0] = L.pix.x;
      out[1] = L.pix.y;
      out[2] = L.pix.z;
    }
    __syncwarp();
  }
  if (P.dbg_times && lane == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    unsigned long long* e = P.dbg_times + (resume ? 3ull * 4096ull + 2ull * 8192ull : 0ull) + 3ull * (blockIdx.x * P.coop.cw + warp);
    e[0] = dbg_t0;
    e[1] = t1;
    e[2] = seg_count;
  }
  if (P.count_segments) {
    unsigned long long box_sum = box_count, test_sum = test_count;
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
      box_sum += __shfl_down_sync(0xffffffffu, box_sum, ofs);
      test_sum += __shfl_down_sync(0xffffffffu, test_sum, ofs);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(P.counters + 0, ray_count);
      atomicAdd(P.counters + 1, seg_count);
      atomicAdd(P.counters + 2, box_sum);
      atomicAdd(P.counters + 3, test_sum);
    }
  }
}

// Runs on the lane kernel's stream just before it: returns once the cooperative CTAs that have pixels are resident
// on their SMs (or after ~1 ms, whatever happens: nothing depends on it but the placement).  Without it the lane
// kernel's persistent grid can take every SM first, and the cooperative pixels — the longest chains of the launch —
// would only start when the lanes run out of work.
__global__ void coop_gate_kernel(const uint32_t* __restrict__ sched, const unsigned int* arrived, CoopLayout lay) {
  const uint32_t n_used = coop_ctas_used(lay, sched[0]);
  const long long t0 = clock64();
  while (*(volatile const unsigned int*)arrived < n_used && clock64() - t0 < 2000000ll) __nanosleep(500);
}

template <int BLOCK, int STAGE, bool CHUNKED>
__global__ void __launch_bounds__(BLOCK, BLOCK <= 256 ? 512 / BLOCK : 1) render_bvh_kernel(const __grid_constant__ BvhRenderParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t stage_bar;

  const int tid = threadIdx.x;
  const BvhView& bv = P.bv;
  // Staged in shared memory: STAGE 2 = nodes + records, STAGE 1 = nodes (stage_plan); the rest comes through L1.
  const StagePlan stg = stage_plan<STAGE>(bv);
  if (tid == 0) {
    mbar_init(&stage_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (stg.bytes0 + stg.bytes1) {
    if (tid == 0) {
      mbar_expect_tx(&stage_bar, stg.bytes0 + stg.bytes1);
      const uint32_t kChunk = 32768;
      for (uint32_t off = 0; off < stg.bytes0; off += kChunk) {
        uint32_t n = stg.bytes0 - off < kChunk ? stg.bytes0 - off : kChunk;
        tma_bulk_g2s(smem + off, P.blob + off, n, &stage_bar);
      }
      for (uint32_t off = 0; off < stg.bytes1; off += kChunk) {
        uint32_t n = stg.bytes1 - off < kChunk ? stg.bytes1 - off : kChunk;
        tma_bulk_g2s(smem + stg.bytes0 + off, P.blob + bv.off_cboxes + off, n, &stage_bar);
      }
    }
    mbar_wait(&stage_bar, 0);
  }
  const float4* __restrict__ nodes = reinterpret_cast<const float4*>((STAGE >= 1 ? smem : P.blob) + bv.off_nodes);
  const double2* __restrict__ recs = reinterpret_cast<const double2*>((STAGE == 2 ? smem : P.blob) + bv.off_objs);
  const uint32_t nodes_sa = smem_u32(smem) + bv.off_nodes;  // meaningful for STAGE >= 1 only

  __shared__ uint32_t cta_ticket;
  __shared__ unsigned long long warp_chunk[CHUNKED ? BLOCK / 32 : 1][2];  // [next, end) slots of each warp's chunk
  unsigned long long* const wchunk = warp_chunk[CHUNKED ? tid >> 5 : 0];
  if (CHUNKED && (tid & 31) == 0) wchunk[0] = wchunk[1] = 0ull;
  __syncwarp();

  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  const double t_min = 0.001;  // render.nim:28
  const unsigned long long total_px = (unsigned long long)P.nsel_rows * (unsigned long long)P.ncols;
  const unsigned long long total_units = total_px << P.sub_log2;
  // Layout of the cost-ranked order (see BvhRenderParams): [dealt: 32 per non-cooperative warp][queue]
  unsigned long long dbg_t0 = 0;
  if (P.dbg_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  const uint32_t n_coop = P.sched ? P.sched[0] : 0u;  // pixels render_coop_kernel traces
  const unsigned long long grid_warps = (unsigned long long)gridDim.x * (unsigned long long)(BLOCK / 32);
  // Dealing rank of this warp.  With cooperative pixels some of this grid's CTAs start late (their SMs are held by
  // render_coop_kernel), so the ranks go out in arrival order: the CTAs that start at once share the dealt pixels.
  uint32_t deal_rank = (uint32_t)(tid >> 5) * gridDim.x + blockIdx.x;
  uint32_t n_deal_warps = gridDim.x * (uint32_t)(BLOCK / 32);
  if (P.sched) {
    if (tid == 0) cta_ticket = atomicAdd(P.deal_ticket, 1u);
    __syncthreads();
    n_deal_warps = deal_warps(P.coop, n_coop);
    deal_rank = cta_ticket * (uint32_t)(BLOCK / 32) + (uint32_t)(tid >> 5);
    if (deal_rank >= n_deal_warps) deal_rank = 0xffffffffu;
  }
  const uint32_t first_wave = P.first_wave ? n_deal_warps * 32u : 0u;
  const uint32_t n_ranked = (uint32_t)total_px - n_coop;  // pixels that go to single lanes
  // length of the shared queue: everything, or what the cost-ranked order leaves after the dealt first wave
  const uint32_t n_dealt_cap = P.first_wave ? n_deal_warps * (uint32_t)P.lanes_per_warp : 0u;  // pixels the dealt wave holds
  const unsigned long long queue_len =
      P.order ? (unsigned long long)(n_ranked - (n_ranked < n_dealt_cap ? n_ranked : n_dealt_cap)) : total_units;
  const int refill = P.refill;

  Lane L;
  L.pix = v3(0, 0, 0);
  L.att = v3(1, 1, 1);
  L.o = v3(0, 0, 0);
  L.d = v3(0, 0, 1);
  L.time = 0.0;
  L.row = L.col = L.sample = L.depth = 0;
  uint32_t pid = 0;      // work unit: pixel index inside the selected rows (<< sub_log2 | sample range)
  uint32_t pix_seg = 0;  // bounce segments of the current unit
  bool active = false, need_pixel = (tid & 31) < P.lanes_per_warp, need_sample = false;
  bool first_fetch = P.first_wave != 0 && deal_rank != 0xffffffffu;  // warps that are dealt nothing go straight to the queue
  bool trav_done = false;  // the current segment's closest hit is final
  bool need_setup = false;  // a new segment needs its traversal state
  bool tail_on = false;     // late hand-off: the launch is in its tail (warp-uniform)
  unsigned int tail_seen = 0;
  unsigned long long seg_count = 0, ray_count = 0;
  uint32_t box_count = 0, test_count = 0;  // per lane: far below 2^32 (a lane sees ~1e5 segments per render)

  // traversal state
  int2 stk[kBvhStackDepth];  // {child reference, float bits of the entry distance}: one 8-byte local access each
  int sp = 0;
  int32_t cur = 0;
  SlabRay R;  // 1/d and o/d in float32
  R.idx = R.idy = R.idz = R.oix = R.oiy = R.oiz = 0.f;
  Closest C;
  C.a = 1.0;
  C.best_t = INF;
  C.best_f = 0.f;
  C.best_orig = 0xffffffffu;
  C.best_rec = -1;
  C.qc.t0 = C.qc.t1 = C.qc.q = 0.0;

  auto test_rec = [&](int32_t ri) { test_record(recs, ri, L.o, L.d, L.time, C); };
  auto ray_setup = [&]() { setup_ray(L.o, L.d, bv.s_limit, C, R); };

  // queue slot -> work unit when there is no cost-ranked order: consecutive slots are the sample ranges of one pixel;
  // the pixels come in row-major or scrambled order
  auto unit_of_slot = [&](unsigned long long slot) -> uint32_t {
    const unsigned long long ps = slot >> P.sub_log2;
    const uint32_t px = P.scramble ? (uint32_t)((ps * (unsigned long long)P.scramble) % total_px) : (uint32_t)ps;
    return (px << P.sub_log2) | ((uint32_t)slot & ((1u << P.sub_log2) - 1u));
  };
  // Starts work unit `pid`: samples [s_begin, s_end) of its pixel (the whole loop of render.nim:62 when
  // sub_log2 == 0).  Returns false for an empty range, whose zero sum is written here.
  auto begin_unit = [&]() -> bool {
    pix_seg = 0;
    const uint32_t px = pid >> P.sub_log2, sub = pid & ((1u << P.sub_log2) - 1u);
    const int32_t s_begin = (int32_t)(((unsigned long long)sub * (unsigned long long)P.spp) >> P.sub_log2);
    const int32_t s_end = (int32_t)(((unsigned long long)(sub + 1u) * (unsigned long long)P.spp) >> P.sub_log2);
    if (s_end > s_begin) {
      int32_t ri = (int32_t)(px / (uint32_t)P.ncols);
      L.col = (int32_t)(px - (uint32_t)ri * (uint32_t)P.ncols);
      L.row = P.row_begin + ri * P.row_step;
      rng_seed_pixel(L.rng, L.row, L.col, sub);  // render.nim:59-60
      L.pix = v3(0, 0, 0);
      L.sample = P.spp - (s_end - s_begin);  // counts up to spp
      active = true;
      need_sample = true;
      return true;
    }
    double* out = P.pixels + 3ull * pid;  // no samples: the zero colour goes through draw() (canvas.nim:49-54)
    out[0] = out[1] = out[2] = 0.0;
    return false;
  };

  for (;;) {
    // =================================================================== phase S
    if (active && trav_done) {
      trav_done = false;
      bool sample_done;
      V3 color = v3(0, 0, 0);
      if (P.max_depth <= 0) {  // render.nim:25 — the bounce loop body never runs
        sample_done = true;
      } else {
        ++seg_count;
        ++pix_seg;
        const bool hit = C.best_rec >= 0;
        const double2* __restrict__ r = recs + kRecStride16 * (hit ? C.best_rec : 0);
        const double2 a2 = r[2];
        const uint32_t kind_mat = (uint32_t)__double_as_longlong(a2.x);
        const uint32_t mat_kind = (kind_mat >> 8) & 0xffu;
        // unit_vector(d) once for every lane that needs it (Metal, Dielectric, sky) instead of once per branch
        V3 ud = v3(0, 0, 0);
        if (!hit || mat_kind != TOR_LAMBERTIAN) ud = unit_vector(L.d);
        if (hit) {
          const double2 a6 = r[6], a7 = r[7];
          Surface S;
          S.center = rec_center(r, kind_mat, L.time, C.qc);  // moving_spheres.nim:61
          S.inv_r = a2.y;
          S.albedo = v3(a6.x, a6.y, a7.x);
          S.fuzz_or_ior = a7.y;
          S.mat_kind = mat_kind;
          sample_done = shade_hit(L, C.best_t, S, P.max_depth, &ud);
        } else {
          color = shade_miss(L, &ud);
          sample_done = true;
        }
      }
      need_setup = !sample_done;
      if (sample_done) {
        L.pix.x += color.x;  // render.nim:67
        L.pix.y += color.y;
        L.pix.z += color.z;
        ++L.sample;
        need_sample = true;
        if (L.sample >= P.spp) {
          if (P.cost) {
            P.cost[pid] = pix_seg;
          } else {
            double* out = P.pixels + 3ull * pid;  // the sum; draw_kernel applies canvas.nim:47-54 afterwards
            out[0] = L.pix.x;
            out[1] = L.pix.y;
            out[2] = L.pix.z;
          }
          need_pixel = true;
          need_sample = false;
          active = false;
        }
      }
    }
    __syncwarp();
    // ---- new work units
    if constexpr (!CHUNKED) {
      // One atomic per lane.  Serves the exact mode: the cost-ranked queue (order != NULL: the lane's dealt pixel
      // first, then the queue) and the row-major / scrambled pixel queue of renders without a cost pre-pass.
      if (need_pixel) {
        need_pixel = false;
        active = false;
        for (;;) {
          if (first_fetch) {  // dealt pixel of this lane, if any
            first_fetch = false;
            const uint32_t gid = deal_rank * 32u + (uint32_t)(tid & 31);
            pid = (deal_rank != 0xffffffffu && gid < first_wave) ? P.order[gid] : 0xffffffffu;
            if (pid == 0xffffffffu) continue;
          } else if (P.first_wave) {
            const unsigned long long slot = atomicAdd(P.work_counter, 1ull);
            if (slot >= queue_len) break;
            pid = P.order[first_wave + slot];
          } else {
            const unsigned long long slot = atomicAdd(P.work_counter, 1ull);
            if (slot >= total_units) break;
            pid = P.order ? P.order[slot] : unit_of_slot(slot);
          }
          if (begin_unit()) break;
        }
      }
    } else {
      // Split-stream mode: the WARP takes `chunk` consecutive queue slots at a time from the global counter and its
      // lanes consume them in order (warp_chunk[] in shared memory).  Consecutive slots are the sample ranges of one
      // pixel, then the next pixel of the row, so the lanes of a warp trace neighbouring rays: the same BVH subtrees,
      // the same few materials.  Warp-uniform loop: each pass hands one slot to every lane that still needs one.
      if (first_fetch) {  // cost-ranked order: the pixel dealt to this lane, if any
        first_fetch = false;
        const uint32_t gid = deal_rank * 32u + (uint32_t)(tid & 31);
        pid = (need_pixel && deal_rank != 0xffffffffu && gid < first_wave) ? P.order[gid] : 0xffffffffu;
        if (pid != 0xffffffffu) need_pixel = !begin_unit();
      }
      for (;;) {
        const unsigned want = __ballot_sync(0xffffffffu, need_pixel);
        if (!want) break;
        const int lane = tid & 31, leader = __ffs(want) - 1;
        unsigned long long base = 0;
        int take = 0;
        if (lane == leader) {
          unsigned long long next = wchunk[0], end = wchunk[1];
          if (next >= end) {  // chunk used up: take the next one (short chunks near the end of the queue)
            const unsigned long long head = *(volatile unsigned long long*)P.work_counter;
            unsigned long long ch = head + P.chunk_guard < queue_len ? (unsigned long long)P.chunk : 32ull;
            if (P.order && P.endgame_min_chunk) {
              // cost-ranked queue with about one pixel per lane left: hand out what remains in equal shares, so that
              // the render does not end on the warps that happened to get a whole chunk of 32 while others got none
              const unsigned long long left = head < queue_len ? queue_len - head : 0ull;
              const unsigned long long share = (left + grid_warps - 1ull) / grid_warps;
              if (share < ch) ch = share < P.endgame_min_chunk ? (unsigned long long)P.endgame_min_chunk : share;
            }
            next = atomicAdd(P.work_counter, ch);
            end = next + ch < queue_len ? next + ch : queue_len;
            if (next > end) next = end;
          }
          const unsigned long long avail = end - next;
          const int n = __popc(want);
          take = avail < (unsigned long long)n ? (int)avail : n;
          base = next;
          wchunk[0] = next + (unsigned long long)take;
          wchunk[1] = end;
        }
        __syncwarp();  // the next pass may have another leader: order the shared-memory update
        base = __shfl_sync(0xffffffffu, base, leader);
        take = __shfl_sync(0xffffffffu, take, leader);
        if (need_pixel) {
          const int rank = __popc(want & ((1u << lane) - 1u));
          if (take == 0) {  // a fresh chunk came back empty: the queue is exhausted
            need_pixel = false;
            active = false;
          } else if (rank < take) {
            const unsigned long long slot = base + (unsigned long long)rank;
            pid = P.order ? P.order[first_wave + slot] : unit_of_slot(slot);
            need_pixel = !begin_unit();  // an empty sample range (spp < ranges): take another slot in the next pass
          }
        }
      }
    }
    if (P.handoff) {
      // Late hand-off (BvhRenderParams::handoff).  Every warp watches the tail flag (lane 0 loads it, the value is
      // used one pass later so that nobody waits for the load); once it is up every lane parks its pixel at its next
      // sample boundary.
      if (!tail_on) {
        tail_on = __shfl_sync(0xffffffffu, tail_seen, 0) != 0;
        if ((tid & 31) == 0) tail_seen = *(volatile const unsigned int*)(P.deal_ticket + 4);
      }
      if (tail_on && active && need_sample && P.spp - L.sample >= (int32_t)P.handoff_min_left) {
        HandoffRec h;
        h.pid = pid;
        h.sample = L.sample;
        h.s0 = L.rng.s0;
        h.s1 = L.rng.s1;
        h.s2 = L.rng.s2;
        h.s3 = L.rng.s3;
        h.px = L.pix.x;
        h.py = L.pix.y;
        h.pz = L.pix.z;
        reinterpret_cast<HandoffRec*>(P.handoff)[P.handoff_cap_a + atomicAdd(P.deal_ticket + 6, 1u)] = h;
        active = false;
        need_sample = false;
      }
      if constexpr (CHUNKED) {
        // The queue slots this warp took as a chunk but has not handed to a lane yet would otherwise be lost when the
        // warp retires: park them as pixels that have not started.
        while (tail_on) {
          const unsigned long long next = wchunk[0], end = wchunk[1];
          if (next >= end) break;
          const unsigned long long slot = next + (unsigned long long)(tid & 31);
          __syncwarp();
          if ((tid & 31) == 0) wchunk[0] = next + 32ull < end ? next + 32ull : end;
          __syncwarp();
          if (slot < end) {
            HandoffRec h;
            h.pid = P.order ? P.order[first_wave + slot] : unit_of_slot(slot);
            const int32_t ri = (int32_t)(h.pid / (uint32_t)P.ncols);
            Rng g;
            rng_seed_pixel(g, P.row_begin + ri * P.row_step, (int32_t)(h.pid - (uint32_t)ri * (uint32_t)P.ncols), 0);
            h.sample = 0;
            h.s0 = g.s0;
            h.s1 = g.s1;
            h.s2 = g.s2;
            h.s3 = g.s3;
            h.px = h.py = h.pz = 0.0;
            reinterpret_cast<HandoffRec*>(P.handoff)[P.handoff_cap_a + atomicAdd(P.deal_ticket + 6, 1u)] = h;
          }
        }
      }
    }
    if (!__any_sync(0xffffffffu, active)) break;

    if (active && need_sample) {
      start_sample(L, P.cam, P.nrows, P.ncols);
      need_sample = false;
      need_setup = true;
      ++ray_count;
    }
    __syncwarp();
    if (active && need_setup) {
      need_setup = false;
      if (P.max_depth <= 0) {
        trav_done = true;
      } else {
        ray_setup();
        for (int32_t ri = bv.n_tree_objs; ri < bv.n_objects; ++ri) {  // objects without a finite box
          test_rec(ri);
          ++test_count;
        }
        cur = 0;
        sp = 0;
      }
    }

    // =================================================================== phase T
    __syncwarp();
    for (;;) {
      const bool trav = active && !trav_done;
      const int n_trav = __popc(__ballot_sync(0xffffffffu, trav));
      if (n_trav == 0) break;
      const int n_wait = __popc(__ballot_sync(0xffffffffu, active && trav_done));
      if (n_wait >= refill) break;
      // ---- inner nodes until a leaf (cur < 0) or the end of the traversal
      if (trav) {
        while (cur >= 0) {
          const int32_t ni = kNodeStride16 * cur;
          const float4 n0 = ld16f<(STAGE >= 1)>(nodes, nodes_sa, ni), n1 = ld16f<(STAGE >= 1)>(nodes, nodes_sa, ni + 1),
                       n2 = ld16f<(STAGE >= 1)>(nodes, nodes_sa, ni + 2);
          const int4 n3 = ld16i<(STAGE >= 1)>(nodes, nodes_sa, ni + 3);
          ++box_count;
          // child 0: lo = (n0.x, n0.y, n0.z), hi = (n0.w, n1.x, n1.y); child 1: lo = (n1.z, n1.w, n2.x), hi = (n2.y, n2.z, n2.w)
          float ax0 = fmaf(n0.x, R.idx, -R.oix), ax1 = fmaf(n0.w, R.idx, -R.oix);
          float ay0 = fmaf(n0.y, R.idy, -R.oiy), ay1 = fmaf(n1.x, R.idy, -R.oiy);
          float az0 = fmaf(n0.z, R.idz, -R.oiz), az1 = fmaf(n1.y, R.idz, -R.oiz);
          float near0 = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.f));
          float far0 = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fminf(fmaxf(az0, az1), C.best_f));
          float bx0 = fmaf(n1.z, R.idx, -R.oix), bx1 = fmaf(n2.y, R.idx, -R.oix);
          float by0 = fmaf(n1.w, R.idy, -R.oiy), by1 = fmaf(n2.z, R.idy, -R.oiy);
          float bz0 = fmaf(n2.x, R.idz, -R.oiz), bz1 = fmaf(n2.w, R.idz, -R.oiz);
          float near1 = fmaxf(fmaxf(fminf(bx0, bx1), fminf(by0, by1)), fmaxf(fminf(bz0, bz1), 0.f));
          float far1 = fminf(fminf(fmaxf(bx0, bx1), fmaxf(by0, by1)), fminf(fmaxf(bz0, bz1), C.best_f));
          const bool h0 = near0 <= far0, h1 = near1 <= far1;
          if (h0 && h1) {
            const bool first0 = near0 <= near1;
            stk[sp] = make_int2(first0 ? n3.y : n3.x, __float_as_int(first0 ? near1 : near0));
            ++sp;
            cur = first0 ? n3.x : n3.y;
          } else if (h0 || h1) {
            cur = h0 ? n3.x : n3.y;
          } else {
            // pop: skip subtrees that start beyond the closest root found since they were pushed
            cur = -1;
            trav_done = true;
            while (sp > 0) {
              --sp;
              const int2 e = stk[sp];
              if (__int_as_float(e.y) <= C.best_f) {
                cur = e.x;
                trav_done = false;
                break;
              }
            }
            if (trav_done) break;
          }
        }
      }
      __syncwarp();
      // ---- leaves: the reference's sphere test on each object, all lanes that hold a leaf in step
      const bool leaf = trav && !trav_done;
      const int32_t lv = ~cur;
      const int32_t first = lv >> 4;
      const int32_t cnt = leaf ? (lv & 15) : 0;
      const int32_t max_cnt = __reduce_max_sync(0xffffffffu, cnt);
      for (int32_t k = 0; k < max_cnt; ++k) {
        if (k < cnt) test_rec(first + k);
        __syncwarp();
      }
      if (leaf) {
        test_count += cnt;
        trav_done = true;
        while (sp > 0) {
          --sp;
          const int2 e = stk[sp];
          if (__int_as_float(e.y) <= C.best_f) {
            cur = e.x;
            trav_done = false;
            break;
          }
        }
      }
    }
    __syncwarp();
  }

  if (P.handoff && (tid & 31) == 0 && deal_rank != 0xffffffffu) {
    // this dealt warp is done: raise the tail flag once handoff_pct % of them are
    const unsigned int done = atomicAdd(P.deal_ticket + 3, 1u) + 1u;
    if ((unsigned long long)done * 100ull >= (unsigned long long)P.handoff_pct * (unsigned long long)n_deal_warps)
      *(volatile unsigned int*)(P.deal_ticket + 4) = 1u;
  }
  if (P.dbg_times && (tid & 31) == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    unsigned long long* e = P.dbg_times + 3ull * 4096ull + 2ull * (blockIdx.x * (BLOCK / 32) + (tid >> 5));
    e[0] = dbg_t0;
    e[1] = t1;
  }
  if (P.count_segments) {
    unsigned long long box_sum = box_count, test_sum = test_count;
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
      seg_count += __shfl_down_sync(0xffffffffu, seg_count, ofs);
      ray_count += __shfl_down_sync(0xffffffffu, ray_count, ofs);
      box_sum += __shfl_down_sync(0xffffffffu, box_sum, ofs);
      test_sum += __shfl_down_sync(0xffffffffu, test_sum, ofs);
    }
    if ((tid & 31) == 0) {
      atomicAdd(P.counters + 0, ray_count);
      atomicAdd(P.counters + 1, seg_count);
      atomicAdd(P.counters + 2, box_sum);
      atomicAdd(P.counters + 3, test_sum);
    }
  }
}

// ------------------------------------------------------------------- longest-pixel-first scheduling
// A pixel's samples are sequential (one RNG stream, render.nim:59-67), so the render cannot end before its most
// expensive pixel does (glass: up to 50 bounces per sample; measured ~95 ms at C2).  A cost pre-pass (the render
// kernel itself on the first few samples, counting segments) and this counting sort put expensive pixels at the
// head of the queue, so they start first and cheap pixels fill the tail.  The order cannot change the image.
static constexpr int kCostBuckets = 1024;

// Cost (bounce segments of the pre-pass samples) -> queue class.  coarse == 0: one class per cost value.  coarse != 0:
// classes of 4 below 64 and of 8 above, so that neighbouring pixels of the same kind (sky, ground, a diffuse sphere)
// share a class although their 8-sample costs differ by Monte-Carlo noise; inside a class the ordered scatter keeps the
// image order, which is what makes the lanes of a warp work on neighbouring pixels.
__device__ __forceinline__ uint32_t cost_class(uint32_t c, uint32_t coarse) {
  if (coarse) c = c < 64u ? c >> 2 : 16u + ((c - 64u) >> 3);
  return c < (uint32_t)kCostBuckets ? c : (uint32_t)kCostBuckets - 1u;
}

// The pre-pass costs are 8-sample estimates of a 500-sample total: a third of the truly expensive pixels come out at
// half their cost or less, and ONE expensive pixel that is ranked as cheap and left to a single lane sets the
// duration of the whole launch.  Expensive pixels come in patches (a glass sphere, an interreflecting corner), so the
// ranking uses  max(own estimate, mean of the 2*kCostWindow+1 estimates around it in its row):  the mean has a
// quarter of the noise, the max keeps isolated spikes and patch edges.
static constexpr int kCostWindow = 3;
__global__ void __launch_bounds__(256) cost_smooth_kernel(const uint32_t* __restrict__ cost, uint32_t* __restrict__ out,
                                                          uint32_t n, uint32_t ncols) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t col = i % ncols;
    const uint32_t lo = col >= (uint32_t)kCostWindow ? col - kCostWindow : 0u;
    const uint32_t hi = col + kCostWindow < ncols ? col + kCostWindow : ncols - 1u;
    uint32_t sum = 0;
    for (uint32_t c = lo; c <= hi; ++c) sum += cost[i - col + c];
    const uint32_t mean = sum / (hi - lo + 1u), own = cost[i];
    out[i] = own > mean ? own : mean;
  }
}

__global__ void __launch_bounds__(256) cost_histogram_kernel(const uint32_t* __restrict__ cost, uint32_t n,
                                                             uint32_t* __restrict__ hist, uint32_t coarse) {
  __shared__ uint32_t h[kCostBuckets];
  for (int i = threadIdx.x; i < kCostBuckets; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    atomicAdd(&h[cost_class(cost[i], coarse)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kCostBuckets; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
}

// Lower bound of the costs that fall into a class (inverse of cost_class).
__device__ __forceinline__ uint32_t class_floor(uint32_t cls, uint32_t coarse) {
  if (!coarse) return cls;
  return cls < 16u ? cls << 2 : 64u + ((cls - 16u) << 3);
}

// hist[b] := number of pixels in buckets above b (queue offset of bucket b, most expensive bucket first).
// Also decides how many of the most expensive pixels are traced warp-cooperatively (sched[0] = n_coop): a pixel
// qualifies when its pre-pass cost exceeds coop_alpha times the mean cost per lane of this launch — such a pixel
// could not finish within the time the bulk of the image takes even if its lane started at once — and at most
// coop_max pixels qualify (one per cooperative warp).  coop_max == 0 switches the mechanism off.
__global__ void __launch_bounds__(kCostBuckets) cost_offsets_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ sched,
                                                                    uint32_t coarse, unsigned long long lanes,
                                                                    float coop_alpha, uint32_t coop_max,
                                                                    uint32_t coop_force) {
  __shared__ uint32_t s[kCostBuckets];
  __shared__ unsigned long long total_cost;
  __shared__ uint32_t n_above;
  const int t = threadIdx.x;            // t = 0 is the most expensive bucket
  const int b = kCostBuckets - 1 - t;
  const uint32_t count = hist[b];
  if (t == 0) {
    total_cost = 0ull;
    n_above = 0u;
  }
  s[t] = count;
  __syncthreads();
  if (count) atomicAdd(&total_cost, (unsigned long long)count * (unsigned long long)(class_floor((uint32_t)b, coarse) + (coarse ? 2u : 0u)));
  for (int ofs = 1; ofs < kCostBuckets; ofs <<= 1) {  // inclusive scan
    uint32_t v = t >= ofs ? s[t - ofs] : 0u;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  hist[b] = s[t] - count;  // exclusive
  if (sched) {
    // mean cost per lane, in the pre-pass's units
    const float per_lane = (float)total_cost / (float)(lanes ? lanes : 1ull);
    if (count && (float)class_floor((uint32_t)b, coarse) > coop_alpha * per_lane) atomicAdd(&n_above, count);
    __syncthreads();
    if (t == 0) {
      uint32_t k = n_above < coop_max ? n_above : coop_max;
      if (coop_force != 0xffffffffu) k = coop_force < coop_max ? coop_force : coop_max;  // developer / test override
      const uint32_t n = s[kCostBuckets - 1];                                            // all pixels
      sched[0] = k < n ? k : n;
      sched[1] = n_above;
      sched[2] = (uint32_t)(total_cost > 0xffffffffull ? 0xffffffffull : total_cost);
    }
  }
}

// pos = rank of the pixel, most expensive first.  Ranks below n_coop = sched[0] go to the cooperative warps
// (coop_list[pos]).  The next ranks are dealt to the remaining warps' lanes, one pixel per lane.  Default
// (CoopLayout::sorted): warp w of the ranking gets the consecutive ranks 32w .. 32w+31 — equally expensive pixels,
// neighbours in the image inside a cost class, so its lanes finish together — and sits in CTA w mod CTAs, so every SM
// holds the same mix of warps.  Otherwise like cards, in tiers of `group` lanes: ranks 0 .. warps*group-1 go to lanes
// 0..group-1 of the warps (rank q -> warp q mod warps), the next warps*group ranks to lanes group..2*group-1, and so
// on: every warp starts with the same mix of costs.  The ranks after the first wave are queued in order.
// warps_all = warps of the render grid (0: no dealing, everything is queued).
struct RankLayout {
  uint32_t n_coop, warps, first_wave, tier, n_first, lanes, wpc, sorted;
};
__device__ __forceinline__ RankLayout rank_layout(const uint32_t* __restrict__ sched, uint32_t n, uint32_t warps_all,
                                                  uint32_t group, const CoopLayout& coop) {
  RankLayout r;
  r.n_coop = sched ? sched[0] : 0u;
  r.warps = warps_all ? (sched ? deal_warps(coop, r.n_coop) : warps_all) : 0u;  // warps that are dealt pixels
  r.first_wave = r.warps * 32u;
  r.tier = r.warps * group;
  const uint32_t n_ranked = n - r.n_coop;
  r.lanes = coop.lanes ? coop.lanes : 32u;
  r.wpc = coop.wpc ? coop.wpc : 1u;
  r.sorted = (coop.sorted && r.warps >= r.wpc && r.warps % r.wpc == 0) ? 1u : 0u;
  const uint32_t cap = r.warps * r.lanes;  // the dealt wave fills lanes 0 .. lanes-1 of every warp
  r.n_first = n_ranked < cap ? n_ranked : cap;
  return r;
}
__device__ __forceinline__ void place_rank(const RankLayout& r, uint32_t pos, uint32_t pixel, uint32_t group,
                                           uint32_t* __restrict__ order, uint32_t* __restrict__ coop_list) {
  if (pos < r.n_coop) {
    coop_list[pos] = pixel;
    return;
  }
  pos -= r.n_coop;
  uint32_t slot;
  if (pos < r.n_first && r.sorted) {
    // warp w of the ranking = `lanes` consecutive ranks; dealing rank = ticket of the CTA * wpc + warp inside the CTA
    // (render_bvh_kernel), so w -> CTA w mod ctas, warp w / ctas: equally expensive warps sit in different CTAs
    const uint32_t w = pos / r.lanes, l = pos - w * r.lanes;
    const uint32_t ctas = r.warps / r.wpc;
    slot = ((w % ctas) * r.wpc + w / ctas) * 32u + l;
  } else if (pos < r.n_first) {
    const uint32_t t = pos / r.tier, q = pos - t * r.tier;
    slot = (q % r.warps) * 32u + t * group + q / r.warps;
  } else {
    slot = r.first_wave + (pos - r.n_first);
  }
  order[slot] = pixel;
}

__global__ void __launch_bounds__(256) cost_scatter_kernel(const uint32_t* __restrict__ cost, uint32_t n,
                                                           uint32_t* __restrict__ offsets, uint32_t* __restrict__ order,
                                                           uint32_t warps_all, uint32_t group,
                                                           const uint32_t* __restrict__ sched,
                                                           uint32_t* __restrict__ coop_list, CoopLayout coop) {
  const RankLayout r = rank_layout(sched, n, warps_all, group, coop);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t c = cost[i];
    uint32_t pos = atomicAdd(&offsets[c < kCostBuckets ? c : kCostBuckets - 1], 1u);
    place_rank(r, pos, i, group, order, coop_list);
  }
}

// Split-stream mode: out[p*3 + ch] = sum over the 2^sub_log2 ranges of partial[(p << sub_log2 | j)*3 + ch], added
// pairwise with the lower index on the left — ((s0+s1)+(s2+s3))+... — which is exactly what the xor butterfly gives
// the lowest lane of every group (IEEE addition is commutative, so both partners of a step hold the same bits).
// One lane per (pixel, channel, range); 32 >> sub_log2 pixel-channels per warp.
__global__ void __launch_bounds__(256) substream_reduce_kernel(const double* __restrict__ partial,
                                                               double* __restrict__ out, unsigned long long n_pc,
                                                               uint32_t sub_log2) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nsub = 1u << sub_log2;
  const unsigned long long pc = t >> sub_log2;
  const uint32_t j = (uint32_t)t & (nsub - 1u);
  double v = 0.0;
  if (pc < n_pc) {
    const unsigned long long p = pc / 3ull;
    const uint32_t ch = (uint32_t)(pc - p * 3ull);
    v = partial[((p << sub_log2) + j) * 3ull + ch];
  }
  for (uint32_t ofs = 1; ofs < nsub; ofs <<= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, ofs);
    v = (j & ofs) ? w + v : v + w;
  }
  if (pc < n_pc && j == 0) out[pc] = v;
}

// The same scatter keeping the image order inside a class (approximately: exactly inside a warp's 32 consecutive
// pixels, and window by window of gridDim * blockDim pixels, because the few blocks of this launch walk the image in
// step).  Lanes of a warp with the same class take consecutive ranks with one atomic (warp-aggregated).
__global__ void __launch_bounds__(1024) cost_scatter_ordered_kernel(const uint32_t* __restrict__ cost, uint32_t n,
                                                                    uint32_t* __restrict__ offsets,
                                                                    uint32_t* __restrict__ order, uint32_t warps_all,
                                                                    uint32_t group, uint32_t coarse,
                                                                    const uint32_t* __restrict__ sched,
                                                                    uint32_t* __restrict__ coop_list, CoopLayout coop) {
  const RankLayout r = rank_layout(sched, n, warps_all, group, coop);
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    const bool in = i < n;
    const uint32_t cls = in ? cost_class(cost[i], coarse) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, cls);
    const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
    uint32_t base = 0;
    if (in && lane == leader) base = atomicAdd(&offsets[cls], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (in) place_rank(r, base + (uint32_t)__popc(peers & ((1u << lane) - 1u)), i, group, order, coop_list);
    __syncthreads();  // keeps the block's warps within one window of the image
  }
}

// canvas.nim:47-54 `draw` over the sums the render kernel left in the framebuffer: one thread per channel,
// pixels[i] = pow(scale * pixels[i], gamma).  A separate pass so that the three pow() per pixel (double-double
// log/exp, ~2000 instructions) run on full warps instead of lane by lane inside the render loop.
__global__ void __launch_bounds__(256) draw_kernel(double* __restrict__ pixels, unsigned long long n, double inv_spp,
                                                   double inv_gamma) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pixels[i] = detmath::pow(inv_spp * pixels[i], inv_gamma);
}

// io/ppm.nim:14-27 quantisation on the device: int(256 * clamp(c, 0, 0.999)) per channel (NaN -> 0, as the host
// helper), packed RGB8 in PPM row order (canvas row nrows-1 first).  One thread per pixel of a full canvas.
__global__ void __launch_bounds__(256) quantise_rgb8_kernel(const double* __restrict__ pixels, int32_t nrows,
                                                            int32_t ncols, uint8_t* __restrict__ rgb8) {
  const unsigned long long n = (unsigned long long)nrows * (unsigned long long)ncols;
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t row = (int32_t)(i / (unsigned long long)ncols);
  const int32_t col = (int32_t)(i - (unsigned long long)row * (unsigned long long)ncols);
  uint8_t* out = rgb8 + 3ull * ((unsigned long long)(nrows - 1 - row) * (unsigned long long)ncols + col);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double c = pixels[3ull * i + k];
    const double cl = c < 0.0 ? 0.0 : (c > 0.999 ? 0.999 : c);  // safe_math.nim:10-14
    out[k] = c != c ? (uint8_t)0 : (uint8_t)(int)(256 * cl);    // ppm.nim:15-16 truncates toward zero
  }
}

// io/rgb.nim:17-31 `toRGB_Raw` + io/color_conversions.nim:180-252 `rgbRaw_to_ycbcr420` (BT.601, video range, 4:2:0)
// fused over the drawn canvas: one thread per 2x2 block quantises its four pixels (uint8(256 * clamp(c, 0, 0.999))),
// forms the reference's 8-bit fixed-point luma per pixel and the chroma of the block from the four (B - Y'), (R - Y')
// differences.  All integer arithmetic below stays inside the reference's uint16 / int16 ranges (|sums| <= 4 * 255,
// products <= 255 * 160 in magnitude only where the reference's int16 does not overflow either: |R - Y'| <= 179), so
// plain int gives the same bits.  `>>` on negative values is an arithmetic shift in CUDA as in Nim.
// Output planes are top row first.  as_written != 0 reproduces rgb.nim:29-31 literally: output row i shows canvas row
// nrows - i (one row off; output row 0, which the reference reads past the end of the buffer, is defined as 0).
// Coefficients: color_conversions.nim:104-114,178 evaluated for BT.601: kr 77, kg 150, kb 29, fb 127, fr 160,
// y_scale 110 (7 bits), y_min 16.
__device__ __forceinline__ int rgb_level(double c) {  // rgb.nim:20-21 (NaN -> 0)
  const double cl = c < 0.0 ? 0.0 : (c > 0.999 ? 0.999 : c);
  return c != c ? 0 : (int)(256 * cl);
}

__global__ void __launch_bounds__(256) ycbcr420_kernel(const double* __restrict__ pixels, int32_t nrows, int32_t ncols,
                                                       int32_t as_written, uint8_t* __restrict__ Y,
                                                       uint8_t* __restrict__ Cb, uint8_t* __restrict__ Cr) {
  const int32_t bw = ncols >> 1, bh = nrows >> 1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)bw * bh) return;
  const int32_t by = (int32_t)(t / bw), bx = (int32_t)(t - (long long)by * bw);
  int tU = 0, tV = 0;
#pragma unroll
  for (int di = 0; di < 2; ++di) {
    const int32_t i = 2 * by + di;  // output row, top first
    const int32_t src = as_written ? nrows - i : nrows - 1 - i;
#pragma unroll
    for (int dj = 0; dj < 2; ++dj) {
      const int32_t j = 2 * bx + dj;
      int r = 0, g = 0, b = 0;
      if (src < nrows) {
        const double* p = pixels + 3ll * ((long long)src * ncols + j);
        r = rgb_level(p[0]);
        g = rgb_level(p[1]);
        b = rgb_level(p[2]);
      }
      const int tY = (77 * r + 150 * g + 29 * b) >> 8;
      tU += b - tY;
      tV += r - tY;
      Y[(long long)i * ncols + j] = (uint8_t)(((tY * 110) >> 7) + 16);
    }
  }
  Cb[(long long)by * bw + bx] = (uint8_t)((((tU >> 2) * 127) >> 8) + 128);
  Cr[(long long)by * bw + bx] = (uint8_t)((((tV >> 2) * 160) >> 8) + 128);
}

}  // namespace tor
