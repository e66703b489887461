// tor_api.cu — the C ABI of include/tor_b200.h over the sm_100a render kernel.
//
// Replaces `proc render*(canvas: var Canvas, cam: Camera, world: HittableList, max_depth: int)`
// (trace_of_radiance/render.nim:49).  There is deliberately NO CPU path in this file: without a
// CUDA device every compute entry point fails with TOR_ERR_NO_DEVICE / TOR_ERR_CUDA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/tor_b200.h"
#include "tor_anim.hpp"
#include "tor_bvh.hpp"
#include "tor_kernels.cuh"
#include "tor_kernels_anim.cuh"
#include "tor_kernels_bvh.cuh"
#include "tor_scene_pack.hpp"

namespace {

constexpr int kBlock = 256;

// Developer tuning knobs of the BVH route, read from the environment ONCE, when a context is created (DESIGN.md
// lists them).  None of them can change an image: they only move work between lanes, warps and launches.
struct Tuning {
  int block = kBlock;          // TOR_BVH_BLOCK: CTA size, 256 (two CTAs per SM) or 512 (one)
  int refill = 20;             // TOR_BVH_REFILL: waiting lanes of a warp that trigger a shade phase (exact mode)
  int refill_fast = 20;        // TOR_BVH_REFILL_FAST: the same in split-stream mode
  int chunk = 0;               // TOR_BVH_CHUNK: queue slots per warp-level fetch (0 = default rule)
  int lanes = 32;              // TOR_BVH_LANES: lanes of each warp that take pixels
  bool deal = true;            // TOR_BVH_EXACT_DEAL=0: queue every ranked pixel, no dealt first wave
  int deal_group = 8;          // TOR_BVH_DEAL_GROUP: lanes per cost tier of the dealt wave
  bool queue_percost = false;  // TOR_BVH_EXACT_QUEUE=percost: one class per cost value, one atomic per lane
  bool no_scramble = false;    // TOR_BVH_NO_SCRAMBLE: row-major queue for renders without a cost pre-pass
  bool fast_scramble = false;  // TOR_BVH_FAST_SCRAMBLE: scattered queue in split-stream mode
  float coop_alpha = 0.7f;     // TOR_BVH_COOP_ALPHA: a pixel is traced by a whole warp when its pre-pass cost
                               //   exceeds alpha x the mean cost per lane of the launch
  int coop_max_pct = 15;       // TOR_BVH_COOP_MAX: at most this percentage of the SMs is set aside for cooperative
                               //   pixels, 8 per SM (0 switches the mechanism off)
  long long coop_force = -1;   // TOR_BVH_COOP_FORCE: trace exactly this many of the most expensive pixels
                               //   cooperatively (tests), still capped by coop_max_pct
  int prepass_min_spp = 256;   // TOR_BVH_PREPASS_SPP: samples per pixel from which the cost pre-pass always runs
  int prepass_full_spp = 64;   // TOR_BVH_PREPASS_FULL_SPP: ... and from which it runs when the launch fills the GPU
  int coop_px_per_lane = 8;    // TOR_BVH_COOP_PXLANE: cooperative pixels only when the launch has fewer pixels per lane
  int deal_sorted = 2;         // TOR_BVH_DEAL_SORTED: dealt wave by consecutive ranks per warp: 1 = launches with few pixels per lane, 2 = always
  int handoff_pct = 75;        // TOR_BVH_HANDOFF: late hand-off once this % of the dealt lane warps are done (0 = off)
  int handoff_min_left = 8;    // TOR_BVH_HANDOFF_MIN_LEFT: pixels with fewer samples left stay where they are
  int handoff_warps = 16;      // TOR_BVH_HANDOFF_WARPS: working warps per CTA of the second cooperative launch
  int handoff_plain = 30;      // TOR_BVH_HANDOFF_PLAIN: the same in renders without a cost pre-pass (few samples per pixel): % of all warps
  int handoff_plain_px = 64;   // TOR_BVH_HANDOFF_PLAIN_PXLANE: ... only with fewer pixels per lane than this
  int handoff_all = 0;         // TOR_BVH_HANDOFF_ALL: also in launches without cooperative CTAs (many pixels per lane)
  int thin_px_per_lane = 0;    // TOR_BVH_THIN_PXLANE: launches with fewer pixels per lane than this / 100 run `thin_lanes`
  int thin_lanes = 16;         // TOR_BVH_THIN_LANES: lanes per warp of the dealt wave (0 = never)
  int coop_fast_pct = 0;       // TOR_BVH_COOP_FAST: this percentage of the cooperative CTAs runs only coop_fast_warps warps
  int coop_fast_warps = 8;     // TOR_BVH_COOP_FAST_WARPS: ... and starts on the most expensive pixels of the launch (off by
                               //   default: measured 46-57 ms against 41 on a GPU's share of C2 on 8 GPUs — the set-aside
                               //   SMs are throughput-bound, not bound by their longest chain)
  int coop_warps = 16;         // TOR_BVH_COOP_WARPS: working warps per cooperative CTA (1 .. 16)
  int endgame_min_chunk = 4;   // TOR_BVH_ENDGAME: smallest share of the cost-ranked queue a warp takes near the end (0: off)
  int coop_queue_factor = 4;   // TOR_BVH_COOP_QUEUE: at most this many cooperative pixels per cooperative warp
  bool debug_times = false;    // TOR_BVH_DEBUG_TIMES: record start / end stamps of every warp (tor_debug_times)
  int anim_grid_divisor = 0;   // TOR_ANIM_GRID_DIV: share of the GPU a frame in flight takes, as a divisor (0 = in_flight / 2)
  int stage_max = 2;           // TOR_BVH_STAGE: 2 = stage nodes + records in shared memory when they fit, else nothing
                               //   (default); 1 = nodes only; 0 = nothing (everything through L1 / L2)

  static Tuning from_env() {
    Tuning t;
    auto geti = [](const char* name, int dflt) {
      const char* e = getenv(name);
      return e ? atoi(e) : dflt;
    };
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    t.block = geti("TOR_BVH_BLOCK", kBlock) == 512 ? 512 : kBlock;
    t.refill = clampi(geti("TOR_BVH_REFILL", 20), 1, 32);
    t.refill_fast = clampi(geti("TOR_BVH_REFILL_FAST", 20), 1, 32);
    {
      int c = geti("TOR_BVH_CHUNK", 0);
      t.chunk = c <= 0 ? 0 : (c < 32 ? 32 : (c > 4096 ? 4096 : c / 32 * 32));
    }
    t.lanes = clampi(geti("TOR_BVH_LANES", 32), 1, 32);
    t.deal = geti("TOR_BVH_EXACT_DEAL", 1) != 0;
    {
      int g = geti("TOR_BVH_DEAL_GROUP", 8);
      t.deal_group = (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) ? g : 8;
    }
    {
      const char* e = getenv("TOR_BVH_EXACT_QUEUE");
      t.queue_percost = e && strcmp(e, "percost") == 0;
    }
    t.no_scramble = getenv("TOR_BVH_NO_SCRAMBLE") != nullptr;
    t.fast_scramble = getenv("TOR_BVH_FAST_SCRAMBLE") != nullptr;
    if (const char* e = getenv("TOR_BVH_COOP_ALPHA")) t.coop_alpha = (float)atof(e);
    if (!(t.coop_alpha > 0.f)) t.coop_alpha = 1.0f;
    t.coop_max_pct = clampi(geti("TOR_BVH_COOP_MAX", 15), 0, 50);
    if (const char* e = getenv("TOR_BVH_COOP_FORCE")) t.coop_force = atoll(e);
    t.prepass_min_spp = clampi(geti("TOR_BVH_PREPASS_SPP", 256), 9, 1 << 30);
    t.prepass_full_spp = clampi(geti("TOR_BVH_PREPASS_FULL_SPP", 64), 9, 1 << 30);
    t.coop_px_per_lane = clampi(geti("TOR_BVH_COOP_PXLANE", 8), 0, 1 << 20);
    t.coop_queue_factor = clampi(geti("TOR_BVH_COOP_QUEUE", 4), 1, 64);
    t.endgame_min_chunk = clampi(geti("TOR_BVH_ENDGAME", 4), 0, 32);
    t.coop_warps = clampi(geti("TOR_BVH_COOP_WARPS", (int)tor::kCoopWarps), 1, 16);
    t.coop_fast_pct = clampi(geti("TOR_BVH_COOP_FAST", 0), 0, 100);
    t.coop_fast_warps = clampi(geti("TOR_BVH_COOP_FAST_WARPS", 8), 1, 16);
    t.deal_sorted = clampi(geti("TOR_BVH_DEAL_SORTED", 2), 0, 2);
    t.handoff_pct = clampi(geti("TOR_BVH_HANDOFF", 75), 0, 100);
    t.handoff_min_left = clampi(geti("TOR_BVH_HANDOFF_MIN_LEFT", 8), 1, 1 << 20);
    t.handoff_warps = clampi(geti("TOR_BVH_HANDOFF_WARPS", 16), 1, 16);
    t.handoff_plain = clampi(geti("TOR_BVH_HANDOFF_PLAIN", 30), 0, 100);
    t.handoff_plain_px = clampi(geti("TOR_BVH_HANDOFF_PLAIN_PXLANE", 64), 0, 1 << 20);
    t.handoff_all = clampi(geti("TOR_BVH_HANDOFF_ALL", 0), 0, 1);
    t.thin_px_per_lane = clampi(geti("TOR_BVH_THIN_PXLANE", 0), 0, 100000);
    t.thin_lanes = clampi(geti("TOR_BVH_THIN_LANES", 16), 1, 32);
    t.stage_max = clampi(geti("TOR_BVH_STAGE", 2), 0, 2);
    t.anim_grid_divisor = clampi(geti("TOR_ANIM_GRID_DIV", 0), 0, 64);
    t.debug_times = getenv("TOR_BVH_DEBUG_TIMES") != nullptr;
    return t;
  }
};

struct DeviceState {
  int dev = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  int max_smem_per_sm = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_busy = nullptr;  // end of the last launch that used this device's scratch buffers (any stream)
  cudaStream_t stream_coop = nullptr;  // render_coop_kernel runs beside the lane kernel, at a higher priority
  cudaEvent_t ev_ranked = nullptr, ev_coop_done = nullptr;
  unsigned long long* d_dbg = nullptr;  // TOR_BVH_DEBUG_TIMES: %globaltimer stamps of the last exact-mode main launch
  uint8_t* d_handoff = nullptr;      // parked pixels of the late hand-off (BvhRenderParams::handoff)
  size_t handoff_cap = 0;
  bool handoff_used = false;         // the last launch ran with the hand-off (its counts are in d_ticket[5..6])
  unsigned int* d_ticket = nullptr;  // [0] arrival counter of the lane kernel's CTAs (BvhRenderParams::deal_ticket),
                                     // [1] cooperative CTAs resident (coop_gate_kernel)
  bool busy = false;
  uint8_t* d_blob = nullptr;  // BVH blob (tor_bvh.hpp)
  size_t blob_cap = 0;
  uint8_t* d_brute = nullptr;  // brute-force route's scene blob (tor_scene_pack.hpp), uploaded on first use
  size_t brute_cap = 0;
  bool brute_current = false;
  double* d_pixels = nullptr;
  size_t pix_cap = 0;  // bytes
  unsigned long long* d_work = nullptr;      // pixel queue heads: [0] render, [1] cost pre-pass
  uint32_t* d_cost = nullptr;                // per-pixel cost of the pre-pass, then the bucket offsets
  uint32_t* d_order = nullptr;               // pixel queue order (most expensive first)
  uint32_t* d_hist = nullptr;                // kCostBuckets counters
  uint32_t* d_sched = nullptr;               // [0] cooperative pixels of the launch (BvhRenderParams::sched)
  uint32_t* d_coop = nullptr;                // the cooperative pixels, most expensive first
  size_t coop_cap = 0;                       // entries
  double* d_partial = nullptr;               // split-stream mode: one partial sum (3 doubles) per (pixel, range)
  size_t partial_cap = 0;                    // bytes
  uint8_t* d_rgb8 = nullptr;                 // packed RGB8 image of tor_render_rgb8
  size_t rgb8_cap = 0;
  size_t order_cap = 0;                      // pixels
  unsigned long long* d_counters = nullptr;  // [0] primary rays, [1] segments, [2] box-pair tests, [3] exact tests
  bool scene_current = false;
  bool timed = false;
};

}  // namespace

struct tor_ctx {
  std::vector<DeviceState> devs;
  std::string err;
  std::vector<tor_hittable> objs;  // the scene as decoded flat records (host copy; the caller's buffer is not kept)
  tor::PackedScene scene;          // brute-force route only: packed on first use (ensure_brute_scene)
  bool brute_packed = false;
  tor::PackedBvh bvh;
  tor_camera cam;
  bool have_scene = false;
  int64_t launches = 0;
  uint64_t counters[3] = {0, 0, 0};
  uint64_t trav_counters[2] = {0, 0};
  bool counters_pending = false;
  Tuning tune;
  // > 1: this context is one of several that render concurrently on the same GPU (frames of an animation in flight):
  // its launches take only 1/grid_divisor of the persistent grid, so that the frames' kernels are resident together
  // and the lanes a frame's cheap pixels free up early are not held idle until its slowest pixel is done.
  int grid_divisor = 1;
};

namespace {

thread_local std::string g_create_err;

cudaError_t create_priority_stream(cudaStream_t* s) {
  int lo = 0, hi = 0;  // numerically lowest value = highest priority
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e != cudaSuccess) return e;
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, hi);
}

int fail(tor_ctx* ctx, int code, const std::string& msg) {
  if (ctx)
    ctx->err = msg;
  else
    g_create_err = msg;
  return code;
}

#define TOR_CUDA(ctx, call)                                                                      \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(ctx, TOR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

using KernelFn = void (*)(const tor::RenderParams);
using BvhKernelFn = void (*)(const tor::BvhRenderParams);

struct LaunchPlan {
  KernelFn fn;
  int stage;
  size_t smem;
};
struct BvhLaunchPlan {
  BvhKernelFn fn;
  int stage;
  size_t smem;
  int block;
};

LaunchPlan plan_for(const tor::SceneView& sv, int max_smem_optin) {
  const size_t cand = (size_t)tor::kMaxCand * kBlock * sizeof(uint16_t);
  // two CTAs per SM want <= ~113 KB each; a bigger scene simply runs one CTA per SM
  if (sv.total_bytes + cand <= (size_t)max_smem_optin)
    return LaunchPlan{tor::render_exact_kernel<kBlock, 2>, 2, sv.total_bytes + cand};
  if (sv.hot_bytes + cand <= (size_t)max_smem_optin)
    return LaunchPlan{tor::render_exact_kernel<kBlock, 1>, 1, sv.hot_bytes + cand};
  return LaunchPlan{tor::render_exact_kernel<kBlock, 0>, 0, cand};
}

// The hierarchy is staged in shared memory only while two CTAs still fit on an SM (the traversal is
// latency-bound and wants the warps); beyond that the nodes and box tables, then nothing, and L1/L2 serve the rest.
template <int B, bool CHUNKED>
BvhLaunchPlan bvh_plan_b(const tor::BvhView& bv, size_t budget, int stage_max) {
  const tor::StagePlan s2 = tor::stage_plan<2>(bv), s1 = tor::stage_plan<1>(bv);
  if (stage_max >= 2 && (size_t)s2.bytes0 + s2.bytes1 <= budget)
    return BvhLaunchPlan{tor::render_bvh_kernel<B, 2, CHUNKED>, 2, (size_t)s2.bytes0 + s2.bytes1, B};
  // Nodes alone in shared memory (STAGE 1) is only taken on request: the records then come through an L1 that the
  // shared-memory carve-out has shrunk to ~28 KB, next to the traversal stacks; measured on 1 938 objects: 68 ms
  // against 59 ms with nothing staged and the whole 256 KB as L1.
  if (stage_max == 1 && (size_t)s1.bytes0 + s1.bytes1 <= budget)
    return BvhLaunchPlan{tor::render_bvh_kernel<B, 1, CHUNKED>, 1, (size_t)s1.bytes0 + s1.bytes1, B};
  return BvhLaunchPlan{tor::render_bvh_kernel<B, 0, CHUNKED>, 0, 0, B};
}

// chunked: the warp-level queue (tor_kernels_bvh.cuh); otherwise one atomic per lane.
BvhLaunchPlan bvh_plan_for(const tor::BvhView& bv, int max_smem_per_sm, int max_smem_optin, bool chunked, int block,
                           int stage_max) {
  // per CTA: the dynamic part + ~1.4 KB of static shared memory + 1 KB the system reserves
  const size_t half = (size_t)max_smem_per_sm / 2 - 4096;
  const size_t whole = (size_t)max_smem_optin - 2048;
  if (block == 512) return chunked ? bvh_plan_b<512, true>(bv, whole, stage_max) : bvh_plan_b<512, false>(bv, whole, stage_max);
  return chunked ? bvh_plan_b<kBlock, true>(bv, half, stage_max) : bvh_plan_b<kBlock, false>(bv, half, stage_max);
}

// A multiplier m coprime to n with m/n near the golden ratio: i -> i*m mod n is a bijection of [0, n) that sends
// neighbours far apart.
uint32_t coprime_near_golden(uint32_t n) {
  if (n < 3) return 1;
  unsigned long long m = (unsigned long long)(0.6180339887 * (double)n) | 1ull;
  auto gcd = [](unsigned long long a, unsigned long long b) {
    while (b) {
      unsigned long long t = a % b;
      a = b;
      b = t;
    }
    return a;
  };
  while (gcd(m, n) != 1) m += 2;
  return (uint32_t)(m % n);
}

size_t device_blob_bytes(const tor_ctx* ctx) { return ctx->bvh.blob.size(); }

int ensure_capacity(tor_ctx* ctx, DeviceState& d, size_t blob_bytes, size_t pix_bytes) {
  TOR_CUDA(ctx, cudaSetDevice(d.dev));
  if (blob_bytes > d.blob_cap) {
    if (d.d_blob) cudaFree(d.d_blob);
    d.d_blob = nullptr;
    d.blob_cap = 0;
    TOR_CUDA(ctx, cudaMalloc(&d.d_blob, blob_bytes));
    d.blob_cap = blob_bytes;
    d.scene_current = false;
  }
  if (pix_bytes > d.pix_cap) {
    if (d.d_pixels) cudaFree(d.d_pixels);
    d.d_pixels = nullptr;
    d.pix_cap = 0;
    TOR_CUDA(ctx, cudaMalloc(&d.d_pixels, pix_bytes));
    d.pix_cap = pix_bytes;
  }
  return TOR_OK;
}

// Renders launched on a caller's stream (tor_render_device_async) may still read this device's scene blobs and
// scratch arrays: wait for the last one before anything is overwritten from the host side.
int wait_idle(tor_ctx* ctx, DeviceState& d) {
  if (d.busy) {
    TOR_CUDA(ctx, cudaSetDevice(d.dev));
    TOR_CUDA(ctx, cudaEventSynchronize(d.ev_busy));
    d.busy = false;
  }
  return TOR_OK;
}

int upload_scene_to(tor_ctx* ctx, DeviceState& d) {
  if (d.scene_current) return TOR_OK;
  int rc = wait_idle(ctx, d);
  if (rc) return rc;
  rc = ensure_capacity(ctx, d, device_blob_bytes(ctx), 0);
  if (rc) return rc;
  TOR_CUDA(ctx, cudaMemcpyAsync(d.d_blob, ctx->bvh.blob.data(), ctx->bvh.blob.size(), cudaMemcpyHostToDevice, d.stream));
  // the blob vector may be rebuilt by the next tor_scene_upload: finish the copy before returning
  TOR_CUDA(ctx, cudaStreamSynchronize(d.stream));
  d.scene_current = true;
  return TOR_OK;
}

// The brute-force route's scene blob (filter records, tor_scene_pack.hpp) is built and uploaded on first use only:
// the default BVH route never pays for it (an animation re-sets the scene every frame).
int upload_brute_scene_to(tor_ctx* ctx, DeviceState& d) {
  if (!ctx->brute_packed) {
    if (ctx->objs.size() > 65535)
      return fail(ctx, TOR_ERR_SCENE_TOO_LARGE, "TOR_FLAG_BRUTE_FORCE: more than 65535 objects (16-bit indices)");
    std::string err;
    if (!tor::pack_scene(ctx->objs.data(), (int64_t)ctx->objs.size(), TOR_STRIDE_FLAT, ctx->cam.shutter_open,
                         ctx->cam.shutter_close, &ctx->scene, &err))
      return fail(ctx, TOR_ERR_INVALID_ARG, err);
    ctx->brute_packed = true;
    for (DeviceState& o : ctx->devs) o.brute_current = false;
  }
  if (d.brute_current) return TOR_OK;
  int rc = wait_idle(ctx, d);
  if (rc) return rc;
  TOR_CUDA(ctx, cudaSetDevice(d.dev));
  if (ctx->scene.blob.size() > d.brute_cap) {
    if (d.d_brute) cudaFree(d.d_brute);
    d.d_brute = nullptr;
    d.brute_cap = 0;
    TOR_CUDA(ctx, cudaMalloc(&d.d_brute, ctx->scene.blob.size()));
    d.brute_cap = ctx->scene.blob.size();
  }
  TOR_CUDA(ctx, cudaMemcpyAsync(d.d_brute, ctx->scene.blob.data(), ctx->scene.blob.size(), cudaMemcpyHostToDevice,
                                d.stream));
  TOR_CUDA(ctx, cudaStreamSynchronize(d.stream));
  d.brute_current = true;
  return TOR_OK;
}

int set_scene(tor_ctx* ctx, const tor_camera* cam, const void* objects, int64_t len, int64_t stride) {
  if (!cam) return fail(ctx, TOR_ERR_INVALID_ARG, "camera is NULL");
  if (!objects || len <= 0)
    return fail(ctx, TOR_ERR_INVALID_ARG, "empty HittableList (hittables_lists.nim:42 asserts len > 0)");
  if (stride != TOR_STRIDE_FLAT && stride != TOR_STRIDE_NIM_VARIANT)
    return fail(ctx, TOR_ERR_LAYOUT, "stride must be 112 (tor_hittable) or 120 (Nim HittableVariant)");
  if (len > tor::kBvhMaxObjects) return fail(ctx, TOR_ERR_SCENE_TOO_LARGE, "more than 16777216 objects");
  ctx->have_scene = false;
  ctx->brute_packed = false;
  ctx->objs.resize((size_t)len);
  for (int64_t i = 0; i < len; ++i)
    if (!tor::decode_object((const uint8_t*)objects + i * stride, stride, &ctx->objs[(size_t)i]))
      return fail(ctx, TOR_ERR_INVALID_ARG, "object " + std::to_string(i) + ": unknown kind / material tag");
  std::string err;
  if (!tor::pack_bvh(ctx->objs, *cam, &ctx->bvh, &err)) return fail(ctx, TOR_ERR_INVALID_ARG, err);
  ctx->cam = *cam;
  ctx->have_scene = true;
  for (DeviceState& d : ctx->devs) d.scene_current = d.brute_current = false;
  return TOR_OK;
}

int check_canvas_dims(tor_ctx* ctx, int32_t nrows, int32_t ncols, int32_t spp, int64_t max_depth, int32_t row_begin,
                      int32_t row_end, int32_t row_step) {
  if (nrows <= 0 || ncols <= 0) return fail(ctx, TOR_ERR_INVALID_ARG, "canvas has no pixels");
  if (spp < 0) return fail(ctx, TOR_ERR_INVALID_ARG, "samples_per_pixel < 0");
  if (max_depth < 0 || max_depth > 0x7fffffff) return fail(ctx, TOR_ERR_INVALID_ARG, "max_depth out of range");
  if (row_step <= 0 || row_begin < 0 || row_end > nrows)
    return fail(ctx, TOR_ERR_INVALID_ARG, "row range outside the canvas");
  return TOR_OK;
}

// TOR_MODE_FAST: number of sample ranges per pixel as log2 (0 = exact mode).  The count depends on the flags and on the
// FULL canvas only, never on the row selection or the device, so any partition of a canvas gives the same image.
int substream_log2(tor_ctx* ctx, uint32_t flags, int32_t nrows, int32_t ncols, int32_t spp, uint32_t* out) {
  *out = 0;
  if (!(flags & TOR_MODE_FAST)) {
    if (flags & 0x00ff0000u) return fail(ctx, TOR_ERR_INVALID_ARG, "TOR_FAST_SUBSTREAMS without TOR_MODE_FAST");
    return TOR_OK;
  }
  if (flags & TOR_FLAG_BRUTE_FORCE)
    return fail(ctx, TOR_ERR_INVALID_ARG, "TOR_MODE_FAST is implemented on the BVH route only");
  uint32_t n = (flags >> 16) & 0xffu;
  if (n == 0) {  // auto: about 2^24 work units per canvas, at most one range per sample, at most 32
    const unsigned long long px = (unsigned long long)nrows * (unsigned long long)ncols;
    unsigned long long want = px ? (1ull << 24) / px : 1ull;
    if (want > 32) want = 32;
    if (want > (unsigned long long)(spp > 0 ? spp : 1)) want = (unsigned long long)(spp > 0 ? spp : 1);
    n = 1;
    while (2ull * n <= want) n *= 2;
  }
  if (n > 32 || (n & (n - 1))) return fail(ctx, TOR_ERR_INVALID_ARG, "TOR_FAST_SUBSTREAMS must be a power of two <= 32");
  uint32_t lg = 0;
  while ((1u << lg) < n) ++lg;
  *out = lg;
  return TOR_OK;
}

// Enqueue one render of rows row_begin, row_begin+row_step, ... < row_end on device `d` into d_out.
int launch_rows(tor_ctx* ctx, DeviceState& d, double* d_out, int32_t nrows, int32_t ncols, int32_t spp, float gamma,
                int64_t max_depth, uint32_t flags, int32_t row_begin, int32_t row_end, int32_t row_step,
                cudaStream_t stream, bool timed) {
  const int32_t nsel = row_end > row_begin ? (row_end - row_begin + row_step - 1) / row_step : 0;
  d.timed = false;
  d.handoff_used = false;
  if (nsel == 0) return TOR_OK;
  TOR_CUDA(ctx, cudaSetDevice(d.dev));

  const bool count = (flags & TOR_FLAG_COUNT_SEGMENTS) != 0;
  const unsigned long long total_px = (unsigned long long)nsel * (unsigned long long)ncols;
  uint32_t sub_log2 = 0;
  {
    int rc = substream_log2(ctx, flags, nrows, ncols, spp, &sub_log2);
    if (rc) return rc;
    if (sub_log2 && (total_px << sub_log2) >= 0xffffffffull)
      return fail(ctx, TOR_ERR_INVALID_ARG, "TOR_MODE_FAST: more than 2^32 (pixel, range) units in one launch");
  }
  const unsigned long long want = (total_px + kBlock - 1) / kBlock;
  if (d.busy) TOR_CUDA(ctx, cudaStreamWaitEvent(stream, d.ev_busy, 0));  // previous launch may be on another stream
  TOR_CUDA(ctx, cudaMemsetAsync(d.d_work, 0, 2 * sizeof(unsigned long long), stream));

  if (flags & TOR_FLAG_BRUTE_FORCE) {
    tor::RenderParams P;
    memset(&P, 0, sizeof(P));
    {
      int rc = upload_brute_scene_to(ctx, d);
      if (rc) return rc;
    }
    P.sv = ctx->scene.view;
    P.blob = d.d_brute;
    P.cam = ctx->cam;
    P.pixels = d_out;
    P.nrows = nrows;
    P.ncols = ncols;
    P.spp = spp;
    P.max_depth = (int32_t)max_depth;
    P.inv_spp = 1.0 / (double)spp;      // canvas.nim:49
    P.inv_gamma = 1.0 / (double)gamma;  // canvas.nim:50 — float32 widened to float64
    P.row_begin = row_begin;
    P.row_step = row_step;
    P.nsel_rows = nsel;
    P.count_segments = count ? 1u : 0u;
    P.work_counter = d.d_work;
    P.counters = d.d_counters;

    LaunchPlan plan = plan_for(P.sv, d.max_smem_optin);
    TOR_CUDA(ctx, cudaFuncSetAttribute(plan.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    int per_sm = 0;
    TOR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plan.fn, kBlock, plan.smem));
    if (per_sm < 1) return fail(ctx, TOR_ERR_CUDA, "render kernel does not fit on an SM");
    // persistent lanes: one grid that exactly fills the GPU; lanes pull pixels from d_work
    unsigned long long cap = (unsigned long long)d.sm_count * (unsigned long long)per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (timed) TOR_CUDA(ctx, cudaEventRecord(d.ev0, stream));
    plan.fn<<<grid, kBlock, plan.smem, stream>>>(P);
  } else {
    const Tuning& tune = ctx->tune;
    tor::BvhRenderParams P;
    memset(&P, 0, sizeof(P));
    P.bv = ctx->bvh.view;
    P.blob = d.d_blob;
    P.cam = ctx->cam;
    P.pixels = d_out;
    P.nrows = nrows;
    P.ncols = ncols;
    P.spp = spp;
    P.max_depth = (int32_t)max_depth;
    P.inv_spp = 1.0 / (double)spp;
    P.inv_gamma = 1.0 / (double)gamma;
    P.row_begin = row_begin;
    P.row_step = row_step;
    P.nsel_rows = nsel;
    P.count_segments = count ? 1u : 0u;
    P.work_counter = d.d_work;
    P.counters = d.d_counters;
    P.refill = tune.refill;
    P.lanes_per_warp = tune.lanes;
    P.chunk = (uint32_t)tune.chunk;

    BvhLaunchPlan plan = bvh_plan_for(P.bv, d.max_smem_per_sm, d.max_smem_optin, /*chunked=*/sub_log2 != 0,
                                      tune.block, tune.stage_max);
    const int block = plan.block;
    BvhLaunchPlan main_plan = plan;  // the cost pre-pass always runs `plan`; the main launch may use another variant
    TOR_CUDA(ctx, cudaFuncSetAttribute(plan.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    int per_sm = 0;
    TOR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plan.fn, block, plan.smem));
    if (per_sm < 1) return fail(ctx, TOR_ERR_CUDA, "render kernel does not fit on an SM");
    unsigned long long cap = (unsigned long long)d.sm_count * (unsigned long long)per_sm;
    if (ctx->grid_divisor > 1) cap = std::max<unsigned long long>(1ull, cap / (unsigned long long)ctx->grid_divisor);
    const unsigned long long want_b = ((total_px << sub_log2) + block - 1) / block;
    int grid = (int)(want_b < cap ? want_b : cap);
    if (timed) TOR_CUDA(ctx, cudaEventRecord(d.ev0, stream));

    // Pixel scheduling.  A pixel's samples are a serial chain, so the order in which pixels start
    // decides when the render ends.  With enough samples per pixel a cost pre-pass (the first `pre` samples of every
    // pixel, only their segment counts kept) ranks the pixels; the few most expensive ones are traced by a whole warp
    // each, the next ones are dealt to the lanes (32 consecutive ranks per warp, every SM the same mix of warps), the rest is
    // queued most-expensive-first.
    const unsigned long long lanes = (unsigned long long)grid * block;
    // Split-stream queue: a warp takes 64 consecutive units at a time when every lane gets plenty of them, 32 (one
    // per lane) otherwise and over the last four units per lane, so that no warp ends the render on a long chunk of
    // neighbouring — i.e. similarly expensive — units.
    P.chunk_guard = lanes * 4ull;
    if (P.chunk == 0) P.chunk = (total_px << sub_log2) >= lanes * 64ull ? 64u : 32u;
    // below 256 spp the pre-pass costs more than the order gains (C1: +1 ms); split-stream units are short and
    // plentiful, so they need no ranking either
    // From 64 spp the ranking pays when the launch has the whole GPU (1200x675: -10 % at 64 spp, -16 % at 100, -21 % at
    // 255; C1: 10.4 -> 9.0 ms) — the queue it builds keeps neighbouring, equally expensive pixels in one warp.  Not for
    // canvases that leave SMs empty (256x144: the hand-off needs the cooperative path's full grid) and not for the
    // animation's frames in flight (C4: 0.97 -> 1.11 s).
    const bool owns_gpu = (unsigned long long)grid == cap && ctx->grid_divisor <= 1;
    const int32_t pre =
        (sub_log2 == 0 && (spp >= tune.prepass_min_spp || (owns_gpu && spp >= tune.prepass_full_spp))) ? 8 : 0;
    if (sub_log2) {
      const size_t need = (size_t)(total_px << sub_log2) * 3 * sizeof(double);
      if (need > d.partial_cap) {
        if (d.d_partial) cudaFree(d.d_partial);
        d.d_partial = nullptr;
        d.partial_cap = 0;
        TOR_CUDA(ctx, cudaMalloc(&d.d_partial, need));
        d.partial_cap = need;
      }
      P.sub_log2 = sub_log2;
      P.pixels = d.d_partial;
    }
    const bool reorder = !(flags & TOR_FLAG_ROW_MAJOR_QUEUE) && total_px > 64 && total_px < 0x7fffffffull &&
                         max_depth > 0;
    if (P.refill > (P.lanes_per_warp * 5) / 8) P.refill = (P.lanes_per_warp * 5) / 8;
    if (sub_log2) {
      // split-stream mode: the lanes of a warp trace neighbouring rays and finish their traversals close together,
      // so waiting for more of them before shading costs little and shades more lanes at once
      P.refill = tune.refill_fast < P.lanes_per_warp ? tune.refill_fast : P.lanes_per_warp;
    }
    if (P.refill < 1) P.refill = 1;
    if (reorder && pre > 0 && (block % 32) == 0) {
      const uint32_t warps_all = (uint32_t)(lanes / 32);
      const uint32_t warps = tune.deal ? warps_all : 0u;  // warps that are dealt a first wave
      const uint32_t n = (uint32_t)total_px;
      const size_t slots = (size_t)warps * 32u + n;  // upper bound of [dealt][queue] for any cooperative count
      if (slots > d.order_cap) {
        if (d.d_cost) cudaFree(d.d_cost);
        if (d.d_order) cudaFree(d.d_order);
        d.d_cost = d.d_order = nullptr;
        d.order_cap = 0;
        TOR_CUDA(ctx, cudaMalloc(&d.d_cost, 2 * slots * sizeof(uint32_t)));  // raw estimates, then the smoothed ones
        TOR_CUDA(ctx, cudaMalloc(&d.d_order, slots * sizeof(uint32_t)));
        d.order_cap = slots;
      }
      // Queue order: most expensive class first.  Default: coarse classes with the image order kept inside a class
      // and warps that take 32 consecutive queue entries at a time (the CHUNKED kernel), so that after the dealt
      // first wave the lanes of a warp work on neighbouring pixels of the same kind.  TOR_BVH_EXACT_QUEUE=percost
      // selects the earlier scheme (one class per cost value, arbitrary order inside, one atomic per lane).
      const bool coherent = !tune.queue_percost;
      // Warp-cooperative pixels: only in the default scheme, with the dealt wave, and never more than half the warps.
      // (TOR_BVH_COOP_FORCE lifts the cap to every pixel of the launch: the parity tests run whole images that way.)
      uint32_t coop_max = 0;
      tor::CoopLayout lay;
      lay.grid = (uint32_t)grid;
      lay.wpc = (uint32_t)(block / 32);
      lay.per_sm = (uint32_t)per_sm;
      lay.coop_grid = 0;
      lay.cw = (uint32_t)tune.coop_warps;
      lay.n_fast = 0;
      lay.cw_fast = (uint32_t)std::min(tune.coop_fast_warps, tune.coop_warps);
      // With about one pixel per lane the warps need not be full: fewer active lanes per warp trace each pixel faster
      // (9.6 us per segment per lane in a full warp, 6.3 with 16 lanes) at a throughput that nobody needs then.
      // The dealt wave hands lanes 0 .. L-1 of every warp a pixel (L a multiple of the dealing group).
      if (tune.lanes == 32 && tune.thin_px_per_lane > 0 &&
          total_px * 100ull < lanes * (unsigned long long)tune.thin_px_per_lane)
        P.lanes_per_warp = tune.thin_lanes;
      if (P.lanes_per_warp % tune.deal_group) P.lanes_per_warp = std::max(tune.deal_group, P.lanes_per_warp / tune.deal_group * tune.deal_group);
      if (P.refill > (P.lanes_per_warp * 5) / 8) P.refill = std::max(1, (P.lanes_per_warp * 5) / 8);
      lay.lanes = (uint32_t)P.lanes_per_warp;
      // Only launches with few pixels per lane can end on a single pixel's chain (C2: a GPU's share in a 4- or
      // 8-GPU render); with many pixels per lane the dealt first wave hides the chains and the SMs are better used
      // by the lanes.  Cooperative CTAs take whole SMs, so the grid must be the full persistent one.
      const bool few_pixels = total_px < lanes * (unsigned long long)tune.coop_px_per_lane;
      const bool full_grid = grid == per_sm * d.sm_count;
      lay.sorted = (tune.deal_sorted == 2 || (tune.deal_sorted == 1 && few_pixels)) ? 1u : 0u;
      if (coherent && warps && tune.coop_max_pct > 0 && ((few_pixels && full_grid) || tune.coop_force >= 0)) {
        lay.coop_grid = std::max(1u, (uint32_t)((unsigned long long)d.sm_count * (unsigned)tune.coop_max_pct / 100ull));
        lay.n_fast = (uint32_t)((unsigned long long)lay.coop_grid * (unsigned)tune.coop_fast_pct / 100ull);
        coop_max = lay.coop_grid * lay.cw * (uint32_t)tune.coop_queue_factor;  // a queue: several per warp
        if (tune.coop_force >= 0) coop_max = n;  // tests: any number of pixels, the warps loop
      }
      if (coop_max > d.coop_cap) {
        if (d.d_coop) cudaFree(d.d_coop);
        d.d_coop = nullptr;
        d.coop_cap = 0;
        TOR_CUDA(ctx, cudaMalloc(&d.d_coop, (size_t)coop_max * sizeof(uint32_t)));
        d.coop_cap = coop_max;
      }
      tor::BvhRenderParams Q = P;
      Q.spp = pre;
      Q.count_segments = 0;
      Q.work_counter = d.d_work + 1;
      Q.cost = d.d_cost;
      Q.scramble = coprime_near_golden(n);
      Q.lanes_per_warp = tune.lanes;  // the pre-pass is throughput work: full warps
      Q.refill = std::max(1, std::min(tune.refill, (tune.lanes * 5) / 8));
      plan.fn<<<grid, block, plan.smem, stream>>>(Q);
      TOR_CUDA(ctx, cudaGetLastError());
      TOR_CUDA(ctx, cudaMemsetAsync(d.d_hist, 0, tor::kCostBuckets * sizeof(uint32_t), stream));
      TOR_CUDA(ctx, cudaMemsetAsync(d.d_order, 0xff, (size_t)warps * 32u * sizeof(uint32_t), stream));
      const int sort_grid = (int)((n + 255) / 256 < (unsigned)(d.sm_count * 8) ? (n + 255) / 256 : d.sm_count * 8);
      const uint32_t group = (uint32_t)tune.deal_group;
      const uint32_t force = tune.coop_force < 0 ? 0xffffffffu
                                                 : (uint32_t)(tune.coop_force > 0x7fffffffll ? 0x7fffffffll : tune.coop_force);
      uint32_t* const d_rank_cost = d.d_cost + d.order_cap;  // what the ranking uses (cost_smooth_kernel)
      tor::cost_smooth_kernel<<<sort_grid, 256, 0, stream>>>(d.d_cost, d_rank_cost, n, (uint32_t)ncols);
      tor::cost_histogram_kernel<<<sort_grid, 256, 0, stream>>>(d_rank_cost, n, d.d_hist, coherent ? 1u : 0u);
      tor::cost_offsets_kernel<<<1, tor::kCostBuckets, 0, stream>>>(d.d_hist, d.d_sched, coherent ? 1u : 0u, lanes,
                                                                    tune.coop_alpha, coop_max, force);
      if (coherent) {
        tor::cost_scatter_ordered_kernel<<<8, 1024, 0, stream>>>(d_rank_cost, n, d.d_hist, d.d_order, warps, group, 1u,
                                                              d.d_sched, d.d_coop, lay);
        main_plan = bvh_plan_for(P.bv, d.max_smem_per_sm, d.max_smem_optin, /*chunked=*/true, tune.block, tune.stage_max);
        TOR_CUDA(ctx, cudaFuncSetAttribute(main_plan.fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)main_plan.smem));
        P.chunk = 32;
        P.chunk_guard = 0;
        P.endgame_min_chunk = (uint32_t)tune.endgame_min_chunk;
      } else {
        tor::cost_scatter_kernel<<<sort_grid, 256, 0, stream>>>(d_rank_cost, n, d.d_hist, d.d_order, warps, group,
                                                                d.d_sched, d.d_coop, lay);
      }
      TOR_CUDA(ctx, cudaGetLastError());
      ctx->launches += 5;
      P.order = d.d_order;
      P.first_wave = warps * 32u;
      P.sched = d.d_sched;  // sched[0] = 0 when coop_max == 0
      P.coop_list = d.d_coop;
      P.coop = lay;
      P.deal_ticket = d.d_ticket;
      if (tune.handoff_pct > 0 && coherent && warps && (lay.coop_grid > 0 || tune.handoff_all)) {
        // late hand-off (BvhRenderParams::handoff): room for every cooperative pixel, one record per lane and the
        // unassigned rest of every warp's last queue chunk (at most 32 slots)
        const size_t need = ((size_t)coop_max + 2 * (size_t)lanes) * sizeof(tor::HandoffRec);
        if (need > d.handoff_cap) {
          if (d.d_handoff) cudaFree(d.d_handoff);
          d.d_handoff = nullptr;
          d.handoff_cap = 0;
          TOR_CUDA(ctx, cudaMalloc(&d.d_handoff, need));
          d.handoff_cap = need;
        }
        P.handoff = d.d_handoff;
        P.handoff_cap_a = coop_max;
        P.handoff_pct = (uint32_t)tune.handoff_pct;
        P.handoff_min_left = (uint32_t)tune.handoff_min_left;
      }
    } else if (reorder && sub_log2 == 0 && !tune.no_scramble) {
      // exact mode without cost information (few samples per pixel): scatter the image over the warps so that
      // expensive neighbours do not share one (BvhRenderParams::scramble).  Split-stream units are short, so there
      // the queue stays in image order and the lanes of a warp work on neighbouring pixels.
      P.scramble = coprime_near_golden((uint32_t)total_px);
      // No ranking here, so expensive pixels start at any time and the launch ends on the last of them: the late
      // hand-off (BvhRenderParams::handoff) takes over once handoff_plain % of ALL warps are done.  C1: 13.5 -> 10.2 ms;
      // 5 % at 10 pixels per lane, 1 % at 27.  (Not for the animation's frames in flight, whose launches share the GPU
      // on purpose.)
      if (tune.handoff_plain > 0 && ctx->grid_divisor <= 1 && spp >= 2 * tune.handoff_min_left &&
          total_px < (unsigned long long)(per_sm * d.sm_count) * block * (unsigned long long)tune.handoff_plain_px) {
        const size_t need = (size_t)lanes * sizeof(tor::HandoffRec);
        if (need > d.handoff_cap) {
          if (d.d_handoff) cudaFree(d.d_handoff);
          d.d_handoff = nullptr;
          d.handoff_cap = 0;
          TOR_CUDA(ctx, cudaMalloc(&d.d_handoff, need));
          d.handoff_cap = need;
        }
        P.handoff = d.d_handoff;
        P.handoff_cap_a = 0;
        P.handoff_pct = (uint32_t)tune.handoff_plain;
        P.handoff_min_left = (uint32_t)tune.handoff_min_left;
        P.deal_ticket = d.d_ticket;
      }
    } else if (sub_log2 && tune.fast_scramble) {
      P.scramble = coprime_near_golden((uint32_t)total_px);
    }
    if (tune.debug_times && !P.cost) {
      const size_t bytes = (3 * 4096 + 2 * 8192 + 3 * 4096) * sizeof(unsigned long long);
      if (!d.d_dbg) TOR_CUDA(ctx, cudaMalloc(&d.d_dbg, bytes));
      TOR_CUDA(ctx, cudaMemsetAsync(d.d_dbg, 0, bytes, stream));
      P.dbg_times = d.d_dbg;
    }
    const bool with_coop = P.sched != nullptr && P.coop.coop_grid > 0;
    if (P.sched || P.handoff) TOR_CUDA(ctx, cudaMemsetAsync(d.d_ticket, 0, 8 * sizeof(unsigned int), stream));  // BvhRenderParams::deal_ticket
    if (with_coop) {
      // render_coop_kernel on its own high-priority stream, launched first so that its CTAs take their SMs before the
      // lane kernel's grid fills the GPU; it reads the ranking the kernels above left on `stream`
      TOR_CUDA(ctx, cudaEventRecord(d.ev_ranked, stream));
      TOR_CUDA(ctx, cudaStreamWaitEvent(d.stream_coop, d.ev_ranked, 0));
      tor::render_coop_kernel<<<P.coop.coop_grid, tor::kCoopBlock, 0, d.stream_coop>>>(P);
      TOR_CUDA(ctx, cudaGetLastError());
      TOR_CUDA(ctx, cudaEventRecord(d.ev_coop_done, d.stream_coop));
      tor::coop_gate_kernel<<<1, 1, 0, stream>>>(d.d_sched, d.d_ticket + 1, P.coop);
      TOR_CUDA(ctx, cudaGetLastError());
      ctx->launches += 2;
    }
    main_plan.fn<<<grid, block, main_plan.smem, stream>>>(P);
    TOR_CUDA(ctx, cudaGetLastError());
    if (with_coop) TOR_CUDA(ctx, cudaStreamWaitEvent(stream, d.ev_coop_done, 0));  // draw needs every pixel's sum
    d.handoff_used = P.handoff != nullptr;
    if (P.handoff) {
      // the pixels that lanes and cooperative warps parked when the launch entered its tail: one warp each, all SMs
      tor::BvhRenderParams H = P;
      H.handoff_mode = 1;
      H.coop.cw = (uint32_t)tune.handoff_warps;
      tor::render_coop_kernel<<<d.sm_count, tor::kCoopBlock, 0, stream>>>(H);
      TOR_CUDA(ctx, cudaGetLastError());
      ctx->launches += 1;
    }
    const unsigned long long nch = total_px * 3ull;
    if (sub_log2) {  // per-range partial sums -> pixel sums, fixed pairwise order
      const unsigned long long nthr = nch << sub_log2;
      tor::substream_reduce_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, stream>>>(d.d_partial, d_out, nch, sub_log2);
      TOR_CUDA(ctx, cudaGetLastError());
      ctx->launches += 1;
    }
    // canvas.nim:47-54 `draw` over the sums the render kernel left behind
    tor::draw_kernel<<<(unsigned)((nch + 255) / 256), 256, 0, stream>>>(d_out, nch, P.inv_spp, P.inv_gamma);
    ctx->launches += 1;
  }
  TOR_CUDA(ctx, cudaGetLastError());
  if (timed) {
    TOR_CUDA(ctx, cudaEventRecord(d.ev1, stream));
    d.timed = true;
  }
  // the per-context scratch (queue heads, cost / order arrays, scene blob) is in use until here, on whatever stream
  // the caller chose: the next launch, the next scene upload and tor_sync order themselves after this event
  TOR_CUDA(ctx, cudaEventRecord(d.ev_busy, stream));
  d.busy = true;
  ctx->launches += 1;
  if (count) ctx->counters_pending = true;
  return TOR_OK;
}

}  // namespace

extern "C" {

int tor_abi_version(void) { return TOR_B200_ABI_VERSION; }

const char* tor_last_error(const tor_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int tor_ctx_create(const int* devices, int ndev, tor_ctx** out) {
  if (!out) return fail(nullptr, TOR_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, TOR_ERR_NO_DEVICE,
                std::string("no CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
  std::vector<int> ids;
  if (!devices || ndev <= 0) {
    ids.push_back(0);
  } else {
    for (int i = 0; i < ndev; ++i) {
      if (devices[i] < 0 || devices[i] >= count)
        return fail(nullptr, TOR_ERR_INVALID_ARG, "device index out of range");
      ids.push_back(devices[i]);
    }
  }
  tor_ctx* ctx = new tor_ctx();
  ctx->tune = Tuning::from_env();
  for (int id : ids) {
    DeviceState d;
    d.dev = id;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(id)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, id)) != cudaSuccess) {
      g_create_err = std::string("cudaSetDevice/GetDeviceProperties: ") + cudaGetErrorString(e);
      tor_ctx_destroy(ctx);
      return TOR_ERR_CUDA;
    }
    if (prop.major < 10) {
      g_create_err = "device is not sm_100-class (the kernels are built for sm_100a only)";
      tor_ctx_destroy(ctx);
      return TOR_ERR_NO_DEVICE;
    }
    d.sm_count = prop.multiProcessorCount;
    d.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    d.max_smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    bool ok = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&d.ev0) == cudaSuccess && cudaEventCreate(&d.ev1) == cudaSuccess &&
              cudaEventCreateWithFlags(&d.ev_busy, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&d.ev_ranked, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&d.ev_coop_done, cudaEventDisableTiming) == cudaSuccess &&
              create_priority_stream(&d.stream_coop) == cudaSuccess &&
              cudaMalloc(&d.d_ticket, 8 * sizeof(unsigned int)) == cudaSuccess &&
              cudaMalloc(&d.d_work, 2 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&d.d_sched, 4 * sizeof(uint32_t)) == cudaSuccess &&
              cudaMemset(d.d_sched, 0, 4 * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&d.d_hist, tor::kCostBuckets * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&d.d_counters, 4 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMemset(d.d_counters, 0, 4 * sizeof(unsigned long long)) == cudaSuccess;
    ctx->devs.push_back(d);
    if (!ok) {
      g_create_err = std::string("context allocation: ") + cudaGetErrorString(cudaGetLastError());
      tor_ctx_destroy(ctx);
      return TOR_ERR_CUDA;
    }
  }
  *out = ctx;
  return TOR_OK;
}

void tor_ctx_destroy(tor_ctx* ctx) {
  if (!ctx) return;
  for (DeviceState& d : ctx->devs) {
    cudaSetDevice(d.dev);
    if (d.stream) cudaStreamSynchronize(d.stream);
    if (d.busy && d.ev_busy) cudaEventSynchronize(d.ev_busy);
    if (d.d_blob) cudaFree(d.d_blob);
    if (d.d_brute) cudaFree(d.d_brute);
    if (d.d_sched) cudaFree(d.d_sched);
    if (d.d_coop) cudaFree(d.d_coop);
    if (d.ev_busy) cudaEventDestroy(d.ev_busy);
    if (d.ev_ranked) cudaEventDestroy(d.ev_ranked);
    if (d.ev_coop_done) cudaEventDestroy(d.ev_coop_done);
    if (d.stream_coop) cudaStreamDestroy(d.stream_coop);
    if (d.d_ticket) cudaFree(d.d_ticket);
    if (d.d_handoff) cudaFree(d.d_handoff);
    if (d.d_dbg) cudaFree(d.d_dbg);
    if (d.d_pixels) cudaFree(d.d_pixels);
    if (d.d_work) cudaFree(d.d_work);
    if (d.d_cost) cudaFree(d.d_cost);
    if (d.d_order) cudaFree(d.d_order);
    if (d.d_hist) cudaFree(d.d_hist);
    if (d.d_rgb8) cudaFree(d.d_rgb8);
    if (d.d_partial) cudaFree(d.d_partial);
    if (d.d_counters) cudaFree(d.d_counters);
    if (d.ev0) cudaEventDestroy(d.ev0);
    if (d.ev1) cudaEventDestroy(d.ev1);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  delete ctx;
}

int tor_fast_substream_count(uint32_t flags, int32_t nrows, int32_t ncols, int32_t samples_per_pixel) {
  if (nrows <= 0 || ncols <= 0 || samples_per_pixel < 0) return TOR_ERR_INVALID_ARG;
  uint32_t lg = 0;
  int rc = substream_log2(nullptr, flags, nrows, ncols, samples_per_pixel, &lg);
  return rc ? rc : (int)(1u << lg);
}

int tor_scene_upload(tor_ctx* ctx, const tor_camera* cam, const void* objects, int64_t len, int64_t stride) {
  if (!ctx) return TOR_ERR_INVALID_ARG;
  int rc = set_scene(ctx, cam, objects, len, stride);
  if (rc) return rc;
  return upload_scene_to(ctx, ctx->devs[0]);
}

int tor_render_device_async(tor_ctx* ctx, double* d_pixels, int32_t nrows, int32_t ncols, int32_t samples_per_pixel,
                            float gamma_correction, int64_t max_depth, uint32_t flags, int32_t row_begin,
                            int32_t row_end, int32_t row_step, void* cuda_stream) {
  if (!ctx) return TOR_ERR_INVALID_ARG;
  if (!ctx->have_scene) return fail(ctx, TOR_ERR_INVALID_ARG, "tor_scene_upload has not been called");
  if (!d_pixels) return fail(ctx, TOR_ERR_INVALID_ARG, "d_pixels is NULL");
  int rc = check_canvas_dims(ctx, nrows, ncols, samples_per_pixel, max_depth, row_begin, row_end, row_step);
  if (rc) return rc;
  DeviceState& d = ctx->devs[0];
  rc = upload_scene_to(ctx, d);
  if (rc) return rc;
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
  return launch_rows(ctx, d, d_pixels, nrows, ncols, samples_per_pixel, gamma_correction, max_depth, flags, row_begin,
                     row_end, row_step, s, /*timed=*/true);
}

int tor_download_rows_async(tor_ctx* ctx, const double* d_rows, double* host_pixels, int32_t ncols, int32_t row_begin,
                            int32_t row_step, int32_t nsel, void* cuda_stream) {
  if (!ctx) return TOR_ERR_INVALID_ARG;
  if (!d_rows || !host_pixels || ncols <= 0 || row_begin < 0 || row_step <= 0 || nsel < 0)
    return fail(ctx, TOR_ERR_INVALID_ARG, "tor_download_rows_async: bad arguments");
  if (nsel == 0) return TOR_OK;
  DeviceState& d = ctx->devs[0];
  TOR_CUDA(ctx, cudaSetDevice(d.dev));
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
  const size_t row_bytes = (size_t)ncols * 3 * sizeof(double);
  // compact device rows -> canvas rows row_begin, row_begin + row_step, ...: one strided copy, no staging buffer
  TOR_CUDA(ctx, cudaMemcpy2DAsync(host_pixels + (size_t)row_begin * ncols * 3, (size_t)row_step * row_bytes, d_rows,
                                  row_bytes, row_bytes, (size_t)nsel, cudaMemcpyDeviceToHost, s));
  return TOR_OK;
}

int tor_sync(tor_ctx* ctx) {
  if (!ctx) return TOR_ERR_INVALID_ARG;
  for (DeviceState& d : ctx->devs) {
    TOR_CUDA(ctx, cudaSetDevice(d.dev));
    TOR_CUDA(ctx, cudaStreamSynchronize(d.stream));
    if (d.busy) {  // also covers renders that tor_render_device_async put on a caller's stream
      TOR_CUDA(ctx, cudaEventSynchronize(d.ev_busy));
      d.busy = false;
    }
  }
  return TOR_OK;
}

int tor_render_rows(tor_ctx* ctx, tor_canvas* canvas, const tor_camera* cam, const void* objects, int64_t len,
                    int64_t stride, int64_t max_depth, uint32_t flags, int32_t row_begin, int32_t row_end,
                    int32_t row_step) {
  if (!ctx) return TOR_ERR_INVALID_ARG;
  if (!canvas || !canvas->pixels) return fail(ctx, TOR_ERR_INVALID_ARG, "canvas or canvas->pixels is NULL");
  int rc = check_canvas_dims(ctx, canvas->nrows, canvas->ncols, canvas->samples_per_pixel, max_depth, row_begin,
                             row_end, row_step);
  if (rc) return rc;
  rc = set_scene(ctx, cam, objects, len, stride);
  if (rc) return rc;

  const int ndev = (int)ctx->devs.size();
  const int32_t ncols = canvas->ncols;
  const size_t row_bytes = (size_t)ncols * 3 * sizeof(double);
  const int32_t nsel_total = row_end > row_begin ? (row_end - row_begin + row_step - 1) / row_step : 0;
  if (nsel_total == 0) return TOR_OK;

  // selected row k (k = 0 .. nsel_total-1) goes to device k mod ndev: cheap sky rows and expensive
  // ground rows interleave across devices.  Seeds use the absolute (row, col), so the image does not
  // depend on ndev (render.nim:59-60).
  // On an error for device g the devices before it already have kernels and copies into canvas->pixels in flight:
  // drain them before the caller gets the error and possibly frees the canvas (the original message is kept).
  auto bail = [&](int code) {
    std::string msg = ctx->err;
    tor_sync(ctx);
    ctx->err = msg;
    return code;
  };
  for (int g = 0; g < ndev; ++g) {
    DeviceState& d = ctx->devs[g];
    if (g >= nsel_total) {
      d.timed = false;
      continue;
    }
    const int32_t rb = row_begin + g * row_step;
    const int32_t rs = row_step * ndev;
    const int32_t nsel = (row_end - rb + rs - 1) / rs;
    rc = ensure_capacity(ctx, d, device_blob_bytes(ctx), (size_t)nsel * row_bytes);
    if (rc) return bail(rc);
    rc = upload_scene_to(ctx, d);
    if (rc) return bail(rc);
    rc = launch_rows(ctx, d, d.d_pixels, canvas->nrows, ncols, canvas->samples_per_pixel, canvas->gamma_correction,
                     max_depth, flags, rb, row_end, rs, d.stream, /*timed=*/true);
    if (rc) return bail(rc);
    // device rows are compact; scatter them back to their canvas rows (pitch = rs rows)
    double* dst = canvas->pixels + (size_t)rb * ncols * 3;
    cudaError_t ce = cudaMemcpy2DAsync(dst, (size_t)rs * row_bytes, d.d_pixels, row_bytes, row_bytes, (size_t)nsel,
                                       cudaMemcpyDeviceToHost, d.stream);
    if (ce != cudaSuccess) return bail(fail(ctx, TOR_ERR_CUDA, std::string("cudaMemcpy2DAsync: ") + cudaGetErrorString(ce)));
  }
  return tor_sync(ctx);
}

// Render the uploaded scene on device d of ctx and deliver the packed RGB8 image (io/ppm.nim quantisation, PPM row
// order) to rgb8_out (host) asynchronously on d.stream.
static int enqueue_rgb8(tor_ctx* ctx, int32_t nrows, int32_t ncols, int32_t spp, float gamma, int64_t max_depth,
                        uint32_t flags, uint8_t* rgb8_out) {
  DeviceState& d = ctx->devs[0];
  const size_t npix = (size_t)nrows * ncols;
  int rc = ensure_capacity(ctx, d, device_blob_bytes(ctx), npix * 3 * sizeof(double));
  if (rc) return rc;
  if (npix * 3 > d.rgb8_cap) {
    if (d.d_rgb8) cudaFree(d.d_rgb8);
    d.d_rgb8 = nullptr;
    d.rgb8_cap = 0;
    TOR_CUDA(ctx, cudaMalloc(&d.d_rgb8, npix * 3));
    d.rgb8_cap = npix * 3;
  }
  rc = upload_scene_to(ctx, d);
  if (rc) return rc;
  rc = launch_rows(ctx, d, d.d_pixels, nrows, ncols, spp, gamma, max_depth, flags, 0, nrows, 1, d.stream, /*timed=*/true);
  if (rc) return rc;
  tor::quantise_rgb8_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, d.stream>>>(d.d_pixels, nrows, ncols, d.d_rgb8);
  TOR_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  // rgb8_out may be host memory (the caller's frame buffer) or device memory (a frame-parallel rank collecting its
  // frames for one gather): unified addressing tells the copy which
  TOR_CUDA(ctx, cudaMemcpyAsync(rgb8_out, d.d_rgb8, npix * 3, cudaMemcpyDefault, d.stream));
  // the copy reads d_rgb8 after launch_rows' own end-of-launch event: move the context's busy mark behind it
  TOR_CUDA(ctx, cudaEventRecord(d.ev_busy, d.stream));
  d.busy = true;
  return TOR_OK;
}

int tor_render_rgb8_async(tor_ctx* ctx, const tor_canvas* canvas, const tor_camera* cam, const void* objects,
                          int64_t len, int64_t stride, int64_t max_depth, uint32_t flags, uint8_t* rgb8_out) {
  if (!ctx) return TOR_ERR_INVALID_ARG;
  if (!canvas || !rgb8_out) return fail(ctx, TOR_ERR_INVALID_ARG, "canvas or rgb8_out is NULL");
  const int32_t nrows = canvas->nrows, ncols = canvas->ncols;
  int rc = check_canvas_dims(ctx, nrows, ncols, canvas->samples_per_pixel, max_depth, 0, nrows, 1);
  if (rc) return rc;
  rc = set_scene(ctx, cam, objects, len, stride);
  if (rc) return rc;
  return enqueue_rgb8(ctx, nrows, ncols, canvas->samples_per_pixel, canvas->gamma_correction, max_depth, flags, rgb8_out);
}

int tor_render_ycbcr420_async(tor_ctx* ctx, const tor_canvas* canvas, const tor_camera* cam, const void* objects,
                              int64_t len, int64_t stride, int64_t max_depth, uint32_t flags, uint8_t* ycbcr_out) {
  if (!ctx) return TOR_ERR_INVALID_ARG;
  if (!canvas || !ycbcr_out) return fail(ctx, TOR_ERR_INVALID_ARG, "canvas or ycbcr_out is NULL");
  const int32_t nrows = canvas->nrows, ncols = canvas->ncols;
  int rc = check_canvas_dims(ctx, nrows, ncols, canvas->samples_per_pixel, max_depth, 0, nrows, 1);
  if (rc) return rc;
  if ((nrows & 1) || (ncols & 1))  // color_conversions.nim:201-202
    return fail(ctx, TOR_ERR_INVALID_ARG, "Y'CbCr 4:2:0 needs an even width and height");
  rc = set_scene(ctx, cam, objects, len, stride);
  if (rc) return rc;
  DeviceState& d = ctx->devs[0];
  const size_t npix = (size_t)nrows * ncols;
  const size_t nbytes = npix + 2 * (npix / 4);
  rc = ensure_capacity(ctx, d, device_blob_bytes(ctx), npix * 3 * sizeof(double));
  if (rc) return rc;
  if (npix * 3 > d.rgb8_cap) {  // the same scratch buffer as the RGB8 output (3 bytes per pixel >= 1.5)
    if (d.d_rgb8) cudaFree(d.d_rgb8);
    d.d_rgb8 = nullptr;
    d.rgb8_cap = 0;
    TOR_CUDA(ctx, cudaMalloc(&d.d_rgb8, npix * 3));
    d.rgb8_cap = npix * 3;
  }
  rc = upload_scene_to(ctx, d);
  if (rc) return rc;
  rc = launch_rows(ctx, d, d.d_pixels, nrows, ncols, canvas->samples_per_pixel, canvas->gamma_correction, max_depth,
                   flags & ~TOR_FLAG_RGB_ROWS_AS_WRITTEN, 0, nrows, 1, d.stream, /*timed=*/true);
  if (rc) return rc;
  uint8_t* dY = d.d_rgb8;
  uint8_t* dCb = dY + npix;
  uint8_t* dCr = dCb + npix / 4;
  tor::ycbcr420_kernel<<<(unsigned)((npix / 4 + 255) / 256), 256, 0, d.stream>>>(
      d.d_pixels, nrows, ncols, (flags & TOR_FLAG_RGB_ROWS_AS_WRITTEN) ? 1 : 0, dY, dCb, dCr);
  TOR_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  TOR_CUDA(ctx, cudaMemcpyAsync(ycbcr_out, d.d_rgb8, nbytes, cudaMemcpyDeviceToHost, d.stream));
  return TOR_OK;
}

int tor_render_ycbcr420(tor_ctx* ctx, const tor_canvas* canvas, const tor_camera* cam, const void* objects, int64_t len,
                        int64_t stride, int64_t max_depth, uint32_t flags, uint8_t* ycbcr_out) {
  int rc = tor_render_ycbcr420_async(ctx, canvas, cam, objects, len, stride, max_depth, flags, ycbcr_out);
  if (rc) return rc;
  return tor_sync(ctx);
}

int tor_render_rgb8(tor_ctx* ctx, const tor_canvas* canvas, const tor_camera* cam, const void* objects, int64_t len,
                    int64_t stride, int64_t max_depth, uint32_t flags, uint8_t* rgb8_out) {
  int rc = tor_render_rgb8_async(ctx, canvas, cam, objects, len, stride, max_depth, flags, rgb8_out);
  if (rc) return rc;
  return tor_sync(ctx);
}

void* tor_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}

void tor_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int tor_render(tor_ctx* ctx, tor_canvas* canvas, const tor_camera* cam, const void* objects, int64_t len,
               int64_t stride, int64_t max_depth, uint32_t flags) {
  if (!ctx) return TOR_ERR_INVALID_ARG;
  if (!canvas) return fail(ctx, TOR_ERR_INVALID_ARG, "canvas is NULL");
  return tor_render_rows(ctx, canvas, cam, objects, len, stride, max_depth, flags, 0, canvas->nrows, 1);
}

int tor_get_counters(tor_ctx* ctx, uint64_t out[3]) {
  if (!ctx || !out) return TOR_ERR_INVALID_ARG;
  uint64_t rays = 0, segs = 0, boxes = 0, tests = 0;
  for (DeviceState& d : ctx->devs) {
    unsigned long long h[4] = {0, 0, 0, 0};
    TOR_CUDA(ctx, cudaSetDevice(d.dev));
    TOR_CUDA(ctx, cudaStreamSynchronize(d.stream));
    TOR_CUDA(ctx, cudaMemcpy(h, d.d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    TOR_CUDA(ctx, cudaMemset(d.d_counters, 0, sizeof(h)));
    rays += h[0];
    segs += h[1];
    boxes += h[2];
    tests += h[3];
  }
  out[0] = rays;
  out[1] = segs;
  out[2] = segs * (uint64_t)(ctx->have_scene ? ctx->objs.size() : 0);
  ctx->trav_counters[0] = boxes;
  ctx->trav_counters[1] = tests;
  ctx->counters_pending = false;
  return TOR_OK;
}

int tor_get_traversal_counters(tor_ctx* ctx, uint64_t out[2]) {
  if (!ctx || !out) return TOR_ERR_INVALID_ARG;
  out[0] = ctx->trav_counters[0];
  out[1] = ctx->trav_counters[1];
  return TOR_OK;
}

int tor_scene_info(tor_ctx* ctx, int64_t out[6]) {
  if (!ctx || !out) return TOR_ERR_INVALID_ARG;
  if (!ctx->have_scene) return fail(ctx, TOR_ERR_INVALID_ARG, "no scene has been set on this context");
  out[0] = (int64_t)ctx->objs.size();
  out[1] = ctx->bvh.view.n_nodes;
  out[2] = ctx->bvh.n_leaves;
  out[3] = ctx->bvh.max_depth;
  out[4] = ctx->bvh.view.n_objects - ctx->bvh.view.n_tree_objs;
  out[5] = (int64_t)ctx->bvh.blob.size();
  return TOR_OK;
}

int tor_last_schedule(tor_ctx* ctx, int64_t out[4]) {
  if (!ctx || !out) return TOR_ERR_INVALID_ARG;
  out[0] = out[1] = out[2] = out[3] = 0;
  for (DeviceState& d : ctx->devs) {
    uint32_t h[4] = {0, 0, 0, 0};
    TOR_CUDA(ctx, cudaSetDevice(d.dev));
    TOR_CUDA(ctx, cudaStreamSynchronize(d.stream));
    int rc = wait_idle(ctx, d);
    if (rc) return rc;
    TOR_CUDA(ctx, cudaMemcpy(h, d.d_sched, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] += h[0];
    out[1] += h[1];
    out[2] += h[2];
  }
  out[3] = (int64_t)ctx->devs.size();
  return TOR_OK;
}

int tor_last_handoffs(tor_ctx* ctx, int64_t out[2]) {
  if (!ctx || !out) return TOR_ERR_INVALID_ARG;
  out[0] = out[1] = 0;
  for (DeviceState& d : ctx->devs) {
    unsigned int h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    TOR_CUDA(ctx, cudaSetDevice(d.dev));
    TOR_CUDA(ctx, cudaStreamSynchronize(d.stream));
    int rc = wait_idle(ctx, d);
    if (rc) return rc;
    if (!d.handoff_used) continue;
    TOR_CUDA(ctx, cudaMemcpy(h, d.d_ticket, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] += h[5];
    out[1] += h[6];
  }
  return TOR_OK;
}

int tor_debug_times(tor_ctx* ctx, uint64_t* out, int64_t n) {
  if (!ctx || !out) return TOR_ERR_INVALID_ARG;
  DeviceState& d = ctx->devs[0];
  if (!d.d_dbg) return fail(ctx, TOR_ERR_INVALID_ARG, "no stamps: create the context with TOR_BVH_DEBUG_TIMES set");
  const int64_t have = 3 * 4096 + 2 * 8192 + 3 * 4096;
  TOR_CUDA(ctx, cudaSetDevice(d.dev));
  TOR_CUDA(ctx, cudaDeviceSynchronize());
  TOR_CUDA(ctx, cudaMemcpy(out, d.d_dbg, (size_t)(n < have ? n : have) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return TOR_OK;
}

int tor_last_kernel_ms(tor_ctx* ctx, float* ms) {
  if (!ctx || !ms) return TOR_ERR_INVALID_ARG;
  float mx = 0.f;
  bool any = false;
  for (DeviceState& d : ctx->devs) {
    if (!d.timed) continue;
    float t = 0.f;
    TOR_CUDA(ctx, cudaSetDevice(d.dev));
    TOR_CUDA(ctx, cudaEventSynchronize(d.ev1));
    TOR_CUDA(ctx, cudaEventElapsedTime(&t, d.ev0, d.ev1));
    if (t > mx) mx = t;
    any = true;
  }
  if (!any) return fail(ctx, TOR_ERR_INVALID_ARG, "no render has been timed on this context");
  *ms = mx;
  return TOR_OK;
}

int64_t tor_launch_count(const tor_ctx* ctx) { return ctx ? ctx->launches : 0; }

int tor_measure_fp64_peak(tor_ctx* ctx, double* dfma_per_second) {
  if (!ctx || !dfma_per_second) return TOR_ERR_INVALID_ARG;
  DeviceState& d = ctx->devs[0];
  TOR_CUDA(ctx, cudaSetDevice(d.dev));
  double* sink = nullptr;
  TOR_CUDA(ctx, cudaMalloc(&sink, sizeof(double)));
  const int iters = 4096, grid = d.sm_count * 8, block = 256;
  cudaEvent_t e0, e1;
  TOR_CUDA(ctx, cudaEventCreate(&e0));
  TOR_CUDA(ctx, cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
    TOR_CUDA(ctx, cudaEventRecord(e0, d.stream));
    tor::fp64_peak_kernel<<<grid, block, 0, d.stream>>>(sink, iters, 1.0000001, 1e-9);
    TOR_CUDA(ctx, cudaEventRecord(e1, d.stream));
    TOR_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    TOR_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    double rate = (double)grid * block * (double)iters * tor::kPeakChains / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
    ctx->launches += 1;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *dfma_per_second = best;
  return TOR_OK;
}

// ------------------------------------------------------------------------------- device-resident animation
struct tor_animation_dev {
  tor_ctx* parent = nullptr;
  tor_anim::Animation* an = nullptr;
  int32_t skip = 6;
  std::vector<tor_ctx*> slots;  // one context (= stream + frame buffers + its own copy of the packed scene) per frame in flight
  std::vector<cudaEvent_t> ev_ready;
  cudaStream_t phys = nullptr;  // physics + refit kernels, in frame order
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  bool timing_open = false;
  double *d_vel = nullptr, *d_pos = nullptr, *d_rest = nullptr, *d_rad = nullptr;
  int32_t *d_rec_dyn = nullptr, *d_node_order = nullptr, *d_level_off = nullptr, *d_flag = nullptr;
  float* d_node_y = nullptr;
  int32_t n_dyn = 0, n_levels = 0;
  double pad = 0.0, s_limit = 0.0;
  int64_t frame = 0;     // frames produced or skipped so far
  int64_t rendered = 0;  // frames rendered by this object
  int dev = 0;
  std::vector<double> vel0, pos0;  // initial physics state (tor_animation_dev_reset)
  float t0 = 0.f;
  double angle0 = 0.0;
};

int tor_animation_dev_create(tor_ctx* ctx, uint64_t seed, int32_t height, int32_t width, float dt, float t_min,
                             float t_max, int32_t skip, int32_t in_flight, tor_animation_dev** out) {
  if (!ctx || !out) return TOR_ERR_INVALID_ARG;
  *out = nullptr;
  if (height <= 0 || width <= 0 || skip < 0 || in_flight < 1 || in_flight > 64 || !(dt > 0.f))
    return fail(ctx, TOR_ERR_INVALID_ARG, "tor_animation_dev_create: bad arguments");
  tor_animation_dev* A = new tor_animation_dev();
  A->parent = ctx;
  A->skip = skip;
  A->dev = ctx->devs[0].dev;
  A->an = tor_anim::animation_create(seed, height, width, dt, t_min, t_max);
  auto bail = [&](int code, const std::string& msg) {
    tor_animation_dev_destroy(A);
    return fail(ctx, code, msg);
  };
  // the first frame's scene (all moving spheres at their start height) fixes the packed layout and the tree topology
  std::vector<tor_hittable> objs;
  tor_anim::animation_scene(*A->an, &objs);
  tor_camera cam;
  tor_anim::animation_camera(*A->an, &cam);
  tor::PackedBvh bvh;
  std::string err;
  if (!tor::pack_bvh(objs, cam, &bvh, &err)) return bail(TOR_ERR_INVALID_ARG, err);
  const tor::BvhView& bv = bvh.view;
  A->n_dyn = (int32_t)A->an->spheres.size();
  // same padding as the builder used (tor_bvh.hpp: 2^-19 of the largest coordinate); recover S from s_limit's source
  {
    double S = 0.0;
    for (int k = 0; k < 3; ++k) {
      double v = fabs(cam.origin[k]) + fabs(cam.lens_radius) * (fabs(cam.u[k]) + fabs(cam.v[k]));
      if (v > S) S = v;
    }
    for (const tor_hittable& h : objs)
      for (int k = 0; k < 3; ++k) {
        const double r = fabs(h.radius), e = r + 1e-9 * (fabs(h.center0[k]) + r);
        const double lo = h.center0[k] - e, hi = h.center0[k] + e;
        if (fabs(lo) < 1e30 && fabs(hi) < 1e30) S = std::max(S, std::max(fabs(lo), fabs(hi)));
      }
    A->pad = ldexp(S, -19);
    A->s_limit = (double)bv.s_limit;
  }
  // per tree record: its moving sphere (objects are [ground, moving spheres..., three big spheres])
  std::vector<int32_t> rec_dyn((size_t)std::max(1, bv.n_tree_objs), -1);
  const tor::ObjRec* recs = (const tor::ObjRec*)(bvh.blob.data() + bv.off_objs);
  for (int32_t j = 0; j < bv.n_tree_objs; ++j) {
    const int64_t o = (int64_t)recs[j].orig - 1;
    if (o >= 0 && o < A->n_dyn) rec_dyn[(size_t)j] = (int32_t)o;
  }
  for (int32_t j = bv.n_tree_objs; j < bv.n_objects; ++j) {
    const int64_t o = (int64_t)recs[j].orig - 1;
    if (o >= 0 && o < A->n_dyn) return bail(TOR_ERR_INVALID_ARG, "internal: a moving sphere sits outside the tree");
  }
  // inner nodes by height
  const tor::BvhNode* nodes = (const tor::BvhNode*)(bvh.blob.data() + bv.off_nodes);
  std::vector<int32_t> node_h((size_t)bv.n_nodes, 0);
  int32_t max_h = 0;
  for (int32_t i = bv.n_nodes - 1; i >= 0; --i) {  // children have larger indices than their parent
    int32_t h = 0;
    for (int32_t c : {nodes[i].child0, nodes[i].child1})
      if (c >= 0) h = std::max(h, node_h[(size_t)c] + 1);
    node_h[(size_t)i] = h;
    max_h = std::max(max_h, h);
  }
  A->n_levels = max_h + 1;
  std::vector<int32_t> level_off((size_t)A->n_levels + 1, 0), order((size_t)bv.n_nodes);
  for (int32_t i = 0; i < bv.n_nodes; ++i) level_off[(size_t)node_h[(size_t)i] + 1]++;
  for (int32_t l = 0; l < A->n_levels; ++l) level_off[(size_t)l + 1] += level_off[(size_t)l];
  {
    std::vector<int32_t> cur(level_off.begin(), level_off.end() - 1);
    for (int32_t i = 0; i < bv.n_nodes; ++i) order[(size_t)cur[(size_t)node_h[(size_t)i]]++] = i;
  }
  std::vector<double> vel, pos, rest, rad;
  for (const tor_anim::AnimSphere& sp : A->an->spheres) {
    vel.push_back(sp.velocity);
    pos.push_back(sp.pos_y);
    rest.push_back(sp.coef_restitution);
    rad.push_back(sp.radius);
  }
  A->vel0 = vel;
  A->pos0 = pos;
  A->t0 = A->an->t;
  A->angle0 = A->an->look_from_angle;
  cudaError_t e = cudaSetDevice(A->dev);
  auto up = [&](auto** dptr, const auto& v) {
    using T = typename std::remove_reference<decltype(v)>::type::value_type;
    const size_t bytes = std::max<size_t>(1, v.size()) * sizeof(T);
    if (e == cudaSuccess) e = cudaMalloc((void**)dptr, bytes);
    if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  };
  up(&A->d_vel, vel);
  up(&A->d_pos, pos);
  up(&A->d_rest, rest);
  up(&A->d_rad, rad);
  up(&A->d_rec_dyn, rec_dyn);
  up(&A->d_node_order, order);
  up(&A->d_level_off, level_off);
  if (e == cudaSuccess) e = cudaMalloc((void**)&A->d_node_y, (size_t)bv.n_nodes * 2 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&A->d_flag, sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemset(A->d_flag, 0, sizeof(int32_t));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&A->phys, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&A->ev_t0);
  if (e == cudaSuccess) e = cudaEventCreate(&A->ev_t1);
  if (e != cudaSuccess) return bail(TOR_ERR_CUDA, std::string("tor_animation_dev_create: ") + cudaGetErrorString(e));
  for (int k = 0; k < in_flight; ++k) {
    tor_ctx* c = nullptr;
    int rc = tor_ctx_create(&A->dev, 1, &c);
    if (rc) return bail(rc, tor_last_error(nullptr));
    A->slots.push_back(c);
    c->grid_divisor = A->parent->tune.anim_grid_divisor > 0 ? A->parent->tune.anim_grid_divisor : std::max(1, in_flight / 2);
    c->objs = objs;
    c->bvh = bvh;
    c->cam = cam;
    c->have_scene = true;
    rc = upload_scene_to(c, c->devs[0]);  // the only host-to-device copy of scene data: once per slot
    if (rc) return bail(rc, c->err);
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(TOR_ERR_CUDA, "cudaEventCreate");
    A->ev_ready.push_back(ev);
  }
  *out = A;
  return TOR_OK;
}

void tor_animation_dev_destroy(tor_animation_dev* A) {
  if (!A) return;
  cudaSetDevice(A->dev);
  if (A->phys) cudaStreamSynchronize(A->phys);
  for (tor_ctx* c : A->slots) tor_ctx_destroy(c);
  for (cudaEvent_t ev : A->ev_ready) cudaEventDestroy(ev);
  for (void* p : {(void*)A->d_vel, (void*)A->d_pos, (void*)A->d_rest, (void*)A->d_rad, (void*)A->d_rec_dyn,
                  (void*)A->d_node_order, (void*)A->d_level_off, (void*)A->d_node_y, (void*)A->d_flag})
    if (p) cudaFree(p);
  if (A->ev_t0) cudaEventDestroy(A->ev_t0);
  if (A->ev_t1) cudaEventDestroy(A->ev_t1);
  if (A->phys) cudaStreamDestroy(A->phys);
  delete A->an;
  delete A;
}

int tor_animation_dev_next(tor_animation_dev* A, int32_t samples_per_pixel, float gamma_correction, int64_t max_depth,
                           uint32_t flags, uint8_t* rgb8_out, int64_t* frame_index) {
  if (!A) return TOR_ERR_INVALID_ARG;
  tor_ctx* ctx = A->parent;
  tor_anim::Animation& an = *A->an;
  // the iterator of scenes_animated.nim:176-225: the clock and the camera angle advance on the host (two scalars),
  // the spheres on the device
  int32_t nsteps = 0;
  if (!an.started) {
    while (an.t < an.t_min) {
      tor_anim::anim_step_clock(an);
      ++nsteps;
    }
    an.started = true;
  } else {
    for (int i = 0; i < A->skip; ++i) tor_anim::anim_step_clock(an);
    nsteps = A->skip;
  }
  if (!(an.t < an.t_max)) return 0;
  TOR_CUDA(ctx, cudaSetDevice(A->dev));
  if (!A->timing_open) {
    TOR_CUDA(ctx, cudaEventRecord(A->ev_t0, A->phys));
    A->timing_open = true;
  }
  if (nsteps > 0 && A->n_dyn > 0) {
    tor::anim_step_kernel<<<(unsigned)((A->n_dyn + 255) / 256), 256, 0, A->phys>>>(
        A->d_vel, A->d_pos, A->d_rest, A->n_dyn, nsteps, (double)an.dt, 9.80665 * (double)an.dt);
    TOR_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
  }
  if (frame_index) *frame_index = A->frame;
  const int64_t f = A->frame++;
  if (!rgb8_out) return 1;  // this frame belongs to somebody else (frame-parallel ranks): physics only
  const size_t k = (size_t)(A->rendered++ % (int64_t)A->slots.size());
  tor_ctx* c = A->slots[k];
  DeviceState& d = c->devs[0];
  int rc = check_canvas_dims(ctx, an.nrows, an.ncols, samples_per_pixel, max_depth, 0, an.nrows, 1);
  if (rc) return rc;
  // the slot's previous frame must be out of its blob before the refit rewrites it
  if (d.busy) TOR_CUDA(ctx, cudaStreamWaitEvent(A->phys, d.ev_busy, 0));
  tor::AnimRefitParams R;
  R.blob = d.d_blob;
  R.bv = c->bvh.view;
  R.pos_y = A->d_pos;
  R.radius = A->d_rad;
  R.rec_dyn = A->d_rec_dyn;
  R.node_order = A->d_node_order;
  R.level_off = A->d_level_off;
  R.n_levels = A->n_levels;
  R.pad = A->pad;
  R.s_limit = A->s_limit;
  R.node_y = A->d_node_y;
  R.flag = A->d_flag;
  tor::anim_refit_kernel<<<1, 1024, 0, A->phys>>>(R);
  TOR_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  TOR_CUDA(ctx, cudaEventRecord(A->ev_ready[k], A->phys));
  TOR_CUDA(ctx, cudaStreamWaitEvent(d.stream, A->ev_ready[k], 0));
  tor_anim::animation_camera(an, &c->cam);  // a kernel parameter: no copy
  (void)f;
  rc = enqueue_rgb8(c, an.nrows, an.ncols, samples_per_pixel, gamma_correction, max_depth, flags, rgb8_out);
  if (rc) return fail(ctx, rc, c->err);
  return 1;
}

int tor_animation_dev_reset(tor_animation_dev* A) {
  if (!A) return TOR_ERR_INVALID_ARG;
  tor_ctx* ctx = A->parent;
  int rc = tor_animation_dev_sync(A, nullptr);
  if (rc) return rc;
  TOR_CUDA(ctx, cudaSetDevice(A->dev));
  // back to the first frame: the initial velocities and heights (two small arrays, once per animation, not per frame)
  if (A->n_dyn > 0) {
    TOR_CUDA(ctx, cudaMemcpyAsync(A->d_vel, A->vel0.data(), A->vel0.size() * sizeof(double), cudaMemcpyHostToDevice, A->phys));
    TOR_CUDA(ctx, cudaMemcpyAsync(A->d_pos, A->pos0.data(), A->pos0.size() * sizeof(double), cudaMemcpyHostToDevice, A->phys));
    TOR_CUDA(ctx, cudaStreamSynchronize(A->phys));
  }
  A->an->t = A->t0;
  A->an->look_from_angle = A->angle0;
  A->an->started = false;
  A->frame = 0;
  A->rendered = 0;
  return TOR_OK;
}

int tor_animation_dev_sync(tor_animation_dev* A, float* elapsed_ms) {
  if (!A) return TOR_ERR_INVALID_ARG;
  tor_ctx* ctx = A->parent;
  TOR_CUDA(ctx, cudaSetDevice(A->dev));
  for (tor_ctx* c : A->slots)
    if (c->devs[0].busy) TOR_CUDA(ctx, cudaStreamWaitEvent(A->phys, c->devs[0].ev_busy, 0));
  if (A->timing_open) TOR_CUDA(ctx, cudaEventRecord(A->ev_t1, A->phys));
  TOR_CUDA(ctx, cudaStreamSynchronize(A->phys));
  for (tor_ctx* c : A->slots) {
    int rc = tor_sync(c);
    if (rc) return fail(ctx, rc, c->err);
  }
  if (elapsed_ms) {
    *elapsed_ms = 0.f;
    if (A->timing_open) TOR_CUDA(ctx, cudaEventElapsedTime(elapsed_ms, A->ev_t0, A->ev_t1));
  }
  A->timing_open = false;
  int32_t flag = 0;
  TOR_CUDA(ctx, cudaMemcpy(&flag, A->d_flag, sizeof(flag), cudaMemcpyDeviceToHost));
  if (flag) return fail(ctx, TOR_ERR_INVALID_ARG, "animation: a sphere left the range the box padding was derived for");
  return TOR_OK;
}

int64_t tor_animation_dev_launch_count(const tor_animation_dev* A) {
  if (!A) return 0;
  int64_t n = 0;
  for (tor_ctx* c : A->slots) n += c->launches;
  return n;
}

}  // extern "C"
