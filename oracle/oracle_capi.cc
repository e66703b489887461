// oracle_capi.cc — C entry points over tor_oracle.hpp for the Python tests / bench (ctypes).
// TEST INFRASTRUCTURE ONLY (see tor_oracle.hpp header).
#include <omp.h>
#include <stdio.h>

#include "tor_oracle.hpp"
#include "tor_oracle_video.hpp"

using namespace oracle;

extern "C" {

// --- RNG (support/rng.nim) -------------------------------------------------------------
void oracle_rng_seed1(uint64_t x, uint64_t* state4) {
  Rng r;
  seed(r, x);
  state4[0] = r.s0; state4[1] = r.s1; state4[2] = r.s2; state4[3] = r.s3;
}
void oracle_rng_seed2(int64_t row, int64_t col, uint64_t* state4) {
  Rng r;
  seed(r, row, col);
  state4[0] = r.s0; state4[1] = r.s1; state4[2] = r.s2; state4[3] = r.s3;
}
uint64_t oracle_rng_next(uint64_t* state4) {
  Rng r{state4[0], state4[1], state4[2], state4[3]};
  uint64_t v = next(r);
  state4[0] = r.s0; state4[1] = r.s1; state4[2] = r.s2; state4[3] = r.s3;
  return v;
}
double oracle_rng_uniform01(uint64_t* state4) {
  Rng r{state4[0], state4[1], state4[2], state4[3]};
  double v = uniform01(r);
  state4[0] = r.s0; state4[1] = r.s1; state4[2] = r.s2; state4[3] = r.s3;
  return v;
}
double oracle_rng_uniform_range(uint64_t* state4, double lo, double hi) {
  Rng r{state4[0], state4[1], state4[2], state4[3]};
  double v = uniform_range(r, lo, hi);
  state4[0] = r.s0; state4[1] = r.s1; state4[2] = r.s2; state4[3] = r.s3;
  return v;
}

// --- deterministic math probes (tests/test_detmath.py) ------------------------------------
void oracle_det_sincos(const double* a, double* s, double* c, int64_t n) {
  for (int64_t i = 0; i < n; ++i) tor::detmath::sincos(a[i], &s[i], &c[i]);
}
void oracle_det_pow(const double* x, const double* y, double* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = tor::detmath::pow(x[i], y[i]);
}
void oracle_det_pow_general(const double* x, const double* y, double* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = tor::detmath::pow_general(x[i], y[i]);
}
void oracle_libm_sincos(const double* a, double* s, double* c, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { s[i] = sin(a[i]); c[i] = cos(a[i]); }
}
void oracle_libm_pow(const double* x, const double* y, double* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = pow(x[i], y[i]);
}

// --- scenes.nim / cameras.nim ---------------------------------------------------------
// Returns the object count; writes min(count, cap) records.
int64_t oracle_random_scene(uint64_t seed_value, int32_t half, Hittable* out, int64_t cap) {
  Rng r;
  seed(r, seed_value);
  std::vector<Hittable> w = random_scene(r, half);
  int64_t n = (int64_t)w.size();
  for (int64_t i = 0; i < n && i < cap; ++i) out[i] = w[i];
  return n;
}

void oracle_camera(const double* lookFrom, const double* lookAt, const double* vup, double vfov_deg,
                   double aspect, double aperture, double focus, double t0, double t1, double* cam24) {
  Camera c = make_camera(vec3(lookFrom[0], lookFrom[1], lookFrom[2]), vec3(lookAt[0], lookAt[1], lookAt[2]),
                         vec3(vup[0], vup[1], vup[2]), vfov_deg, aspect, aperture, focus, t0, t1);
  memcpy(cam24, &c, sizeof(c));
}

// --- scenes_animated.nim ------------------------------------------------------------------
void* oracle_anim_create(uint64_t seed_value, int32_t height, int32_t width, float dt, float t_min, float t_max) {
  Rng r;
  seed(r, seed_value);
  Animation* an = new Animation(random_moving_spheres(r, height, width, dt, t_min, t_max));
  return an;
}
void oracle_anim_destroy(void* h) { delete (Animation*)h; }
int64_t oracle_anim_num_spheres(void* h) { return (int64_t)((Animation*)h)->spheres.size(); }
// Advance to the next yielded frame; returns object count (0 when the iterator is exhausted).
int64_t oracle_anim_next_frame(void* h, int32_t skip, int32_t first, double* cam24, Hittable* out, int64_t cap) {
  Animation* an = (Animation*)h;
  Camera cam;
  std::vector<Hittable> scene;
  if (!anim_next_frame(*an, skip, first != 0, cam, scene)) return 0;
  memcpy(cam24, &cam, sizeof(cam));
  int64_t n = (int64_t)scene.size();
  for (int64_t i = 0; i < n && i < cap; ++i) out[i] = scene[i];
  return n;
}

// --- render.nim -----------------------------------------------------------------------
// math_mode: 0 = glibc libm (what the reference links), 1 = tor_detmath (bit-comparable with GPU).
// nthreads <= 0: all cores.  counters3 (optional): primary rays, segments, sphere tests (added to).
int oracle_render(double* pixels, int32_t nrows, int32_t ncols, int32_t spp, float gamma_correction,
                  const double* cam24, const Hittable* world, int64_t n, int64_t max_depth,
                  int32_t row_begin, int32_t row_end, int32_t row_step, int32_t math_mode, int32_t nthreads,
                  uint64_t* counters3) {
  if (!pixels || !cam24 || !world || n <= 0 || row_step <= 0) return -1;
  Camera cam;
  memcpy(&cam, cam24, sizeof(cam));
  Counters c;
  int prev = omp_get_max_threads();
  if (nthreads > 0) omp_set_num_threads(nthreads);
  if (math_mode == 0)
    render<LibmMath>(pixels, nrows, ncols, spp, gamma_correction, cam, world, n, max_depth, row_begin, row_end,
                     row_step, &c);
  else
    render<DetMath>(pixels, nrows, ncols, spp, gamma_correction, cam, world, n, max_depth, row_begin, row_end,
                    row_step, &c);
  if (nthreads > 0) omp_set_num_threads(prev);
  if (counters3) {
    counters3[0] += c.primary_rays;
    counters3[1] += c.segments;
    counters3[2] += c.sphere_tests;
  }
  return 0;
}

// Split-stream ("fast") mode of include/tor_b200.h, restated (tor_oracle.hpp render_split).  nsub must be a power
// of two in 1..64; nsub = 1 is oracle_render.  pixels / linear_sum / sum_sq are optional canvas-sized outputs.
int oracle_render_split(double* pixels, int32_t nrows, int32_t ncols, int32_t spp, float gamma_correction,
                        const double* cam24, const Hittable* world, int64_t n, int64_t max_depth, int32_t row_begin,
                        int32_t row_end, int32_t row_step, int32_t math_mode, int32_t nthreads, uint32_t nsub,
                        double* linear_sum, double* sum_sq, uint64_t* counters3) {
  if (!cam24 || !world || n <= 0 || row_step <= 0) return -1;
  if (nsub < 1 || nsub > 64 || (nsub & (nsub - 1))) return -1;
  Camera cam;
  memcpy(&cam, cam24, sizeof(cam));
  Counters c;
  int prev = omp_get_max_threads();
  if (nthreads > 0) omp_set_num_threads(nthreads);
  if (math_mode == 0)
    render_split<LibmMath>(pixels, nrows, ncols, spp, gamma_correction, cam, world, n, max_depth, row_begin, row_end,
                           row_step, nsub, linear_sum, sum_sq, &c);
  else
    render_split<DetMath>(pixels, nrows, ncols, spp, gamma_correction, cam, world, n, max_depth, row_begin, row_end,
                          row_step, nsub, linear_sum, sum_sq, &c);
  if (nthreads > 0) omp_set_num_threads(prev);
  if (counters3) {
    counters3[0] += c.primary_rays;
    counters3[1] += c.segments;
    counters3[2] += c.sphere_tests;
  }
  return 0;
}

int32_t oracle_num_threads(void) { return omp_get_max_threads(); }

// --- io/ppm.nim -----------------------------------------------------------------------
// 8-bit quantisation of the canvas in PPM order (top row first), 3 bytes per pixel.
void oracle_quantise_rgb8(const double* pixels, int32_t nrows, int32_t ncols, uint8_t* out) {
  int64_t k = 0;
  for (int32_t i = nrows - 1; i >= 0; --i)
    for (int32_t j = 0; j < ncols; ++j) {
      const double* p = pixels + 3 * ((int64_t)i * ncols + j);
      out[k++] = (uint8_t)ppm_conv(p[0]);
      out[k++] = (uint8_t)ppm_conv(p[1]);
      out[k++] = (uint8_t)ppm_conv(p[2]);
    }
}
int oracle_export_ppm(const double* pixels, int32_t nrows, int32_t ncols, const char* path) {
  std::string s = export_ppm(pixels, nrows, ncols);
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  fwrite(s.data(), 1, s.size(), f);
  fclose(f);
  return 0;
}

// --- io/rgb.nim, io/color_conversions.nim, io/h264.nim (tor_oracle_video.hpp) -------------------------
void oracle_to_rgb_raw(const double* pixels, int32_t nrows, int32_t ncols, int32_t as_written, uint8_t* out) {
  to_rgb_raw(pixels, nrows, ncols, as_written != 0, out);
}
int oracle_rgb_to_ycbcr420(int32_t width, int32_t height, const uint8_t* rgb, uint8_t* Y, uint8_t* U, uint8_t* V) {
  return rgb_to_ycbcr420(width, height, rgb, Y, U, V) ? 0 : -1;
}
void oracle_bt601_coefs(uint8_t* out7) {
  YCbCrCoefs c = bt601_coefs();
  out7[0] = c.kr; out7[1] = c.kg; out7[2] = c.kb; out7[3] = c.fb; out7[4] = c.fr; out7[5] = c.y_scale; out7[6] = c.y_min;
}
// Both return the number of bytes the stream piece has; they write min(size, cap) bytes.
int64_t oracle_h264_header(int32_t width, int32_t height, uint8_t* out, int64_t cap) {
  std::vector<uint8_t> v;
  h264_header(width, height, v);
  memcpy(out, v.data(), (size_t)((int64_t)v.size() < cap ? (int64_t)v.size() : cap));
  return (int64_t)v.size();
}
int64_t oracle_h264_frame(int32_t width, int32_t height, const uint8_t* Y, const uint8_t* Cb, const uint8_t* Cr,
                          uint8_t* out, int64_t cap) {
  std::vector<uint8_t> v;
  h264_frame(width, height, Y, Cb, Cr, v);
  memcpy(out, v.data(), (size_t)((int64_t)v.size() < cap ? (int64_t)v.size() : cap));
  return (int64_t)v.size();
}

}  // extern "C"
