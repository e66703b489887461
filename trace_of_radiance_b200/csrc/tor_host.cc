// tor_host.cc — host-side helpers of include/tor_b200.h that restate the reference's constructors and
// output format (camera(), random_scene(), exportToPPM).  They are inputs / outputs either side of the
// render path, not the path itself; nothing here touches the GPU.
//
// Built with -ffp-contract=off so the float64 results are the reference's (scalar SSE2, no FMA).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/tor_b200.h"
#include "tor_anim.hpp"

namespace {

using namespace tor_anim;

int ppm_level(double c) {  // io/ppm.nim:15-16 + safe_math.nim:10-14; NaN (UB in the reference) -> 0
  if (c != c) return 0;
  double cl = c < 0.0 ? 0.0 : (c > 0.999 ? 0.999 : c);
  return (int)(256 * cl);
}

}  // namespace

extern "C" {

// physics/cameras.nim:24-45
void tor_camera_make(tor_camera* out, const double look_from[3], const double look_at[3], const double view_up[3],
                     double vfov_degrees, double aspect_ratio, double aperture, double focus_distance,
                     double shutter_open, double shutter_close) {
  camera_make(out, look_from, look_at, view_up, vfov_degrees, aspect_ratio, aperture, focus_distance, shutter_open,
              shutter_close);
}

// scenes.nim:13-50 (draw order: SURVEY.md appendix B)
int64_t tor_random_scene(uint64_t seed, int32_t half, tor_hittable* out, int64_t cap) {
  HostRng rng(seed);
  std::vector<tor_hittable> w;
  w.push_back(mk_sphere(V{0, -1000, 0}, 1000, TOR_LAMBERTIAN, V{0.5, 0.5, 0.5}, 0));
  for (int a = -half; a < half; ++a)
    for (int b = -half; b < half; ++b) {
      double cx = (double)a + 0.9 * rng.u01();
      double cz = (double)b + 0.9 * rng.u01();
      V center{cx, 0.2, cz};
      if (len(sub(center, V{4, 0.2, 0})) > 0.9) {
        double choose = rng.u01();
        if (choose < 0.8) {  // diffuse: a *moving* sphere at this commit (scenes.nim:32-36)
          V a1{0, 0, 0}, a2{0, 0, 0};
          a1.x = rng.u01(); a1.y = rng.u01(); a1.z = rng.u01();
          a2.x = rng.u01(); a2.y = rng.u01(); a2.z = rng.u01();
          tor_hittable h = mk_sphere(center, 0.2, TOR_LAMBERTIAN, V{a1.x * a2.x, a1.y * a2.y, a1.z * a2.z}, 0);
          h.kind = TOR_MOVING_SPHERE;
          put(h.center1, V{center.x + 0, center.y + rng.umax(0.5), center.z + 0});
          h.time0 = 0.0;
          h.time1 = 1.0;
          w.push_back(h);
        } else if (choose < 0.95) {
          V albedo{0, 0, 0};
          albedo.x = rng.urange(0.5, 1); albedo.y = rng.urange(0.5, 1); albedo.z = rng.urange(0.5, 1);
          double fuzz = rng.umax(0.5);
          w.push_back(mk_sphere(center, 0.2, TOR_METAL, albedo, fuzz <= 1.0 ? fuzz : 1.0));  // materials.nim:35-37
        } else {
          w.push_back(mk_sphere(center, 0.2, TOR_DIELECTRIC, V{0, 0, 0}, 1.5));
        }
      }
    }
  w.push_back(mk_sphere(V{0, 1, 0}, 1.0, TOR_DIELECTRIC, V{0, 0, 0}, 1.5));
  w.push_back(mk_sphere(V{-4, 1, 0}, 1.0, TOR_LAMBERTIAN, V{0.4, 0.2, 0.1}, 0));
  w.push_back(mk_sphere(V{4, 1, 0}, 1.0, TOR_METAL, V{0.7, 0.6, 0.5}, 0.0));
  int64_t n = (int64_t)w.size();
  if (out)
    for (int64_t i = 0; i < n && i < cap; ++i) out[i] = w[(size_t)i];
  return n;
}

// scenes_animated.nim:90-154 `random_moving_spheres` with rng.seed(seed) (trace_of_radiance_animation.nim:61-63)
void* tor_animation_create(uint64_t seed, int32_t height, int32_t width, float dt, float t_min, float t_max) {
  return animation_create(seed, height, width, dt, t_min, t_max);
}

void tor_animation_destroy(void* h) { delete (Animation*)h; }

// One iteration of `iterator scenes(anim, skip)` (scenes_animated.nim:176-225): the frame is produced, then the
// physics advances `skip` steps.  Returns the object count of the frame (written to out[0 .. min(count, cap))),
// or 0 when t >= t_max (the iterator is exhausted).
int64_t tor_animation_next_frame(void* h, int32_t skip, tor_camera* cam, tor_hittable* out, int64_t cap) {
  if (!h || !cam) return TOR_ERR_INVALID_ARG;
  Animation& an = *(Animation*)h;
  if (!an.started) {
    while (an.t < an.t_min) anim_step(an);
    an.started = true;
  } else {
    for (int i = 0; i < skip; ++i) anim_step(an);
  }
  if (!(an.t < an.t_max)) return 0;
  animation_camera(an, cam);
  std::vector<tor_hittable> w;
  animation_scene(an, &w);
  const int64_t n = (int64_t)w.size();
  if (out)
    for (int64_t i = 0; i < n && i < cap; ++i) out[i] = w[(size_t)i];
  return n;
}

// io/ppm.nim:14-27
int tor_quantise_rgb8(const tor_canvas* canvas, uint8_t* out) {
  if (!canvas || !canvas->pixels || !out) return TOR_ERR_INVALID_ARG;
  int64_t k = 0;
  for (int32_t i = canvas->nrows - 1; i >= 0; --i)
    for (int32_t j = 0; j < canvas->ncols; ++j) {
      const double* p = canvas->pixels + 3 * ((int64_t)i * canvas->ncols + j);
      out[k++] = (uint8_t)ppm_level(p[0]);
      out[k++] = (uint8_t)ppm_level(p[1]);
      out[k++] = (uint8_t)ppm_level(p[2]);
    }
  return TOR_OK;
}

int tor_export_ppm(const tor_canvas* canvas, const char* path) {
  if (!canvas || !canvas->pixels || !path) return TOR_ERR_INVALID_ARG;
  FILE* f = fopen(path, "wb");
  if (!f) return TOR_ERR_INVALID_ARG;
  fprintf(f, "P3\n%d %d\n255\n", canvas->ncols, canvas->nrows);
  std::string line;
  for (int32_t i = canvas->nrows - 1; i >= 0; --i) {
    line.clear();
    for (int32_t j = 0; j < canvas->ncols; ++j) {
      const double* p = canvas->pixels + 3 * ((int64_t)i * canvas->ncols + j);
      char buf[48];
      int n = snprintf(buf, sizeof(buf), "%d %d %d\n", ppm_level(p[0]), ppm_level(p[1]), ppm_level(p[2]));
      line.append(buf, (size_t)n);
    }
    fwrite(line.data(), 1, line.size(), f);
  }
  fclose(f);
  return TOR_OK;
}

}  // extern "C"
