"""Developer tool: kernel / call timings of BASELINE.json's other configs on one GPU -> gpurun_out/configs.json.
C1 and C2 in full; C4 on the first frames of the animation (per-frame tor_render calls, host buffers);
C5 (10 004 spheres, 3840x2160, 2000 spp) on every 20th row; and the object-count sweep of SURVEY.md 8(d)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T  # noqa: E402

out = {}
ctx = T.Context()
cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)


def timed(cv, c, w, depth=50, flags=0, rows=None, reps=2):
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        ctx.render(cv, c, w, depth, flags=flags, rows=rows)
        wall = time.perf_counter() - t
        ms = ctx.last_kernel_ms()
        if best is None or ms < best[0]:
            best = (ms, wall)
    return best


world = T.random_scene().list()
for name, (h, w, spp) in {"c1": (216, 384, 100), "c2": (675, 1200, 500)}.items():
    cv = T.newCanvas(h, w, spp, 2.2)
    ms, wall = timed(cv, cam, world)
    out[name] = {"kernel_ms": ms, "call_wall_ms": wall * 1e3, "mray_s_kernel": h * w * spp / ms / 1e3,
                 "mray_s_call": h * w * spp / wall / 1e6}
    print(name, out[name], flush=True)

# ---- C4: scenes_animated, 256x144, 100 spp, per-frame calls
frames = 24
cv = T.newCanvas(144, 256, 100, 2.2)
kms, t0 = [], None
for i, (c, w) in enumerate(T.Animation(height=144, width=256, t_max=9.0).scenes(skip=6)):
    if i == 2:
        t0 = time.perf_counter()  # first two frames warm up
    if i >= frames + 2:
        break
    ctx.render(cv, c, w, 50)
    kms.append(ctx.last_kernel_ms())
wall = time.perf_counter() - t0
out["c4"] = {"frames_timed": frames, "objects_per_frame": len(w), "kernel_ms_per_frame": sum(kms[2:]) / frames,
             "wall_ms_per_frame_incl_scene_step_bvh_h2d_d2h": wall / frames * 1e3,
             "mray_s_call": 144 * 256 * 100 * frames / wall / 1e6,
             "extrapolated_300_frames_s": wall / frames * 300, "scene_info": ctx.scene_info()}
print("c4", out["c4"], flush=True)

# ---- C4 through the frame pipeline (4 frames in flight, RGB8 output, pinned host buffers)
for inflight in (1, 2, 4, 8):
    t0 = time.perf_counter()
    n = T.render_animation(T.Animation(height=144, width=256, t_max=9.0), samples_per_pixel=100, in_flight=inflight,
                           max_frames=48)
    wall = time.perf_counter() - t0
    out[f"c4_pipeline_{inflight}_in_flight"] = {"frames": n, "wall_ms_per_frame": wall / n * 1e3,
                                                "mray_s": 144 * 256 * 100 * n / wall / 1e6,
                                                "extrapolated_300_frames_s": wall / n * 300}
    print("c4 pipeline", inflight, out[f"c4_pipeline_{inflight}_in_flight"], flush=True)

# ---- C5: 10 004 spheres, 3840x2160, 2000 spp, every 20th row
big = T.random_scene(0xFACADE, 50).list()
cv = T.newCanvas(2160, 3840, 2000, 2.2)
rows = (0, 2160, 20)
ms, wall = timed(cv, cam, big, rows=rows, reps=1)
nrows_sel = len(range(*rows))
out["c5_sample"] = {"objects": len(big), "rows": nrows_sel, "of_rows": 2160, "kernel_ms": ms,
                    "mray_s_kernel": nrows_sel * 3840 * 2000 / ms / 1e3,
                    "extrapolated_full_frame_s_1gpu": ms / 1e3 * 2160 / nrows_sel, "scene_info": ctx.scene_info()}
print("c5", out["c5_sample"], flush=True)

# ---- object-count sweep at 1200x675 / 50 spp
sweep = []
for half in (11,) if "--quick" in sys.argv else (0, 1, 2, 4, 8, 11, 16, 23, 32, 50):
    w = T.random_scene(0xFACADE, half).list()
    cv = T.newCanvas(675, 1200, 50, 2.2)
    row = {"half": half, "objects": len(w)}
    ctx.render(cv, cam, w, 50, flags=T.api.TOR_FLAG_COUNT_SEGMENTS)
    cnt = ctx.counters()
    ms, _ = timed(cv, cam, w, reps=1)
    row.update(bvh_kernel_ms=ms, bvh_mray_s=675 * 1200 * 50 / ms / 1e3, segments=cnt["segments"],
               node_visits_per_segment=cnt["bvh_node_visits"] / max(1, cnt["segments"]),
               sphere_tests_per_segment=cnt["bvh_sphere_tests"] / max(1, cnt["segments"]))
    if len(w) <= 2200:
        ms, _ = timed(cv, cam, w, flags=T.api.TOR_FLAG_BRUTE_FORCE, reps=1)
        row.update(brute_kernel_ms=ms, brute_mray_s=675 * 1200 * 50 / ms / 1e3,
                   brute_sphere_tests_per_s=cnt["segments"] * len(w) / ms * 1e3)
    sweep.append(row)
    print(row, flush=True)
out["object_count_sweep_1200x675x50spp"] = sweep
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
