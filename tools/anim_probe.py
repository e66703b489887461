"""Developer tool: C4 frame pipeline timings for several in-flight counts and flag choices."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T
for inflight in (1, 2, 3, 4):
    for fl in (0, T.api.TOR_FLAG_FULL_WARPS):
        t0 = time.perf_counter()
        n = T.render_animation(T.Animation(height=144, width=256, t_max=9.0), samples_per_pixel=100, in_flight=inflight,
                               max_frames=36, flags=fl)
        wall = time.perf_counter() - t0
        print(f"in_flight {inflight} flags {fl:#x}: {wall / n * 1e3:.2f} ms/frame", flush=True)
# host-only cost of a frame: scene step + packing + BVH build (no GPU work is needed for this part)
an = T.Animation(height=144, width=256, t_max=9.0)
t0 = time.perf_counter(); k = 0
for cam, world in an.scenes(skip=6):
    k += 1
    if k == 50: break
print(f"scene iterator: {(time.perf_counter() - t0) / k * 1e3:.3f} ms/frame", flush=True)
ctx = T.Context()
t0 = time.perf_counter()
for _ in range(20):
    ctx.scene_upload(cam, world)
print(f"scene_upload (pack + BVH build + H2D + sync): {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms", flush=True)
