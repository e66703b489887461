"""Developer tool: config C4 (scenes_animated.nim, 256x144 / 100 spp / 300 frames) end to end on one GPU, exact and
split-stream mode, through the frame pipeline (RGB8 frames to pinned host memory) and through the MP4 export path
-> gpurun_out/c4.json."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T  # noqa: E402

A = T.api
out = {}
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300
T.render_animation(T.Animation(height=144, width=256), samples_per_pixel=100, in_flight=2, max_frames=8)  # warm-up
for mode, fl in (("exact", 0), ("split_stream", A.TOR_MODE_FAST)):
    for inflight in (1, 2, 4):
        t0 = time.perf_counter()
        n = T.render_animation(T.Animation(height=144, width=256, t_max=9.0), samples_per_pixel=100, in_flight=inflight,
                               max_frames=frames, flags=fl)
        wall = time.perf_counter() - t0
        out[f"{mode}_in_flight_{inflight}"] = {"frames": n, "wall_s": wall, "ms_per_frame": wall / n * 1e3,
                                                "mray_s": 144 * 256 * 100 * n / wall / 1e6}
        print(mode, inflight, json.dumps(out[f"{mode}_in_flight_{inflight}"]), flush=True)
with tempfile.TemporaryDirectory() as d:
    for mode, fl in (("exact", 0), ("split_stream", A.TOR_MODE_FAST)):
        t0 = time.perf_counter()
        n = T.render_animation_mp4(T.Animation(height=144, width=256, t_max=9.0), os.path.join(d, "a.264"),
                                   os.path.join(d, "a.mp4"), samples_per_pixel=100, max_frames=frames, flags=fl)
        wall = time.perf_counter() - t0
        out[f"{mode}_mp4"] = {"frames": n, "wall_s": wall, "ms_per_frame": wall / n * 1e3,
                              "mp4_bytes": os.path.getsize(os.path.join(d, "a.mp4"))}
        print(mode, "mp4", json.dumps(out[f"{mode}_mp4"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c4.json"), "w"), indent=1)
