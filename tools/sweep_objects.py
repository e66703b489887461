"""Developer tool: the object-count sweep of SURVEY.md 8(d) at ONE stated resolution (1200x675, 100 spp, depth 50, the
book camera): scenes of the scenes.nim recipe with 1 .. 10 002 objects, exact and split-stream mode, BVH route (and the
brute-force route where it finishes quickly).  Per scene: Mray/s, bounce segments per primary ray, node visits and
sphere tests per segment, what the kernel staged in shared memory.  Writes gpurun_out/<tag>_object_sweep.json and
prints a markdown table for DESIGN.md."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T  # noqa: E402

TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
H, W, SPP = 675, 1200, 100
ctx = T.Context()
cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
scenes = [("1 (the ground sphere)", T.Scene(T.random_scene(0xFACADE, 0).list().objects[:1]).list())]
for half in (0, 1, 2, 4, 8, 11, 22, 50):
    w = T.random_scene(0xFACADE, half).list()
    scenes.append((f"{len(w)} (grid -{half}..<{half})", w))
out, rows = {}, []
cv = T.newCanvas(H, W, SPP, 2.2)
for name, world in scenes:
    rec = {"objects": len(world)}
    for mode, fl in (("exact", 0), ("split", T.api.TOR_MODE_FAST)):
        ms = []
        for _ in range(3):
            ctx.render(cv, cam, world, 50, flags=fl)
            ms.append(ctx.last_kernel_ms())
        ctx.render(cv, cam, world, 50, flags=fl | T.api.TOR_FLAG_COUNT_SEGMENTS)
        cnt = ctx.counters()
        rec[mode] = {"kernel_ms": min(ms), "mray_s": H * W * SPP / min(ms) / 1e3,
                     "segments_per_ray": cnt["segments"] / cnt["primary_rays"],
                     "node_visits_per_segment": cnt["bvh_node_visits"] / cnt["segments"],
                     "sphere_tests_per_segment": cnt["bvh_sphere_tests"] / cnt["segments"]}
    if len(world) <= 500:
        ms = []
        for _ in range(2):
            ctx.render(cv, cam, world, 50, flags=T.api.TOR_FLAG_BRUTE_FORCE)
            ms.append(ctx.last_kernel_ms())
        rec["brute_force"] = {"kernel_ms": min(ms), "mray_s": H * W * SPP / min(ms) / 1e3}
    rec["scene_info"] = ctx.scene_info()
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
    e, s = rec["exact"], rec["split"]
    rows.append(f"| {name} | {rec['scene_info']['bvh_nodes']} | {rec['scene_info']['bvh_bytes'] / 1024:.0f} KB | "
                f"{e['mray_s']:.0f} | {s['mray_s']:.0f} | {rec.get('brute_force', {}).get('mray_s', float('nan')):.0f} | "
                f"{e['segments_per_ray']:.2f} | {e['node_visits_per_segment']:.1f} | {e['sphere_tests_per_segment']:.2f} |")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{TAG}_object_sweep.json"), "w"), indent=1)
print("| objects | BVH nodes | blob | exact Mray/s | split-stream Mray/s | brute force Mray/s | segments / ray | node visits / segment | "
      "sphere tests / segment |")
print("|---|---|---|---|---|---|---|---|---|")
print("\n".join(rows))
