"""Quick GPU sanity run (developer tool): small parity check vs the oracle + kernel timings, both routes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import trace_of_radiance_b200 as T  # noqa: E402


def main():
    out = {}
    ctx = T.Context()
    scene = T.random_scene()
    world = scene.list()
    cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
    routes = {"bvh": 0, "brute": T.api.TOR_FLAG_BRUTE_FORCE}
    for (h, w, spp) in [(36, 64, 10), (216, 384, 100)]:
        ocnt = {}
        t = time.time()
        ref = O.render(h, w, spp, cam.as_array(), world.objects, math="det", counters=ocnt)
        owall = time.time() - t
        for name, fl in routes.items():
            cv = T.newCanvas(h, w, spp, 2.2)
            t = time.time()
            ctx.render(cv, cam, world, 50, flags=T.api.TOR_FLAG_COUNT_SEGMENTS | fl)
            wall = time.time() - t
            ms = ctx.last_kernel_ms()
            cnt = ctx.counters()
            same = ref.tobytes() == cv.pixels.tobytes()
            ndiff = int((ref != cv.pixels).sum())
            key = f"{name}:{w}x{h}x{spp}"
            out[key] = dict(kernel_ms=ms, wall_s=wall, oracle_s=owall, bit_exact=same, ndiff=ndiff, counters=cnt,
                            oracle_counters=ocnt, mray_s=h * w * spp / ms / 1e3)
            print(key, json.dumps(out[key]), flush=True)
    print("scene_info", ctx.scene_info(), flush=True)
    if "--c2" in sys.argv:
        imgs = {}
        for name, fl in routes.items():
            cv = T.newCanvas(675, 1200, 500, 2.2)
            t = time.time()
            ctx.render(cv, cam, world, 50, flags=fl)
            wall = time.time() - t
            ms = ctx.last_kernel_ms()
            out["c2:" + name] = dict(kernel_ms=ms, wall_s=wall, mray_s=675 * 1200 * 500 / ms / 1e3)
            print("c2:" + name, json.dumps(out["c2:" + name]), flush=True)
            imgs[name] = cv.pixels.copy()
        out["c2:routes_identical"] = imgs["bvh"].tobytes() == imgs["brute"].tobytes()
        print("c2 routes identical:", out["c2:routes_identical"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_quick.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
