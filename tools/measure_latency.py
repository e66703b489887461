"""Developer tool: per-segment latency of one lane alone in its warp (TOR_BVH_LANES=1) and of a cooperative warp, on a
render small enough that the GPU is nearly idle (2 368 pixels = 10 CTAs = 80 warps; two warps per scheduler)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T  # noqa: E402
from tools.measure_coop import ctx_with  # noqa: E402


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    world = T.random_scene(0xFACADE, 11).list()
    cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
    h, w, spp = 37, 64, 64
    out = {}
    for name, env in (("solo_lane", {"TOR_BVH_LANES": 1, "TOR_BVH_PREPASS_SPP": 1 << 20}),
                      ("full_warps", {"TOR_BVH_PREPASS_SPP": 1 << 20}),
                      ("coop_warp", {"TOR_BVH_PREPASS_SPP": 9, "TOR_BVH_COOP_FORCE": 10 ** 9})):
        if only and name != only:
            continue
        ctx = ctx_with(env)
        cv = T.newCanvas(h, w, spp, 2.2)
        ms = []
        for _ in range(3):
            ctx.render(cv, cam, world, 50, flags=T.api.TOR_FLAG_COUNT_SEGMENTS)
            ms.append(ctx.last_kernel_ms())
        cnt = ctx.counters()
        segs = cnt["segments"] / 3
        warps = 80
        units = warps if name != "full_warps" else warps * 32
        out[name] = {"kernel_ms": ms, "segments": segs, "us_per_segment_per_unit": min(ms) * 1e3 * units / segs,
                     "sched": ctx.last_schedule()}
        print(name, json.dumps(out[name]), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
