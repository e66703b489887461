"""Developer tool: kernel time of C2 and of its per-GPU shares (every 2nd / 4th / 8th row on ONE GPU = the load one
rank has in a 2 / 4 / 8-GPU render) for several settings of the warp-cooperative pixel policy.  Knobs are read when a
context is created, so one process compares them.  Writes gpurun_out/<tag>_coop.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T  # noqa: E402

TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
QUICK = "--quick" in sys.argv


def ctx_with(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    c = T.Context()
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return c


def main():
    world = T.random_scene(0xFACADE, 11).list()
    cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
    h, w, spp = 675, 1200, 500
    cv = T.newCanvas(h, w, spp, 2.2)
    settings = [("off", {"TOR_BVH_COOP_MAX": 0})]
    if QUICK:
        settings.append(("default", {}))
    else:
        for fast, fw, mx in ((0, 8, 15), (50, 8, 15), (50, 4, 15), (30, 4, 15), (70, 8, 15), (50, 8, 20), (100, 8, 20), (50, 6, 18)):
            settings.append((f"fast{fast}_fw{fw}_max{mx}", {"TOR_BVH_COOP_FAST": fast, "TOR_BVH_COOP_FAST_WARPS": fw,
                                                            "TOR_BVH_COOP_MAX": mx}))
    out = {}
    for name, env in settings:
        ctx = ctx_with(env)
        for step in (8, 4, 2, 1):
            ms = []
            for _ in range(3):
                ctx.render(cv, cam, world, 50, rows=(0, h, step))
                ms.append(ctx.last_kernel_ms())
            sched = ctx.last_schedule()
            out[f"{name}:step{step}"] = {"kernel_ms": ms, "min_ms": min(ms), "sched": sched}
            print(name, "step", step, ["%.2f" % m for m in ms], sched, flush=True)
        ctx.close()
    # the whole of C1 through the cooperative route (cost per segment of a cooperative warp) against the lanes
    c1 = T.newCanvas(216, 384, 100, 2.2)
    for name, env in (("c1_lanes", {}), ("c1_all_coop", {"TOR_BVH_PREPASS_SPP": 9, "TOR_BVH_COOP_FORCE": 10 ** 9})):
        ctx = ctx_with(env)
        ms = []
        for _ in range(3):
            ctx.render(c1, cam, world, 50, flags=T.api.TOR_FLAG_COUNT_SEGMENTS)
            ms.append(ctx.last_kernel_ms())
        cnt = ctx.counters()
        out[name] = {"kernel_ms": ms, "counters": cnt, "sched": ctx.last_schedule()}
        print(name, ms, cnt, flush=True)
        ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{TAG}_coop.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
