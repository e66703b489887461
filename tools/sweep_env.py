"""Developer tool: kernel time of C2 and of its per-GPU shares (every 2nd / 4th / 8th row on ONE GPU = the load one
rank has in a 2 / 4 / 8-GPU render) for any settings of the TOR_* tuning knobs (read once per context, so one process
compares them).

  python tools/sweep_env.py TAG [--steps 8,4,2,1] [--reps 3] [--split] name=KNOB:VALUE,KNOB:VALUE ...

`name=` alone is the default policy.  Writes gpurun_out/<TAG>_sweep.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T  # noqa: E402


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
    args = sys.argv[1:]
    tag = args.pop(0)
    steps, reps, split = (8, 4, 2, 1), 3, False
    settings = []
    while args:
        a = args.pop(0)
        if a == "--steps":
            steps = tuple(int(x) for x in args.pop(0).split(","))
        elif a == "--reps":
            reps = int(args.pop(0))
        elif a == "--split":
            split = True
        else:
            name, _, kv = a.partition("=")
            env = {}
            for item in filter(None, kv.split(",")):
                k, _, v = item.partition(":")
                env[k] = v
            settings.append((name, env))
    world = T.random_scene(0xFACADE, 11).list()
    cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
    h, w, spp = 675, 1200, 500
    cv = T.newCanvas(h, w, spp, 2.2)
    out = {}
    for name, env in settings:
        ctx = ctx_with(env)
        line = []
        for step in steps:
            ms = []
            for _ in range(reps):
                ctx.render(cv, cam, world, 50, rows=(0, h, step), flags=T.api.TOR_MODE_FAST if split else 0)
                ms.append(ctx.last_kernel_ms())
            out[f"{name}:step{step}"] = {"env": env, "kernel_ms": ms, "min_ms": min(ms), "sched": ctx.last_schedule()}
            line.append("step%d %.2f" % (step, min(ms)))
        print("%-28s %s" % (name, "  ".join(line)), flush=True)
        ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
