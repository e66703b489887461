"""Developer tool: split-stream mode (TOR_MODE_FAST) against the exact mode on one GPU -> gpurun_out/fast.json.
Kernel times (CUDA events inside the library) for C1, C2, the per-GPU row shares of C2 on 2/4/8 GPUs, a C4 frame, and
the substream-count sweep on C2."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T  # noqa: E402

A = T.api
out = {}
ctx = T.Context()
cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
world = T.random_scene().list()


def best_ms(cv, c, w, flags, rows=None, reps=3):
    ms = []
    for _ in range(reps):
        ctx.render(cv, c, w, 50, flags=flags, rows=rows)
        ms.append(ctx.last_kernel_ms())
    return min(ms)


def fast(n):
    return A.TOR_MODE_FAST | A.TOR_FAST_SUBSTREAMS(n)


for name, (h, w, spp) in {"c1": (216, 384, 100), "c2": (675, 1200, 500)}.items():
    cv = T.newCanvas(h, w, spp, 2.2)
    rays = h * w * spp
    r = {"exact_ms": best_ms(cv, cam, world, 0)}
    r["exact_mray_s"] = rays / r["exact_ms"] / 1e3
    for n in (0, 2, 4, 8, 16, 32):
        ms = best_ms(cv, cam, world, fast(n))
        r[f"fast_n{n if n else 'auto'}_ms"] = ms
        r[f"fast_n{n if n else 'auto'}_mray_s"] = rays / ms / 1e3
    out[name] = r
    print(name, json.dumps(r), flush=True)

# per-GPU share of C2 on G GPUs (rows g, g+G, ...): what strong scaling has to work with
h, w, spp = 675, 1200, 500
cv = T.newCanvas(h, w, spp, 2.2)
for G in (2, 4, 8):
    e = best_ms(cv, cam, world, 0, rows=(0, h, G), reps=2)
    f = best_ms(cv, cam, world, A.TOR_MODE_FAST, rows=(0, h, G), reps=2)
    out[f"c2_share_of_{G}_gpus"] = {"exact_ms": e, "fast_auto_ms": f, "rows": len(range(0, h, G))}
    print("share", G, json.dumps(out[f"c2_share_of_{G}_gpus"]), flush=True)

# C4: one animation frame (1 601 static spheres, 256x144, 100 spp)
cv = T.newCanvas(144, 256, 100, 2.2)
for i, (c, wld) in enumerate(T.Animation(height=144, width=256).scenes(skip=6)):
    if i == 10:
        r = {"exact_ms": best_ms(cv, c, wld, 0), "fast_auto_ms": best_ms(cv, c, wld, A.TOR_MODE_FAST)}
        for n in (4, 8, 16, 32):
            r[f"fast_n{n}_ms"] = best_ms(cv, c, wld, fast(n))
        out["c4_frame10"] = r
        print("c4", json.dumps(r), flush=True)
        break

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fast.json"), "w"), indent=1)
