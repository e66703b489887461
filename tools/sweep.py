"""Developer tool: time the C2 render kernel (device time) for the current env knobs."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trace_of_radiance_b200 as T
h, w, spp = (675, 1200, 500) if "--c1" not in sys.argv else (216, 384, 100)
reps = 3
if "--dims" in sys.argv:
    i = sys.argv.index("--dims")
    h, w, spp, reps = (int(x) for x in sys.argv[i + 1:i + 5])
ctx = T.Context()
half = int(sys.argv[sys.argv.index("--half") + 1]) if "--half" in sys.argv else 11  # 50 = the 10 002-sphere scene of C5
world = T.random_scene(0xFACADE, half).list()
cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
fl = T.api.TOR_FLAG_BRUTE_FORCE if "--brute" in sys.argv else 0
if "--rowmajor" in sys.argv:
    fl |= T.api.TOR_FLAG_ROW_MAJOR_QUEUE
if "--fast" in sys.argv:  # split-stream mode; optional substream count after the flag
    i = sys.argv.index("--fast")
    n = int(sys.argv[i + 1]) if len(sys.argv) > i + 1 and sys.argv[i + 1].isdigit() else 0
    fl |= T.api.TOR_MODE_FAST | T.api.TOR_FAST_SUBSTREAMS(n)
cv = T.newCanvas(h, w, spp, 2.2)
rows = None
if "--rowstep" in sys.argv:
    rows = (0, h, int(sys.argv[sys.argv.index("--rowstep") + 1]))
ms = []
for _ in range(reps):
    ctx.render(cv, cam, world, 50, flags=fl, rows=rows)
    ms.append(ctx.last_kernel_ms())
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("TOR_")}, "kernel_ms": ms, "rows": rows,
                  "mray_s": h * w * spp / min(ms) / 1e3}), flush=True)
