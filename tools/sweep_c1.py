"""Developer tool: kernel time of C1 (384x216 / 100 spp, about one pixel per lane) under TOR_* knobs, with the frozen
digest of the oracle's framebuffer checked for every setting.  python tools/sweep_c1.py name=KNOB:VALUE,... ..."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import trace_of_radiance_b200 as T  # noqa: E402
from sweep_env import ctx_with  # noqa: E402

want = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_oracle_digest.json")))["det"]["f64_sha256"]
world = T.random_scene(0xFACADE, 11).list()
cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
for a in sys.argv[1:]:
    name, _, kv = a.partition("=")
    env = dict(item.split(":") for item in kv.split(",") if item)
    ctx = ctx_with(env)
    ms = []
    for _ in range(5):
        cv = T.newCanvas(216, 384, 100, 2.2)
        ctx.render(cv, cam, world, 50)
        ms.append(ctx.last_kernel_ms())
    ok = hashlib.sha256(cv.pixels.tobytes()).hexdigest() == want
    print("%-24s min %.3f ms  %s  digest %s  parked %s" % (name, min(ms), ["%.2f" % m for m in ms], "ok" if ok else "WRONG", ctx.last_handoffs()), flush=True)
    ctx.close()
