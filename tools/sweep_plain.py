"""Developer tool: renders with few samples per pixel at several canvas sizes: scrambled queue, the same with the late
hand-off (TOR_BVH_HANDOFF_PLAIN), the cost-ranked path forced on, and the default policy; images compared with each
other bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import trace_of_radiance_b200 as T  # noqa: E402
from sweep_env import ctx_with  # noqa: E402

world = T.random_scene(0xFACADE, 11).list()
cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
sizes = [(144, 256, 100), (216, 384, 100), (216, 384, 32), (216, 384, 64), (675, 1200, 32), (675, 1200, 64), (270, 480, 100), (360, 640, 100), (450, 800, 100), (540, 960, 100), (675, 1200, 100), (1080, 1920, 100), (675, 1200, 255)]
ctxs = {name: ctx_with(env) for name, env in (
    ("plain", {"TOR_BVH_HANDOFF_PLAIN": 0, "TOR_BVH_PREPASS_FULL_SPP": 100000}),
    ("plain+handoff", {"TOR_BVH_PREPASS_FULL_SPP": 100000}), ("ranked", {"TOR_BVH_PREPASS_SPP": 16}), ("default", {}))}
for h, w, spp in sizes:
    ref, line = None, []
    for name, ctx in ctxs.items():
        ms = []
        for _ in range(4):
            cv = T.newCanvas(h, w, spp, 2.2)
            ctx.render(cv, cam, world, 50)
            ms.append(ctx.last_kernel_ms())
        if ref is None:
            ref = cv.pixels.tobytes()
        same = cv.pixels.tobytes() == ref
        line.append("%s %.3f%s (%d)" % (name, min(ms), "" if same else " WRONG", sum(ctx.last_handoffs().values())))
    print("%dx%d/%d  %.2f px/lane: " % (w, h, spp, h * w / 75776.0) + "  ".join(line), flush=True)
