#!/bin/bash
# compute-sanitizer memcheck over a small render on both routes + the multi-device context test
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import trace_of_radiance_b200 as T
ctx = T.Context()
world = T.random_scene().list()
cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
for fl in (0, T.api.TOR_FLAG_BRUTE_FORCE, T.api.TOR_FLAG_ROW_MAJOR_QUEUE):
    cv = T.newCanvas(24, 40, 64, 2.2)
    ctx.render(cv, cam, world, 50, flags=fl | T.api.TOR_FLAG_COUNT_SEGMENTS)
    print(fl, ctx.counters(), float(cv.pixels.sum()))
big = T.random_scene(0xFACADE, 50).list()
cv = T.newCanvas(12, 20, 4, 2.2)
ctx.render(cv, cam, big, 50)
print("rgb8", ctx.render_rgb8(cv, cam, world, 50).sum())
for n in (1, 4, 32):  # split-stream mode: (pixel, range) units + substream_reduce_kernel
    cv = T.newCanvas(24, 40, 10, 2.2)
    ctx.render(cv, cam, world, 50, flags=T.api.TOR_MODE_FAST | T.api.TOR_FAST_SUBSTREAMS(n))
    print("split", n, float(cv.pixels.sum()))
# cost-ranked path with cooperative warps and the late hand-off forced onto a small render (records parked by lanes
# and by cooperative warps, second launch of render_coop_kernel), and the plain path's hand-off
for env in ({"TOR_BVH_PREPASS_SPP": "9", "TOR_BVH_COOP_FORCE": "40", "TOR_BVH_HANDOFF": "1", "TOR_BVH_HANDOFF_MIN_LEFT": "2"},
            {"TOR_BVH_PREPASS_SPP": "9", "TOR_BVH_COOP_FORCE": "1000000000", "TOR_BVH_HANDOFF": "1"},
            {"TOR_BVH_HANDOFF_PLAIN": "5", "TOR_BVH_HANDOFF_MIN_LEFT": "1"}):
    os.environ.update(env)
    c2 = T.Context()
    for k in env:
        del os.environ[k]
    ref = T.newCanvas(36, 64, 24, 2.2)
    ctx.render(ref, cam, world, 50, flags=T.api.TOR_FLAG_ROW_MAJOR_QUEUE)
    for _ in range(2):
        cv = T.newCanvas(36, 64, 24, 2.2)
        c2.render(cv, cam, world, 50)
        print("handoff", sorted(env), c2.last_schedule()["cooperative_pixels"], c2.last_handoffs(), cv.pixels.tobytes() == ref.pixels.tobytes())
    c2.close()
cv = T.newCanvas(24, 40, 4, 2.2)
y, cb, cr = ctx.render_ycbcr420(cv, cam, world, 50, flags=T.api.TOR_MODE_FAST)  # ycbcr420_kernel
print("ycbcr", int(y.sum()), int(cb.sum()), int(cr.sum()))
PY
compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -14 gpurun_out/sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 3 python /tmp/san.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/sanitizer_racecheck.log
