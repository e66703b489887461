"""Developer tool: when do the cooperative warps and the lane-kernel warps of a C2 share start and end?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TOR_BVH_DEBUG_TIMES"] = "1"
import trace_of_radiance_b200 as T
step = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = T.Context()
world = T.random_scene(0xFACADE, 11).list()
cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
cv = T.newCanvas(675, 1200, 500, 2.2)
for _ in range(2):
    ctx.render(cv, cam, world, 50, rows=(0, 675, step))
print("kernel ms", ctx.last_kernel_ms(), ctx.last_schedule())
coop, lanes, tail = ctx.debug_times()
print('handoffs', ctx.last_handoffs())
c = coop[coop[:, 1] > 0].astype(np.int64)
l = lanes[lanes[:, 1] > 0].astype(np.int64)
t0 = min(c[:, 0].min() if len(c) else 1 << 62, l[:, 0].min())
print("lane warps", len(l), "start ms: min %.2f median %.2f max %.2f" % tuple((np.percentile(l[:, 0], q) - t0) / 1e6 for q in (0, 50, 100)),
      "end ms: p50 %.2f p90 %.2f p99 %.2f max %.2f" % tuple((np.percentile(l[:, 1], q) - t0) / 1e6 for q in (50, 90, 99, 100)))
if len(c):
    d = (c[:, 1] - c[:, 0]) / 1e3
    print("coop warps", len(c), "start ms: min %.2f max %.2f" % ((c[:, 0].min() - t0) / 1e6, (c[:, 0].max() - t0) / 1e6),
          "end ms: p50 %.2f max %.2f" % ((np.percentile(c[:, 1], 50) - t0) / 1e6, (c[:, 1].max() - t0) / 1e6),
          "segments: max %d" % c[:, 2].max(), "us/segment: median %.3f min %.3f max %.3f" % (np.median(d / np.maximum(1, c[:, 2])), (d / np.maximum(1, c[:, 2])).min(), (d / np.maximum(1, c[:, 2])).max()))
    k = np.argsort(-c[:, 1])[:5]
    for i in k:
        print("  late coop warp: start %.2f end %.2f segs %d us/seg %.3f" % ((c[i, 0] - t0) / 1e6, (c[i, 1] - t0) / 1e6, c[i, 2], (c[i, 1] - c[i, 0]) / 1e3 / max(1, c[i, 2])))
late = np.argsort(-l[:, 1])[:5]
for i in late:
    print("  late lane warp: start %.2f end %.2f" % ((l[i, 0] - t0) / 1e6, (l[i, 1] - t0) / 1e6))
tl = tail[tail[:, 1] > 0].astype(np.int64)
if len(tl):
    d = (tl[:, 1] - tl[:, 0]) / 1e3
    busy = tl[tl[:, 2] > 0]
    print("hand-off warps", len(tl), "with work", len(busy), "start ms: min %.2f max %.2f" % ((tl[:, 0].min() - t0) / 1e6, (tl[:, 0].max() - t0) / 1e6),
          "end ms: p50 %.2f p90 %.2f max %.2f" % tuple((np.percentile(busy[:, 1], q) - t0) / 1e6 for q in (50, 90, 100)) if len(busy) else "",
          "segments: total %d max %d" % (tl[:, 2].sum(), tl[:, 2].max()))
    for i in np.argsort(-tl[:, 1])[:5]:
        print("  late hand-off warp: start %.2f end %.2f segs %d us/seg %.3f" % ((tl[i, 0] - t0) / 1e6, (tl[i, 1] - t0) / 1e6, tl[i, 2], (tl[i, 1] - tl[i, 0]) / 1e3 / max(1, tl[i, 2])))
print("lane warp end histogram (ms):", np.histogram((l[:, 1] - t0) / 1e6, bins=12)[0].tolist(), "up to %.1f" % ((l[:, 1].max() - t0) / 1e6))
