"""Developer tool: one process per setting of the host BVH builder's knobs (TOR_BVH_LEAF, TOR_BVH_ISECT are read once
per process): kernel time of the C2 scene (485 objects) and of the C5 scene (10 002) at 1200x675 / 100 spp."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import trace_of_radiance_b200 as T
    ctx = T.Context()
    cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
    out = []
    for half in (11, 50):
        world = T.random_scene(0xFACADE, half).list()
        cv = T.newCanvas(675, 1200, 100, 2.2)
        ms = []
        for _ in range(3):
            ctx.render(cv, cam, world, 50)
            ms.append(ctx.last_kernel_ms())
        ctx.render(cv, cam, world, 50, flags=T.api.TOR_FLAG_COUNT_SEGMENTS)
        c = ctx.counters()
        out.append("%d objs %.2f ms (%.1f visits, %.2f tests per segment)" % (len(world), min(ms), c["bvh_node_visits"] / c["segments"], c["bvh_sphere_tests"] / c["segments"]))
    print("LEAF=%s ISECT=%s: " % (os.environ.get("TOR_BVH_LEAF", "4"), os.environ.get("TOR_BVH_ISECT", "0.7")) + "; ".join(out), flush=True)
else:
    for leaf, isect in (("4", "0.7"), ("2", "0.7"), ("3", "0.7"), ("6", "0.7"), ("8", "0.7"), ("4", "0.4"), ("4", "1.0"), ("4", "1.5"),
                        ("6", "1.0"), ("8", "1.5"), ("3", "0.4"), ("2", "0.4")):
        env = dict(os.environ, TOR_BVH_LEAF=leaf, TOR_BVH_ISECT=isect)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env)
