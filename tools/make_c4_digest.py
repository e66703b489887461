"""Generates tests/golden/c4_rgb8_digest.json on a GPU box: sha256 over the RGB8 frames of config C4 (scenes_animated,
256x144 / 100 spp / 300 frames) from the device-resident pipeline, after checking frames 0, 150 and 299 byte for byte
against the CPU oracle's render of the oracle's own scene iterator (the whole animation on the CPU would take hours:
4.6e12 sphere tests).  bench.py --workload c4 compares its frames with this digest at every GPU count."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import trace_of_radiance_b200 as T  # noqa: E402

h, w, spp = 144, 256, 100
ctx = T.Context()
an = T.DeviceAnimation(ctx, height=h, width=w, t_max=9.0, in_flight=8)
frames = np.zeros((300, h, w, 3), dtype=np.uint8)
n, ms = an.render_all(samples_per_pixel=spp, on_frame=lambda i, rgb: frames.__setitem__(i, rgb))
assert n == 300
ref = O.Animation(height=h, width=w, t_max=9.0)
checked = {}
for i in range(300):
    cam, objs = ref.next_frame(skip=6)
    if i in (0, 150, 299):
        want = O.quantise_rgb8(O.render(h, w, spp, cam, objs, math="det"))
        assert np.array_equal(frames[i], want), i
        checked[str(i)] = hashlib.sha256(want.tobytes()).hexdigest()
assert ref.next_frame(skip=6) is None
out = {"config": "scenes_animated random_moving_spheres seed 0xFACADE, 256x144, 100 spp, depth 50, 300 frames (dt 0.005, "
                 "skip 6, t_max 9.0), RGB8 in PPM row order", "rgb8_sha256_all_frames": hashlib.sha256(frames.tobytes()).hexdigest(),
       "frames_equal_to_the_cpu_oracle": checked, "device_ms": ms}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c4_rgb8_digest.json"), "w"), indent=1)
print(json.dumps(out))
