"""Generates tests/golden/c2_oracle_digest.json: sha256 of the oracle's float64 framebuffer (tor_detmath build, the
one the CUDA path is bit-compared with) for config C2 — random_scene seed 0xFACADE, 1200x675, 500 spp, depth 50 —
with its deterministic counters and per-row digests.  One full CPU render of C2 (about 5 minutes on 8 cores); the
result pins the full-size image for tests/test_gpu_parity.py and for bench.py's image check at every GPU count."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402


def main():
    world = O.random_scene()
    cam = O.book_camera()
    h, w, spp = 675, 1200, 500
    cnt = {}
    t = time.time()
    img = O.render(h, w, spp, cam, world, math="det", counters=cnt)
    secs = time.time() - t
    q = O.quantise_rgb8(img)
    digest = {
        "config": "random_scene seed 0xFACADE, 1200x675, 500 spp, depth 50, gamma float32(2.2); oracle built with "
                  "tor_detmath.h",
        "f64_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
        "rgb8_sha256": hashlib.sha256(q.tobytes()).hexdigest(),
        "counters": cnt,
        "row_sha256_every_25": {str(r): hashlib.sha256(img[r].tobytes()).hexdigest() for r in range(0, h, 25)},
        "oracle_seconds": secs,
        "oracle_threads": O.num_threads(),
    }
    with open(os.path.join(ROOT, "tests", "golden", "c2_oracle_digest.json"), "w") as f:
        json.dump(digest, f, indent=1)
    print(json.dumps(digest)[:600])


if __name__ == "__main__":
    main()
