"""Generates tests/golden/* from the reference's own media and from the oracle.  Run in the build
container (needs /root/reference); the outputs are committed so nothing reads /root/reference at test time.

  book2_motion_blur_rgb8.npz : the 8-bit RGB raster of /root/reference/media/book2_motion_blur.png
                               (the reference's own render of random_scene, 384x216) — decoded, not re-rendered.
  c1_oracle_digest.json      : sha256 of the oracle's float64 framebuffer for config C1 in both math modes,
                               plus its deterministic counters (primary rays / segments).
  book1_animation_gif_frames.npz : frames 0, 1, 50, 199 of /root/reference/media/book1_animation.gif, decoded.
  c1_split32_digest.json     : the same for the split-stream mode (32 substreams per pixel) and its PSNR against the PNG.
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    png = np.asarray(Image.open("/root/reference/media/book2_motion_blur.png").convert("RGB"))
    assert png.shape == (216, 384, 3)
    np.savez_compressed(os.path.join(GOLD, "book2_motion_blur_rgb8.npz"), rgb8=png)

    world = O.random_scene()
    cam = O.book_camera()
    digest = {"config": "random_scene seed 0xFACADE, 384x216, 100 spp, depth 50, gamma float32(2.2)"}
    for math in ("libm", "det"):
        cnt = {}
        img = O.render(216, 384, 100, cam, world, math=math, counters=cnt)
        q = O.quantise_rgb8(img)
        digest[math] = {
            "f64_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
            "rgb8_sha256": hashlib.sha256(q.tobytes()).hexdigest(),
            "rgb8_equals_reference_png": bool(np.array_equal(q, png)),
            "counters": cnt,
        }
    with open(os.path.join(GOLD, "c1_oracle_digest.json"), "w") as f:
        json.dump(digest, f, indent=1)
    print(json.dumps(digest, indent=1))

    # the reference's own animation (media/book1_animation.gif: 256x144, 200 frames = t_max 6.0, 30 ms per frame):
    # four decoded frames, the weak golden (palette-quantised) that pins scenes_animated.nim — scene, physics steps and
    # camera orbit — in tests/test_oracle_pins.py
    gif = Image.open("/root/reference/media/book1_animation.gif")
    assert gif.size == (256, 144) and gif.n_frames == 200
    idx = [0, 1, 50, 199]
    frames = []
    for k in idx:
        gif.seek(k)
        frames.append(np.asarray(gif.convert("RGB")).copy())
    np.savez_compressed(os.path.join(GOLD, "book1_animation_gif_frames.npz"), index=np.array(idx), rgb8=np.stack(frames))

    # split-stream mode (TOR_MODE_FAST) at C1 with 32 sample ranges per pixel: digest of the oracle's restatement and
    # its distance to the reference's PNG (an independent 100-spp estimate of the same image)
    cnt = {}
    img = O.render_split(216, 384, 100, cam, world, 32, math="det", counters=cnt)
    q = O.quantise_rgb8(img).astype(np.float64)
    psnr = float(10.0 * np.log10(255.0**2 / np.mean((q - png.astype(np.float64)) ** 2)))
    split = {
        "config": digest["config"] + ", TOR_MODE_FAST with 32 substreams",
        "f64_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
        "counters": cnt,
        "psnr_vs_reference_png_db": psnr,
        "psnr_vs_reference_png_floor_db": float(np.floor(psnr - 0.5)),
    }
    with open(os.path.join(GOLD, "c1_split32_digest.json"), "w") as f:
        json.dump(split, f, indent=1)
    print(json.dumps(split, indent=1))


if __name__ == "__main__":
    main()
