"""The C-ABI library loads without a GPU and exports every symbol include/tor_b200.h declares; host-side
helpers (camera(), random_scene(), PPM) agree with the oracle; compute entry points fail loudly (no CPU
fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "tor_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tor_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(tor):
    L = tor.load_library()
    names = _declared_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tor_b200.h but not exported"
    assert sorted(tor.api.EXPORTED_SYMBOLS) == names
    assert L.tor_abi_version() == 1


def test_struct_layouts_match_the_reference_objects(tor):
    # canvas.nim:21-28 is 24 bytes; cameras.nim:15-22 is 24 float64; flat hittable 112 bytes
    assert C.sizeof(tor.api._CCanvas) == 24
    assert C.sizeof(tor.api._CCamera) == 192
    assert tor.HITTABLE_DTYPE.itemsize == 112


def test_camera_helper_equals_oracle(tor, oracle):  # cameras.nim:24-45
    cam = tor.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
    assert cam.as_array().tobytes() == oracle.book_camera().tobytes()
    cam2 = tor.camera((-3, 4, 9), (4, 1, 0), (0, 1, 0), 35.0, 256 / 144, 0.0, 7.5)
    assert cam2.as_array().tobytes() == oracle.camera((-3, 4, 9), (4, 1, 0), (0, 1, 0), 35.0, 256 / 144, 0.0, 7.5).tobytes()


def test_random_scene_helper_equals_oracle(tor, oracle):  # scenes.nim:13-50
    for seed, half in [(0xFACADE, 11), (1, 3), (42, 50)]:
        mine = tor.random_scene(seed, half).list().objects
        assert mine.tobytes() == oracle.random_scene(seed, half).tobytes()
    assert len(tor.random_scene(0xFACADE, 50)) > 9000


def test_scene_builders(tor, oracle):  # spheres.nim:20-26, moving_spheres.nim:22-37, materials.nim:21,35,52
    s = tor.Scene()
    s.add(tor.sphere((0, -1000, 0), 1000, tor.lambertian((0.5, 0.5, 0.5))))
    s.add(tor.movingSphere((1, 0.2, 2), 0.0, (1, 0.5, 2), 1.0, 0.2, tor.metal((0.9, 0.8, 0.7), 3.0)))
    s.add(tor.sphere((0, 1, 0), 1.0, tor.dielectric(1.5)))
    w = s.list()
    assert len(w) == 3
    assert w.objects[1]["fuzz_or_ior"] == 1.0  # metal() clamps fuzz
    assert w.objects[1]["kind"] == 1 and w.objects[2]["mat_kind"] == 2
    with pytest.raises(AssertionError):
        tor.Scene().list()  # hittables_lists.nim:42


def test_ppm_helpers_equal_oracle(tor, oracle, tmp_path):  # io/ppm.nim:14-27
    rng = np.random.default_rng(0)
    cv = tor.newCanvas(5, 7, 1, 2.2)
    cv.pixels[:] = rng.uniform(-0.1, 1.1, cv.pixels.shape)
    cv.pixels[0, 0, 0] = float("nan")
    assert np.array_equal(cv.toRGB8(), oracle.quantise_rgb8(cv.pixels))
    a, b = tmp_path / "a.ppm", tmp_path / "b.ppm"
    tor.exportToPPM(cv, str(a))
    oracle.export_ppm(cv.pixels, str(b))
    assert a.read_bytes() == b.read_bytes()


def test_no_cpu_fallback(tor):
    """Without a CUDA device the context cannot be created and nothing renders."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tor.TorError) as e:
        tor.Context()
    assert e.value.code == -5  # TOR_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "trace_of_radiance_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "tor_oracle" not in text and "liboracle" not in text, f


def test_animation_helper_equals_oracle(tor, oracle):  # scenes_animated.nim:90-225
    mine = tor.Animation(height=36, width=64, t_max=0.2).scenes(skip=6)
    ref = oracle.Animation(height=36, width=64, t_max=0.2)
    frames = 0
    for cam, world in mine:
        cam_ref, objs_ref = ref.next_frame(skip=6)
        assert cam.as_array().tobytes() == np.ascontiguousarray(cam_ref).tobytes()
        assert world.objects.tobytes() == np.ascontiguousarray(objs_ref).tobytes()
        assert len(world) == 1601
        frames += 1
    assert frames == 7  # t = 0, 0.03, ..., 0.18 in float32 steps of 6 * 0.005
    # the frame count of the full C4 animation (t_max = 9.0) comes from float32 accumulation: 300
    n = sum(1 for _ in tor.Animation(height=8, width=8, t_max=9.0).scenes(skip=6))
    assert n == 300


def test_header_is_plain_c_and_links_from_c(tor, tmp_path):
    """The boundary is a C ABI: include/tor_b200.h must compile as C99 and a C program (what Nim's importc emits)
    must link against libtor_b200.so.  Without a device the program sees TOR_ERR_NO_DEVICE from tor_ctx_create and
    the device-free helpers still work; with one it creates and destroys a context."""
    import shutil
    import subprocess

    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not cc:
        pytest.skip("no C compiler")
    src = tmp_path / "consumer.c"
    src.write_text(r'''
#include <stdio.h>
#include "tor_b200.h"
int main(void) {
  tor_ctx* ctx = NULL;
  tor_camera cam;
  double from[3] = {13, 2, 3}, at[3] = {0, 0, 0}, up[3] = {0, 1, 0};
  tor_camera_make(&cam, from, at, up, 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0);
  if (sizeof(tor_camera) != 192 || sizeof(tor_canvas) != 24 || sizeof(tor_hittable) != 112) return 10;
  if (tor_random_scene(0xFACADE, 11, NULL, 0) != 485) return 11;
  if (tor_fast_substream_count(TOR_MODE_FAST, 675, 1200, 500) != 16) return 12;
  int rc = tor_ctx_create(NULL, 0, &ctx);
  printf("%d %d %s\n", tor_abi_version(), rc, rc ? tor_last_error(NULL) : "ok");
  if (rc == TOR_OK) {
    /* with a device: the drop-in call itself, the way the Nim shim makes it (render.nim:49) */
    static tor_hittable world[485];
    static double pixels[8 * 8 * 3];
    tor_canvas canvas;
    long long n = tor_random_scene(0xFACADE, 11, world, 485);
    int i, bad = 0;
    canvas.pixels = pixels;
    canvas.nrows = 8;
    canvas.ncols = 8;
    canvas.samples_per_pixel = 4;
    canvas.gamma_correction = 2.2f;
    rc = tor_render(ctx, &canvas, &cam, world, n, TOR_STRIDE_FLAT, 50, TOR_MODE_EXACT);
    for (i = 0; i < 8 * 8 * 3; ++i) bad += !(pixels[i] >= 0.0 && pixels[i] <= 1.0);
    printf("render %d bad %d top-left %.17g\n", rc, bad, pixels[(7 * 8 + 0) * 3]);
    tor_ctx_destroy(ctx);
    if (rc != TOR_OK || bad) return 14;
    return 0;
  }
  return rc == TOR_ERR_NO_DEVICE ? 0 : 13;
}
''')
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(tor.lib_path())
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-ltor_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[0] == "1"
