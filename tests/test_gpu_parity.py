"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Bar: the float64 framebuffer is
BIT-EXACT against the oracle built with the same deterministic sin/cos/pow (tor_detmath.h), and the 8-bit
PPM image equals the reference's own media/book2_motion_blur.png and the oracle built with glibc libm."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _book_cam(tor, aspect=16.0 / 9.0, t0=0.0, t1=1.0):
    return tor.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, aspect, 0.1, 10.0, t0, t1)


def _check(tor, oracle, ctx, world, cam, h, w, spp, depth=50, gamma=2.2, rows=None):
    """Brute-force scan, BVH with the row-major pixel queue, BVH with the longest-pixel-first queue (default; the
    pre-pass only exists for spp >= 64): each against the oracle, bit for bit."""
    ocnt = {}
    ref = np.full((h, w, 3), -7.0)
    oracle.render(h, w, spp, cam.as_array(), world.objects, max_depth=depth, gamma=gamma, rows=rows, math="det",
                  counters=ocnt, out=ref)
    for route in (tor.api.TOR_FLAG_BRUTE_FORCE, tor.api.TOR_FLAG_ROW_MAJOR_QUEUE, 0):
        cv = tor.newCanvas(h, w, spp, gamma)
        cv.pixels[:] = -7.0  # rows that are not selected must stay untouched
        ctx.render(cv, cam, world, depth, flags=tor.api.TOR_FLAG_COUNT_SEGMENTS | route, rows=rows)
        cnt = ctx.counters()
        assert cv.pixels.tobytes() == ref.tobytes(), f"route {route:#x}: {int((cv.pixels != ref).sum())} float64 values differ"
        assert cnt["primary_rays"] == ocnt["primary_rays"]
        if depth > 0:
            assert cnt["segments"] == ocnt["segments"]
    return cv


def test_small_random_scene_bit_exact(tor, oracle, gpu_ctx):
    _check(tor, oracle, gpu_ctx, tor.random_scene().list(), _book_cam(tor), 36, 64, 10)


def test_c1_bit_exact_and_equals_reference_png(tor, oracle, gpu_ctx):
    """Config C1 = trace_of_radiance.nim:27-57 (384x216, 100 spp, depth 50, gamma 2.2)."""
    cv = _check(tor, oracle, gpu_ctx, tor.random_scene().list(), _book_cam(tor), 216, 384, 100)
    assert gpu_ctx.last_handoffs()["from_lanes"] > 0  # about one pixel per lane: the tail goes to one warp per pixel
    digest = json.load(open(os.path.join(GOLD, "c1_oracle_digest.json")))
    assert hashlib.sha256(cv.pixels.tobytes()).hexdigest() == digest["det"]["f64_sha256"]
    png = np.load(os.path.join(GOLD, "book2_motion_blur_rgb8.npz"))["rgb8"]
    assert np.array_equal(cv.toRGB8(), png)  # the reference's own output, every channel of every pixel
    # and the glibc-libm oracle (what the Nim binary links) at the PPM's 8 bits
    lib = oracle.render(216, 384, 100, _book_cam(tor).as_array(), tor.random_scene().list().objects, math="libm")
    assert np.array_equal(cv.toRGB8(), oracle.quantise_rgb8(lib))
    # float64 distance to the libm image: sin/cos/pow differ in the last bit and a bounce chain amplifies it
    # (measured 2.4e-12); stated tolerance 1e-9, i.e. ~4e-7 of one 8-bit PPM step
    assert np.max(np.abs(cv.pixels - lib)) < 1e-9


def test_row_partitions_are_bit_identical(tor, oracle, gpu_ctx):
    world, cam = tor.random_scene().list(), _book_cam(tor)
    full = tor.newCanvas(54, 96, 6, 2.2)
    gpu_ctx.render(full, cam, world, 50)
    for step in (2, 3, 8):
        parts = tor.newCanvas(54, 96, 6, 2.2)
        for g in range(step):
            gpu_ctx.render(parts, cam, world, 50, rows=(g, 54, step))
        assert parts.pixels.tobytes() == full.pixels.tobytes()
    _check(tor, oracle, gpu_ctx, world, cam, 54, 96, 6, rows=(5, 41, 7))
    _check(tor, oracle, gpu_ctx, world, cam, 54, 96, 6, rows=(10, 10, 1))  # empty selection


@pytest.mark.parametrize("depth,spp", [(1, 3), (2, 5), (0, 2), (50, 1), (50, 0)])
def test_depth_and_spp_edges(tor, oracle, gpu_ctx, depth, spp):
    cv = tor.newCanvas(20, 30, spp, 2.2)
    world, cam = tor.random_scene().list(), _book_cam(tor)
    gpu_ctx.render(cv, cam, world, depth)
    ref = oracle.render(20, 30, spp, cam.as_array(), world.objects, max_depth=depth, math="det")
    assert cv.pixels.tobytes() == ref.tobytes()  # includes the NaN image of spp == 0 (canvas.nim:49-54)


def test_degenerate_canvases(tor, oracle, gpu_ctx):
    world, cam = tor.random_scene().list(), _book_cam(tor)
    # one column / one row: u or v divides by zero exactly as render.nim:64-65 does
    for (h, w) in [(1, 1), (7, 1), (1, 9), (3, 257)]:
        cv = tor.newCanvas(h, w, 3, 2.2)
        gpu_ctx.render(cv, cam, world, 10)
        ref = oracle.render(h, w, 3, cam.as_array(), world.objects, max_depth=10, math="det")
        assert cv.pixels.tobytes() == ref.tobytes(), (h, w)


def _handmade_scenes(tor):
    lam, met, die = tor.lambertian, tor.metal, tor.dielectric
    scenes = {}
    scenes["single_static"] = [tor.sphere((0, 0, 0), 1.0, lam((0.8, 0.3, 0.3)))]
    scenes["static_only"] = [
        tor.sphere((0, -100.5, -1), 100, lam((0.8, 0.8, 0.0))),
        tor.sphere((0, 0, -1), 0.5, lam((0.1, 0.2, 0.5))),
        tor.sphere((-1, 0, -1), 0.5, die(1.5)),
        tor.sphere((-1, 0, -1), -0.45, die(1.5)),  # hollow glass: negative radius flips the normal
        tor.sphere((1, 0, -1), 0.5, met((0.8, 0.6, 0.2), 0.0)),
        tor.sphere((1, 1.2, -1), 0.5, met((0.8, 0.8, 0.8), 1.0)),
    ]
    scenes["general_movers_two_time_classes"] = [
        tor.sphere((0, -1000, 0), 1000, lam((0.5, 0.5, 0.5))),
        tor.movingSphere((0, 1, 0), 0.0, (2, 1.5, 1), 1.0, 1.0, lam((0.4, 0.2, 0.1))),
        tor.movingSphere((-4, 1, 0), 0.0, (-4, 1, 2), 1.0, 1.0, met((0.7, 0.6, 0.5), 0.1)),
        tor.movingSphere((4, 1, 0), -1.0, (4, 3, 0), 3.0, 1.0, die(1.5)),
        tor.movingSphere((2, 0.4, 3), 0.25, (3, 0.4, 3), 0.5, 0.4, lam((0.1, 0.7, 0.1))),
        tor.movingSphere((6, 0.5, 2), 0.25, (6, 1.0, 2), 0.5, 0.5, lam((0.1, 0.1, 0.7))),
    ]
    # identical spheres: exact t ties must go to the lowest index (hittables_lists.nim:48-55)
    scenes["exact_ties"] = [
        tor.sphere((0, -1000, 0), 1000, lam((0.5, 0.5, 0.5))),
        tor.sphere((0, 1, 0), 1.0, lam((0.9, 0.1, 0.1))),
        tor.sphere((0, 1, 0), 1.0, met((0.1, 0.9, 0.1), 0.0)),
        tor.sphere((0, 1, 0), 1.0, die(1.5)),
    ]
    # degenerate mover: time0 == time1 divides by zero in moving_spheres.nim:41-42
    scenes["zero_length_time_interval"] = [
        tor.sphere((0, -1000, 0), 1000, lam((0.5, 0.5, 0.5))),
        tor.movingSphere((0, 1, 0), 0.5, (0, 2, 0), 0.5, 1.0, lam((0.4, 0.2, 0.1))),
        tor.sphere((4, 1, 0), 1.0, met((0.7, 0.6, 0.5), 0.0)),
    ]
    return {k: tor.Scene(np.array(v, dtype=tor.HITTABLE_DTYPE)).list() for k, v in scenes.items()}


@pytest.mark.parametrize("name", ["single_static", "static_only", "general_movers_two_time_classes", "exact_ties",
                                  "zero_length_time_interval"])
def test_handmade_scenes(tor, oracle, gpu_ctx, name):
    world = _handmade_scenes(tor)[name]
    _check(tor, oracle, gpu_ctx, world, _book_cam(tor), 45, 80, 16)
    # camera inside the scene looking along -z, shutter beyond the movers' time range, no defocus
    cam = tor.camera((0, 0.3, 2.5), (0, 0.2, -1), (0, 1, 0), 60.0, 80 / 45, 0.0, 1.0, -0.5, 2.5)
    _check(tor, oracle, gpu_ctx, world, cam, 45, 80, 16)


def test_camera_inside_a_glass_sphere(tor, oracle, gpu_ctx):
    world = tor.Scene(np.array([tor.sphere((0, 0, 0), 5.0, tor.dielectric(1.5)),
                                tor.sphere((0, -1000, 0), 990, tor.lambertian((0.5, 0.5, 0.5))),
                                tor.sphere((3, 0, 0), 1.0, tor.metal((0.9, 0.9, 0.9), 0.3))],
                               dtype=tor.HITTABLE_DTYPE)).list()
    cam = tor.camera((0, 0, 0.5), (3, 0, 0), (0, 1, 0), 70.0, 1.0, 0.2, 2.0, 0.0, 0.0)
    _check(tor, oracle, gpu_ctx, world, cam, 40, 40, 20)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_randomised_scenes(tor, oracle, gpu_ctx, seed):
    """Random mixes of static / y-moving / freely moving spheres, all three materials, several time classes."""
    rng = np.random.default_rng(seed)
    objs = [tor.sphere((0, -1000, 0), 1000, tor.lambertian((0.5, 0.5, 0.5)))]
    n = int(rng.integers(5, 120))
    for _ in range(n):
        c = rng.uniform(-6, 6, 3) * (1, 0.15, 1) + (0, 0.6, 0)
        r = float(rng.uniform(0.1, 0.7))
        m = rng.integers(0, 3)
        mat = (tor.lambertian(rng.uniform(0, 1, 3)), tor.metal(rng.uniform(0.5, 1, 3), rng.uniform(0, 1.2)),
               tor.dielectric(float(rng.uniform(1.1, 2.4))))[m]
        k = rng.integers(0, 3)
        if k == 0:
            objs.append(tor.sphere(c, r, mat))
        elif k == 1:
            objs.append(tor.movingSphere(c, 0.0, c + (0, rng.uniform(0, 0.5), 0), 1.0, r, mat))
        else:
            t0 = float(rng.choice([0.0, 0.1, -1.0]))
            objs.append(tor.movingSphere(c, t0, c + rng.uniform(-0.5, 0.5, 3), t0 + float(rng.choice([1.0, 0.5])), r, mat))
    world = tor.Scene(np.array(objs, dtype=tor.HITTABLE_DTYPE)).list()
    cam = tor.camera(rng.uniform(-9, 9, 3) * (1, 0, 1) + (0, 2, 0), (0, 0.5, 0), (0, 1, 0), 40.0, 1.5, 0.05, 8.0, 0.0, 1.0)
    _check(tor, oracle, gpu_ctx, world, cam, 40, 60, 12)


def _encode_nim_variants(objs):
    """HittableVariant array as Nim lays it out (include/tor_b200.h, encoding ii; stride 120)."""
    buf = np.zeros((len(objs), 120), dtype=np.uint8)
    for i, o in enumerate(objs):
        rec = buf[i]
        rec[0] = o["kind"]
        mat = np.zeros(40, dtype=np.uint8)
        mat[0] = o["mat_kind"]
        if o["mat_kind"] == 0:
            mat[8:32] = np.frombuffer(o["albedo"].tobytes(), dtype=np.uint8)
        elif o["mat_kind"] == 1:
            mat[8:32] = np.frombuffer(o["albedo"].tobytes(), dtype=np.uint8)
            mat[32:40] = np.frombuffer(np.float64(o["fuzz_or_ior"]).tobytes(), dtype=np.uint8)
        else:
            mat[8:16] = np.frombuffer(np.float64(o["fuzz_or_ior"]).tobytes(), dtype=np.uint8)
        u = rec[8:]
        if o["kind"] == 0:
            u[0:24] = np.frombuffer(o["center0"].tobytes(), dtype=np.uint8)
            u[24:32] = np.frombuffer(np.float64(o["radius"]).tobytes(), dtype=np.uint8)
            u[32:72] = mat
        else:
            u[0:24] = np.frombuffer(o["center0"].tobytes(), dtype=np.uint8)
            u[24:48] = np.frombuffer(o["center1"].tobytes(), dtype=np.uint8)
            u[48:56] = np.frombuffer(np.float64(o["time0"]).tobytes(), dtype=np.uint8)
            u[56:64] = np.frombuffer(np.float64(o["time1"]).tobytes(), dtype=np.uint8)
            u[64:72] = np.frombuffer(np.float64(o["radius"]).tobytes(), dtype=np.uint8)
            u[72:112] = mat
    return buf


def test_nim_variant_encoding_equals_flat(tor, gpu_ctx):
    world, cam = tor.random_scene().list(), _book_cam(tor)
    flat = tor.newCanvas(27, 48, 5, 2.2)
    gpu_ctx.render(flat, cam, world, 50)
    enc = _encode_nim_variants(world.objects)
    nimv = tor.newCanvas(27, 48, 5, 2.2)
    gpu_ctx.render_raw(nimv, cam, enc.ctypes.data, len(world), 120, 50)
    assert nimv.pixels.tobytes() == flat.pixels.tobytes()


def test_error_behaviour(tor, gpu_ctx):
    world, cam = tor.random_scene().list(), _book_cam(tor)
    cv = tor.newCanvas(4, 4, 1, 2.2)
    with pytest.raises(tor.TorError) as e:  # empty list: hittables_lists.nim:42 asserts len > 0
        gpu_ctx.render_raw(cv, cam, world.objects.ctypes.data, 0, 112, 50)
    assert e.value.code == -1
    with pytest.raises(tor.TorError) as e:
        gpu_ctx.render_raw(cv, cam, world.objects.ctypes.data, len(world), 96, 50)
    assert e.value.code == -4 and "stride" in str(e.value)
    with pytest.raises(tor.TorError) as e:
        gpu_ctx.render(cv, cam, world, 50, rows=(0, 9, 1))
    assert e.value.code == -1
    bad = world.objects.copy()
    bad["mat_kind"][3] = 7
    with pytest.raises(tor.TorError) as e:
        gpu_ctx.render(cv, cam, tor.HittableList(bad), 50)
    assert "object 3" in str(e.value)
    gpu_ctx.render(cv, cam, world, 50)  # the context stays usable after errors


def test_animation_frames(tor, oracle, gpu_ctx):
    """scenes_animated.nim: 1 601 static spheres per frame, shutter 0..0, orbiting camera."""
    an = oracle.Animation(height=36, width=64, t_max=9.0)
    for f in range(3):
        cam_arr, objs = an.next_frame(skip=6)
        if f == 1:
            continue
        _check(tor, oracle, gpu_ctx, tor.HittableList(objs), tor.Camera.from_array(cam_arr), 36, 64, 6)


def test_ten_thousand_spheres(tor, oracle, gpu_ctx):
    """C5's scene (random_scene_grid half = 50): too large for shared-memory staging of the whole blob."""
    world = tor.random_scene(0xFACADE, 50).list()
    assert len(world) > 9000
    _check(tor, oracle, gpu_ctx, world, _book_cam(tor), 18, 32, 3)


def test_c2_full_size_properties(tor, oracle, gpu_ctx):
    """Config C2 (1200x675, 500 spp) is ~50x C1 on the CPU, so: sampled rows against the oracle, and
    size-independent properties on the full image (partition invariance, ray count, finite range)."""
    world, cam = tor.random_scene().list(), _book_cam(tor)
    h, w, spp = 675, 1200, 500
    full = tor.newCanvas(h, w, spp, 2.2)
    gpu_ctx.render(full, cam, world, 50, flags=tor.api.TOR_FLAG_COUNT_SEGMENTS)
    cnt = gpu_ctx.counters()
    assert cnt["primary_rays"] == h * w * spp
    assert np.isfinite(full.pixels).all() and full.pixels.min() >= 0.0
    ref = np.zeros((h, w, 3))
    for r in (40, 333, 674):
        oracle.render(h, w, spp, cam.as_array(), world.objects, rows=(r, r + 1, 1), math="det", out=ref)
        assert full.pixels[r].tobytes() == ref[r].tobytes(), r
    halves = tor.newCanvas(h, w, spp, 2.2)
    gpu_ctx.render(halves, cam, world, 50, rows=(0, h, 2))
    gpu_ctx.render(halves, cam, world, 50, rows=(1, h, 2))
    assert halves.pixels.tobytes() == full.pixels.tobytes()


def test_device_resident_entry_point(tor, gpu_ctx):
    import torch

    world, cam = tor.random_scene().list(), _book_cam(tor)
    host = tor.newCanvas(30, 50, 4, 2.2)
    gpu_ctx.render(host, cam, world, 50)
    gpu_ctx.scene_upload(cam, world)
    dev = torch.zeros((10, 50, 3), dtype=torch.float64, device="cuda:0")
    stream = torch.cuda.current_stream()
    gpu_ctx.render_device_async(dev.data_ptr(), 30, 50, 4, 2.2, 50, rows=(2, 30, 3), stream=stream.cuda_stream)
    stream.synchronize()
    assert dev.cpu().numpy().tobytes() == host.pixels[2:30:3].tobytes()
    assert gpu_ctx.launch_count() > 0 and gpu_ctx.last_kernel_ms() > 0


def test_rgb8_output_equals_ppm_quantisation(tor, oracle, gpu_ctx):
    """tor_render_rgb8: io/ppm.nim:14-27 quantisation on the device, PPM row order."""
    world, cam = tor.random_scene().list(), _book_cam(tor)
    cv = tor.newCanvas(45, 80, 8, 2.2)
    gpu_ctx.render(cv, cam, world, 50)
    rgb = gpu_ctx.render_rgb8(cv, cam, world, 50)
    assert rgb.dtype == np.uint8 and rgb.shape == (45, 80, 3)
    assert np.array_equal(rgb, cv.toRGB8())
    ref = oracle.render(45, 80, 8, cam.as_array(), world.objects, math="det")
    assert np.array_equal(rgb, oracle.quantise_rgb8(ref))
    nan = tor.newCanvas(5, 7, 0, 2.2)  # spp == 0: a NaN image (canvas.nim:49-54) quantises to 0
    assert not gpu_ctx.render_rgb8(nan, cam, world, 50).any()


def test_animation_frame_pipeline(tor, oracle):
    """render_animation: several frames in flight on separate contexts; every frame equals the oracle's."""
    frames = {}
    n = tor.render_animation(tor.Animation(height=27, width=48, t_max=9.0), samples_per_pixel=4, in_flight=3,
                             on_frame=lambda i, rgb: frames.__setitem__(i, rgb.copy()), max_frames=7)
    assert n == 7 and sorted(frames) == list(range(7))
    an = oracle.Animation(height=27, width=48, t_max=9.0)
    for i in range(7):
        cam_arr, objs = an.next_frame(skip=6)
        ref = oracle.render(27, 48, 4, cam_arr, objs, math="det")
        assert np.array_equal(frames[i], oracle.quantise_rgb8(ref)), i


def test_multi_device_context(tor, oracle):
    """tor_ctx_create with several devices: rows interleave over the devices inside one process; same bits."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = tor.Context([0, 1])
    world, cam = tor.random_scene().list(), _book_cam(tor)
    _check(tor, oracle, ctx, world, cam, 45, 80, 8)
    _check(tor, oracle, ctx, world, cam, 45, 80, 8, rows=(3, 44, 5))
    _check(tor, oracle, ctx, world, cam, 1, 9, 3)  # fewer rows than devices
    ctx.close()


@pytest.mark.parametrize("h,w,spp,rows", [(20, 30, 256, None), (45, 80, 300, None), (45, 80, 256, (1, 45, 4)),
                                           (3, 700, 256, None), (90, 160, 256, None)])
def test_cost_ranked_pixel_queue(tor, oracle, gpu_ctx, h, w, spp, rows):
    """spp >= 256 turns on the cost pre-pass + dealt / longest-first pixel queue (tor_api.cu): canvases with fewer
    pixels than lanes, pixel counts that are not a multiple of the CTA size, row subsets, and more pixels than one
    CTA wave."""
    _check(tor, oracle, gpu_ctx, tor.random_scene().list(), _book_cam(tor), h, w, spp, rows=rows)


# ---------------------------------------------------------------------------------- warp-cooperative pixels
class _EnvCtx:
    """A context created under developer knobs (they are read once, at tor_ctx_create)."""

    def __init__(self, tor, **env):
        self.tor, self.env, self.ctx, self.old = tor, {k: str(v) for k, v in env.items()}, None, {}

    def __enter__(self):
        for k, v in self.env.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v
        self.ctx = self.tor.Context()
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        return self.ctx

    def __exit__(self, *a):
        self.ctx.close()


@pytest.mark.parametrize("force", [10 ** 9, 37])
def test_cooperative_pixels_bit_exact(tor, oracle, force):
    """The warp-cooperative route (one warp traces one expensive pixel, tor_kernels_bvh.cuh) forced onto every pixel
    of the launch (force = 10^9) or onto the 37 most expensive ones beside the ordinary lanes: same bits as the
    oracle for the book scene, hand-made scenes (ties, degenerate movers, hollow glass), random mixes, 1 601 and
    10 002 spheres, row subsets and canvases smaller than a warp's worth of clusters."""
    with _EnvCtx(tor, TOR_BVH_PREPASS_SPP=9, TOR_BVH_COOP_FORCE=force) as ctx:
        world, cam = tor.random_scene().list(), _book_cam(tor)
        _check(tor, oracle, ctx, world, cam, 36, 64, 10)
        _check(tor, oracle, ctx, world, cam, 54, 96, 12, rows=(5, 41, 7))
        _check(tor, oracle, ctx, world, cam, 9, 11, 9)
        for name, w in _handmade_scenes(tor).items():
            _check(tor, oracle, ctx, w, cam, 30, 40, 10)
            inside = tor.camera((0, 0.3, 2.5), (0, 0.2, -1), (0, 1, 0), 60.0, 4 / 3, 0.0, 1.0, -0.5, 2.5)
            _check(tor, oracle, ctx, w, inside, 30, 40, 10)
        an = oracle.Animation(height=36, width=64, t_max=9.0)
        cam_arr, objs = an.next_frame(skip=6)
        _check(tor, oracle, ctx, tor.HittableList(objs), tor.Camera.from_array(cam_arr), 36, 64, 9)
        big = tor.random_scene(0xFACADE, 50).list()
        _check(tor, oracle, ctx, big, cam, 18, 32, 9)
        # depth edges inside the cooperative loop
        for depth in (1, 2):
            cv = tor.newCanvas(20, 30, 9, 2.2)
            ctx.render(cv, cam, world, depth)
            ref = oracle.render(20, 30, 9, cam.as_array(), world.objects, max_depth=depth, math="det")
            assert cv.pixels.tobytes() == ref.tobytes()


def test_cooperative_pixels_c1_digest(tor):
    """All of C1 through the cooperative route: the frozen sha256 of the oracle's float64 framebuffer."""
    with _EnvCtx(tor, TOR_BVH_PREPASS_SPP=9, TOR_BVH_COOP_FORCE=10 ** 9) as ctx:
        cv = tor.newCanvas(216, 384, 100, 2.2)
        ctx.render(cv, _book_cam(tor), tor.random_scene().list(), 50)
        digest = json.load(open(os.path.join(GOLD, "c1_oracle_digest.json")))
        assert hashlib.sha256(cv.pixels.tobytes()).hexdigest() == digest["det"]["f64_sha256"]


def test_cooperative_pixels_default_policy_on_a_gpu_share_of_c2(tor, oracle, gpu_ctx):
    """An eighth and a half of C2's rows (one GPU's share of an 8- and of a 2-GPU render) with the default policy —
    the cost pre-pass picks the cooperative pixels itself, and the tail of the launch is handed to one warp per pixel —
    against the same rows through the plain row-major pixel queue, and one row against the oracle."""
    world, cam = tor.random_scene().list(), _book_cam(tor)
    h, w, spp = 675, 1200, 500
    for first, step in ((3, 8), (1, 2)):
        a = tor.newCanvas(h, w, spp, 2.2)
        b = tor.newCanvas(h, w, spp, 2.2)
        gpu_ctx.render(a, cam, world, 50, rows=(first, h, step))
        assert gpu_ctx.last_schedule()["cooperative_pixels"] > 0
        parked = gpu_ctx.last_handoffs()
        assert parked["from_lanes"] + parked["from_cooperative_warps"] > 0, parked
        gpu_ctx.render(b, cam, world, 50, rows=(first, h, step), flags=tor.api.TOR_FLAG_ROW_MAJOR_QUEUE)
        assert a.pixels.tobytes() == b.pixels.tobytes(), step
    ref = np.zeros((h, w, 3))
    oracle.render(h, w, spp, cam.as_array(), world.objects, rows=(203, 204, 1), math="det", out=ref)
    assert a.pixels[203].tobytes() == ref[203].tobytes()


# ---------------------------------------------------------------------------------- late hand-off
@pytest.mark.parametrize("env", [
    dict(TOR_BVH_HANDOFF=1, TOR_BVH_HANDOFF_MIN_LEFT=1),                          # tail flag up at once: parked at sample 0 or 1
    dict(TOR_BVH_HANDOFF=1, TOR_BVH_HANDOFF_MIN_LEFT=4, TOR_BVH_COOP_FORCE=10 ** 9),  # every pixel cooperative: parked by those warps
    dict(TOR_BVH_HANDOFF=93, TOR_BVH_HANDOFF_MIN_LEFT=2),                         # flag goes up while pixels are in flight
    dict(TOR_BVH_HANDOFF=93, TOR_BVH_HANDOFF_MIN_LEFT=2, TOR_BVH_COOP_FORCE=300, TOR_BVH_HANDOFF_WARPS=3),
])
def test_late_handoff_bit_exact(tor, oracle, env):
    """BvhRenderParams::handoff: pixels parked at a sample boundary by lanes and by cooperative warps (generator state,
    samples done, colour sum) and finished by the second launch of render_coop_kernel give the oracle's bits — the
    knobs force the hand-off onto small renders, at sample 0 and in mid-flight."""
    with _EnvCtx(tor, TOR_BVH_PREPASS_SPP=9, TOR_BVH_HANDOFF_ALL=1, TOR_BVH_DEAL_SORTED=2, **env) as ctx:
        world, cam = tor.random_scene().list(), _book_cam(tor)
        first = _check(tor, oracle, ctx, world, cam, 90, 160, 24)
        # once more: the first launch of a kernel in a context loads its code, which shifts who sees the tail flag when
        cv = tor.newCanvas(90, 160, 24, 2.2)
        ctx.render(cv, cam, world, 50)
        assert cv.pixels.tobytes() == first.pixels.tobytes()
        parked = ctx.last_handoffs()
        assert parked["from_lanes"] + parked["from_cooperative_warps"] > 0, parked
        if env.get("TOR_BVH_COOP_FORCE") == 10 ** 9:
            assert parked["from_cooperative_warps"] > 0, parked
        _check(tor, oracle, ctx, world, cam, 54, 96, 12, rows=(5, 41, 7))
        _check(tor, oracle, ctx, world, cam, 9, 11, 9)
        for name, w in _handmade_scenes(tor).items():
            _check(tor, oracle, ctx, w, cam, 30, 40, 10)
        an = oracle.Animation(height=36, width=64, t_max=9.0)  # 1 601 movers with their own shutter interval
        cam_arr, objs = an.next_frame(skip=6)
        _check(tor, oracle, ctx, tor.HittableList(objs), tor.Camera.from_array(cam_arr), 36, 64, 9)
        _check(tor, oracle, ctx, tor.random_scene(0xFACADE, 50).list(), cam, 18, 32, 9)  # 10 002 spheres, nothing staged
        for depth in (0, 1, 2):
            cv = tor.newCanvas(20, 30, 9, 2.2)
            ctx.render(cv, cam, world, depth)
            ref = oracle.render(20, 30, 9, cam.as_array(), world.objects, max_depth=depth, math="det")
            assert cv.pixels.tobytes() == ref.tobytes()


# ---------------------------------------------------------------------------------- device-resident animation
def test_device_animation_frames_equal_the_oracle(tor, oracle, gpu_ctx):
    """tor_animation_dev_*: physics, record rebuild and BVH re-fit on the device (no per-frame H2D).  Every frame
    equals the oracle's render of the oracle's own scene iterator, including frames late in the animation (spheres
    high above the ground: the re-fitted boxes must still contain them) and frames that are only stepped."""
    an = tor.DeviceAnimation(gpu_ctx, height=27, width=48, t_max=9.0, in_flight=3)
    frames = {}
    n, ms = an.render_all(samples_per_pixel=4, on_frame=lambda i, rgb: frames.__setitem__(i, rgb.copy()), max_frames=9)
    assert n == 9 and sorted(frames) == list(range(9)) and ms > 0
    ref = oracle.Animation(height=27, width=48, t_max=9.0)
    for i in range(9):
        cam_arr, objs = ref.next_frame(skip=6)
        want = oracle.quantise_rgb8(oracle.render(27, 48, 4, cam_arr, objs, math="det"))
        assert np.array_equal(frames[i], want), i
    again = {}
    an.reset()  # back to the first frame: the same frames once more
    n, _ = an.render_all(samples_per_pixel=4, on_frame=lambda i, rgb: again.__setitem__(i, rgb.copy()), max_frames=4)
    assert n == 4 and all(np.array_equal(again[i], frames[i]) for i in range(4))
    an.close()
    # every 20th frame of the whole animation on "rank 1 of 20": 285 frames are physics-only steps in between
    an = tor.DeviceAnimation(gpu_ctx, height=18, width=32, t_max=9.0, in_flight=2)
    frames = {}
    n, _ = an.render_all(samples_per_pixel=3, on_frame=lambda i, rgb: frames.__setitem__(i, rgb.copy()), rank=1, world=20)
    assert n == 300 and sorted(frames) == list(range(1, 300, 20))
    ref = oracle.Animation(height=18, width=32, t_max=9.0)
    for i in range(300):
        got = ref.next_frame(skip=6)
        assert got is not None
        if i in frames:
            want = oracle.quantise_rgb8(oracle.render(18, 32, 3, got[0], got[1], math="det"))
            assert np.array_equal(frames[i], want), i
    an.close()


def test_device_animation_split_stream_equals_host_pipeline(tor, gpu_ctx):
    """The same frames through the host iterator (scene re-packed and BVH rebuilt per frame) and through the device
    pipeline (re-fitted boxes), in split-stream mode: identical bytes — the hierarchy is only a filter."""
    fl = tor.api.TOR_MODE_FAST
    a = {}
    tor.render_animation(tor.Animation(height=36, width=64, t_max=9.0), samples_per_pixel=8, in_flight=2, flags=fl,
                         on_frame=lambda i, rgb: a.__setitem__(i, rgb.copy()), max_frames=6)
    b = {}
    an = tor.DeviceAnimation(gpu_ctx, height=36, width=64, t_max=9.0, in_flight=2)
    an.render_all(samples_per_pixel=8, flags=fl, on_frame=lambda i, rgb: b.__setitem__(i, rgb.copy()), max_frames=6)
    an.close()
    assert sorted(a) == sorted(b) == list(range(6))
    for i in range(6):
        assert np.array_equal(a[i], b[i]), i


def test_c_program_calls_tor_render(tor, oracle, tmp_path):
    """A C99 program (what Nim's importc emits for the shim of INTEGRATION.md) renders an 8x8 canvas through
    tor_render; its top-left value equals the oracle's."""
    import shutil
    import subprocess

    import test_abi

    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not cc:
        pytest.skip("no C compiler")
    test_abi.test_header_is_plain_c_and_links_from_c(tor, tmp_path)  # builds tmp_path/consumer and checks rc == 0
    out = subprocess.run([str(tmp_path / "consumer")], capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("render")][0].split()
    assert line[1] == "0" and line[3] == "0"
    world, cam = tor.random_scene().list(), _book_cam(tor)
    ref = oracle.render(8, 8, 4, cam.as_array(), world.objects, math="det")
    assert float(line[5]) == ref[7, 0, 0]
