"""Split-stream mode (TOR_MODE_FAST, include/tor_b200.h): the pixel's sample loop of render.nim:62-67 cut into n
RNG substreams that run in parallel.

Two bars, both stated here:
  * against the CPU restatement of the same definition (oracle render_split): BIT-EXACT float64, any n, any
    partition — the mode is deterministic;
  * against the reference's render (the exact mode): a different Monte-Carlo estimate of the same integrals with the
    same number of samples.  Stated tolerance: per pixel and channel, |fast - exact| of the linear (pre-gamma) mean
    <= 4 * sqrt(2) * sigma / sqrt(spp) (sigma = that pixel's sample standard deviation, sqrt(2) because both sides
    are estimates) for >= 99.9 % of the values, and image means within 0.5 %; n = 1 is the exact mode bit for bit.
"""
import hashlib
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
M64 = (1 << 64) - 1


def _book_cam_array(oracle):
    return oracle.book_camera()


# ------------------------------------------------------------------------------------ CPU: the definition itself
def _splitmix_sic(state):
    """support/rng.nim:31-36 (the first multiplier is used twice, sic)."""
    state = (state + 0x9E3779B97F4A7C15) & M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0xBF58476D1CE4E5B9) & M64
    return state, z ^ (z >> 31)


def test_substream_zero_is_the_reference_seed(oracle):
    """Range 0 of every pixel uses seed(row, col) of rng.nim:46-53 (KAT from SURVEY.md §8c); range j continues the
    same SplitMix64 sequence at output 4j."""
    st = (5 << 32) ^ 7
    outs = []
    for _ in range(12):
        st, z = _splitmix_sic(st)
        outs.append(z)
    assert outs[:4] == [0x9AE9738C8C29FE95, 0x8C85084125631F66, 0xC23C9B1B61A77FAA, 0x10C0A27E9A9FAA3C]
    assert oracle.Rng().seed(5, 7).words() == outs[:4]
    # a 1x1-sample render of one pixel with n = 2: the second range must start from outs[4:8]; checked through the
    # first jitter draw, which fixes u of the only sample of range 1 (rng.nim:58-74,129-133)
    s0, s3 = outs[4], outs[7]
    first = (((s0 + s3) & M64) >> 12 | 0x3FF0000000000000).to_bytes(8, "little")
    u01 = np.frombuffer(first, dtype="<f8")[0] - 1.0
    assert 0.0 <= u01 < 1.0


def test_one_substream_is_the_exact_mode(oracle):
    world, cam = oracle.random_scene(), _book_cam_array(oracle)
    a = oracle.render(27, 48, 7, cam, world, math="det")
    b = oracle.render_split(27, 48, 7, cam, world, 1, math="det")
    assert a.tobytes() == b.tobytes()


def test_split_render_is_deterministic_and_partition_invariant(oracle):
    world, cam = oracle.random_scene(), _book_cam_array(oracle)
    full = oracle.render_split(27, 48, 10, cam, world, 4)
    one = oracle.render_split(27, 48, 10, cam, world, 4, nthreads=1)
    assert full.tobytes() == one.tobytes()
    parts = np.zeros_like(full)
    for g in range(3):
        p = oracle.render_split(27, 48, 10, cam, world, 4, rows=(g, 27, 3))
        parts[g::3] = p[g::3]
    assert parts.tobytes() == full.tobytes()


def test_first_range_shares_the_reference_samples(oracle):
    """With spp = n every range holds one sample, and range 0's sample is the reference's first sample of the pixel:
    an spp = 1 exact render equals range 0 alone."""
    world, cam = oracle.random_scene(), _book_cam_array(oracle)
    _, lin_exact, _ = oracle.render_split(18, 32, 1, cam, world, 1, stats=True)
    _, lin4, _ = oracle.render_split(18, 32, 4, cam, world, 4, stats=True)
    _, lin4_seq, _ = oracle.render_split(18, 32, 4, cam, world, 1, stats=True)
    assert not np.array_equal(lin4, lin4_seq)  # different estimates ...
    # ... whose first sample is common: remove it and the remaining three samples are non-negative sums
    assert np.all(lin4 - lin_exact >= -1e-12) and np.all(lin4_seq - lin_exact >= -1e-12)


@pytest.mark.parametrize("nsub", [4, 32])
def test_stated_tolerance_against_the_exact_render(oracle, nsub):
    world, cam = oracle.random_scene(), _book_cam_array(oracle)
    h, w, spp = 54, 96, 128
    _, lin_e, sq_e = oracle.render_split(h, w, spp, cam, world, 1, stats=True)
    _, lin_f, sq_f = oracle.render_split(h, w, spp, cam, world, nsub, stats=True)
    mean_e, mean_f = lin_e / spp, lin_f / spp
    var = 0.5 * ((sq_e / spp - mean_e**2) + (sq_f / spp - mean_f**2)) * spp / (spp - 1)
    sigma = np.sqrt(np.maximum(var, 0.0))
    bound = 4.0 * np.sqrt(2.0) * sigma / np.sqrt(spp) + 1e-12
    inside = np.abs(mean_f - mean_e) <= bound
    assert inside.mean() >= 0.999, f"{(~inside).sum()} of {inside.size} values outside 4*sqrt(2)*sigma/sqrt(spp)"
    for ch in range(3):
        assert abs(mean_f[..., ch].mean() / mean_e[..., ch].mean() - 1.0) <= 0.005


def test_automatic_substream_count(tor):
    """The count depends on the flags, the FULL canvas and spp only (never on a row selection or the device), so every
    partition of a canvas renders the same image: min(32, spp, 2^24 / pixels) rounded down to a power of two."""
    A = tor.api
    n = A.fast_substream_count
    assert n(0, 675, 1200, 500) == 1  # exact mode
    assert n(A.TOR_MODE_FAST, 216, 384, 100) == 32  # C1
    assert n(A.TOR_MODE_FAST, 675, 1200, 500) == 16  # C2 / C3
    assert n(A.TOR_MODE_FAST, 144, 256, 100) == 32  # C4
    assert n(A.TOR_MODE_FAST, 2160, 3840, 2000) == 2  # C5
    assert n(A.TOR_MODE_FAST, 216, 384, 5) == 4 and n(A.TOR_MODE_FAST, 216, 384, 1) == 1 and n(A.TOR_MODE_FAST, 8, 8, 0) == 1
    assert n(A.TOR_MODE_FAST, 8192, 8192, 64) == 1
    assert n(A.TOR_MODE_FAST | A.TOR_FAST_SUBSTREAMS(8), 675, 1200, 500) == 8
    for bad in (A.TOR_MODE_FAST | A.TOR_FAST_SUBSTREAMS(3), A.TOR_MODE_FAST | A.TOR_FAST_SUBSTREAMS(64),
                A.TOR_FAST_SUBSTREAMS(4), A.TOR_MODE_FAST | A.TOR_FLAG_BRUTE_FORCE):
        with pytest.raises(A.TorError):
            n(bad, 10, 10, 10)


# ---------------------------------------------------------------------------- GPU: the CUDA path against the oracle
def _book_cam(tor, aspect=16.0 / 9.0, t0=0.0, t1=1.0):
    return tor.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, aspect, 0.1, 10.0, t0, t1)


def _fast(tor, n):
    return tor.api.TOR_MODE_FAST | tor.api.TOR_FAST_SUBSTREAMS(n)


@pytest.mark.gpu
@pytest.mark.parametrize("nsub", [1, 2, 8, 32])
def test_gpu_split_stream_bit_exact(tor, oracle, gpu_ctx, nsub):
    """spp = 10 < 32 also covers empty sample ranges."""
    world, cam = tor.random_scene().list(), _book_cam(tor)
    h, w, spp = 36, 64, 10
    ocnt = {}
    ref = oracle.render_split(h, w, spp, cam.as_array(), world.objects, nsub, counters=ocnt)
    cv = tor.newCanvas(h, w, spp, 2.2)
    gpu_ctx.render(cv, cam, world, 50, flags=_fast(tor, nsub) | tor.api.TOR_FLAG_COUNT_SEGMENTS)
    cnt = gpu_ctx.counters()
    assert cv.pixels.tobytes() == ref.tobytes(), f"{int((cv.pixels != ref).sum())} float64 values differ"
    assert cnt["primary_rays"] == ocnt["primary_rays"] and cnt["segments"] == ocnt["segments"]
    if nsub == 1:
        exact = tor.newCanvas(h, w, spp, 2.2)
        gpu_ctx.render(exact, cam, world, 50)
        assert exact.pixels.tobytes() == cv.pixels.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("depth,spp,nsub", [(0, 4, 2), (1, 3, 4), (50, 1, 8), (50, 0, 4), (50, 33, 16)])
def test_gpu_split_stream_edges(tor, oracle, gpu_ctx, depth, spp, nsub):
    world, cam = tor.random_scene().list(), _book_cam(tor)
    ref = oracle.render_split(20, 30, spp, cam.as_array(), world.objects, nsub, max_depth=depth)
    cv = tor.newCanvas(20, 30, spp, 2.2)
    gpu_ctx.render(cv, cam, world, depth, flags=_fast(tor, nsub))
    assert cv.pixels.tobytes() == ref.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,spp,nsub", [(1, 1, 5, 2), (2, 3, 4, 4), (1, 40, 7, 8), (33, 1, 3, 32), (7, 9, 64, 1)])
def test_gpu_split_stream_tiny_canvases(tor, oracle, gpu_ctx, h, w, spp, nsub):
    """Fewer work units than one warp's chunk, 1-pixel-wide / 1-pixel-high canvases (u or v divide by zero as in
    render.nim:64-65): the warp-level queue hands out partial chunks and runs dry mid-pass."""
    world, cam = tor.random_scene().list(), _book_cam(tor)
    ref = oracle.render_split(h, w, spp, cam.as_array(), world.objects, nsub)
    cv = tor.newCanvas(h, w, spp, 2.2)
    gpu_ctx.render(cv, cam, world, 50, flags=_fast(tor, nsub))
    assert cv.pixels.tobytes() == ref.tobytes()


@pytest.mark.gpu
def test_gpu_split_stream_partitions_and_rgb8(tor, oracle, gpu_ctx):
    world, cam = tor.random_scene().list(), _book_cam(tor)
    h, w, spp, fl = 54, 96, 12, _fast(tor, 4)
    full = tor.newCanvas(h, w, spp, 2.2)
    gpu_ctx.render(full, cam, world, 50, flags=fl)
    for step in (2, 5):
        parts = tor.newCanvas(h, w, spp, 2.2)
        for g in range(step):
            gpu_ctx.render(parts, cam, world, 50, flags=fl, rows=(g, h, step))
        assert parts.pixels.tobytes() == full.pixels.tobytes()
    rgb = gpu_ctx.render_rgb8(full, cam, world, 50, flags=fl)
    assert np.array_equal(rgb, oracle.quantise_rgb8(full.pixels))
    # 1 601 static spheres of an animation frame (scenes_animated.nim), shutter 0..0
    for i, (acam, aworld) in enumerate(tor.Animation(height=36, width=64).scenes(skip=6)):
        ref = oracle.render_split(36, 64, 8, acam.as_array(), aworld.objects, 8)
        cv = tor.newCanvas(36, 64, 8, 2.2)
        gpu_ctx.render(cv, acam, aworld, 50, flags=_fast(tor, 8))
        assert cv.pixels.tobytes() == ref.tobytes()
        if i == 1:
            break


@pytest.mark.gpu
def test_gpu_split_stream_c1_auto_and_reference_png(tor, oracle, gpu_ctx):
    """C1 with the automatic range count (2^24 / 82 944 pixels -> capped at 32): bit-exact against the oracle's
    restatement, and as close to the reference's own PNG as an independent 100-spp estimate can be."""
    world, cam = tor.random_scene().list(), _book_cam(tor)
    cv = tor.newCanvas(216, 384, 100, 2.2)
    gpu_ctx.render(cv, cam, world, 50, flags=tor.api.TOR_MODE_FAST)
    ref = oracle.render_split(216, 384, 100, cam.as_array(), world.objects, 32)
    assert cv.pixels.tobytes() == ref.tobytes()
    digest = json.load(open(os.path.join(GOLD, "c1_split32_digest.json")))
    assert hashlib.sha256(cv.pixels.tobytes()).hexdigest() == digest["f64_sha256"]
    png = np.load(os.path.join(GOLD, "book2_motion_blur_rgb8.npz"))["rgb8"].astype(np.float64)
    mse = np.mean((cv.toRGB8().astype(np.float64) - png) ** 2)
    psnr = 10.0 * np.log10(255.0**2 / mse)
    assert psnr >= digest["psnr_vs_reference_png_floor_db"], psnr


@pytest.mark.gpu
def test_gpu_split_stream_flag_errors(tor, gpu_ctx):
    world, cam = tor.random_scene().list(), _book_cam(tor)
    cv = tor.newCanvas(8, 8, 2, 2.2)
    for bad in (_fast(tor, 3), _fast(tor, 64), tor.api.TOR_FAST_SUBSTREAMS(4),
                _fast(tor, 4) | tor.api.TOR_FLAG_BRUTE_FORCE):
        with pytest.raises(tor.api.TorError) as e:
            gpu_ctx.render(cv, cam, world, 50, flags=bad)
        assert e.value.code == -1  # TOR_ERR_INVALID_ARG
    gpu_ctx.render(cv, cam, world, 50, flags=_fast(tor, 2))  # the context stays usable


@pytest.mark.gpu
def test_gpu_split_stream_c2_full_size(tor, oracle, gpu_ctx):
    """C2 (1200x675, 500 spp) in split-stream mode with the automatic range count (2^24 / 810 000 -> 16): sampled
    rows against the oracle, partition invariance and the ray count on the full image; the stated statistical
    tolerance against the exact render on the same rows."""
    world, cam = tor.random_scene().list(), _book_cam(tor)
    h, w, spp = 675, 1200, 500
    full = tor.newCanvas(h, w, spp, 2.2)
    gpu_ctx.render(full, cam, world, 50, flags=tor.api.TOR_MODE_FAST | tor.api.TOR_FLAG_COUNT_SEGMENTS)
    cnt = gpu_ctx.counters()
    assert cnt["primary_rays"] == h * w * spp
    assert np.isfinite(full.pixels).all() and full.pixels.min() >= 0.0
    for r in (40, 333):
        ref = oracle.render_split(h, w, spp, cam.as_array(), world.objects, 16, rows=(r, r + 1, 1))
        assert full.pixels[r].tobytes() == ref[r].tobytes(), r
    thirds = tor.newCanvas(h, w, spp, 2.2)
    for g in range(3):
        gpu_ctx.render(thirds, cam, world, 50, flags=tor.api.TOR_MODE_FAST, rows=(g, h, 3))
    assert thirds.pixels.tobytes() == full.pixels.tobytes()
    exact = tor.newCanvas(h, w, spp, 2.2)
    gpu_ctx.render(exact, cam, world, 50)
    a, b = full.pixels ** 2.2, exact.pixels ** 2.2  # back to (approximately) linear means
    assert abs(a.mean() / b.mean() - 1.0) <= 0.005
    assert np.mean(np.abs(full.toRGB8().astype(int) - exact.toRGB8().astype(int)) <= 8) >= 0.99
