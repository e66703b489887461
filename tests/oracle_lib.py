"""ctypes binding for the CPU oracle (oracle/build/liboracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (trace_of_radiance_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "build", "liboracle.so")

# Flat hittable record: identical field order to include/tor_b200.h `tor_hittable` (112 bytes).
HITTABLE_DTYPE = np.dtype(
    [
        ("kind", "<u4"),
        ("mat_kind", "<u4"),
        ("center0", "<f8", (3,)),
        ("center1", "<f8", (3,)),
        ("time0", "<f8"),
        ("time1", "<f8"),
        ("radius", "<f8"),
        ("albedo", "<f8", (3,)),
        ("fuzz_or_ior", "<f8"),
    ],
    align=False,
)
assert HITTABLE_DTYPE.itemsize == 112

K_SPHERE, K_MOVING_SPHERE = 0, 1
K_LAMBERTIAN, K_METAL, K_DIELECTRIC = 0, 1, 2

_lib = None


def build(force=False):
    """Compile the oracle with its Makefile (g++ is present wherever this runs on CPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src_newer = False
    if os.path.exists(LIB_PATH):
        t = os.path.getmtime(LIB_PATH)
        for f in ("oracle_capi.cc", "tor_oracle.hpp", "tor_oracle_video.hpp", "../trace_of_radiance_b200/csrc/tor_detmath.h"):
            p = os.path.join(ORACLE_DIR, f)
            if os.path.exists(p) and os.path.getmtime(p) > t:
                src_newer = True
    build(force=src_newer)
    L = C.CDLL(LIB_PATH)
    u64p = C.POINTER(C.c_uint64)
    dp = C.POINTER(C.c_double)
    L.oracle_rng_seed1.argtypes = [C.c_uint64, u64p]
    L.oracle_rng_seed2.argtypes = [C.c_int64, C.c_int64, u64p]
    L.oracle_rng_next.argtypes = [u64p]
    L.oracle_rng_next.restype = C.c_uint64
    L.oracle_rng_uniform01.argtypes = [u64p]
    L.oracle_rng_uniform01.restype = C.c_double
    L.oracle_rng_uniform_range.argtypes = [u64p, C.c_double, C.c_double]
    L.oracle_rng_uniform_range.restype = C.c_double
    for name in ("oracle_det_sincos", "oracle_libm_sincos"):
        getattr(L, name).argtypes = [dp, dp, dp, C.c_int64]
    for name in ("oracle_det_pow", "oracle_det_pow_general", "oracle_libm_pow"):
        getattr(L, name).argtypes = [dp, dp, dp, C.c_int64]
    L.oracle_random_scene.argtypes = [C.c_uint64, C.c_int32, C.c_void_p, C.c_int64]
    L.oracle_random_scene.restype = C.c_int64
    L.oracle_camera.argtypes = [dp, dp, dp] + [C.c_double] * 6 + [dp]
    L.oracle_anim_create.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float]
    L.oracle_anim_create.restype = C.c_void_p
    L.oracle_anim_destroy.argtypes = [C.c_void_p]
    L.oracle_anim_num_spheres.argtypes = [C.c_void_p]
    L.oracle_anim_num_spheres.restype = C.c_int64
    L.oracle_anim_next_frame.argtypes = [C.c_void_p, C.c_int32, C.c_int32, dp, C.c_void_p, C.c_int64]
    L.oracle_anim_next_frame.restype = C.c_int64
    L.oracle_render.argtypes = [
        dp, C.c_int32, C.c_int32, C.c_int32, C.c_float, dp, C.c_void_p, C.c_int64, C.c_int64,
        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, u64p,
    ]
    L.oracle_render.restype = C.c_int
    L.oracle_render_split.argtypes = [
        dp, C.c_int32, C.c_int32, C.c_int32, C.c_float, dp, C.c_void_p, C.c_int64, C.c_int64,
        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, dp, dp, u64p,
    ]
    L.oracle_render_split.restype = C.c_int
    L.oracle_num_threads.restype = C.c_int32
    L.oracle_quantise_rgb8.argtypes = [dp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
    L.oracle_export_ppm.argtypes = [dp, C.c_int32, C.c_int32, C.c_char_p]
    L.oracle_export_ppm.restype = C.c_int
    u8p = C.POINTER(C.c_uint8)
    L.oracle_to_rgb_raw.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32, u8p]
    L.oracle_rgb_to_ycbcr420.argtypes = [C.c_int32, C.c_int32, u8p, u8p, u8p, u8p]
    L.oracle_rgb_to_ycbcr420.restype = C.c_int
    L.oracle_bt601_coefs.argtypes = [u8p]
    L.oracle_h264_header.argtypes = [C.c_int32, C.c_int32, u8p, C.c_int64]
    L.oracle_h264_header.restype = C.c_int64
    L.oracle_h264_frame.argtypes = [C.c_int32, C.c_int32, u8p, u8p, u8p, u8p, C.c_int64]
    L.oracle_h264_frame.restype = C.c_int64
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ------------------------------------------------------------------------------------- RNG
class Rng:
    """support/rng.nim `Rng` driven through the oracle."""

    def __init__(self):
        self.state = (C.c_uint64 * 4)()

    def seed(self, x, y=None):
        if y is None:
            lib().oracle_rng_seed1(x, self.state)
        else:
            lib().oracle_rng_seed2(x, y, self.state)
        return self

    def words(self):
        return [int(v) for v in self.state]

    def next(self):
        return int(lib().oracle_rng_next(self.state))

    def uniform(self, lo=None, hi=None):
        if lo is None:
            return float(lib().oracle_rng_uniform01(self.state))
        return float(lib().oracle_rng_uniform_range(self.state, lo, hi))


# ----------------------------------------------------------------------------- scene, camera
def random_scene(seed=0xFACADE, half=11):
    cap = 4 * half * half + 8
    buf = np.zeros(cap, dtype=HITTABLE_DTYPE)
    n = lib().oracle_random_scene(seed, half, buf.ctypes.data, cap)
    return buf[:n].copy()


def camera(look_from, look_at, vup, vfov_deg, aspect, aperture, focus, t0=0.0, t1=0.0):
    out = np.zeros(24, dtype=np.float64)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (look_from, look_at, vup)]
    lib().oracle_camera(_dp(a[0]), _dp(a[1]), _dp(a[2]), vfov_deg, aspect, aperture, focus, t0, t1, _dp(out))
    return out


def book_camera(aspect=16.0 / 9.0):
    """trace_of_radiance.nim:38-51."""
    return camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, aspect, 0.1, 10.0, 0.0, 1.0)


class Animation:
    """scenes_animated.nim `random_moving_spheres` + `iterator scenes`."""

    def __init__(self, seed=0xFACADE, height=144, width=256, dt=0.005, t_min=0.0, t_max=9.0):
        self.h = lib().oracle_anim_create(seed, height, width, dt, t_min, t_max)
        self.first = True
        self.cap = int(lib().oracle_anim_num_spheres(self.h)) + 8

    def num_spheres(self):
        return int(lib().oracle_anim_num_spheres(self.h))

    def next_frame(self, skip=6):
        cam = np.zeros(24, dtype=np.float64)
        buf = np.zeros(self.cap, dtype=HITTABLE_DTYPE)
        n = lib().oracle_anim_next_frame(self.h, skip, 1 if self.first else 0, _dp(cam), buf.ctypes.data, self.cap)
        self.first = False
        if n == 0:
            return None
        return cam, buf[:n].copy()

    def __del__(self):
        try:
            lib().oracle_anim_destroy(self.h)
        except Exception:
            pass


# ------------------------------------------------------------------------------------ render
def render(nrows, ncols, spp, cam, world, max_depth=50, gamma=2.2, rows=None, math="libm", nthreads=0,
           counters=None, out=None):
    """render.nim:49-68.  rows = (begin, end, step) or None for all.  Returns (nrows, ncols, 3) f64,
    row 0 = bottom of the image; rows not selected are left as they were in `out` (zeros if None)."""
    if out is None:
        out = np.zeros((nrows, ncols, 3), dtype=np.float64)
    rb, re, rs = rows if rows is not None else (0, nrows, 1)
    world = np.ascontiguousarray(world)
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    cnt = (C.c_uint64 * 3)()
    rc = lib().oracle_render(_dp(out), nrows, ncols, spp, gamma, _dp(cam), world.ctypes.data, len(world),
                             max_depth, rb, re, rs, 0 if math == "libm" else 1, nthreads, cnt)
    if rc != 0:
        raise RuntimeError("oracle_render failed")
    if counters is not None:
        counters["primary_rays"] = counters.get("primary_rays", 0) + int(cnt[0])
        counters["segments"] = counters.get("segments", 0) + int(cnt[1])
        counters["sphere_tests"] = counters.get("sphere_tests", 0) + int(cnt[2])
    return out


def render_split(nrows, ncols, spp, cam, world, nsub, max_depth=50, gamma=2.2, rows=None, math="det", nthreads=0,
                 counters=None, stats=False):
    """The split-stream ("fast") mode of include/tor_b200.h restated on the CPU: nsub substreams per pixel.
    Returns the drawn canvas, or (canvas, linear_sum, sum_sq) with stats=True."""
    out = np.zeros((nrows, ncols, 3), dtype=np.float64)
    lin = np.zeros_like(out) if stats else None
    sq = np.zeros_like(out) if stats else None
    rb, re, rs = rows if rows is not None else (0, nrows, 1)
    world = np.ascontiguousarray(world)
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    cnt = (C.c_uint64 * 3)()
    rc = lib().oracle_render_split(_dp(out), nrows, ncols, spp, gamma, _dp(cam), world.ctypes.data, len(world),
                                   max_depth, rb, re, rs, 0 if math == "libm" else 1, nthreads, nsub,
                                   _dp(lin) if stats else None, _dp(sq) if stats else None, cnt)
    if rc != 0:
        raise RuntimeError("oracle_render_split failed")
    if counters is not None:
        counters["primary_rays"] = counters.get("primary_rays", 0) + int(cnt[0])
        counters["segments"] = counters.get("segments", 0) + int(cnt[1])
    return (out, lin, sq) if stats else out


def num_threads():
    return int(lib().oracle_num_threads())


def quantise_rgb8(pixels):
    """io/ppm.nim conv: returns (nrows, ncols, 3) uint8 in PPM order (top row first)."""
    nrows, ncols, _ = pixels.shape
    pixels = np.ascontiguousarray(pixels, dtype=np.float64)
    out = np.zeros((nrows, ncols, 3), dtype=np.uint8)
    lib().oracle_quantise_rgb8(_dp(pixels), nrows, ncols, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def export_ppm(pixels, path):
    nrows, ncols, _ = pixels.shape
    pixels = np.ascontiguousarray(pixels, dtype=np.float64)
    if lib().oracle_export_ppm(_dp(pixels), nrows, ncols, path.encode()) != 0:
        raise OSError(path)


def det_sincos(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    s = np.empty_like(a)
    c = np.empty_like(a)
    lib().oracle_det_sincos(_dp(a), _dp(s), _dp(c), a.size)
    return s, c


def libm_sincos(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    s = np.empty_like(a)
    c = np.empty_like(a)
    lib().oracle_libm_sincos(_dp(a), _dp(s), _dp(c), a.size)
    return s, c


def _pow(fn, x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(np.broadcast_to(y, x.shape), dtype=np.float64)
    o = np.empty_like(x)
    fn(_dp(x), _dp(y), _dp(o), x.size)
    return o


def det_pow(x, y):
    return _pow(lib().oracle_det_pow, x, y)


def det_pow_general(x, y):
    return _pow(lib().oracle_det_pow_general, x, y)


def libm_pow(x, y):
    return _pow(lib().oracle_libm_pow, x, y)


# ------------------------------------------------------------------ io/rgb.nim, color_conversions.nim, h264.nim
def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def to_rgb_raw(pixels, as_written=False):
    """io/rgb.nim:17-31.  as_written=True keeps the reference's row indexing (one row off, top row undefined -> 0)."""
    nrows, ncols, _ = pixels.shape
    pixels = np.ascontiguousarray(pixels, dtype=np.float64)
    out = np.zeros((nrows, ncols, 3), dtype=np.uint8)
    lib().oracle_to_rgb_raw(_dp(pixels), nrows, ncols, 1 if as_written else 0, _u8p(out))
    return out


def rgb_to_ycbcr420(rgb):
    """io/color_conversions.nim:180-252 (BT.601): (h, w, 3) uint8 -> Y (h, w), Cb, Cr (h/2, w/2)."""
    h, w, _ = rgb.shape
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    y = np.zeros((h, w), dtype=np.uint8)
    cb = np.zeros(((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
    cr = np.zeros_like(cb)
    if lib().oracle_rgb_to_ycbcr420(w, h, _u8p(rgb), _u8p(y), _u8p(cb), _u8p(cr)) != 0:
        raise ValueError("width and height must be even (color_conversions.nim:201-202)")
    return y, cb, cr


def bt601_coefs():
    out = np.zeros(7, dtype=np.uint8)
    lib().oracle_bt601_coefs(_u8p(out))
    return dict(zip(("kr", "kg", "kb", "fb", "fr", "y_scale", "y_min"), (int(v) for v in out)))


def h264_header(width, height):
    """io/h264.nim:159-168: SPS + PPS."""
    buf = np.zeros(64, dtype=np.uint8)
    n = lib().oracle_h264_header(width, height, _u8p(buf), buf.size)
    return bytes(buf[:n])


def h264_frame(y, cb, cr):
    """io/h264.nim:249-259: one I_PCM slice."""
    h, w = y.shape
    y, cb, cr = (np.ascontiguousarray(a, dtype=np.uint8) for a in (y, cb, cr))
    cap = 16 + (h // 16) * (w // 16) * 386 + 8
    buf = np.zeros(cap, dtype=np.uint8)
    n = lib().oracle_h264_frame(w, h, _u8p(y), _u8p(cb), _u8p(cr), _u8p(buf), cap)
    assert n <= cap
    return bytes(buf[:n])
