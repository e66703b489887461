"""Host-side mirror of the reference's Nim API over the C ABI (ctypes).

Reference interface mirrored here (paths under /root/reference/trace_of_radiance/):
  render(canvas, cam, world, max_depth)            render.nim:49
  camera(lookFrom, lookAt, view_up, vfov, ...)     physics/cameras.nim:24-45
  Scene.add / Scene.list() -> HittableList         physics/hittables/hittables_lists.nim:15-46
  sphere / movingSphere                            physics/hittables/spheres.nim:20-26, moving_spheres.nim:22-37
  lambertian / metal / dielectric                  physics/materials.nim:21,35,52
  newCanvas / Canvas                               primitives/canvas.nim:21-45
  exportToPPM                                      io/ppm.nim:14-27
  random_scene                                     scenes.nim:13-50
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libtor_b200.so"

# tor_hittable (include/tor_b200.h), 112 bytes
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

TOR_SPHERE, TOR_MOVING_SPHERE = 0, 1
TOR_LAMBERTIAN, TOR_METAL, TOR_DIELECTRIC = 0, 1, 2
TOR_MODE_EXACT = 0x0
TOR_MODE_FAST = 0x1  # split-stream mode (include/tor_b200.h): n RNG substreams per pixel, deterministic, float64


def TOR_FAST_SUBSTREAMS(n):
    """Number of sample ranges per pixel for TOR_MODE_FAST (power of two <= 32; 0 = chosen from canvas size and spp)."""
    return (int(n) & 0xFF) << 16


TOR_FLAG_COUNT_SEGMENTS = 0x100
TOR_FLAG_ROW_MAJOR_QUEUE = 0x400  # BVH route without the longest-pixel-first pre-pass (same image)
TOR_FLAG_RGB_ROWS_AS_WRITTEN = 0x800  # video export: keep io/rgb.nim:29-31's off-by-one row indexing
TOR_FLAG_BRUTE_FORCE = 0x200  # scan every object like hittables_lists.nim:48-55 instead of the BVH (same image)

EXPORTED_SYMBOLS = [
    "tor_abi_version", "tor_ctx_create", "tor_ctx_destroy", "tor_last_error", "tor_render", "tor_render_rows",
    "tor_scene_upload", "tor_render_device_async", "tor_sync", "tor_get_counters", "tor_last_kernel_ms",
    "tor_launch_count", "tor_measure_fp64_peak", "tor_get_traversal_counters", "tor_scene_info", "tor_camera_make", "tor_random_scene", "tor_export_ppm", "tor_quantise_rgb8", "tor_animation_create",
    "tor_animation_next_frame", "tor_animation_destroy", "tor_render_rgb8", "tor_render_rgb8_async", "tor_host_alloc",
    "tor_host_free", "tor_render_ycbcr420", "tor_render_ycbcr420_async", "tor_h264_open", "tor_h264_frame_buffer",
    "tor_h264_flush_frame", "tor_h264_finish", "tor_mp4_mux_h264_file", "tor_fast_substream_count", "tor_last_schedule", "tor_last_handoffs", "tor_download_rows_async", "tor_animation_dev_create",
    "tor_animation_dev_next", "tor_animation_dev_sync", "tor_animation_dev_launch_count", "tor_animation_dev_destroy",
    "tor_animation_dev_reset", "tor_debug_times",
]


class TorError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tor_b200 error {code}: {msg}")
        self.code = code


class _CCamera(C.Structure):  # tor_camera == cameras.nim:15-22
    _fields_ = [(n, C.c_double * 3) for n in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w")] + [
        ("lens_radius", C.c_double), ("shutter_open", C.c_double), ("shutter_close", C.c_double)]


class _CCanvas(C.Structure):  # tor_canvas == canvas.nim:21-28
    _fields_ = [("pixels", C.POINTER(C.c_double)), ("nrows", C.c_int32), ("ncols", C.c_int32),
                ("samples_per_pixel", C.c_int32), ("gamma_correction", C.c_float)]


assert C.sizeof(_CCamera) == 192 and C.sizeof(_CCanvas) == 24

_lib = None


def lib_path():
    return os.path.join(_HERE, "lib", _LIB_NAME)


def load_library():
    """Load libtor_b200.so.  Raises (never falls back) when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make -C trace_of_radiance_b200/csrc` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.tor_abi_version.restype = C.c_int
    L.tor_ctx_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.tor_ctx_destroy.argtypes = [vp]
    L.tor_ctx_destroy.restype = None
    L.tor_last_error.argtypes = [vp]
    L.tor_last_error.restype = C.c_char_p
    L.tor_render.argtypes = [vp, C.POINTER(_CCanvas), C.POINTER(_CCamera), vp, C.c_int64, C.c_int64, C.c_int64, C.c_uint32]
    L.tor_render_rows.argtypes = L.tor_render.argtypes + [C.c_int32, C.c_int32, C.c_int32]
    L.tor_render_rgb8.argtypes = [vp, C.POINTER(_CCanvas), C.POINTER(_CCamera), vp, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_uint32, vp]
    L.tor_render_rgb8_async.argtypes = L.tor_render_rgb8.argtypes
    L.tor_render_ycbcr420.argtypes = L.tor_render_rgb8.argtypes
    L.tor_render_ycbcr420_async.argtypes = L.tor_render_rgb8.argtypes
    L.tor_h264_open.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.tor_h264_frame_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64)]
    L.tor_h264_flush_frame.argtypes = [vp]
    L.tor_h264_finish.argtypes = [vp]
    L.tor_mp4_mux_h264_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
    L.tor_fast_substream_count.argtypes = [C.c_uint32, C.c_int32, C.c_int32, C.c_int32]
    L.tor_fast_substream_count.restype = C.c_int
    L.tor_host_alloc.argtypes = [C.c_size_t]
    L.tor_host_alloc.restype = vp
    L.tor_host_free.argtypes = [vp]
    L.tor_host_free.restype = None
    L.tor_scene_upload.argtypes = [vp, C.POINTER(_CCamera), vp, C.c_int64, C.c_int64]
    L.tor_render_device_async.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int64, C.c_uint32,
                                          C.c_int32, C.c_int32, C.c_int32, vp]
    L.tor_sync.argtypes = [vp]
    L.tor_download_rows_async.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    L.tor_get_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.tor_get_traversal_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.tor_scene_info.argtypes = [vp, C.POINTER(C.c_int64)]
    L.tor_last_schedule.argtypes = [vp, C.POINTER(C.c_int64)]
    L.tor_last_handoffs.argtypes = [vp, C.POINTER(C.c_int64)]
    L.tor_debug_times.argtypes = [vp, vp, C.c_int64]
    L.tor_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.tor_launch_count.argtypes = [vp]
    L.tor_launch_count.restype = C.c_int64
    L.tor_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.tor_camera_make.argtypes = [C.POINTER(_CCamera)] + [C.POINTER(C.c_double)] * 3 + [C.c_double] * 6
    L.tor_camera_make.restype = None
    L.tor_random_scene.argtypes = [C.c_uint64, C.c_int32, vp, C.c_int64]
    L.tor_random_scene.restype = C.c_int64
    L.tor_animation_create.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float]
    L.tor_animation_create.restype = vp
    L.tor_animation_next_frame.argtypes = [vp, C.c_int32, C.POINTER(_CCamera), vp, C.c_int64]
    L.tor_animation_next_frame.restype = C.c_int64
    L.tor_animation_destroy.argtypes = [vp]
    L.tor_animation_destroy.restype = None
    L.tor_animation_dev_create.argtypes = [vp, C.c_uint64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32,
                                           C.c_int32, C.POINTER(vp)]
    L.tor_animation_dev_next.argtypes = [vp, C.c_int32, C.c_float, C.c_int64, C.c_uint32, vp, C.POINTER(C.c_int64)]
    L.tor_animation_dev_sync.argtypes = [vp, C.POINTER(C.c_float)]
    L.tor_animation_dev_reset.argtypes = [vp]
    L.tor_animation_dev_launch_count.argtypes = [vp]
    L.tor_animation_dev_launch_count.restype = C.c_int64
    L.tor_animation_dev_destroy.argtypes = [vp]
    L.tor_animation_dev_destroy.restype = None
    L.tor_export_ppm.argtypes = [C.POINTER(_CCanvas), C.c_char_p]
    L.tor_quantise_rgb8.argtypes = [C.POINTER(_CCanvas), C.POINTER(C.c_uint8)]
    _lib = L
    return L


# ------------------------------------------------------------------------------- materials
def lambertian(albedo):
    """materials.nim:21"""
    return (TOR_LAMBERTIAN, tuple(float(a) for a in albedo), 0.0)


def metal(albedo, fuzz):
    """materials.nim:35-37 — fuzz is clamped to <= 1."""
    fuzz = float(fuzz)
    return (TOR_METAL, tuple(float(a) for a in albedo), fuzz if fuzz <= 1.0 else 1.0)


def dielectric(refraction_index):
    """materials.nim:52"""
    return (TOR_DIELECTRIC, (0.0, 0.0, 0.0), float(refraction_index))


# ------------------------------------------------------------------------------- hittables
def sphere(center, radius, material):
    """spheres.nim:20-26"""
    h = np.zeros((), dtype=HITTABLE_DTYPE)
    h["kind"] = TOR_SPHERE
    h["mat_kind"], h["albedo"], h["fuzz_or_ior"] = material
    h["center0"] = center
    h["radius"] = radius
    return h


def movingSphere(center0, time0, center1, time1, radius, material):
    """moving_spheres.nim:22-37"""
    h = sphere(center0, radius, material)
    h["kind"] = TOR_MOVING_SPHERE
    h["center1"] = center1
    h["time0"] = time0
    h["time1"] = time1
    return h


class HittableList:
    """hittables_lists.nim:20-24 — a non-owning (len, pointer) view; here a contiguous record array."""

    def __init__(self, objects):
        self.objects = np.ascontiguousarray(objects, dtype=HITTABLE_DTYPE)

    def __len__(self):
        return int(self.objects.shape[0])


class Scene:
    """hittables_lists.nim:15-46"""

    def __init__(self, objects=None):
        self._objs = [] if objects is None else [o for o in np.asarray(objects, dtype=HITTABLE_DTYPE)]

    def add(self, h):
        self._objs.append(np.asarray(h, dtype=HITTABLE_DTYPE))

    def clear(self):
        self._objs = []

    def __len__(self):
        return len(self._objs)

    def list(self):
        assert len(self._objs) > 0  # hittables_lists.nim:42
        return HittableList(np.array(self._objs, dtype=HITTABLE_DTYPE))


def random_scene(seed=0xFACADE, half=11):
    """scenes.nim:13-50 with `worldRNG.seed(seed)` (trace_of_radiance.nim:34-36)."""
    L = load_library()
    n = L.tor_random_scene(seed, half, None, 0)
    buf = np.zeros(n, dtype=HITTABLE_DTYPE)
    L.tor_random_scene(seed, half, buf.ctypes.data, n)
    return Scene(buf)


class Animation:
    """scenes_animated.nim:57-225: `random_moving_spheres` + `iterator scenes(anim, skip)`.

        for cam, world in tor.Animation(height=144, width=256).scenes(skip=6): tor.render(canvas, cam, world, 50)
    """

    def __init__(self, seed=0xFACADE, height=144, width=256, dt=0.005, t_min=0.0, t_max=9.0):
        self.L = load_library()
        self.height, self.width = int(height), int(width)
        self.h = self.L.tor_animation_create(seed, height, width, dt, t_min, t_max)

    def scenes(self, skip=6):
        cap = 40 * 40 + 4
        while True:
            cam = _CCamera()
            buf = np.zeros(cap, dtype=HITTABLE_DTYPE)
            n = self.L.tor_animation_next_frame(self.h, skip, C.byref(cam), buf.ctypes.data, cap)
            if n <= 0:
                return
            yield Camera(cam), HittableList(buf[:n])

    def __del__(self):
        try:
            if self.h:
                self.L.tor_animation_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ---------------------------------------------------------------------------------- camera
class Camera:
    """cameras.nim:15-22 as 24 float64."""

    def __init__(self, c=None):
        self.c = c if c is not None else _CCamera()

    def as_array(self):
        return np.frombuffer(bytes(self.c), dtype=np.float64).copy()

    @staticmethod
    def from_array(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == 24
        return Camera(_CCamera.from_buffer_copy(a.tobytes()))


def camera(lookFrom, lookAt, view_up, vertical_field_of_view, aspect_ratio, aperture, focus_distance,
           shutterOpen=0.0, shutterClose=0.0):
    """cameras.nim:24-45"""
    L = load_library()
    v = [(C.c_double * 3)(*[float(x) for x in p]) for p in (lookFrom, lookAt, view_up)]
    cam = _CCamera()
    L.tor_camera_make(C.byref(cam), v[0], v[1], v[2], vertical_field_of_view, aspect_ratio, aperture, focus_distance,
                      shutterOpen, shutterClose)
    return Camera(cam)


# ---------------------------------------------------------------------------------- canvas
class Canvas:
    """canvas.nim:21-28.  pixels: (nrows, ncols, 3) float64, row 0 = bottom of the image."""

    def __init__(self, height, width, samples_per_pixel, gamma_correction):
        self.nrows, self.ncols = int(height), int(width)
        self.samples_per_pixel = int(samples_per_pixel)
        self.gamma_correction = float(np.float32(gamma_correction))
        self.pixels = np.zeros((self.nrows, self.ncols, 3), dtype=np.float64)

    def _c(self):
        return _CCanvas(self.pixels.ctypes.data_as(C.POINTER(C.c_double)), self.nrows, self.ncols,
                        self.samples_per_pixel, self.gamma_correction)

    def __getitem__(self, rc):  # canvas.nim:56-57
        return self.pixels[rc[0], rc[1]]

    def toRGB8(self):
        """io/ppm.nim quantisation, PPM row order (top row first): (nrows, ncols, 3) uint8."""
        out = np.zeros((self.nrows, self.ncols, 3), dtype=np.uint8)
        c = self._c()
        rc = load_library().tor_quantise_rgb8(C.byref(c), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        if rc:
            raise TorError(rc, "tor_quantise_rgb8")
        return out


def newCanvas(height, width, samples_per_pixel, gamma_correction):
    """canvas.nim:31-41"""
    return Canvas(height, width, samples_per_pixel, gamma_correction)


def exportToPPM(canvas, path):
    """io/ppm.nim:14-27 (to a file path instead of a Nim File)."""
    c = canvas._c()
    rc = load_library().tor_export_ppm(C.byref(c), os.fsencode(path))
    if rc:
        raise TorError(rc, f"cannot write {path}")


# --------------------------------------------------------------------------------- context
class Context:
    """tor_ctx: device buffers + streams for one calling thread."""

    def __init__(self, devices=None):
        self.L = load_library()
        self.h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.L.tor_ctx_create(arr, len(devices), C.byref(self.h))
        else:
            rc = self.L.tor_ctx_create(None, 0, C.byref(self.h))
        if rc:
            raise TorError(rc, (self.L.tor_last_error(None) or b"").decode())

    def _check(self, rc):
        if rc:
            raise TorError(rc, (self.L.tor_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.L.tor_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, canvas, cam, world, max_depth, flags=0, rows=None):
        c = canvas._c()
        objs = world.objects
        if rows is None:
            self._check(self.L.tor_render(self.h, C.byref(c), C.byref(cam.c), objs.ctypes.data, len(objs),
                                          objs.dtype.itemsize, max_depth, flags))
        else:
            rb, re, rs = rows
            self._check(self.L.tor_render_rows(self.h, C.byref(c), C.byref(cam.c), objs.ctypes.data, len(objs),
                                               objs.dtype.itemsize, max_depth, flags, rb, re, rs))

    def render_rgb8(self, canvas, cam, world, max_depth, flags=0, out=None, wait=True):
        """render + io/ppm.nim quantisation on the device: returns the (nrows, ncols, 3) uint8 image in PPM row
        order.  wait=False enqueues only (pass a pinned `out`, see PinnedBuffer, and call sync())."""
        if out is None:
            out = np.empty((canvas.nrows, canvas.ncols, 3), dtype=np.uint8)
        c = _CCanvas(None, canvas.nrows, canvas.ncols, canvas.samples_per_pixel, canvas.gamma_correction)
        objs = world.objects
        fn = self.L.tor_render_rgb8 if wait else self.L.tor_render_rgb8_async
        self._check(fn(self.h, C.byref(c), C.byref(cam.c), objs.ctypes.data, len(objs), objs.dtype.itemsize, max_depth,
                       flags, out.ctypes.data))
        return out

    def render_ycbcr420(self, canvas, cam, world, max_depth, flags=0, out=None, wait=True):
        """render + io/rgb.nim quantisation + io/color_conversions.nim BT.601 4:2:0 conversion on the device.  `out`:
        uint8 array of nrows*ncols*3/2 bytes (Y' plane, Cb, Cr), e.g. H264Encoder.frame; returns (Y, Cb, Cr) views."""
        h, w = canvas.nrows, canvas.ncols
        if out is None:
            out = np.empty(h * w * 3 // 2, dtype=np.uint8)
        c = _CCanvas(None, h, w, canvas.samples_per_pixel, canvas.gamma_correction)
        objs = world.objects
        fn = self.L.tor_render_ycbcr420 if wait else self.L.tor_render_ycbcr420_async
        self._check(fn(self.h, C.byref(c), C.byref(cam.c), objs.ctypes.data, len(objs), objs.dtype.itemsize, max_depth,
                       flags, out.ctypes.data))
        return split_ycbcr420(out, h, w)

    def render_raw(self, canvas, cam, objects_ptr, length, stride, max_depth, flags=0):
        """Same call with an explicit (pointer, len, stride) — e.g. the 120-byte Nim variant encoding."""
        c = canvas._c()
        self._check(self.L.tor_render(self.h, C.byref(c), C.byref(cam.c), objects_ptr, length, stride, max_depth, flags))

    def scene_upload(self, cam, world):
        objs = world.objects
        self._check(self.L.tor_scene_upload(self.h, C.byref(cam.c), objs.ctypes.data, len(objs), objs.dtype.itemsize))

    def render_device_async(self, d_pixels_ptr, nrows, ncols, spp, gamma, max_depth, flags=0, rows=None, stream=None):
        """stream: a cudaStream_t handle as int (e.g. torch.cuda.current_stream().cuda_stream) or None for the
        context's own stream.  Handle 0 is CUDA's legacy default stream and is passed as cudaStreamLegacy (0x1),
        because the C ABI reserves NULL for "the context's stream"."""
        rb, re, rs = rows if rows is not None else (0, nrows, 1)
        if stream is not None and int(stream) == 0:
            stream = 1  # cudaStreamLegacy
        self._check(self.L.tor_render_device_async(self.h, d_pixels_ptr, nrows, ncols, spp, float(np.float32(gamma)),
                                                   max_depth, flags, rb, re, rs, stream))

    def download_rows_async(self, d_rows_ptr, host_pixels, ncols, row_begin, row_step, nsel, stream=None):
        """Compact device rows -> rows row_begin, row_begin + row_step, ... of the host canvas array (one strided copy)."""
        if stream is not None and int(stream) == 0:
            stream = 1  # cudaStreamLegacy
        self._check(self.L.tor_download_rows_async(self.h, d_rows_ptr, host_pixels.ctypes.data, ncols, row_begin, row_step,
                                                   nsel, stream))

    def sync(self):
        self._check(self.L.tor_sync(self.h))

    def counters(self):
        out = (C.c_uint64 * 3)()
        self._check(self.L.tor_get_counters(self.h, out))
        tr = (C.c_uint64 * 2)()
        self._check(self.L.tor_get_traversal_counters(self.h, tr))
        return {"primary_rays": int(out[0]), "segments": int(out[1]), "sphere_tests": int(out[2]),
                "bvh_node_visits": int(tr[0]), "bvh_sphere_tests": int(tr[1])}

    def scene_info(self):
        out = (C.c_int64 * 6)()
        self._check(self.L.tor_scene_info(self.h, out))
        keys = ("objects", "bvh_nodes", "bvh_leaves", "bvh_depth", "objects_outside_tree", "bvh_bytes")
        return dict(zip(keys, (int(v) for v in out)))

    def last_schedule(self):
        """Pixel scheduling of the last exact-mode render with a cost pre-pass (tor_last_schedule)."""
        out = (C.c_int64 * 4)()
        self._check(self.L.tor_last_schedule(self.h, out))
        return {"cooperative_pixels": int(out[0]), "qualifying_pixels": int(out[1]), "prepass_segments": int(out[2]),
                "devices": int(out[3])}

    def last_handoffs(self):
        """Pixels parked in the tail of the last cost-ranked render and finished by one warp each (tor_last_handoffs)."""
        out = (C.c_int64 * 2)()
        self._check(self.L.tor_last_handoffs(self.h, out))
        return {"from_cooperative_warps": int(out[0]), "from_lanes": int(out[1])}

    def debug_times(self):
        """(coop (4096, 3): start, end, segments; lanes (8192, 2): start, end; hand-off launch (4096, 3) like coop) in
        ns (tor_debug_times)."""
        out = np.zeros(3 * 4096 + 2 * 8192 + 3 * 4096, dtype=np.uint64)
        self._check(self.L.tor_debug_times(self.h, out.ctypes.data, out.size))
        a, b = 3 * 4096, 3 * 4096 + 2 * 8192
        return out[:a].reshape(4096, 3), out[a:b].reshape(8192, 2), out[b:].reshape(4096, 3)

    def last_kernel_ms(self):
        ms = C.c_float()
        self._check(self.L.tor_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def measure_fp64_peak(self):
        """Sustained FP64 FMA instructions per second per GPU (lane-level), from a register-only DFMA loop."""
        v = C.c_double()
        self._check(self.L.tor_measure_fp64_peak(self.h, C.byref(v)))
        return float(v.value)

    def launch_count(self):
        return int(self.L.tor_launch_count(self.h))


class PinnedBuffer:
    """Page-locked host memory (tor_host_alloc) viewed as a numpy array."""

    def __init__(self, shape, dtype=np.uint8):
        self.L = load_library()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.L.tor_host_alloc(n)
        if not self.ptr:
            raise MemoryError("tor_host_alloc failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * n).from_address(self.ptr)).view(dtype).reshape(shape)

    def __del__(self):
        try:
            if self.ptr:
                self.L.tor_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def fast_substream_count(flags, height, width, samples_per_pixel):
    """Sample ranges per pixel that a render with these flags would use (1 in exact mode); needs no device."""
    n = load_library().tor_fast_substream_count(flags, height, width, samples_per_pixel)
    if n < 0:
        raise TorError(n, "invalid split-stream flags")
    return n


def split_ycbcr420(buf, height, width):
    """Views of a contiguous Y'CbCr 4:2:0 frame buffer: Y (h, w), Cb and Cr (h/2, w/2)."""
    n, q = height * width, (height // 2) * (width // 2)
    flat = buf.reshape(-1)
    return (flat[:n].reshape(height, width), flat[n:n + q].reshape(height // 2, width // 2),
            flat[n + q:n + 2 * q].reshape(height // 2, width // 2))


class H264Encoder:
    """io/h264.nim `H264Encoder`: Baseline SPS/PPS + one I_PCM slice per frame, written to `path`.

        enc = H264Encoder.init(width, height, path); Y, Cb, Cr = enc.getFrameBuffers(); ...; enc.flushFrame(); enc.finish()
    """

    def __init__(self):
        self.L = load_library()
        self.h = None

    @classmethod
    def init(cls, width, height, path):
        """h264.nim:159-168 (takes a path instead of a Nim File)."""
        self = cls()
        h = C.c_void_p()
        rc = self.L.tor_h264_open(os.fsencode(path), width, height, C.byref(h))
        if rc:
            raise TorError(rc, f"tor_h264_open({path}, {width}x{height}): width and height must be even")
        self.h, self.width, self.height = h, width, height
        p, n = C.c_void_p(), C.c_int64()
        self.L.tor_h264_frame_buffer(self.h, C.byref(p), C.byref(n))
        self.frame = np.ctypeslib.as_array((C.c_uint8 * n.value).from_address(p.value))
        return self

    def getFrameBuffers(self):
        """h264.nim:207-224"""
        return split_ycbcr420(self.frame, self.height, self.width)

    def getFrameBuffer(self):
        """h264.nim:226-236"""
        return self.frame

    def flushFrame(self):
        """h264.nim:249-259"""
        rc = self.L.tor_h264_flush_frame(self.h)
        if rc:
            raise TorError(rc, "tor_h264_flush_frame")

    def finish(self):
        """h264.nim:170-173 (also closes the file)."""
        if self.h is not None:
            self.frame = None
            rc = self.L.tor_h264_finish(self.h)
            self.h = None
            if rc:
                raise TorError(rc, "tor_h264_finish")


class MP4Muxer:
    """io/mp4.nim `MP4Muxer` (:104-163)."""

    def initialize(self, path, width, height):
        self.path, self.width, self.height = path, int(width), int(height)
        return self

    def writeMP4_from(self, src, fps=30):
        rc = load_library().tor_mp4_mux_h264_file(os.fsencode(src), os.fsencode(self.path), self.width, self.height, fps)
        if rc:
            raise TorError(rc, f"tor_mp4_mux_h264_file({src} -> {self.path})")

    def close(self):
        pass


def render_animation_mp4(animation, dest_264, dest_mp4, samples_per_pixel=10, max_depth=50, gamma_correction=2.2, skip=6,
                         max_frames=None, flags=0, ctx=None, fps=30):
    """`main_animation_mp4` of trace_of_radiance_animation.nim:101-214: for every (camera, scene) of the animation
    render, convert to Y'CbCr 4:2:0 (both on the device, straight into the encoder's page-locked frame buffer), flush
    one I_PCM frame; then mux the .264 into an .mp4.  Returns the number of frames."""
    h, w = animation_dims(animation)
    canvas = Canvas.__new__(Canvas)
    canvas.nrows, canvas.ncols, canvas.samples_per_pixel = h, w, int(samples_per_pixel)
    canvas.gamma_correction = float(np.float32(gamma_correction))
    canvas.pixels = None
    own = ctx is None
    ctx = ctx or Context()
    enc = H264Encoder.init(w, h, dest_264)
    n = 0
    try:
        for i, (cam, world) in enumerate(animation.scenes(skip=skip)):
            if max_frames is not None and i >= max_frames:
                break
            ctx.render_ycbcr420(canvas, cam, world, max_depth, flags=flags, out=enc.getFrameBuffer())
            enc.flushFrame()
            n += 1
    finally:
        enc.finish()
        if own:
            ctx.close()
    mux = MP4Muxer().initialize(dest_mp4, w, h)
    mux.writeMP4_from(dest_264, fps=fps)
    mux.close()
    return n


def render_animation(animation, samples_per_pixel=100, max_depth=50, gamma_correction=2.2, skip=6, in_flight=4,
                     devices=None, on_frame=None, max_frames=None, flags=None):
    """The frame loop of trace_of_radiance_animation.nim:173-199 (for camera, scene in scenes(animation, skip):
    canvas.render(camera, scene.list(), max_depth); export) with `in_flight` frames enqueued at once, each on its own
    context (= its own stream), so that small frames (256x144) keep the GPU full.  on_frame(index, rgb8) is called in
    frame order with the PPM-order uint8 image; returns the number of frames."""
    h, w = animation_dims(animation)
    canvas = Canvas.__new__(Canvas)
    canvas.nrows, canvas.ncols, canvas.samples_per_pixel = h, w, int(samples_per_pixel)
    canvas.gamma_correction = float(np.float32(gamma_correction))
    canvas.pixels = None
    ctxs = [Context(devices) for _ in range(in_flight)]
    bufs = [PinnedBuffer((h, w, 3)) for _ in range(in_flight)]
    pending = [None] * in_flight
    n = 0
    for i, (cam, world) in enumerate(animation.scenes(skip=skip)):
        if max_frames is not None and i >= max_frames:
            break
        k = i % in_flight
        if pending[k] is not None:
            ctxs[k].sync()
            if on_frame:
                on_frame(pending[k], bufs[k].array)
        ctxs[k].render_rgb8(canvas, cam, world, max_depth, out=bufs[k].array, wait=False, flags=flags or 0)
        pending[k] = i
        n += 1
    for j in sorted((p, k) for k, p in enumerate(pending) if p is not None):
        ctxs[j[1]].sync()
        if on_frame:
            on_frame(j[0], bufs[j[1]].array)
    for c in ctxs:
        c.close()
    return n


class DeviceAnimation:
    """The animation with physics, scene rebuild and BVH re-fit on the device (tor_animation_dev_*): no per-frame
    host-to-device copy.  Same frames, bit for bit, as `Animation` + `render_rgb8`.

        an = DeviceAnimation(ctx, height=144, width=256, in_flight=4)
        n = an.render_all(samples_per_pixel=100, on_frame=lambda i, rgb8: ...)
    """

    def __init__(self, ctx=None, seed=0xFACADE, height=144, width=256, dt=0.005, t_min=0.0, t_max=9.0, skip=6,
                 in_flight=4):
        self.ctx = ctx or default_context()
        self.L = self.ctx.L
        self.height, self.width, self.in_flight = int(height), int(width), int(in_flight)
        self.h = C.c_void_p()
        self.ctx._check(self.L.tor_animation_dev_create(self.ctx.h, seed, height, width, dt, t_min, t_max, skip, in_flight,
                                                         C.byref(self.h)))
        self.bufs = [PinnedBuffer((self.height, self.width, 3)) for _ in range(self.in_flight)]

    def next(self, samples_per_pixel, max_depth=50, gamma_correction=2.2, flags=0, out=None, render=True):
        """Enqueues the next frame.  Returns its index, or None when the iterator is exhausted.  render=False only
        advances the physics (the frame belongs to another rank)."""
        idx = C.c_int64(-1)
        ptr = out.ctypes.data if (render and out is not None) else None
        if render and out is None:
            raise ValueError("render=True needs a (pinned) output array")
        rc = self.L.tor_animation_dev_next(self.h, samples_per_pixel, float(np.float32(gamma_correction)), max_depth, flags,
                                           ptr, C.byref(idx))
        if rc < 0:
            self.ctx._check(rc)
        return int(idx.value) if rc == 1 else None

    def sync(self):
        """Completes every frame enqueued so far; returns the device time (ms) they took."""
        ms = C.c_float()
        self.ctx._check(self.L.tor_animation_dev_sync(self.h, C.byref(ms)))
        return float(ms.value)

    def reset(self):
        """Back to the first frame."""
        self.ctx._check(self.L.tor_animation_dev_reset(self.h))

    def launch_count(self):
        return int(self.L.tor_animation_dev_launch_count(self.h)) + self.ctx.launch_count()

    def render_all(self, samples_per_pixel=100, max_depth=50, gamma_correction=2.2, flags=0, on_frame=None, max_frames=None,
                   rank=0, world=1, out=None):
        """The frame loop of trace_of_radiance_animation.nim:173-199.  Frame f is rendered when f % world == rank (frame-
        parallel ranks step the physics of every frame themselves).  on_frame(index, rgb8) is called in frame order;
        `out` (optional, (frames_of_this_rank, h, w, 3) uint8 pinned array) receives the frames directly instead of the
        rotating buffers.  Returns (frames seen, device ms)."""
        pending = []  # (frame index, buffer index) in flight
        n_mine = 0
        total_ms = 0.0
        f = 0
        while max_frames is None or f < max_frames:
            mine = f % world == rank
            spare = False
            if mine:
                if out is not None:
                    try:
                        buf = out[n_mine]
                    except IndexError:  # `out` is full: this call can only be the one that finds the iterator exhausted
                        buf, spare = self.bufs[0].array, True
                else:
                    k = n_mine % self.in_flight
                    if len(pending) == self.in_flight:  # the buffer is still owned by a frame in flight
                        total_ms += self.sync()
                        for (pf, pk) in pending:
                            if on_frame:
                                on_frame(pf, self.bufs[pk].array)
                        pending = []
                    buf = self.bufs[k].array
                idx = self.next(samples_per_pixel, max_depth, gamma_correction, flags, out=buf)
            else:
                idx = self.next(samples_per_pixel, max_depth, gamma_correction, flags, render=False)
            if idx is None:
                break
            if spare:
                raise ValueError("render_all: `out` holds fewer frames than this rank renders")
            if mine:
                if out is None:
                    pending.append((idx, n_mine % self.in_flight))
                n_mine += 1
            f += 1
        total_ms += self.sync()
        for (pf, pk) in pending:
            if on_frame:
                on_frame(pf, self.bufs[pk].array)
        return f, total_ms

    def close(self):
        if self.h:
            self.L.tor_animation_dev_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def animation_dims(animation):
    return animation.height, animation.width


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def render(canvas, cam, world, max_depth, ctx=None, flags=0):
    """render.nim:49 — `canvas.render(cam, world.list(), max_depth)`.  Synchronous."""
    (ctx or default_context()).render(canvas, cam, world, max_depth, flags=flags)
