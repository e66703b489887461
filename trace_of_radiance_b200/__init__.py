"""trace_of_radiance_b200 — B200-native render path behind trace-of-radiance's render() API.

The product is `lib/libtor_b200.so` (C ABI: include/tor_b200.h; CUDA kernels: csrc/).  This package
is the Python host-side mirror of the reference's Nim interface for the path (same names and
argument meaning as the Nim procs, cited per function) used by the tests and bench.py.  There is
no CPU fallback: importing works anywhere, but every compute call fails loudly without the
sm_100a library and a CUDA device.
"""
from .api import (  # noqa: F401
    Animation,
    DeviceAnimation,
    PinnedBuffer,
    H264Encoder,
    MP4Muxer,
    render_animation_mp4,
    render_animation,
    HITTABLE_DTYPE,
    Camera,
    Canvas,
    Context,
    HittableList,
    Scene,
    TorError,
    camera,
    default_context,
    dielectric,
    exportToPPM,
    lambertian,
    lib_path,
    load_library,
    metal,
    movingSphere,
    newCanvas,
    random_scene,
    render,
    sphere,
)
