"""Host-side BVH builder (trace_of_radiance_b200/csrc/tor_bvh.hpp), checked on the CPU: structure invariants, and a
host restatement of the kernel's float32 traversal that must reach every object the reference's float64 scan hits
(tests/native/bvh_check.cc).  The GPU parity tests check the kernel itself; this one needs no device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "bvh_check.cc")
OUT = os.path.join(ROOT, "tests", "native", "build", "libbvh_check.so")


@pytest.fixture(scope="module")
def chk():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "trace_of_radiance_b200", "csrc", "tor_bvh.hpp")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-o", OUT, SRC])
    L = C.CDLL(OUT)
    L.bvh_check_structure.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64)]
    L.bvh_check_rays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.bvh_check_rays.restype = C.c_int64
    L.bvh_check_rays_coop.argtypes = L.bvh_check_rays.argtypes
    L.bvh_check_rays_coop.restype = C.c_int64
    L.bvh_blob_hash.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64)]
    L.bvh_blob_hash.restype = C.c_uint64
    return L


def _cam(oracle, t0=0.0, t1=1.0):
    return np.ascontiguousarray(oracle.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16 / 9, 0.1, 10.0, t0, t1))


def _rays(rng, objs, n, cam_origin=(13.0, 2.0, 3.0)):
    """Rays like the renderer's: from the camera towards the scene, and from points on sphere surfaces in random
    directions; times in [0, 1] and exactly 0.0 (rays.nim:19)."""
    rays = np.zeros((n, 7))
    k = n // 3
    rays[:k, 0:3] = np.asarray(cam_origin) + rng.uniform(-0.05, 0.05, (k, 3))
    target = rng.uniform(-12, 12, (k, 3)) * (1, 0.1, 1)
    rays[:k, 3:6] = target - rays[:k, 0:3]
    idx = rng.integers(0, len(objs), n - k)
    u = rng.normal(size=(n - k, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    rays[k:, 0:3] = objs["center0"][idx] + u * np.abs(objs["radius"][idx])[:, None]
    v = rng.normal(size=(n - k, 3))
    rays[k:, 3:6] = u + v / np.linalg.norm(v, axis=1, keepdims=True)
    rays[:, 6] = np.where(rng.uniform(size=n) < 0.3, 0.0, rng.uniform(0, 1, n))
    return np.ascontiguousarray(rays)


@pytest.mark.parametrize("half", [11, 3, 30])
def test_structure_random_scene(chk, oracle, half):
    objs = oracle.random_scene(0xFACADE, half)
    info = (C.c_int64 * 4)()
    cam = _cam(oracle)
    assert chk.bvh_check_structure(objs.ctypes.data, len(objs), cam.ctypes.data, info) == 0
    assert info[3] == 1  # the r = 1000 ground sphere spans the scene: tested up front, not in the tree
    assert info[1] >= (len(objs) - 1) / 4 and info[2] <= 40


def test_rays_reach_every_hit_object(chk, oracle):
    rng = np.random.default_rng(7)
    objs = oracle.random_scene(0xFACADE, 11)
    cam = _cam(oracle)
    rays = _rays(rng, objs, 60000)
    stats = (C.c_int64 * 2)()
    assert chk.bvh_check_rays(objs.ctypes.data, len(objs), cam.ctypes.data, rays.ctypes.data, len(rays), stats) == 0
    assert stats[1] < 0.05 * len(rays) * len(objs)  # the hierarchy actually prunes
    # the box tables of the warp-cooperative search reach the same hits, with few candidates per ray
    cstats = (C.c_int64 * 2)()
    assert chk.bvh_check_rays_coop(objs.ctypes.data, len(objs), cam.ctypes.data, rays.ctypes.data, len(rays), cstats) == 0
    assert cstats[1] < 0.05 * len(rays) * len(objs)


def test_degenerate_inputs(chk, oracle):
    rng = np.random.default_rng(3)
    objs = oracle.random_scene(5, 2)[:12].copy()
    objs["kind"][3] = 1  # a mover with time0 == time1 (division by zero): must sit outside the tree
    objs["center1"][3] = objs["center0"][3] + (0, 0.3, 0)
    objs["time0"][3] = objs["time1"][3] = 0.5
    objs["center0"][5] = objs["center0"][4]  # identical spheres: exact ties
    objs["radius"][5] = objs["radius"][4]
    objs["radius"][6] = -abs(objs["radius"][6])  # negative radius
    cam = _cam(oracle, -0.5, 2.5)
    info = (C.c_int64 * 4)()
    assert chk.bvh_check_structure(objs.ctypes.data, len(objs), cam.ctypes.data, info) == 0
    assert info[3] >= 2  # ground + the degenerate mover
    rays = _rays(rng, objs, 20000)
    rays[::7, 3] = 0.0  # axis-parallel and tiny direction components
    rays[::11, 4] = 1e-300
    rays[::13, 3:6] *= 1e-20  # outside the float32 range the padding was derived for: "test everything" route
    rays[::17, 3:6] = np.nan
    assert chk.bvh_check_rays(objs.ctypes.data, len(objs), cam.ctypes.data, rays.ctypes.data, len(rays), None) == 0
    assert chk.bvh_check_rays_coop(objs.ctypes.data, len(objs), cam.ctypes.data, rays.ctypes.data, len(rays), None) == 0
    # single object and two objects: degenerate trees
    for n in (1, 2):
        small = objs[1:1 + n].copy()
        assert chk.bvh_check_structure(small.ctypes.data, n, cam.ctypes.data, info) == 0
        assert chk.bvh_check_rays(small.ctypes.data, n, cam.ctypes.data, rays.ctypes.data, 5000, None) == 0
        assert chk.bvh_check_rays_coop(small.ctypes.data, n, cam.ctypes.data, rays.ctypes.data, 5000, None) == 0


# (scene, node count, FNV-1a of the packed blob) recorded from the builder the GPU kernel was tuned and profiled with
FROZEN_C2_TREE = 0x1BC729B594CFFB45  # random_scene(0xFACADE, 11): 485 objects, 327 nodes
FROZEN_TREES = {
    0: (2, 0xE0BFDF8683103940), 1: (4, 0xA4651621F9916A30), 2: (10, 0x10374B49BA599AF2), 4: (42, 0xAB9D623490AB08B9),
}


def test_builder_output_is_frozen(chk, oracle):
    """Speed work on the host-side builder must not change the trees: node and record blobs are compared by digest
    for the small random_scene sizes, and for the larger ones against a second build (determinism)."""
    cam = _cam(oracle)
    for half, (nodes, digest) in FROZEN_TREES.items():
        objs = oracle.random_scene(0xFACADE, half)
        n = C.c_int64()
        h = chk.bvh_blob_hash(objs.ctypes.data, len(objs), cam.ctypes.data, C.byref(n))
        assert (n.value, h) == (nodes, digest), (half, n.value, hex(h))
    objs = oracle.random_scene(0xFACADE, 11)
    n = C.c_int64()
    h11 = chk.bvh_blob_hash(objs.ctypes.data, len(objs), cam.ctypes.data, C.byref(n))
    assert n.value == 327 and h11 == chk.bvh_blob_hash(objs.ctypes.data, len(objs), cam.ctypes.data, None)
    assert h11 == FROZEN_C2_TREE
