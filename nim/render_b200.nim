# render_b200.nim — drop-in for trace_of_radiance/render.nim:49 over libtor_b200.so (include/tor_b200.h).
#
# NOT COMPILED IN THIS REPO'S ENVIRONMENT (no Nim toolchain there): shipped as source for the maintainer.
# Usage in trace_of_radiance.nim / trace_of_radiance_animation.nim:
#     import ./trace_of_radiance/render          ->   import ./render_b200
# The call sites stay as they are:  canvas.render(cam, world.list(), max_depth)
#
# Build:  nim c -d:danger --threads:on --passL:"-L<repo>/trace_of_radiance_b200/lib -ltor_b200" ...
# Needs Nim >= 1.6 for std/importutils (private field access to Camera / HittableList / MovingSphere).

import std/importutils
import ./trace_of_radiance/[primitives, physics/cameras, physics/core, physics/hittables]

type
  TorCtx = distinct pointer
  TorHittable {.bycopy.} = object        # tor_hittable, 112 bytes
    kind, mat_kind: uint32
    center0, center1: array[3, float64]
    time0, time1, radius: float64
    albedo: array[3, float64]
    fuzz_or_ior: float64

const
  TOR_STRIDE_FLAT = 112
  TOR_MODE_EXACT = 0'u32
  TOR_MODE_FAST = 1'u32      # split-stream mode (include/tor_b200.h): deterministic, float64, NOT the reference's image
  # nim c -d:torSplitStream ... selects the split-stream mode for every render() call of the program
  torFlags = when defined(torSplitStream): TOR_MODE_FAST else: TOR_MODE_EXACT

{.push importc, cdecl, dynlib: "libtor_b200.so".}
proc tor_ctx_create(devices: ptr cint, ndev: cint, ctx: ptr TorCtx): cint
proc tor_ctx_destroy(ctx: TorCtx)
proc tor_last_error(ctx: TorCtx): cstring
proc tor_render(ctx: TorCtx, canvas: ptr Canvas, cam: ptr Camera, objects: pointer,
                len, stride, max_depth: int64, flags: uint32): cint
proc tor_render_ycbcr420(ctx: TorCtx, canvas: ptr Canvas, cam: ptr Camera, objects: pointer,
                         len, stride, max_depth: int64, flags: uint32, ycbcr_out: ptr UncheckedArray[uint8]): cint
{.pop.}

static:
  # the C side reinterprets these two objects in place
  doAssert sizeof(Canvas) == 24   # primitives/canvas.nim:21-28
  doAssert sizeof(Camera) == 192  # physics/cameras.nim:15-22: 24 contiguous float64
  doAssert sizeof(TorHittable) == 112

var ctx {.threadvar.}: TorCtx

func toFlat(mat: Material, h: var TorHittable) =
  h.mat_kind = uint32 ord(mat.kind)                 # kLambertian, kMetal, kDielectric = 0, 1, 2
  case mat.kind
  of kLambertian:
    h.albedo = [mat.fLambertian.albedo.x, mat.fLambertian.albedo.y, mat.fLambertian.albedo.z]
  of kMetal:
    h.albedo = [mat.fMetal.albedo.x, mat.fMetal.albedo.y, mat.fMetal.albedo.z]
    h.fuzz_or_ior = mat.fMetal.fuzz
  of kDielectric:
    h.fuzz_or_ior = mat.fDielectric.refraction_index

proc toFlatList(world: HittableList): seq[TorHittable] =
  privateAccess(HittableList)   # len, objects        (hittables_lists.nim:20-24)
  privateAccess(MovingSphere)   # center0, time0, time1 (moving_spheres.nim:15-20)
  var flat = newSeq[TorHittable](world.len)
  for i in 0 ..< world.len:
    let o = world.objects[i]
    flat[i].kind = uint32 ord(o.kind)               # kSphere, kMovingSphere = 0, 1
    case o.kind
    of kSphere:
      let s = o.fSphere
      flat[i].center0 = [s.center.x, s.center.y, s.center.z]
      flat[i].radius = s.radius
      s.material.toFlat(flat[i])
    of kMovingSphere:
      let s = o.fMovingSphere
      flat[i].center0 = [s.center0.x, s.center0.y, s.center0.z]
      flat[i].center1 = [s.center1.x, s.center1.y, s.center1.z]
      flat[i].time0 = float64 s.time0
      flat[i].time1 = float64 s.time1
      flat[i].radius = s.radius
      s.material.toFlat(flat[i])

  flat

proc render*(canvas: var Canvas, cam: Camera, world: HittableList, max_depth: int) =
  ## Same signature as trace_of_radiance/render.nim:49.  Synchronous: the image is complete on
  ## return (the reference completes at syncRoot/exit(Weave)); Weave does not need to be initialised.
  if pointer(ctx).isNil:
    doAssert tor_ctx_create(nil, 0, addr ctx) == 0, $tor_last_error(TorCtx(nil))
  var flat = world.toFlatList()
  var camCopy = cam
  let rc = tor_render(ctx, addr canvas, addr camCopy, addr flat[0], int64 flat.len,
                      TOR_STRIDE_FLAT, int64 max_depth, torFlags)
  doAssert rc == 0, $tor_last_error(ctx)

proc renderYCbCr420*(canvas: var Canvas, cam: Camera, world: HittableList, max_depth: int,
                     frame: ptr UncheckedArray[uint8]) =
  ## For trace_of_radiance_animation.nim:181-195: replaces
  ##   canvas.render(...); syncRoot(Weave); let rgb = canvas.toRGB_Raw(); rgbRaw_to_ycbcr420(...)
  ## by one call that renders and converts on the device and writes Y', Cb, Cr straight into the encoder's frame
  ## buffer (`encoder.getFrameBuffer()`, io/h264.nim:226-236); `encoder.flushFrame()` follows as before.
  ## canvas.pixels is not written.  Rows come out top first (io/rgb.nim:29-31 is one row off; pass
  ## TOR_FLAG_RGB_ROWS_AS_WRITTEN = 0x800 in the flags to keep that).
  if pointer(ctx).isNil:
    doAssert tor_ctx_create(nil, 0, addr ctx) == 0, $tor_last_error(TorCtx(nil))
  var flat = world.toFlatList()
  var camCopy = cam
  let rc = tor_render_ycbcr420(ctx, addr canvas, addr camCopy, addr flat[0], int64 flat.len,
                               TOR_STRIDE_FLAT, int64 max_depth, torFlags, frame)
  doAssert rc == 0, $tor_last_error(ctx)

  # Alternative without the per-object copy: pass the raw variant array,
  #   tor_render(ctx, addr canvas, addr camCopy, world.objects, world.len, sizeof(HittableVariant), ...)
  # The library accepts it only when sizeof(HittableVariant) == 120 (the layout derived in
  # include/tor_b200.h); any other stride is rejected with TOR_ERR_LAYOUT.
