#!/usr/bin/env python
"""bench.py — Mray/s of the render path on BASELINE.json's workload (config C2: book-1 random_scene,
1200x675, 500 spp, depth 50), and with --workload on the other configurations (c1, c4 animation, c5 stress).

  python bench.py --gpus N --steps K --warmup W            our sm_100a path (N>1: under torchrun, NCCL)
  python bench.py --impl reference --gpus N ...             the reference's CPU algorithm on the host cores

One "step" = one full render of the workload (rows interleaved over the N ranks + one gather to rank 0; for the
animation c4: all 300 frames, frame f on rank f mod N, + one gather of the RGB8 frames).
`value`   : primary rays / CUDA-event time, scene already resident in HBM, image left in rank 0's HBM.
`e2e`     : the same metric through the public host-buffer call (tor_render for N=1; DistributedRenderer.render
            for N>1): scene packing + H2D, kernel, gather, D2H into the caller's canvas, every step.
`roofline`: the binding roof of this path is the FP64 pipe: FP64 arithmetic instructions the step
            executes (per-opcode counts of the committed ncu capture of the same launch, profiles/fp64_ops.json)
            / device time, against the DFMA issue rate measured on the same GPU in the same run.  The HBM roofline the
            metric's wording asks for (algorithmic bytes / time against MEASURED_PEAKS.json) is kept beside it under
            roofline.hbm — it is ~1e-5 by construction (a megakernel writes the framebuffer once).
`image_check`: sha256 of the image that arrived on rank 0 against the CPU oracle's digest of the full frame
            (tests/golden/c2_oracle_digest.json) — at every GPU count the same bits.
`schedule`: what the device-side pixel scheduling of rank 0's last launch did: pixels traced by
            a whole warp, pixels whose pre-pass cost qualified, pixels parked in the launch's tail and finished by the
            hand-off launch.  Scheduling cannot change the image.
`cpu_baseline`: the oracle (C++ restatement of the reference, glibc libm, OpenMP over all host cores) timed on
            a bounded row sample of the same workload.  The Nim/Weave binary cannot be built (no nim).
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nrows, ncols, spp, max_depth, scene half-grid)
    "c1": (216, 384, 100, 50, 11),     # trace_of_radiance.nim:27-32
    "c2": (675, 1200, 500, 50, 11),    # BASELINE.json configs[1] (and [2] at N = 8)
    "c5": (2160, 3840, 2000, 50, 50),  # BASELINE.json configs[4]: 10 002 spheres (random_scene grid -50 ..< 50)
}
ANIMATION = {"c4": dict(nrows=144, ncols=256, spp=100, depth=50, t_max=9.0, skip=6, frames=300)}  # configs[3]
GAMMA = 2.2
METRICS = {
    "c1": "Mray/s (primary rays) at 384x216 / 100 spp random_scene",
    "c2": "Mray/s (primary rays) at 1200x675 / 500 spp random_scene",
    "c5": "Mray/s (primary rays) at 3840x2160 / 2000 spp, 10 002 spheres",
    "c4": "Mray/s (primary rays) over 300 frames of scenes_animated at 256x144 / 100 spp",
}
GOLDEN = {"c1": ("c1_oracle_digest.json", lambda d: d["det"]["f64_sha256"]),
          "c2": ("c2_oracle_digest.json", lambda d: d["f64_sha256"])}


def golden_digest(workload):
    if workload not in GOLDEN:
        return None
    name, get = GOLDEN[workload]
    p = os.path.join(ROOT, "tests", "golden", name)
    return get(json.load(open(p))) if os.path.exists(p) else None


def fp64_ops(workload, mode):
    """FP64 arithmetic thread-instructions of one step on ONE GPU (render launches incl. the cost pre-pass), from the
    committed ncu per-opcode capture (profiles/fp64_ops.json, tools/ncu_fp64_ops.py); None if not captured."""
    p = os.path.join(ROOT, "profiles", "fp64_ops.json")
    if os.path.exists(p):
        return json.load(open(p)).get(f"{workload}:{mode}")
    return None


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on C2, from the committed ncu capture
    (profiles/ncu_traffic.json, written from an `ncu --set full` run of tools/sweep.py); None if not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel)
    return None


def ncu_pipe_summary(variant="bvh"):
    """FP64 pipe / issue / SIMD utilisation of the render kernel from the committed `ncu --set full` capture
    (profiles/r01k_ncu_<variant>_800x450x128.json: same scene and camera at 800x450 / 128 spp; variant "bvh" = exact
    mode with the row-major queue, "bvh_split" = split-stream mode); None if absent."""
    name = f"r01k_ncu_{variant}_800x450x128.json"
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    d = json.load(open(p))[0]

    def v(k):
        return d.get(k, {}).get("value")

    return {"fp64_pipe_active_pct": v("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": v("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "active_lanes_per_instruction": v("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "source": f"profiles/{name} (captured under ncu, not a timing)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------- CPU oracle legs
def cpu_sample_rows(nrows, n):
    """n rows spread evenly over the image (sky and ground rows in proportion)."""
    n = max(1, min(nrows, n))
    return sorted(set(int((i + 0.5) * nrows / n) for i in range(n)))


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the CPU arm
    is meant to measure, so the count is passed to the oracle explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_time_rows(O, wl, rows, cam, world):
    nrows, ncols, spp, depth, _ = wl
    img = np.zeros((nrows, ncols, 3))
    nt = host_threads()
    t = time.perf_counter()
    for r in rows:
        O.render(nrows, ncols, spp, cam, world, max_depth=depth, gamma=GAMMA, rows=(r, r + 1, 1), math="libm", out=img,
                 nthreads=nt)
    return time.perf_counter() - t


def cpu_baseline(wl, target_s):
    """Times the oracle on a bounded sample of the workload: whole rows, evenly spread, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    nrows, ncols, spp, depth, half = wl
    world = O.random_scene(0xFACADE, half)
    cam = O.book_camera(16.0 / 9.0)
    probe = cpu_sample_rows(nrows, 3)
    t_probe = oracle_time_rows(O, wl, probe, cam, world)
    n = int(max(3, min(nrows, target_s / (t_probe / len(probe)))))
    rows = cpu_sample_rows(nrows, n)
    t = oracle_time_rows(O, wl, rows, cam, world)
    rays = len(rows) * ncols * spp
    return {"value": rays / t / 1e6, "unit": "Mray/s", "cores": host_threads(), "kind": "port",
            "sample": f"{len(rows)} of {nrows} rows evenly spread ({rays / 1e6:.1f} M primary rays, {t:.1f} s), "
                      "C++/OpenMP restatement of render.nim with glibc libm; not the Nim/Weave binary",
            "seconds": t}


def run_reference(args, wl, rank):
    """--impl reference: the reference's CPU algorithm (oracle port) on this box's host cores.

    Same config as our arm: the FIRST timed step renders the whole frame (every row) when the probe says that fits
    the budget (--ref-full-seconds, default 240 s: C2 takes ~150-190 s on 16 threads); the remaining steps and the
    warm-up render an evenly spread row sample of a few seconds each, so that the driver's --steps K run still ends in
    minutes.  The value is all timed rays / all timed seconds."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    nrows, ncols, spp, depth, half = wl
    world = O.random_scene(0xFACADE, half)
    cam = O.book_camera(16.0 / 9.0)
    probe = cpu_sample_rows(nrows, 3)
    t_probe = oracle_time_rows(O, wl, probe, cam, world)
    per_row = t_probe / len(probe)
    est_full = per_row * nrows
    full_first = est_full <= args.ref_full_seconds
    per_step = max(2.0, min(15.0, 60.0 / max(1, args.steps + args.warmup)))
    n = int(max(1, min(nrows, per_step / per_row)))
    rows = cpu_sample_rows(nrows, n)
    for _ in range(args.warmup):
        oracle_time_rows(O, wl, rows, cam, world)
    times, rays = [], 0
    for k in range(args.steps):
        sel = list(range(nrows)) if (full_first and k == 0) else rows
        times.append(oracle_time_rows(O, wl, sel, cam, world))
        rays += len(sel) * ncols * spp
    total = sum(times)
    value = rays / total / 1e6
    if full_first:
        sample = (f"step 1 = the whole frame ({nrows} rows, {nrows * ncols * spp / 1e6:.0f} M primary rays, "
                  f"{times[0]:.1f} s); steps 2..{args.steps} = {len(rows)} of {nrows} rows evenly spread each; ")
    else:
        sample = (f"each step = {len(rows)} of {nrows} rows evenly spread ({len(rows) * ncols * spp / 1e6:.1f} M primary "
                  f"rays): a whole frame would take ~{est_full:.0f} s on these cores; ")
    sample += "C++/OpenMP restatement of render.nim (oracle/), glibc libm, all host threads"
    line = {
        "impl": "reference", "metric": METRICS[args.workload], "value": value, "unit": "Mray/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, wl, len(world)), "sample": sample,
                   "full_frame_in_step_1": full_first},
        "cpu_baseline": {"value": value, "unit": "Mray/s", "cores": host_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mray/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def workload_name(name, wl, n_obj):
    nrows, ncols, spp, depth, half = wl
    return (f"{name}: random_scene(seed 0xFACADE, grid -{half}..<{half}, {n_obj} objects) {ncols}x{nrows} / {spp} spp / "
            f"depth {depth}, gamma float32(2.2)")


# ---------------------------------------------------------------------------------- our arm
def run_ours(args, wl, rank, world_size, local_rank):
    import torch
    import torch.distributed as dist

    import trace_of_radiance_b200 as T
    from trace_of_radiance_b200 import distributed as D

    nrows, ncols, spp, depth, half = wl
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    # everything below runs on ONE explicit (non-default) stream: the kernel is launched on it through the C ABI,
    # the events that time it are recorded on it, and NCCL picks it up as torch's current stream
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        _run_ours_on_stream(args, wl, rank, world_size, local_rank, dev, torch, dist, T, D)
    if world_size > 1:
        dist.destroy_process_group()


def _run_ours_on_stream(args, wl, rank, world_size, local_rank, dev, torch, dist, T, D):
    nrows, ncols, spp, depth, half = wl
    ctx = T.Context([local_rank])
    scene = T.random_scene(0xFACADE, half).list()
    cam = T.camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16.0 / 9.0, 0.1, 10.0, 0.0, 1.0)
    R = D.DistributedRenderer(ctx, device=dev)
    R.upload(cam, scene)  # inputs resident in HBM before the timed region
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    route = T.api.TOR_FLAG_BRUTE_FORCE if args.route == "brute" else 0
    if args.mode == "fast":
        route |= T.api.TOR_MODE_FAST

    def step(flags=route):
        return R.render_device(nrows, ncols, spp, GAMMA, depth, flags=flags)

    def timed_steps(flags):
        """K steps, device time per step (CUDA events on the launching stream), L2 flushed between them.  Returns
        (sum of step times in ms, max over ranks; sum of render-kernel times in ms, max over ranks; wall seconds)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        kernel_ms = []
        barrier()
        t_wall = time.perf_counter()
        for a, b in ev:
            flush.zero_()
            a.record()
            step(flags)
            b.record()
            b.synchronize()
            kernel_ms.append(ctx.last_kernel_ms())
        barrier()
        t_wall = time.perf_counter() - t_wall
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev), sum(kernel_ms)], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), t_wall, statistics.mean(kernel_ms)

    for _ in range(max(3, args.warmup) if args.warmup >= 0 else 0):
        step()
    barrier()

    # ---- timed region
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count()
    dev_ms, kern_ms, t_wall, my_kernel_ms = timed_steps(route)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - launches0
    sched = ctx.last_schedule() if args.mode == "exact" and args.route == "bvh" else None
    if sched is not None:
        sched["parked_in_tail"] = ctx.last_handoffs()  # rank 0's launch: pixels finished by the hand-off launch
    if dev_ms < 0.98 * kern_ms:
        raise SystemExit(f"timing events ({dev_ms:.3f} ms) do not bracket the render kernel ({kern_ms:.3f} ms)")
    rays_per_step = nrows * ncols * spp
    value = rays_per_step * args.steps / (dev_ms * 1e-3) / 1e6

    # ---- image check: what arrived on rank 0 in the last timed step, bit for bit
    image_check = None
    slots = step()
    torch.cuda.synchronize(dev)
    if rank == 0:
        img = R.image(slots, nrows).cpu().numpy()
        digest = hashlib.sha256(img.tobytes()).hexdigest()
        want = golden_digest(args.workload) if args.mode == "exact" else None
        image_check = {"sha256": digest, "finite": bool(np.isfinite(img).all())}
        if want is not None:
            image_check["oracle_sha256"] = want
            image_check["result"] = "ok" if digest == want else "MISMATCH"
            image_check["against"] = f"tests/golden/{GOLDEN[args.workload][0]} (CPU oracle, full frame)"
        else:
            image_check["result"] = "no committed digest for this workload / mode"
        del img
    if world_size > 1:  # independent of any golden file: rank 0 renders a few rows alone and compares
        rows = sorted(set(int((i + 0.5) * nrows / 5) for i in range(5)))
        if rank == 0:
            mine = torch.zeros((len(rows), ncols, 3), dtype=torch.float64, device=dev)
            for k, r in enumerate(rows):
                ctx.render_device_async(mine[k].data_ptr(), nrows, ncols, spp, GAMMA, depth, route, rows=(r, r + 1, 1),
                                        stream=torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.synchronize(dev)
            full = R.image(slots, nrows)
            same = all(torch.equal(full[r], mine[k]) for k, r in enumerate(rows))
            image_check["gathered_rows_equal_single_rank_render"] = bool(same)
            image_check["rows_checked"] = rows
            if not same:
                image_check["result"] = "MISMATCH"
        barrier()

    # ---- e2e: host buffers through the public call, every step: pack + H2D + render (+ gather) + D2H
    pinned = torch.empty((nrows, ncols, 3), dtype=torch.float64, pin_memory=True)
    canvas = T.newCanvas(nrows, ncols, spp, GAMMA)
    canvas.pixels = pinned.numpy()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step(flags=route):
        if world_size == 1:
            ctx.render(canvas, cam, scene, depth, flags=flags)  # tor_render: the drop-in for render.nim:49
        else:
            R.render(canvas, cam, scene, depth, flags=flags)

    def timed_e2e(flags):
        e2e_step(flags)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step(flags)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return rays_per_step * e2e_steps / float(te[0]) / 1e6

    e2e_value = timed_e2e(route)
    if rank == 0 and image_check is not None and image_check.get("oracle_sha256"):
        image_check["e2e_canvas_matches"] = hashlib.sha256(canvas.pixels.tobytes()).hexdigest() == image_check["oracle_sha256"]
    h2d = int(scene.objects.nbytes) + 192  # what the caller hands over (every rank, for N > 1)
    d2h = nrows * ncols * 24               # rank 0 only

    # ---- rooflines of the render kernel (the only kernel of ours that matters in the step)
    hbm_peak, peak_src = peaks()
    my_rows = D.partition_rows(nrows, rank, world_size)[1]
    alg_bytes = my_rows * ncols * 24 + len(scene) * 112 + 192  # framebuffer write + scene + camera
    kern_s = my_kernel_ms * 1e-3
    hbm_achieved = alg_bytes / kern_s / 1e9
    # ALU side: work counters (instrumented, untimed) — segments are deterministic
    step(T.api.TOR_FLAG_COUNT_SEGMENTS | route)
    torch.cuda.synchronize(dev)
    cnt = ctx.counters()
    ct = torch.tensor([cnt["primary_rays"], cnt["segments"], cnt["bvh_node_visits"], cnt["bvh_sphere_tests"]],
                      dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(ct, op=dist.ReduceOp.SUM)
    segments = float(ct[1])
    fp64_peak = ctx.measure_fp64_peak() if rank == 0 else 0.0
    n_obj = len(scene)
    ops = fp64_ops(args.workload, args.mode if args.route == "bvh" else "brute")

    # ---- split-stream mode (TOR_MODE_FAST) beside the headline: same workload, same float64 arithmetic, the pixel's
    #      sample loop cut into RNG substreams (deterministic, bit-exact against the oracle's restatement, a different
    #      Monte-Carlo estimate than the reference's image) — reported, never substituted for `value`
    split = None
    if args.mode == "exact" and args.route == "bvh" and not args.no_split_stream:
        fl = route | T.api.TOR_MODE_FAST
        for _ in range(2):
            step(fl)
        f_dev_ms, _, _, f_kernel_ms = timed_steps(fl)
        fms = f_dev_ms / args.steps
        f_e2e = timed_e2e(fl)
        # algorithmic HBM bytes of the split-stream step: the exact mode's + one 24-byte partial sum per (pixel, range)
        # written by the render kernel and read by substream_reduce_kernel
        nsub = T.api.fast_substream_count(fl, nrows, ncols, spp)
        split_bytes = alg_bytes + 2 * 24 * my_rows * ncols * nsub
        split = {"value": rays_per_step / (fms * 1e-3) / 1e6, "unit": "Mray/s", "ms_per_step": fms,
                 "substreams_per_pixel": nsub,
                 "hbm": {"achieved": split_bytes / (f_kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": split_bytes / (f_kernel_ms * 1e-3) / 1e9 / hbm_peak,
                         "algorithmic_bytes_per_launch": split_bytes},
                 "e2e": f_e2e,
                 "flags": "TOR_MODE_FAST (automatic substream count: 2^24 / pixels, <= spp, <= 32)",
                 "parity": "bit-exact vs the oracle's render_split; vs the reference image: within 4*sqrt(2)*sigma/"
                           "sqrt(spp) per pixel (tests/test_split_stream.py)"}
        sops = fp64_ops(args.workload, "fast")
        if sops and rank == 0 and fp64_peak > 0:
            ach = sops["fp64_arith_thread_inst"] / world_size / (f_kernel_ms * 1e-3)
            split["roofline"] = {"bound": "fp64", "achieved": ach / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                                 "frac": ach / fp64_peak, "source": sops.get("source")}

    if rank == 0:
        base = cpu_baseline(wl, args.cpu_seconds) if (world_size == 1 and not args.no_cpu_baseline) else None
        roof = {"kernel": "render_bvh_kernel" if args.route == "bvh" else "render_exact_kernel",
                "kernel_ms": my_kernel_ms,
                "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
                        "traffic": (ncu_traffic("render_bvh_kernel" if args.route == "bvh" else "render_exact_kernel")
                                    if args.workload == "c2" and world_size == 1 else None),
                        "algorithmic_bytes_per_launch": alg_bytes,
                        "note": "megakernel: HBM sees the framebuffer write + one scene read per CTA (from L2)"},
                "work": {"segments_per_step": segments, "bvh_node_visits_per_step": float(ct[2]),
                         "sphere_tests_per_step": float(ct[3]), "route": args.route,
                         "full_scan_sphere_tests_avoided_per_step": max(0.0, segments * n_obj - float(ct[3]))}}
        if ops and fp64_peak > 0:
            # the capture is of ONE GPU rendering the whole frame; N ranks split the same instructions (the image and
            # every path in it are identical), up to the pre-pass's fixed 8 samples per pixel
            per_rank = ops["fp64_arith_thread_inst"] / world_size
            ach = per_rank / kern_s
            roof.update({"bound": "fp64", "achieved": ach / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                         "frac": ach / fp64_peak,
                         "definition": "FP64 arithmetic instructions executed per second, lane level (DADD + DMUL + DFMA, "
                                       "one operation each: the reference's arithmetic is non-fused, -fmad=false), "
                                       "against the DFMA issue rate measured on this GPU in this run "
                                       "(tor_measure_fp64_peak)",
                         "fp64_arith_thread_inst_per_step": ops["fp64_arith_thread_inst"], "per_opcode": ops.get("per_opcode"),
                         "source": ops.get("source"), "traffic": roof["hbm"]["traffic"]})
        else:
            roof.update({"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": hbm_achieved / hbm_peak, "traffic": roof["hbm"]["traffic"],
                         "note": "no FP64 instruction capture committed for this workload / mode (profiles/fp64_ops.json): "
                                 "only the HBM roofline can be stated, and it is not the binding one",
                         "measured_dfma_per_s": fp64_peak})
        line = {
            "metric": METRICS[args.workload], "value": value, "unit": "Mray/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, wl, n_obj) + ", " +
                                   ("exact mode (bit-identical image)" if args.mode == "exact" else
                                    "split-stream mode (TOR_MODE_FAST)"),
                       "partition": f"rows interleaved over {world_size} rank(s)" +
                                    (" + one gather to rank 0 (NCCL send/recv)" if world_size > 1 else ""),
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step event pairs)",
                       "wall_s_timed_region": t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mray/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "call": "tor_render (host canvas, pinned)" if world_size == 1 else
                    "DistributedRenderer.render (scene upload + kernel + gather to rank 0 + strided D2H on rank 0)"},
            "gpu_launches": launches,
            "image_check": image_check,
            "roofline": roof,
        }
        if sched:
            line["schedule"] = sched
        if split:
            line["split_stream_mode"] = split
        if base:
            line["cpu_baseline"] = base
        print(json.dumps(line), file=RESULT_OUT, flush=True)


RESULT_OUT = sys.stdout


def keep_stdout_for_the_result_line():
    """stdout must carry exactly one JSON line.  Libraries write to file descriptor 1 as well (NCCL prints its version
    banner there when the first communicator is created), so fd 1 is pointed at stderr for the rest of the process and
    the result line goes to a private copy of the original stdout."""
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    keep_stdout_for_the_result_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + sorted(ANIMATION))
    ap.add_argument("--route", default="bvh", choices=["bvh", "brute"],
                    help="closest-hit search: BVH in front of the reference's sphere test (default) or the full scan")
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"],
                    help="exact: the reference's per-pixel RNG stream (bit-identical image; the headline).  fast: "
                         "TOR_MODE_FAST split-stream mode as the measured arm")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-split-stream", action="store_true", help="skip the split_stream_mode leg")
    ap.add_argument("--in-flight", type=int, default=8, help="c4: animation frames enqueued at once per GPU")
    ap.add_argument("--ref-full-seconds", type=float, default=240.0,
                    help="--impl reference renders the whole frame in its first step when the probe predicts at most "
                         "this many seconds")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in ANIMATION:
        import bench_animation

        bench_animation.main(args, ANIMATION[args.workload], METRICS[args.workload], RESULT_OUT)
        return
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if world_size != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world_size})")
    run_ours(args, wl, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
