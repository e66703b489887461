#!/usr/bin/env python
"""bench.py — Mray/s of the render path on BASELINE.json's workload (config C2: book-1 random_scene,
1200x675, 500 spp, depth 50).

  python bench.py --gpus N --steps K --warmup W            our sm_100a path (N>1: under torchrun, NCCL)
  python bench.py --impl reference --gpus N ...             the reference's CPU algorithm on the host cores

One "step" = one full render of the workload (rows interleaved over the N ranks + one all_gather).
`value`   : primary rays / CUDA-event time, scene already resident in HBM, image left in HBM.
`e2e`     : the same metric through the public host-buffer call (tor_render for N=1; DistributedRenderer.render
            for N>1): scene packing + H2D, kernel, gather, D2H into the caller's canvas, every step.
`roofline`: algorithmic HBM bytes of the render kernel / its CUDA-event time against the measured copy peak
            (MEASURED_PEAKS.json) — tiny by construction: the kernel is FP64-issue bound, so the
            binding roof is reported beside it under roofline.fp64.
`cpu_baseline`: the oracle (C++ restatement of the reference, glibc libm, OpenMP over all host cores) timed on
            a bounded row sample of the same workload.  The Nim/Weave binary cannot be built (no nim).
"""
import argparse
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
    "c1": (216, 384, 100, 50, 11),   # trace_of_radiance.nim:27-32
    "c2": (675, 1200, 500, 50, 11),  # BASELINE.json configs[1]
}
GAMMA = 2.2
METRIC = "Mray/s (primary rays) at 1200x675 / 500 spp random_scene"


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
    """--impl reference: the reference's CPU algorithm (oracle port) on this box's host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    nrows, ncols, spp, depth, half = wl
    world = O.random_scene(0xFACADE, half)
    cam = O.book_camera(16.0 / 9.0)
    per_step = max(2.0, min(15.0, 150.0 / max(1, args.steps + args.warmup)))
    probe = cpu_sample_rows(nrows, 3)
    t_probe = oracle_time_rows(O, wl, probe, cam, world)
    n = int(max(1, min(nrows, per_step / (t_probe / len(probe)))))
    rows = cpu_sample_rows(nrows, n)
    for _ in range(args.warmup):
        oracle_time_rows(O, wl, rows, cam, world)
    times = [oracle_time_rows(O, wl, rows, cam, world) for _ in range(args.steps)]
    rays = len(rows) * ncols * spp
    total = sum(times)
    value = rays * args.steps / total / 1e6
    sample = (f"each step = {len(rows)} of {nrows} rows evenly spread ({rays / 1e6:.1f} M primary rays) of the same "
              "workload; C++/OpenMP restatement of render.nim (oracle/), glibc libm, all host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mray/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: random_scene(seed 0xFACADE) {ncols}x{nrows} / {spp} spp / depth {depth}",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mray/s", "cores": host_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mray/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


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

    def step():
        return R.render_device(nrows, ncols, spp, GAMMA, depth, flags=route)

    for _ in range(max(3, args.warmup) if args.warmup >= 0 else 0):
        step()
    barrier()

    # ---- timed region: K steps, device time per step (CUDA events on the launching stream), L2 flushed between
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    t_wall = time.perf_counter()
    for a, b in ev:
        flush.zero_()
        a.record()
        step()
        b.record()
        b.synchronize()
        kernel_ms.append(ctx.last_kernel_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, sum(kernel_ms)], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, kern_ms = float(t[0]), float(t[1])
    if dev_ms < 0.98 * kern_ms:
        raise SystemExit(f"timing events ({dev_ms:.3f} ms) do not bracket the render kernel ({kern_ms:.3f} ms)")
    rays_per_step = nrows * ncols * spp
    value = rays_per_step * args.steps / (dev_ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the public call, every step: pack + H2D + render (+ gather) + D2H
    pinned = torch.empty((nrows, ncols, 3), dtype=torch.float64, pin_memory=True)
    canvas = T.newCanvas(nrows, ncols, spp, GAMMA)
    canvas.pixels = pinned.numpy()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        if world_size == 1:
            ctx.render(canvas, cam, scene, depth, flags=route)  # tor_render: the drop-in for render.nim:49
        else:
            R.render(canvas, cam, scene, depth, flags=route)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays_per_step * e2e_steps / float(te[0]) / 1e6
    h2d = int(scene.objects.nbytes) + 192  # what the caller hands over; the packed blob on the wire is reported below
    d2h = nrows * ncols * 24

    # ---- roofline of the render kernel (the only kernel of ours in the step)
    hbm_peak, peak_src = peaks()
    my_rows = D.partition_rows(nrows, rank, world_size)[1]
    alg_bytes = my_rows * ncols * 24 + len(scene) * 112 + 192  # framebuffer write + scene + camera
    kern_s = (kernel_ms and statistics.mean(kernel_ms) or 0.0) * 1e-3
    achieved = alg_bytes / kern_s / 1e9
    # ALU side: counted once (instrumented, untimed) — segments are deterministic
    R.render_device(nrows, ncols, spp, GAMMA, depth, flags=T.api.TOR_FLAG_COUNT_SEGMENTS | route)
    torch.cuda.synchronize(dev)
    cnt = ctx.counters()
    ct = torch.tensor([cnt["primary_rays"], cnt["segments"]], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(ct, op=dist.ReduceOp.SUM)
    segments = float(ct[1])
    fp64_peak = ctx.measure_fp64_peak() if rank == 0 else 0.0
    n_obj = len(scene)
    tests_per_s = segments * n_obj / (dev_ms / args.steps * 1e-3)

    # ---- split-stream mode (TOR_MODE_FAST) beside the headline: same workload, same float64 arithmetic, the pixel's
    #      sample loop cut into RNG substreams (deterministic, bit-exact against the oracle's restatement, a different
    #      Monte-Carlo estimate than the reference's image) — reported, never substituted for `value`
    split = None
    if args.mode == "exact" and args.route == "bvh":
        fl = route | T.api.TOR_MODE_FAST
        for _ in range(2):
            R.render_device(nrows, ncols, spp, GAMMA, depth, flags=fl)
        barrier()
        fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in fev:
            flush.zero_()
            a.record()
            R.render_device(nrows, ncols, spp, GAMMA, depth, flags=fl)
            b.record()
            b.synchronize()
        barrier()
        ft = torch.tensor([sum(a.elapsed_time(b) for a, b in fev)], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        fms = float(ft[0]) / args.steps
        # the same through the host-buffer call
        def fast_e2e_step():
            if world_size == 1:
                ctx.render(canvas, cam, scene, depth, flags=fl)
            else:
                R.render(canvas, cam, scene, depth, flags=fl)

        fast_e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fast_e2e_step()
        barrier()
        tf = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        # algorithmic HBM bytes of the split-stream step: the exact mode's + one 24-byte partial sum per (pixel, range)
        # written by the render kernel and read by substream_reduce_kernel
        nsub = T.api.fast_substream_count(fl, nrows, ncols, spp)
        split_bytes = alg_bytes + 2 * 24 * my_rows * ncols * nsub
        split = {"value": rays_per_step / (fms * 1e-3) / 1e6, "unit": "Mray/s", "ms_per_step": fms,
                 "substreams_per_pixel": nsub,
                 "roofline": {"bound": "hbm", "achieved": split_bytes / (fms * 1e-3) / 1e9, "peak": hbm_peak,
                              "unit": "GB/s", "frac": split_bytes / (fms * 1e-3) / 1e9 / hbm_peak,
                              "algorithmic_bytes_per_step": split_bytes},
                 "e2e": rays_per_step * e2e_steps / float(tf[0]) / 1e6, "ncu": ncu_pipe_summary("bvh_split"),
                 "flags": "TOR_MODE_FAST (automatic substream count: 2^24 / pixels, <= spp, <= 32)",
                 "parity": "bit-exact vs the oracle's render_split; vs the reference image: within 4*sqrt(2)*sigma/"
                           "sqrt(spp) per pixel (tests/test_split_stream.py)"}

    if rank == 0:
        base = cpu_baseline(wl, args.cpu_seconds) if (world_size == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": METRIC, "value": value, "unit": "Mray/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: random_scene(seed 0xFACADE, {n_obj} objects) {ncols}x{nrows} / "
                                   f"{spp} spp / depth {depth}, gamma float32(2.2), " +
                                   ("exact mode (bit-identical image)" if args.mode == "exact" else
                                    "split-stream mode (TOR_MODE_FAST)"),
                       "partition": f"rows interleaved over {world_size} rank(s) + one all_gather",
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step event pairs)",
                       "wall_s_timed_region": t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mray/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "call": "tor_render (host canvas, pinned)" if world_size == 1 else
                    "DistributedRenderer.render (scene upload + kernel + all_gather + D2H)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak,
                         "traffic": (ncu_traffic("render_bvh_kernel" if args.route == "bvh" else "render_exact_kernel")
                                     if args.workload == "c2" and world_size == 1 else None),
                         "peak_source": peak_src,
                         "kernel": "render_bvh_kernel" if args.route == "bvh" else "render_exact_kernel", "kernel_ms": statistics.mean(kernel_ms),
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "megakernel: HBM sees only the framebuffer write + one scene read; the binding roof is "
                                 "FP64 issue, see fp64",
                         "fp64": {"segments_per_step": segments, "sphere_tests_per_s": tests_per_s,
                                  "measured_dfma_per_s": fp64_peak, "route": args.route,
                                  "bvh_node_visits_per_step": cnt.get("bvh_node_visits"),
                                  "bvh_sphere_tests_per_step": cnt.get("bvh_sphere_tests"),
                                  "ncu": ncu_pipe_summary() if args.route == "bvh" else None,
                                  "reference_flop_per_test": 32.8,
                                  "reference_equivalent_flop_per_s": tests_per_s * 32.8}},
        }
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
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--route", default="bvh", choices=["bvh", "brute"],
                    help="closest-hit search: BVH in front of the reference's sphere test (default) or the full scan")
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"],
                    help="exact: the reference's per-pixel RNG stream (bit-identical image; the headline).  fast: "
                         "TOR_MODE_FAST split-stream mode as the measured arm")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if world_size != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world_size})")
    run_ours(args, wl, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
