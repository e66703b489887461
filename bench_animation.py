"""bench.py --workload c4: BASELINE.json configs[3] — scenes_animated.nim's bouncing spheres, 256x144 / 100 spp,
300 frames (t_max = 9.0, skip = 6), frame f on rank f mod N, one gather of the RGB8 frames to rank 0.

One step = the whole animation.  Our arm runs the device-resident pipeline (tor_animation_dev_*: physics, scene
rebuild and BVH re-fit as kernels, zero per-frame H2D); `value` = primary rays of all frames / device time
(animation streams + gather), `e2e` = the same through DeviceAnimation.render_all into pinned host frames on rank 0
(wall clock).  The reference arm times the oracle on a bounded sample of frames."""
import hashlib
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _reference(args, cfg, metric, out, rank):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from bench import host_threads

    h, w, spp, depth = cfg["nrows"], cfg["ncols"], cfg["spp"], cfg["depth"]
    an = O.Animation(height=h, width=w, t_max=cfg["t_max"])
    frames = []
    for f in range(cfg["frames"]):
        got = an.next_frame(skip=cfg["skip"])
        assert got is not None
        frames.append(got)
    assert an.next_frame(skip=cfg["skip"]) is None
    nt = host_threads()

    def render(f):
        t = time.perf_counter()
        O.render(h, w, spp, frames[f][0], frames[f][1], max_depth=depth, math="libm", nthreads=nt)
        return time.perf_counter() - t

    t_probe = render(0)
    per_step = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    n = int(max(1, min(cfg["frames"], per_step / t_probe)))
    sel = sorted(set(int((i + 0.5) * cfg["frames"] / n) for i in range(n)))
    for _ in range(args.warmup):
        for f in sel:
            render(f)
    times = [sum(render(f) for f in sel) for _ in range(args.steps)]
    rays = len(sel) * h * w * spp
    total = sum(times)
    value = rays * args.steps / total / 1e6
    sample = (f"each step = frames {sel} of {cfg['frames']} ({rays / 1e6:.1f} M primary rays); C++/OpenMP restatement of "
              "render.nim + scenes_animated.nim (oracle/), glibc libm, all host threads")
    line = {"impl": "reference", "metric": metric, "value": value, "unit": "Mray/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": _name(cfg), "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mray/s", "cores": nt, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mray/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)


def _name(cfg):
    return (f"c4: scenes_animated random_moving_spheres(seed 0xFACADE, 1 601 objects per frame) {cfg['ncols']}x{cfg['nrows']} / "
            f"{cfg['spp']} spp / depth {cfg['depth']}, {cfg['frames']} frames (dt 0.005, skip {cfg['skip']}, t_max {cfg['t_max']})")


def main(args, cfg, metric, out):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return _reference(args, cfg, metric, out, rank)
    if world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    import torch
    import torch.distributed as dist

    import trace_of_radiance_b200 as T
    from bench import ClockSampler, cpu_baseline  # noqa: F401

    h, w, spp, depth, nframes = cfg["nrows"], cfg["ncols"], cfg["spp"], cfg["depth"], cfg["frames"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(dev)
    ctx = T.Context([local_rank])
    fpr = (nframes + world - 1) // world
    mine = len(range(rank, nframes, world))
    flags = T.api.TOR_MODE_FAST if args.mode == "fast" else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    an = T.DeviceAnimation(ctx, height=h, width=w, t_max=cfg["t_max"], skip=cfg["skip"], in_flight=args.in_flight)
    local = torch.zeros((fpr, h, w, 3), dtype=torch.uint8, device=dev)

    def one_step(fl, to_host):
        """All frames of the animation; returns (device ms of the animation streams + gather, gathered device frames on
        rank 0, pinned host frames on rank 0 if asked, kernel launches)."""
        an.reset()  # first frame again: initial velocities and heights re-sent (two small arrays per ANIMATION)
        l0 = an.launch_count()
        n, ms = an.render_all(samples_per_pixel=spp, max_depth=depth, flags=fl, rank=rank, world=world,
                              out=_DevFrames(local))
        assert n == nframes, n
        launches = an.launch_count() - l0
        gathered = None
        with torch.cuda.stream(stream):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if world > 1:
                slots = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
                dist.gather(local, gather_list=slots, dst=0)
                if rank == 0:
                    gathered = torch.stack(slots, dim=1).reshape((fpr * world, h, w, 3))[:nframes]  # frame f = slot (f % G, f // G)
            else:
                gathered = local[:nframes]
            host = None
            if to_host and rank == 0:
                host = torch.empty((nframes, h, w, 3), dtype=torch.uint8, pin_memory=True)
                host.copy_(gathered, non_blocking=True)
            b.record()
            b.synchronize()
        return ms + a.elapsed_time(b), gathered, host, launches

    def timed(fl, steps):
        barrier()
        t_wall = time.perf_counter()
        ms, launches = [], 0
        for _ in range(steps):
            m, _, _, l = one_step(fl, to_host=False)
            ms.append(m)
            launches += l
        barrier()
        t_wall = time.perf_counter() - t_wall
        t = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), t_wall, launches

    for _ in range(max(3, args.warmup) if args.warmup >= 0 else 0):
        one_step(flags, to_host=False)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count()
    dev_ms, t_wall, launches = timed(flags, args.steps)
    clocks = sampler.stop() if sampler else None
    rays = nframes * h * w * spp
    value = rays * args.steps / (dev_ms * 1e-3) / 1e6

    # e2e: every step creates the animation, renders, gathers and lands the frames in pinned host memory on rank 0
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, gathered, host, _ = one_step(flags, to_host=True)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays * e2e_steps / float(te[0]) / 1e6

    split = None
    if args.mode == "exact" and not args.no_split_stream:
        fl = T.api.TOR_MODE_FAST
        one_step(fl, to_host=False)
        f_ms, _, _ = timed(fl, args.steps)
        split = {"value": rays * args.steps / (f_ms * 1e-3) / 1e6, "unit": "Mray/s", "ms_per_step": f_ms / args.steps,
                 "flags": "TOR_MODE_FAST", "parity": "bit-exact vs the oracle's render_split (tests/test_split_stream.py)"}

    if rank == 0:
        frames = host.numpy()
        digest = hashlib.sha256(frames.tobytes()).hexdigest()
        image_check = {"rgb8_sha256_all_frames": digest, "frames": int(frames.shape[0])}
        gold = os.path.join(ROOT, "tests", "golden", "c4_rgb8_digest.json")
        if args.mode == "exact" and os.path.exists(gold):
            want = json.load(open(gold))["rgb8_sha256_all_frames"]
            image_check["golden_sha256"] = want
            image_check["result"] = "ok" if want == digest else "MISMATCH"
            image_check["against"] = "tests/golden/c4_rgb8_digest.json (single-GPU run whose sampled frames equal the oracle's)"
        else:
            image_check["result"] = "no committed digest for this mode"
        line = {
            "metric": metric, "value": value, "unit": "Mray/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": _name(cfg) + (", exact mode" if args.mode == "exact" else ", split-stream mode"),
                       "partition": f"frame f on rank f mod {world}" + (" + one gather of the RGB8 frames to rank 0" if world > 1 else ""),
                       "pipeline": f"device-resident animation (physics + scene rebuild + BVH re-fit kernels, zero per-frame H2D), "
                                   f"{args.in_flight} frames in flight per GPU",
                       "l2": "every frame rewrites the scene blob and a different framebuffer; 300 frames x 16 kernels per step",
                       "wall_s_timed_region": t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mray/s", "h2d_bytes_per_step": int(2 * 8 * 1597),
                    "d2h_bytes_per_step": int(nframes * h * w * 3), "steps": e2e_steps,
                    "call": "DeviceAnimation.reset + render_all + gather + D2H of the RGB8 frames (pinned) on rank 0"},
            "gpu_launches": launches,
            "image_check": image_check,
            "roofline": {"bound": "fp64", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                         "note": "no FP64 instruction capture for this workload; see the c2 line for the kernel's roofline"},
        }
        if split:
            line["split_stream_mode"] = split
        print(json.dumps(line), file=out, flush=True)
    an.close()
    if world > 1:
        dist.destroy_process_group()


class _DevFrames:
    """Lets DeviceAnimation.render_all write frame k of this rank straight into a torch CUDA tensor (the C ABI accepts
    device memory for the RGB8 output)."""

    def __init__(self, t):
        self.t = t

    def __getitem__(self, k):
        return _Ptr(self.t[k].data_ptr())


class _Ptr:
    def __init__(self, p):
        class _C:
            data = p
        self.ctypes = _C()
