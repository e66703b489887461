"""Host-side logic of the N>1 path (row partition + one gather) with world_size 2 and 3 over gloo on CPU.
The per-rank rows come from the oracle here (no GPU); the -m gpu suite checks the CUDA rows themselves."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nrows, ncols, spp, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from trace_of_radiance_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene, cam = O.random_scene(), O.book_camera()
        rows, n = D.partition_rows(nrows, rank, world)
        full = O.render(nrows, ncols, spp, cam, scene, rows=rows, math="det", nthreads=2)
        rpr = D.rows_per_rank(nrows, world)
        local = torch.zeros((rpr, ncols, 3), dtype=torch.float64)
        local[:n] = torch.from_numpy(full[rows[0]:rows[1]:rows[2]])
        img = D.gather_rows(local, nrows)
        # the single-receiver form used by DistributedRenderer: rank 0 gets the slots, nobody else gets anything
        slots = D.gather_rows_to(local, nrows, dst=0)
        assert (slots is None) == (rank != 0)
        if rank == 0:
            canvas = np.zeros((nrows, ncols, 3))
            D.scatter_slots_to_canvas(slots.numpy(), canvas)
            assert canvas.tobytes() == img.numpy().tobytes()
        q.put((rank, img.numpy().tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nrows", [(2, 11), (3, 10), (2, 8)])
def test_gather_reassembles_the_single_process_image(world, nrows):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    ncols, spp = 16, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nrows, ncols, spp, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = O.render(nrows, ncols, spp, O.book_camera(), O.random_scene(), math="det").tobytes()
    for rank, b in got:
        assert b == want, f"rank {rank}"


def test_partition_covers_every_row_once():
    from trace_of_radiance_b200 import distributed as D

    for nrows in (1, 7, 8, 675, 2160):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(nrows, dtype=int)
            for r in range(world):
                (rb, re, rs), n = D.partition_rows(nrows, r, world)
                sel = np.arange(rb, re, rs) if re > rb else np.arange(0)
                assert len(sel) == n <= D.rows_per_rank(nrows, world)
                seen[sel] += 1
            assert (seen == 1).all()


def test_uninterleave():
    from trace_of_radiance_b200 import distributed as D

    nrows, world = 7, 3
    rpr = D.rows_per_rank(nrows, world)
    g = torch.full((world, rpr, 2, 3), -1.0, dtype=torch.float64)
    for r in range(nrows):
        g[r % world, r // world] = r
    full = D.uninterleave(g, nrows)
    assert full.shape == (nrows, 2, 3)
    assert all((full[r] == r).all() for r in range(nrows))


def _anim_worker(rank, world, port, nframes, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import trace_of_radiance_b200 as T
    from trace_of_radiance_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def render_frame(cam, world_list):  # the oracle stands in for the GPU render on this CPU-only test
            img = O.render(9, 16, 2, cam.as_array(), world_list.objects, math="det", nthreads=2)
            return O.quantise_rgb8(img)

        frames = D.render_animation_distributed(lambda: T.Animation(height=9, width=16, t_max=9.0), render_frame,
                                                nframes, (9, 16, 3))
        q.put((rank, frames.numpy().tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nframes", [(2, 5), (3, 4)])
def test_animation_frames_are_dealt_to_ranks_and_gathered(world, nframes):
    """scenes_animated.nim frames f -> rank f mod G, one all_gather; equals the single-process frame sequence."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_anim_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    an = O.Animation(height=9, width=16, t_max=9.0)
    want = []
    for _ in range(nframes):
        cam, objs = an.next_frame(skip=6)
        want.append(O.quantise_rgb8(O.render(9, 16, 2, cam, objs, math="det", nthreads=2)))
    want = np.stack(want).tobytes()
    for rank, b in got:
        assert b == want, f"rank {rank}"
