"""One process per GPU: interleaved row partition + ONE gather.

The reference parallelises render.nim:55-58 over pixels inside one process (Weave).  Here rank g renders
rows {r : r mod G == g} — cheap sky rows and expensive ground rows interleave — into a compact device
buffer, and a single all_gather over NCCL (NVLink/NVSwitch) assembles the framebuffer.  Seeds use the
absolute (row, col) (render.nim:59-60), so the image is bit-identical for every G.  There is no other
exchange step on this path, so no other collective.

The partition / un-interleave logic is backend-agnostic (gloo + CPU tensors in tests/test_distributed.py).
"""
import torch
import torch.distributed as dist


def partition_rows(nrows, rank, world):
    """(row_begin, row_end, row_step) of this rank and its row count."""
    rb, re, rs = rank, nrows, world
    n = (re - rb + rs - 1) // rs if re > rb else 0
    return (rb, re, rs), n


def rows_per_rank(nrows, world):
    """Padded row count of the gather slots (rank 0 always owns the most rows)."""
    return (nrows + world - 1) // world


def gather_rows(local, nrows, group=None, out=None):
    """local: (rows_per_rank(nrows, G), ncols, 3) float64 on every rank (the tail row is padding on ranks that
    own one row fewer).  Returns the (nrows, ncols, 3) framebuffer on every rank; row r = slot (r mod G, r div G).
    One collective."""
    world = dist.get_world_size(group)
    rpr = rows_per_rank(nrows, world)
    assert local.shape[0] == rpr
    gathered = out if out is not None else torch.empty((world,) + tuple(local.shape), dtype=local.dtype,
                                                       device=local.device)
    # the collective wants the slots concatenated along dim 0: (G*rpr, ncols, 3) is the same memory
    dist.all_gather_into_tensor(gathered.view((world * rpr,) + tuple(local.shape[1:])), local.contiguous(),
                                group=group)
    return uninterleave(gathered, nrows)


def uninterleave(gathered, nrows):
    """(G, rpr, ncols, 3) -> (nrows, ncols, 3): row r lives at [r mod G, r div G]."""
    world, rpr = gathered.shape[0], gathered.shape[1]
    full = gathered.transpose(0, 1).reshape((world * rpr,) + tuple(gathered.shape[2:]))
    return full[:nrows]


class DistributedRenderer:
    """`render()` for a torchrun job: every rank calls it with the same arguments; each returns the full image."""

    def __init__(self, ctx, device=None, group=None):
        self.ctx = ctx
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._local = None
        self._gathered = None

    def _buffers(self, nrows, ncols):
        rpr = rows_per_rank(nrows, self.world)
        shape = (rpr, ncols, 3)
        if self._local is None or tuple(self._local.shape) != shape:
            self._local = torch.zeros(shape, dtype=torch.float64, device=self.device)
            self._gathered = torch.empty((self.world,) + shape, dtype=torch.float64, device=self.device)
        return self._local, self._gathered

    def upload(self, cam, world_list):
        self.ctx.scene_upload(cam, world_list)

    def render_device(self, nrows, ncols, spp, gamma, max_depth, flags=0):
        """Scene already uploaded.  Enqueues kernel (+ gather) on torch's current stream; returns the device image."""
        local, gathered = self._buffers(nrows, ncols)
        rows, _ = partition_rows(nrows, self.rank, self.world)
        stream = torch.cuda.current_stream(self.device)
        self.ctx.render_device_async(local.data_ptr(), nrows, ncols, spp, gamma, max_depth, flags, rows=rows,
                                     stream=stream.cuda_stream)
        if self.world == 1:
            return local[:nrows]
        return gather_rows(local, nrows, group=self.group, out=gathered)

    def render(self, canvas, cam, world_list, max_depth, flags=0):
        """Host-to-host: upload the scene, render this rank's rows, gather, copy the image into canvas.pixels."""
        self.upload(cam, world_list)
        img = self.render_device(canvas.nrows, canvas.ncols, canvas.samples_per_pixel, canvas.gamma_correction,
                                 max_depth, flags)
        host = torch.from_numpy(canvas.pixels)
        host.copy_(img, non_blocking=False)
        return canvas


# ------------------------------------------------------------------------------------ animation
def frames_of_rank(nframes, rank, world):
    """Frame f goes to rank f mod G (frames are the independent unit of the animation)."""
    return list(range(rank, nframes, world))


def frames_per_rank(nframes, world):
    return (nframes + world - 1) // world


def gather_frames(local, nframes, group=None):
    """local: (frames_per_rank, h, w, 3) uint8 on every rank, frame f of this rank at slot f div G (the tail slot
    is padding on ranks that own one frame fewer).  Returns (nframes, h, w, 3) on every rank.  One collective."""
    world = dist.get_world_size(group)
    fpr = frames_per_rank(nframes, world)
    assert local.shape[0] == fpr
    gathered = torch.empty((world * fpr,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)
    return uninterleave(gathered.view((world, fpr) + tuple(local.shape[1:])), nframes)


def render_animation_distributed(make_animation, render_frame, nframes, frame_shape, skip=6, device=None, group=None):
    """The frame loop of trace_of_radiance_animation.nim:173-199 across ranks.  Every rank steps the (cheap,
    deterministic) animation itself and renders only its frames f = rank mod G; ONE all_gather assembles the
    RGB8 frames.  make_animation() -> object with .scenes(skip) (tor.Animation); render_frame(cam, world) -> (h, w, 3)
    uint8 array (e.g. ctx.render_rgb8).  Returns a (nframes, h, w, 3) uint8 tensor on every rank."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    fpr = frames_per_rank(nframes, world)
    local = torch.zeros((fpr,) + tuple(frame_shape), dtype=torch.uint8, device=device or "cpu")
    for f, (cam, scene) in enumerate(make_animation().scenes(skip=skip)):
        if f >= nframes:
            break
        if f % world == rank:
            local[f // world] = torch.as_tensor(render_frame(cam, scene)).to(local.device)
    if world == 1:
        return local[:nframes]
    return gather_frames(local, nframes, group=group)
