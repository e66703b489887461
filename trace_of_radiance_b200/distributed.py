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


def gather_rows_to(local, nrows, dst=0, group=None, out=None):
    """The same exchange with ONE receiver: rank `dst` gets the slots tensor (G, rows_per_rank, ncols, 3) — slot g holds
    rows g, g + G, ... compactly — and every other rank gets None.  Nothing is replicated and nothing is
    un-interleaved on the device: scatter_slots_to_canvas / DistributedRenderer.render land the slots in the caller's
    canvas with one strided copy each.  One collective (NCCL: grouped send / recv)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rpr = rows_per_rank(nrows, world)
    assert local.shape[0] == rpr
    if rank == dst:
        slots = out if out is not None else torch.empty((world,) + tuple(local.shape), dtype=local.dtype,
                                                        device=local.device)
        dist.gather(local.contiguous(), gather_list=list(slots.unbind(0)), dst=dst, group=group)
        return slots
    dist.gather(local.contiguous(), gather_list=None, dst=dst, group=group)
    return None


def scatter_slots_to_canvas(slots, pixels):
    """Host-side un-interleave for CPU tensors / arrays: slot g's rows go to canvas rows g, g + G, ..."""
    world = slots.shape[0]
    nrows = pixels.shape[0]
    for g in range(world):
        n = (nrows - g + world - 1) // world if nrows > g else 0
        pixels[g::world] = slots[g][:n]
    return pixels


def uninterleave(gathered, nrows):
    """(G, rpr, ncols, 3) -> (nrows, ncols, 3): row r lives at [r mod G, r div G]."""
    world, rpr = gathered.shape[0], gathered.shape[1]
    full = gathered.transpose(0, 1).reshape((world * rpr,) + tuple(gathered.shape[2:]))
    return full[:nrows]


class DistributedRenderer:
    """`render()` for a torchrun job: every rank calls it with the same arguments; rank 0 ends up with the image."""

    def __init__(self, ctx, device=None, group=None):
        self.ctx = ctx
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._local = None
        self._slots = None

    def _buffers(self, nrows, ncols):
        rpr = rows_per_rank(nrows, self.world)
        shape = (rpr, ncols, 3)
        if self._local is None or tuple(self._local.shape) != shape:
            self._local = torch.zeros(shape, dtype=torch.float64, device=self.device)
            self._slots = (torch.empty((self.world,) + shape, dtype=torch.float64, device=self.device)
                           if self.rank == 0 and self.world > 1 else None)
        return self._local, self._slots

    def upload(self, cam, world_list):
        self.ctx.scene_upload(cam, world_list)

    def render_device(self, nrows, ncols, spp, gamma, max_depth, flags=0):
        """Scene already uploaded.  Enqueues kernel (+ the gather to rank 0) on torch's current stream.  Returns the
        slots tensor (G, rows_per_rank, ncols, 3) on rank 0 — for G == 1 that is the image itself with a leading
        axis of 1 — and None on the other ranks."""
        local, slots = self._buffers(nrows, ncols)
        rows, _ = partition_rows(nrows, self.rank, self.world)
        stream = torch.cuda.current_stream(self.device)
        self.ctx.render_device_async(local.data_ptr(), nrows, ncols, spp, gamma, max_depth, flags, rows=rows,
                                     stream=stream.cuda_stream)
        if self.world == 1:
            return local[:nrows].unsqueeze(0)
        return gather_rows_to(local, nrows, dst=0, group=self.group, out=slots)

    def image(self, slots, nrows):
        """The (nrows, ncols, 3) device image from render_device's slots (a copy when G > 1; for checks, not for the
        timed path)."""
        return uninterleave(slots, nrows).contiguous()

    def render(self, canvas, cam, world_list, max_depth, flags=0):
        """Host-to-host: upload the scene, render this rank's rows, gather to rank 0, and on rank 0 copy slot g
        straight into canvas rows g, g + G, ... (one strided device-to-host copy per slot).  canvas.pixels is only
        written on rank 0."""
        self.upload(cam, world_list)
        nrows, ncols = canvas.nrows, canvas.ncols
        slots = self.render_device(nrows, ncols, canvas.samples_per_pixel, canvas.gamma_correction, max_depth, flags)
        stream = torch.cuda.current_stream(self.device)
        if slots is not None:
            for g in range(self.world):
                n = (nrows - g + self.world - 1) // self.world if nrows > g else 0
                self.ctx.download_rows_async(slots[g].data_ptr(), canvas.pixels, ncols, g, self.world, n,
                                             stream=stream.cuda_stream)
        stream.synchronize()
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
