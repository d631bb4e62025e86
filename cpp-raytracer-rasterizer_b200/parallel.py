"""Row-band / frame partitioning across ranks (one process per GPU, torch.distributed for the plumbing).

The render paths shard without any data-path exchange: the scene is replicated and pixels
are independent.  Two partitionings:
  * frames: rank r renders frames r, r+world, ... of an animation -- no collective at all;
  * row bands: one frame, rank r renders rows [y0,y1) into the full-frame offsets of its buffer, then one
    in-place all-gather assembles the frame on every rank (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def row_band(rank, world, height):
    """Contiguous rows [y0,y1) of `height` for `rank`; the first height % world ranks get one extra row."""
    if not (0 <= rank < world) or height < 0:
        raise ValueError("bad rank/world/height")
    base, extra = divmod(height, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


def frames_for_rank(rank, world, nframes):
    """Frame indices rendered by `rank` when an animation is partitioned round-robin."""
    return list(range(rank, nframes, world))


def gather_bands(frame, rank, world, group=None):
    """All-gather the row bands of a full-frame tensor [H, W, ...] in place.

    Every rank has filled only its own band (row_band(rank, world, H)).  Equal bands use one
    all_gather_into_tensor on views of the same storage (no staging copy); ragged bands fall back to
    one broadcast per band.
    """
    h = frame.shape[0]
    bands = [row_band(r, world, h) for r in range(world)]
    sizes = {b[1] - b[0] for b in bands}
    flat = frame.view(h, -1)
    if len(sizes) == 1 and flat.is_contiguous():
        y0, y1 = bands[rank]
        dist.all_gather_into_tensor(flat.view(-1), flat[y0:y1].reshape(-1), group=group)
    else:
        for r, (y0, y1) in enumerate(bands):
            if y1 > y0:
                dist.broadcast(flat[y0:y1], src=r, group=group)
    return frame
