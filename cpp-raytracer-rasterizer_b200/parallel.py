"""Row-band / frame partitioning across ranks (one process per GPU, torch.distributed for the plumbing).

The render paths shard without any data-path exchange: the scene is replicated and pixels
are independent.  Two partitionings:
  * frames: rank r renders frames r, r+world, ... of an animation -- no collective at all;
  * row bands: one frame, rank r renders rows [y0,y1) into the full-frame offsets of its buffer, then one
    in-place all-gather assembles the frame on every rank (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import mmap
import os

import numpy as np
import torch
import torch.distributed as dist


def row_band(rank, world, height):
    """Contiguous rows [y0,y1) of `height` for `rank`; the first height % world ranks get one extra row."""
    if not (0 <= rank < world) or height < 0:
        raise ValueError("bad rank/world/height")
    base, extra = divmod(height, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


def tile_rows_for_part(part, nparts, height, tile_rows=8):
    """Pixel rows of `height` drawn by `part` of a frame split `nparts` ways: tile rows (8 pixel rows) part,
    part + nparts, ... -- interleaved, so that every part carries the same mix of cheap and expensive rows
    (b2r_rt_frame_part / b2r_rt_frame_gather_device_async).  Returns a list of (y0, y1) row ranges."""
    if not (0 <= part < nparts) or height < 0:
        raise ValueError("bad part/nparts/height")
    ntiles = (height + tile_rows - 1) // tile_rows
    return [(t * tile_rows, min((t + 1) * tile_rows, height)) for t in range(part, ntiles, nparts)]


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


class SharedHostFrame:
    """One full-frame uint32 host surface shared by the ranks of a box (POSIX shared memory under /dev/shm).

    The host side of a single-frame split: every rank page-locks the mapping (Context.pin_host_buffer) and copies only
    its own tile rows into it (Context.rt_frame_part), each over its own PCIe link.  After the frame array come 64
    int64 progress words, one per rank, for a host-side "all parts delivered" check without a collective.
    Rank 0 creates the file; the others open it after a barrier (dist.barrier or any other rendezvous).
    """

    def __init__(self, width, height, rank, world, tag="0"):
        self.path = f"/dev/shm/b2r_frame_{tag}_{width}x{height}"
        self.rank, self.world = rank, world
        nbytes = width * height * 4 + 64 * 8
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        if world > 1 and dist.is_initialized():
            dist.barrier()
        self._f = open(self.path, "r+b")
        self._mm = mmap.mmap(self._f.fileno(), nbytes)
        self.frame = np.frombuffer(self._mm, np.uint32, width * height).reshape(height, width)
        self.progress = np.frombuffer(self._mm, np.int64, 64, offset=width * height * 4)

    def publish(self, step):
        """This rank's part of frame `step` is in the frame."""
        self.progress[self.rank] = step

    def wait_all(self, step):
        """Spin until every rank has published `step` (host memory, no collective)."""
        while int(self.progress[:self.world].min()) < step:
            pass

    def close(self):
        self.frame = self.progress = None
        try:
            self._mm.close()
        except BufferError:
            pass
        self._f.close()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass
