"""Worker for tests/test_multigpu.py (run under torch.distributed.run, one rank per GPU).

Single-frame split: every rank renders its row band of the same frame, resolves it with ONE kernel straight
into every rank's surface (peer-mapped buffers, stores over NVLink), barrier, and every rank compares the
assembled frame with a frame it rendered alone.  Also checks the one-kernel split (interleaved tile rows, trace
kernel stores to peers), the NCCL all-gather variant, the frame-partitioned orbit and the rasteriser's sort-first
row-band split."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = g.load_package()
    par = pkg.parallel
    w, h = 640, 360
    ctx = pkg.Context(w, h, device=local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_triangles(pkg.cornell_box())
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = 1, 2
    ctx.set_frame(fp)
    col = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    # reference: this rank renders the whole frame alone
    ref_surf = torch.zeros((h, w), dtype=torch.int32, device=dev)
    ctx.rt_draw_device_async(0, h, col.data_ptr())
    ctx.resolve_surface_device_async(0, h, col.data_ptr(), 0, ref_surf.data_ptr())
    ctx.synchronize()

    # --- fused exchange: shared surfaces, one per rank
    mine, handle = ctx.shared_alloc(w * h * 4)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    ptrs = [mine if r == rank else ctx.shared_open(handles[r]) for r in range(world)]
    y0, y1 = par.row_band(rank, world, h)
    col.zero_()
    ctx.rt_draw_device_async(y0, y1, col.data_ptr())
    ctx.resolve_surface_multi_device_async(y0, y1, col.data_ptr(), 0, ptrs)
    ctx.synchronize()
    dist.barrier()
    torch.cuda.synchronize(dev)
    # read the shared buffer back through torch: device-to-device copy from the raw pointer
    tmp = torch.zeros((h, w), dtype=torch.int32, device=dev)
    ctx.copy_device_async(tmp.data_ptr(), mine, w * h * 4)
    ctx.synchronize()
    assert torch.equal(tmp, ref_surf), f"rank {rank}: fused band exchange differs from the single-GPU frame"

    # --- exchange inside the trace kernel: interleaved tile rows, every pixel stored into every rank's surface
    dist.barrier()
    ctx.copy_device_async(mine, torch.full((h, w), -1, dtype=torch.int32, device=dev).data_ptr(), w * h * 4)
    ctx.synchronize()
    dist.barrier()
    ctx.rt_frame_split_device_async(rank, world, [mine] + [p for r, p in enumerate(ptrs) if r != rank])
    ctx.synchronize()
    dist.barrier()
    torch.cuda.synchronize(dev)
    ctx.copy_device_async(tmp.data_ptr(), mine, w * h * 4)
    ctx.synchronize()
    assert torch.equal(tmp, ref_surf), f"rank {rank}: split frame (trace kernel + peer stores) differs from the single-GPU frame"

    # --- gather to the root: every rank stores its tile rows into rank 0's surface only; the last thread block of
    # each launch bumps an arrival word in rank 0's memory and rank 0's stream waits for it on the GPU (no collective)
    from oracle import portbind as oracle  # checker only
    want = oracle.resolve_surface(oracle.rt_draw(pkg.cornell_box(), fp, w, h)["pixelColours"], None)
    assert np.array_equal(ref_surf.cpu().numpy().view(np.uint32), want), "single-GPU frame differs from the oracle"
    root_bytes = w * h * 4
    if rank == 0:
        gmine, ghandle = ctx.shared_alloc(root_bytes + 256)
        ctx.copy_device_async(gmine, torch.full((h * w + 64,), -1, dtype=torch.int32, device=dev).data_ptr(), root_bytes + 256)
        ctx.copy_device_async(gmine + root_bytes, torch.zeros(64, dtype=torch.int32, device=dev).data_ptr(), 256)
        ctx.synchronize()
    else:
        gmine, ghandle = 0, None
    box = [ghandle]
    dist.broadcast_object_list(box, src=0)
    groot = gmine if rank == 0 else ctx.shared_open(box[0])
    dist.barrier()
    for frame in range(1, 4):
        ctx.rt_frame_gather_device_async(rank, world, groot, groot + root_bytes)
        if rank == 0:
            ctx.stream_wait_value32(groot + root_bytes, world * frame)
            ctx.copy_device_async(tmp.data_ptr(), groot, root_bytes)
            ctx.synchronize()
            assert np.array_equal(tmp.cpu().numpy().view(np.uint32), want), f"gather to root differs from the oracle (frame {frame})"
        ctx.synchronize()
        dist.barrier()
    if rank != 0:
        ctx.shared_close(groot)
    dist.barrier()
    if rank == 0:
        ctx.shared_free(gmine)

    # --- host side of the split: every rank copies its own tile rows into ONE page-locked host frame (POSIX shm)
    shm = par.SharedHostFrame(w, h, rank, world, tag=os.environ.get("MASTER_PORT", "0"))
    ctx.pin_host_buffer(shm.frame)
    dist.barrier()
    ctx.rt_frame_part(rank, world, shm.frame)
    dist.barrier()
    assert np.array_equal(shm.frame, want), f"rank {rank}: shared host frame differs from the oracle"
    dist.barrier()
    ctx.unpin_host_buffer(shm.frame)
    shm.close()

    # --- NCCL variant (parallel.gather_bands)
    surf = torch.zeros((h, w), dtype=torch.int32, device=dev)
    ctx.resolve_surface_device_async(y0, y1, col.data_ptr(), 0, surf.data_ptr())
    ctx.synchronize()
    par.gather_bands(surf, rank, world)
    torch.cuda.synchronize(dev)
    assert torch.equal(surf, ref_surf), f"rank {rank}: NCCL band gather differs"

    # --- frame partition of an orbit: no collective on the data path; checksum of checksums over all frames
    nframes = 8
    sums = torch.zeros(nframes, dtype=torch.int64, device=dev)
    for fidx in par.frames_for_rank(rank, world, nframes):
        pos, rot = pkg.orbit_camera(fidx, nframes)
        fp.set_camera(pos, rot, h / 2)
        ctx.set_frame(fp)
        ctx.rt_draw_device_async(0, h, col.data_ptr())
        ctx.resolve_surface_device_async(0, h, col.data_ptr(), 0, surf.data_ptr())
        ctx.synchronize()
        sums[fidx] = surf.to(torch.int64).sum()
    dist.all_reduce(sums)
    if rank == 0:
        alone = []
        for fidx in range(nframes):
            pos, rot = pkg.orbit_camera(fidx, nframes)
            fp.set_camera(pos, rot, h / 2)
            ctx.set_frame(fp)
            ctx.rt_draw_device_async(0, h, col.data_ptr())
            ctx.resolve_surface_device_async(0, h, col.data_ptr(), 0, surf.data_ptr())
            ctx.synchronize()
            alone.append(int(surf.to(torch.int64).sum()))
        assert alone == sums.tolist(), (alone, sums.tolist())
    # --- rasteriser, sort-first: triangle list replicated, every rank rasterises and shades its row band only;
    # exchange by peer stores of the resolve kernel and, separately, by the NCCL band gather
    rctx = pkg.Context(w, h, device=local)
    rctx.set_stream(stream.cuda_stream)
    rtris = pkg.tessellate(pkg.cornell_box(), 6)  # 1,080 triangles
    rctx.set_triangles(rtris)
    rctx.set_frame(pkg.default_frame_params(1, w, h))
    rctx.ras_cull()
    ras_ref = torch.zeros((h, w), dtype=torch.int32, device=dev)
    rctx.ras_frame_device_async(0, h, ras_ref.data_ptr())
    rctx.synchronize()
    dist.barrier()
    rctx.ras_draw_device_async(y0, y1, 0, col.data_ptr(), 0, 0)
    rctx.resolve_surface_multi_device_async(y0, y1, col.data_ptr(), 0, ptrs)
    rctx.synchronize()
    dist.barrier()
    torch.cuda.synchronize(dev)
    ctx.copy_device_async(tmp.data_ptr(), mine, w * h * 4)
    ctx.synchronize()
    assert torch.equal(tmp, ras_ref), f"rank {rank}: rasteriser band split (peer stores) differs from the single-GPU frame"
    surf.zero_()
    rctx.ras_frame_device_async(y0, y1, surf.data_ptr())
    rctx.synchronize()
    par.gather_bands(surf, rank, world)
    torch.cuda.synchronize(dev)
    assert torch.equal(surf, ras_ref), f"rank {rank}: rasteriser band split (NCCL gather) differs"
    rctx.close()

    dist.barrier()
    for r in range(world):
        if r != rank:
            ctx.shared_close(ptrs[r])
    dist.barrier()
    ctx.shared_free(mine)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU_OK world=%d" % world)


if __name__ == "__main__":
    main()
