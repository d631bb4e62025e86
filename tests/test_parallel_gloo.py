"""Multi-rank host logic on CPU: gloo, world size 2 (and 3 for ragged bands).

Each rank renders its row band with the CPU oracle (standing in for its GPU), the bands are gathered with the
same code bench.py uses over NCCL, and the assembled frame must equal the single-rank frame.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, h, w, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as g
        pkg = g.load_package()
        from oracle import portbind as oracle
        par = pkg.parallel
        tris = pkg.cornell_box()
        fp = pkg.default_frame_params(0, w, h)
        y0, y1 = par.row_band(rank, world, h)
        band = oracle.rt_draw(tris, fp, w, h, y0, y1, threads=2)["pixelColours"]
        frame = torch.zeros((h, w, 3), dtype=torch.float32)
        frame[y0:y1] = torch.from_numpy(band[y0:y1])
        par.gather_bands(frame, rank, world)
        # frame partitioning: every frame index is owned by exactly one rank
        mine = torch.zeros(10, dtype=torch.int32)
        mine[par.frames_for_rank(rank, world, 10)] = 1
        dist.all_reduce(mine)
        if rank == 0:
            q.put((frame.numpy().copy(), mine.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 48), (3, 50)])
def test_row_band_gather(pkg, oracle, world, h):
    w = 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    frame, owners = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = oracle.rt_draw(pkg.cornell_box(), pkg.default_frame_params(0, w, h), w, h)["pixelColours"]
    assert np.array_equal(frame.view(np.uint32), want.view(np.uint32))
    assert (owners == 1).all()


def test_band_arithmetic(pkg):
    par = pkg.parallel
    for h in (0, 1, 7, 270, 2160, 2161):
        for world in (1, 2, 3, 8):
            bands = [par.row_band(r, world, h) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in bands) - min(b[1] - b[0] for b in bands) <= 1
    assert par.row_band(3, 8, 2160) == (810, 1080)
    with pytest.raises(ValueError):
        par.row_band(2, 2, 10)


def _shm_worker(rank, world, port, h, w, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as g
        pkg = g.load_package()
        from oracle import portbind as oracle
        par = pkg.parallel
        tris = pkg.cornell_box()
        fp = pkg.default_frame_params(0, w, h)
        shm = par.SharedHostFrame(w, h, rank, world, tag=str(port))
        for step in (1, 2):  # two frames: the progress words order them without a collective
            surf = oracle.resolve_surface(oracle.rt_draw(tris, fp, w, h, threads=2)["pixelColours"], None)
            for y0, y1 in par.tile_rows_for_part(rank, world, h):  # this rank's interleaved tile rows only
                shm.frame[y0:y1] = surf[y0:y1] + np.uint32(step - 1)
            shm.publish(step)
            if rank == 0:
                shm.wait_all(step)
                q.put(shm.frame.copy())
            dist.barrier()
        shm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 50), (3, 41)])
def test_shared_host_frame_interleaved_tile_rows(pkg, oracle, world, h):
    """Host side of the single-frame split (bench.py e2e at N > 1): every rank writes only its interleaved tile rows
    into ONE shared host frame; rank 0 sees the whole frame once every rank has published the step."""
    w = 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shm_worker, args=(r, world, port, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    frames = [q.get(timeout=120), q.get(timeout=120)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = oracle.resolve_surface(oracle.rt_draw(pkg.cornell_box(), pkg.default_frame_params(0, w, h), w, h)["pixelColours"], None)
    assert np.array_equal(frames[0], want)
    assert np.array_equal(frames[1], want + np.uint32(1))


def test_tile_row_partition(pkg):
    par = pkg.parallel
    for h in (1, 8, 9, 150, 2160, 2161):
        for n in (1, 2, 3, 8):
            rows = sorted(r for p in range(n) for r in par.tile_rows_for_part(p, n, h))
            assert rows[0][0] == 0 and rows[-1][1] == h
            assert all(rows[i][1] == rows[i + 1][0] for i in range(len(rows) - 1))
    assert par.tile_rows_for_part(1, 8, 2160)[:2] == [(8, 16), (72, 80)]
