"""Several GPUs behind one Draw(): the b2r_group_* C ABI, the per-part host frame and the gather-to-root form of the
single-frame split (include/b2r.h).  Everything is compared with the CPU oracle.  With one GPU the group has one
member and the parts of a split are drawn one after the other on it; on a multi-GPU box (gpurun --gpus N) the same
tests run over 2..8 devices."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _group_sizes():
    return sorted({1, min(2, _ndev()), min(4, _ndev()), min(8, _ndev())})


def _oracle_surface(oracle, tris, fp, w, h):
    want = oracle.rt_draw(tris, fp, w, h)
    return oracle.resolve_surface(want["pixelColours"], None)


def test_group_rt_frame_matches_oracle(pkg, oracle):
    w, h = 200, 150  # 19 tile rows: not a multiple of any group size, the last tile row is short
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = 1, 2
    want = _oracle_surface(oracle, tris, fp, w, h)
    for n in _group_sizes():
        g = pkg.Group(w, h, list(range(n)))
        g.set_triangles(tris)
        g.set_frame(fp)
        got = g.rt_frame()
        assert np.array_equal(got, want), f"group of {n}"
        # a second frame into the same (now page-locked) surface, different camera
        pos, rot = pkg.orbit_camera(1, 8)
        fp2 = pkg.default_frame_params(0, w, h)
        fp2.set_camera(pos, rot, h / 2)
        g.set_frame(fp2)
        g.rt_frame(got)
        assert np.array_equal(got, _oracle_surface(oracle, tris, fp2, w, h)), f"group of {n}, second frame"
        g.close()


def test_group_ras_frame_matches_oracle(pkg, oracle):
    w, h = 160, 121
    tris = pkg.tessellate(pkg.cornell_box(), 4)
    fp = pkg.default_frame_params(1, w, h)
    culled = oracle.ras_cull(tris, fp, w, h)
    want = oracle.resolve_surface(oracle.ras_draw(tris, culled, fp, w, h)["pixelColours"], None)
    for n in _group_sizes():
        g = pkg.Group(w, h, list(range(n)))
        g.set_triangles(tris)
        g.set_frame(fp)
        for i in range(n):
            g.member_ras_cull(i)
        assert np.array_equal(g.ras_frame(), want), f"group of {n}"
        g.close()


def test_group_rt_frames_surfaces_and_bmp(pkg, oracle, tmp_path):
    """Animation frames handed out round-robin; both destinations: the caller's array and BMP files."""
    w, h, nframes = 96, 64, 7
    tris = pkg.cornell_box()
    frames = []
    for f in range(nframes):
        fp = pkg.default_frame_params(0, w, h)
        pos, rot = pkg.orbit_camera(f, nframes)
        fp.set_camera(pos, rot, h / 2)
        frames.append(fp)
    want = [_oracle_surface(oracle, tris, fp, w, h) for fp in frames]
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    for n in _group_sizes():
        g = pkg.Group(w, h, list(range(n)))
        g.set_triangles(tris)
        out = np.zeros((nframes, h, w), np.uint32)
        g.rt_frames(frames, surfaces=out)
        for f in range(nframes):
            assert np.array_equal(out[f], want[f]), (n, f)
        pattern = str(tmp_path / f"g{n}_%03d.bmp")
        g.rt_frames(frames, bmp_pattern=pattern)
        for f in range(nframes):
            ctx.set_frame(frames[f])
            ctx.rt_frame()
            ref = str(tmp_path / "ref.bmp")
            pkg.write_bmp(ref, ctx.resolve_bgr8(), w, h)
            assert open(pattern % f, "rb").read() == open(ref, "rb").read(), (n, f)
        g.close()
    ctx.close()


@pytest.mark.parametrize("nparts", [1, 3, 8])
def test_rt_frame_part_assembles_the_frame(pkg, oracle, nparts):
    """b2r_rt_frame_part: every part copies exactly its own tile rows into one host surface."""
    w, h = 200, 150
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    want = _oracle_surface(oracle, tris, fp, w, h)
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    surf = np.full((h, w), 0xDEADBEEF, np.uint32)
    for part in range(nparts):
        before = surf.copy()
        ctx.rt_frame_part(part, nparts, surf)
        rows = np.array([((y // 8) % nparts) == part for y in range(h)])
        assert np.array_equal(surf[~rows], before[~rows]), "a part wrote rows of another part"
    assert np.array_equal(surf, want)
    ctx.close()


@pytest.mark.parametrize("nparts", [2, 5])
def test_gather_to_root_with_arrival_counter(pkg, oracle, nparts):
    """b2r_rt_frame_gather_device_async + b2r_stream_wait_value32: all parts store into one surface; the word counts
    the launches whose pixels have landed and the consumer's stream waits for it on the GPU."""
    import torch
    w, h = 200, 150
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = 1, 3
    want = _oracle_surface(oracle, tris, fp, w, h)
    dev = torch.device("cuda:0")
    root = torch.full((h, w), -1, dtype=torch.int32, device=dev)
    arrive = torch.zeros(4, dtype=torch.int32, device=dev)
    out = torch.zeros((h, w), dtype=torch.int32, device=dev)
    producer = pkg.Context(w, h)
    consumer = pkg.Context(w, h)
    for c in (producer, consumer):
        c.set_triangles(tris)
        c.set_frame(fp)
    torch.cuda.synchronize()
    for frame in range(1, 3):
        # the consumer enqueues its wait + copy FIRST: it must not run before every part has arrived
        consumer.stream_wait_value32(arrive.data_ptr(), nparts * frame)
        consumer.copy_device_async(out.data_ptr(), root.data_ptr(), w * h * 4)
        for part in range(nparts):
            producer.rt_frame_gather_device_async(part, nparts, root.data_ptr(), arrive.data_ptr())
        consumer.synchronize()
        producer.synchronize()
        assert int(arrive[0].item()) == nparts * frame
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want), frame
        root.fill_(-1)
        torch.cuda.synchronize()
    producer.close()
    consumer.close()


def test_headless_binary_multi_gpu(pkg, tmp_path):
    """The reference's main() loop without SDL, over every GPU of the box: same BMP as on one GPU."""
    import subprocess
    from util import ROOT
    exe = os.path.join(ROOT, "cpp-raytracer-rasterizer_b200", "lib", "b2r_headless")
    if not os.path.exists(exe):
        pytest.skip("host binaries not built")
    outs = []
    for n in sorted({1, min(2, _ndev()), _ndev()}):
        prefix = str(tmp_path / f"rt_n{n}")
        r = subprocess.run([exe, "raytracer", "320", "200", "--aa", "2", "--out", prefix, "--gpus", str(n)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(open(prefix + ".bmp", "rb").read())
    assert all(o == outs[0] for o in outs)
