"""CalculateDOF + PutPixelSDL + BMP payload parity (CUDA vs oracle)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("dof", [0, 1])
def test_resolve_surface_and_bmp(pkg, oracle, which, dof, tmp_path):
    w, h = 160, 120
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(which, w, h)
    fp.dofEnabled = dof
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    if which == 0:
        out = ctx.rt_draw()
    else:
        ctx.ras_cull()
        out = ctx.ras_draw()
    surf = ctx.resolve_surface()
    want = oracle.resolve_surface(out["pixelColours"], out["focalDistances"], bool(dof), 8)
    assert np.array_equal(surf, want)
    assert not surf[0].any() and not surf[-1].any() and not surf[:, 0].any() and not surf[:, -1].any()
    bgr = ctx.resolve_bgr8()
    assert np.array_equal(bgr, oracle.surface_to_bgr8(want))
    frame = ctx.rt_frame() if which == 0 else ctx.ras_frame()
    assert np.array_equal(frame, want)
    # without depth of field the frame call is fused (no pixelColours kept): resolving again starts from its surface
    assert np.array_equal(ctx.resolve_bgr8(), bgr)
    assert np.array_equal(ctx.resolve_surface(), want)
    path = str(tmp_path / "frame.bmp")
    pkg.write_bmp(path, bgr, w, h)
    raw = open(path, "rb").read()
    assert raw[:2] == b"BM" and len(raw) == 54 + len(bgr) and raw[54:] == bgr.tobytes()
    assert int.from_bytes(raw[18:22], "little") == w and int.from_bytes(raw[22:26], "little") == h
    ctx.close()


@pytest.mark.parametrize("w", [33, 34, 35, 36, 37, 38, 39, 40, 1, 2, 3, 4, 5])
def test_bmp_row_padding(pkg, oracle, w):
    """Every residue of W modulo 4 (the conversion kernel packs 4 pixels into 3 words) and of the row length modulo 4
    (33 pixels: 99 bytes per row -> padded to 100)."""
    h = 17
    fp = pkg.default_frame_params(0, w, h)
    ctx = pkg.Context(w, h)
    ctx.set_triangles(pkg.cornell_box())
    ctx.set_frame(fp)
    out = ctx.rt_draw()
    bgr = ctx.resolve_bgr8()
    assert len(bgr) == ((3 * w + 3) // 4 * 4) * h
    assert np.array_equal(bgr, oracle.surface_to_bgr8(oracle.resolve_surface(out["pixelColours"], None)))
    ctx.close()


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("dof", [0, 1])
@pytest.mark.parametrize("aa", [0, 1])
def test_fused_frame_device_equals_draw_then_resolve(pkg, which, dof, aa):
    """b2r_{rt,ras}_frame_device_async (surface written by the trace / shade kernel itself when depth of field is
    off) == draw + b2r_resolve_surface_device_async, also band by band and with the optional outputs left out."""
    import torch
    w, h = 200, 136
    dev = torch.device("cuda:0")
    fp = pkg.default_frame_params(which, w, h)
    fp.dofEnabled = dof
    if which == 0:
        fp.aaEnabled, fp.aaSamples = aa, 3
    ctx = pkg.Context(w, h)
    ctx.set_triangles(pkg.cornell_box())
    ctx.set_frame(fp)
    if which == 1:
        ctx.ras_cull()
    col = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    foc = torch.zeros((h, w), dtype=torch.float32, device=dev)
    ref = torch.zeros((h, w), dtype=torch.int32, device=dev)
    if which == 0:
        ctx.rt_draw_device_async(0, h, col.data_ptr(), 0, foc.data_ptr())
    else:
        ctx.ras_draw_device_async(0, h, 0, col.data_ptr(), foc.data_ptr(), 0)
    ctx.resolve_surface_device_async(0, h, col.data_ptr(), foc.data_ptr(), ref.data_ptr())
    ctx.synchronize()
    col2, foc2 = torch.zeros_like(col), torch.zeros_like(foc)
    got = torch.full((h, w), -1, dtype=torch.int32, device=dev)
    frame = ctx.rt_frame_device_async if which == 0 else ctx.ras_frame_device_async
    if dof:   # needs the whole frame's colours before any row is resolved: one band
        bands = [(0, h)]
    else:
        bands = [(0, 40), (40, 41), (41, h)]
    for y0, y1 in bands:
        if which == 0:
            frame(y0, y1, got.data_ptr(), col2.data_ptr(), 0, foc2.data_ptr())
        else:
            frame(y0, y1, got.data_ptr(), 0, col2.data_ptr(), foc2.data_ptr(), 0)
    ctx.synchronize()
    assert torch.equal(got, ref)
    assert torch.equal(col2.view(torch.int32), col.view(torch.int32))
    if not dof:  # surface only
        got2 = torch.full((h, w), -1, dtype=torch.int32, device=dev)
        if which == 0:
            frame(0, h, got2.data_ptr())
        else:
            frame(0, h, got2.data_ptr())
        ctx.synchronize()
        assert torch.equal(got2, ref)
    else:
        with pytest.raises(pkg.B2RError):
            frame(0, h, got.data_ptr())
    ctx.close()


@pytest.mark.parametrize("nparts", [1, 3, 8])
def test_split_frame_interleaved_tile_rows(pkg, nparts):
    """b2r_rt_frame_split_device_async: part p draws tile rows p, p+nparts, ... and stores every pixel into all
    destination surfaces; after all parts ran, every surface equals the frame one launch draws (on several GPUs the
    extra destinations are peer-mapped, tests/test_multigpu.py)."""
    import torch
    w, h = 200, 141  # ragged last tile row
    dev = torch.device("cuda:0")
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = 1, 2
    ctx = pkg.Context(w, h)
    ctx.set_triangles(pkg.cornell_box())
    ctx.set_frame(fp)
    ref = torch.zeros((h, w), dtype=torch.int32, device=dev)
    col_ref = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    ctx.rt_frame_device_async(0, h, ref.data_ptr(), col_ref.data_ptr())
    a = torch.full((h, w), -1, dtype=torch.int32, device=dev)
    b = torch.full((h, w), -1, dtype=torch.int32, device=dev)
    col = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    for part in range(nparts):
        ctx.rt_frame_split_device_async(part, nparts, [a.data_ptr(), b.data_ptr()], col.data_ptr())
    ctx.synchronize()
    assert torch.equal(a, ref) and torch.equal(b, ref)
    assert torch.equal(col.view(torch.int32), col_ref.view(torch.int32))
    fp.dofEnabled = 1
    ctx.set_frame(fp)
    with pytest.raises(pkg.B2RError):
        ctx.rt_frame_split_device_async(0, 2, [a.data_ptr()])
    ctx.close()


@pytest.mark.parametrize("size", [(200, 141), (131, 9), (640, 360)])
def test_dof_tiled_kernel_equals_generic(pkg, oracle, size):
    """Depth of field, 8x8 window: the shared-memory tiled kernel (default) against the generic one-thread-per-pixel
    kernel (B2R_OPT_DOF_VARIANT=1) and the oracle, at sizes with ragged tiles; band by band too (the window reads
    rows outside the band)."""
    w, h = size
    fp = pkg.default_frame_params(0, w, h)
    fp.dofEnabled = 1
    ctx = pkg.Context(w, h)
    ctx.set_triangles(pkg.cornell_box())
    ctx.set_frame(fp)
    out = ctx.rt_draw()
    want = oracle.resolve_surface(out["pixelColours"], out["focalDistances"], True, 8)
    tiled = ctx.resolve_surface()
    ctx.set_option(pkg.capi.OPT_DOF_VARIANT, 1)
    generic = ctx.resolve_surface()
    ctx.set_option(pkg.capi.OPT_DOF_VARIANT, 0)
    assert np.array_equal(generic, want)
    assert np.array_equal(tiled, want)
    import torch
    dev = torch.device("cuda:0")
    col = torch.from_numpy(out["pixelColours"]).to(dev)
    foc = torch.from_numpy(out["focalDistances"]).to(dev)
    surf = torch.full((h, w), -1, dtype=torch.int32, device=dev)
    cuts = sorted({0, min(3, h), min(h // 2 + 1, h), h})
    for y0, y1 in zip(cuts[:-1], cuts[1:]):
        ctx.resolve_surface_device_async(y0, y1, col.data_ptr(), foc.data_ptr(), surf.data_ptr())
    ctx.synchronize()
    assert np.array_equal(surf.cpu().numpy().view(np.uint32), want)
    ctx.close()
