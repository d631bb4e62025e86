"""C-ABI checks that need no GPU: the library loads, exports every declared symbol, structs match, host helpers work."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from util import ROOT, fnv1a32


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    header = open(os.path.join(ROOT, "include", "b2r.h")).read()
    declared = sorted(set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"libb2r.so does not export {name}"
    assert sorted(pkg.capi.SYMBOLS) == declared, "capi.SYMBOLS out of sync with include/b2r.h"
    assert lib.b2r_abi_version() == 1


def test_struct_layouts_match_the_header(pkg, tmp_path):
    """sizeof/offsetof as the C compiler sees them == the ctypes mirror."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b2r.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(b2r_frame_params),sizeof(b2r_light),sizeof(b2r_intersection),offsetof(b2r_frame_params,randomPositions),"
                   "offsetof(b2r_frame_params,aaEnabled),offsetof(b2r_frame_params,currentReflectance));return 0;}\n")
    exe = tmp_path / "sz"
    # the header is plain C: no C++ (or torch) types in any signature
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    FP = pkg.FrameParams
    assert got == [C.sizeof(FP), C.sizeof(pkg.capi.Light), 20, FP.randomPositions.offset, FP.aaEnabled.offset,
                   FP.currentReflectance.offset]
    assert C.sizeof(pkg.capi.Light) == 28  # == reference Light (TestModel.h:35-45)


def test_cornell_box_fingerprint(pkg):
    t = pkg.cornell_box()
    assert t.shape == (30, 15)
    assert "%08x" % fnv1a32(t.tobytes()) == "b715a8a2"  # FNV-1a32 over the 30 x 60-byte Triangles
    # 64-byte flavour: same 60 bytes + isCulled = 0
    buf = np.zeros(30 * 64, np.uint8)
    assert pkg.load_library().b2r_scene_cornell_box(buf.ctypes.data_as(C.c_void_p), 30, 64) == 30
    rec = buf.reshape(30, 64)
    assert np.array_equal(rec[:, :60].copy().view(np.float32).reshape(30, 15), t) and not rec[:, 60:].any()
    assert pkg.load_library().b2r_scene_cornell_box(buf.ctypes.data_as(C.c_void_p), 29, 64) < 0


def test_tessellation(pkg):
    t = pkg.cornell_box()
    assert np.array_equal(pkg.tessellate(t, 1), t)
    for k in (2, 5):
        tt = pkg.tessellate(t, k)
        assert len(tt) == 30 * k * k
        # colours inherited, normals unit length, total area preserved
        assert np.array_equal(np.unique(tt[:, 12:15], axis=0), np.unique(t[:, 12:15], axis=0))
        assert np.allclose(np.linalg.norm(tt[:, 9:12], axis=1), 1, atol=1e-5)
        area = lambda a: 0.5 * np.linalg.norm(np.cross(a[:, 3:6] - a[:, 0:3], a[:, 6:9] - a[:, 0:3]), axis=1).sum()
        assert abs(area(tt) - area(t)) < 1e-3
    assert pkg.load_library().b2r_scene_tessellate(t.ctypes.data_as(C.c_void_p), 30, 60, 183, None, 60) == 1004670


def test_default_params_and_cameras(pkg):
    fp = pkg.default_frame_params(0, 500, 500)
    assert (fp.focalLength, fp.cameraPos[2], fp.numLights, fp.lights[0].intensity) == (250.0, -2.0, 1, 14.0)
    assert list(fp.cameraRot) == [1, 0, 0, 0, 1, 0, 0, 0, 1] and fp.aaSamples == 3 and fp.softShadowsSamples == 16
    fp = pkg.default_frame_params(1, 500, 500)
    assert (fp.focalLength, fp.cameraPos[2]) == (500.0, -3.0) and abs(fp.cameraRot[4] - 1.01) < 1e-7
    assert fp.backfaceCulling == 1 and fp.frustumCulling == 1
    pos, rot = pkg.orbit_camera(90, 360)
    assert np.allclose(pos, [2, 0, 0], atol=1e-6) and np.allclose(rot.reshape(3, 3)[2], [-1, 0, 0], atol=1e-6)
    tab = pkg.jitter_table(1, [0, -0.5, -0.7])
    assert np.abs(tab[:16] - np.array([0, -0.5, -0.7], np.float32)).max() <= 0.04 + 1e-6 and not tab[16:].any()


def test_bmp_writer(pkg, tmp_path):
    w, h = 5, 3  # 15 bytes per row -> pitch 16
    payload = np.arange(16 * 3, dtype=np.uint8)
    path = str(tmp_path / "t.bmp")
    pkg.write_bmp(path, payload, w, h)
    raw = open(path, "rb").read()
    assert len(raw) == 54 + 48 and raw[:2] == b"BM"
    assert int.from_bytes(raw[2:6], "little") == 102 and int.from_bytes(raw[10:14], "little") == 54
    assert int.from_bytes(raw[28:30], "little") == 24 and raw[54:] == payload.tobytes()
    assert pkg.load_library().b2r_bmp_payload_bytes(500, 500) == 750000  # the shipped screenshots: 750,054 bytes


def test_no_cpu_fallback(pkg):
    """Without a GPU the product refuses to run instead of silently rendering on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.B2RError, match="no CPU fallback"):
        pkg.Context(64, 64)


def test_product_never_touches_the_oracle():
    """Nothing under the package directory may import, link or mention the oracle as a dependency."""
    pkg_dir = os.path.join(ROOT, "cpp-raytracer-rasterizer_b200")
    for d, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(d, f)).read()
                assert "liboracle" not in text and "portbind" not in text and "refbind" not in text, os.path.join(d, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), os.path.join(d, f)


def test_stl_loader(pkg, tmp_path):
    """ASCII STL ingestion on a small mesh of our own, incl. the loader's quirks (tokens split on
    single spaces, the word 'vertex' dropped, everything scaled by -0.05f)."""
    p = tmp_path / "m.stl"
    p.write_text("solid t\n facet normal 0 0 1\n  outer loop\n  vertex 1 2 3\n  vertex  4.5  -6  7e0\n   vertex 8 9 10\n  endloop\n"
                 " endfacet\n facet normal 0 1 0\n  outer loop\n  vertex 0 0 0\n  vertex 20 0 0\n  vertex 0 0 20\n  endloop\n endfacet\nendsolid\n")
    t = pkg.load_stl(str(p))
    assert t.shape == (2, 15)
    s = np.float32(-0.05)
    assert np.array_equal(t[0, :9], (np.array([1, 2, 3, 4.5, -6, 7, 8, 9, 10], np.float32) * s))
    assert np.array_equal(t[:, 12:15], np.full((2, 3), 0.5, np.float32))
    e1, e2 = t[1, 3:6] - t[1, 0:3], t[1, 6:9] - t[1, 0:3]
    n = np.cross(e2, e1)
    assert np.allclose(t[1, 9:12], n / np.linalg.norm(n), atol=1e-6)
    with pytest.raises(pkg.B2RError):
        pkg.load_stl(str(tmp_path / "missing.stl"))


def test_stl_loader_equals_reference_loader(pkg):
    """Bit-exact against the reference's own LoadSTL on its enemy1.stl (where the reference is mounted)."""
    from oracle import refbind
    ref_dir = "/root/reference/rasteriser"
    if not os.path.isdir(ref_dir) or not refbind.available("ras", 96, 64):
        pytest.skip("reference sources / oracle/_ref not available")
    ra = refbind.RefRasteriser(96, 64)
    want = ra.load_stl(ref_dir)
    got = pkg.load_stl(os.path.join(ref_dir, "Source", "enemy1.stl"))
    assert got.shape == want.shape == (9028, 15)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_no_packed_fma_in_the_reference_order_code(pkg):
    """ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false; the reference-order helpers
    (csrc/exact.cuh) are written so that this pattern never arises.  The only FFMA2 in the library are the ones the
    raytracer's conservative per-ray filters ask for (__ffma2_rn): every kernel without those filters -- the
    FILTER = false variants of the trace kernel, which run exactly the same reference-order code, the rasteriser,
    the resolve and the sub-stage kernels -- must be free of FFMA2, and the packed adds/multiplies that are used on
    purpose must be there."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "cpp-raytracer-rasterizer_b200", "lib", "libb2r.so")
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    assert len(funcs) > 20
    seen_filtered = 0
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        n = f.count("FFMA2")
        m = re.search(r"rt_trace_shade_kernelILb[01]ELb[01]ELb([01])E", name)
        if m and m.group(1) == "1":
            seen_filtered += 1
            assert n >= 5, (name, n)       # 2 for the primary forms, 3 for the shadow forms (per copy of the loop)
        else:
            assert n == 0, (name, n)
    assert seen_filtered >= 8
    assert "FADD2" in sass and "FMUL2" in sass
