"""The host C++ drop-ins (host/): the reference's Draw() interface on top of libb2r.so."""
import os
import subprocess

import numpy as np
import pytest

from util import ROOT

HOST = os.path.join(ROOT, "cpp-raytracer-rasterizer_b200", "host")
LIBDIR = os.path.join(ROOT, "cpp-raytracer-rasterizer_b200", "lib")


def test_reference_type_layouts_match_the_reference_headers(tmp_path):
    """Compile reference_types.h next to the reference's own TestModel.h (when mounted) and compare sizeof."""
    if not os.path.isdir("/root/reference/rasteriser/Source"):
        pytest.skip("reference sources not mounted")
    src = tmp_path / "layout.cpp"
    src.write_text('#include <cstdio>\n#include <cstddef>\n#include <omp.h>\n#include "TestModel.h"\n#include "reference_types.h"\n'
                   "int main(){\n"
                   "static_assert(sizeof(Triangle)==sizeof(b2rhost::TriangleRA),\"Triangle\");\n"
                   "static_assert(offsetof(Triangle,isCulled)==offsetof(b2rhost::TriangleRA,isCulled),\"isCulled\");\n"
                   "static_assert(offsetof(Triangle,color)==offsetof(b2rhost::TriangleRA,color),\"color\");\n"
                   "static_assert(sizeof(Pixel)==sizeof(b2rhost::Pixel)&&offsetof(Pixel,pos3d)==offsetof(b2rhost::Pixel,pos3d),\"Pixel\");\n"
                   "static_assert(sizeof(Light)==sizeof(b2rhost::Light)&&sizeof(Vertex)==sizeof(b2rhost::Vertex),\"Light\");\n"
                   'std::puts("ok");return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["g++", "-w", "-fopenmp", "-DB2R_WITH_GLM", "-I/root/reference/rasteriser/Source", "-I/root/reference/raytracer",
                           "-I" + HOST, str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)]).strip() == b"ok"


def test_host_library_builds():
    subprocess.check_call(["make", "-s", "-C", HOST])
    assert os.path.exists(os.path.join(LIBDIR, "libb2r_host.so")) and os.path.exists(os.path.join(LIBDIR, "b2r_headless"))


@pytest.mark.gpu
@pytest.mark.parametrize("prog", ["raytracer", "rasteriser"])
def test_headless_frame_equals_oracle(pkg, oracle, prog, tmp_path):
    """Update(); Draw(); SaveBMP through the drop-in == the oracle's frame, byte for byte."""
    subprocess.check_call(["make", "-s", "-C", HOST])
    w, h = 160, 120
    out = str(tmp_path / prog)
    subprocess.check_call([os.path.join(LIBDIR, "b2r_headless"), prog, str(w), str(h), "--out", out])
    raw = open(out + ".bmp", "rb").read()
    tris = pkg.cornell_box()
    if prog == "raytracer":
        col = oracle.rt_draw(tris, pkg.default_frame_params(0, w, h), w, h)["pixelColours"]
    else:
        fp = pkg.default_frame_params(1, w, h)
        col = oracle.ras_draw(tris, oracle.ras_cull(tris, fp, w, h), fp, w, h)["pixelColours"]
    want = oracle.surface_to_bgr8(oracle.resolve_surface(col, None))
    assert len(raw) == 54 + len(want) and raw[54:] == want.tobytes()


@pytest.mark.gpu
def test_headless_orbit_animation(tmp_path):
    out = str(tmp_path / "orbit")
    subprocess.check_call([os.path.join(LIBDIR, "b2r_headless"), "raytracer", "96", "64", "--frames", "4", "--out", out])
    frames = [open(f"{out}_{i:04d}.bmp", "rb").read() for i in range(4)]
    assert len(set(frames)) == 4


@pytest.mark.gpu
def test_stage_functions_reassemble_draw():
    """host/stage_check.cpp: the reference's Draw() loops rebuilt from ClosestIntersection/DirectLight and
    VertexShader/ComputePolygonRows/PixelShader (reference signatures, GPU-backed) equal the fused Draw() bit for bit."""
    subprocess.check_call(["make", "-s", "-C", HOST])
    out = subprocess.run([os.path.join(LIBDIR, "b2r_stage_check")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "raytracer stages vs Draw(): 0 of" in out.stdout and "rasteriser stages vs Draw(): 0 of" in out.stdout
