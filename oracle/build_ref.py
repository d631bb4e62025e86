#!/usr/bin/env python3
"""Build oracle/_ref/*.so from the reference sources WHERE THEY LIE (/root/reference).

TEST INFRASTRUCTURE ONLY.  Nothing produced here is linked into, imported by
or executed from the product library; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may load these objects.

What it does:
  * reads raytracer/Source/raytracer.cpp and rasteriser/Source/rasteriser.cpp,
  * applies the mechanical patches P1..P4 below with exact-count checks (a
    patch that matches a different number of sites than expected aborts),
  * writes the patched text to a TEMPORARY directory (never into the repo),
  * compiles oracle/ref_harness_{rt,ras}.cpp -- which #include the patched file --
    with the reference Makefile's flags (raytracer/Makefile:13-15: g++ -fopenmp
    -O3) plus -ffp-contract=off and no -march, against oracle/sdl_stub/SDL.h and
    the GLM 0.9.7.2 vendored in the reference (raytracer/glm),
  * leaves only shared objects in oracle/_ref/ (git-ignored, NOT gpurun-ignored,
    so the prebuilt objects travel to the GPU box where /root/reference is absent).

One object per (program, screen size): the reference sizes static arrays with
its compile-time SCREEN_WIDTH/SCREEN_HEIGHT.
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B2R_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (W, H) variants.  500x500 is the reference default; 3840x2160 is the 4K
# configs; the small ones keep the CPU parity tests fast and exercise W != H.
SIZES = [(500, 500), (3840, 2160), (96, 64), (64, 96), (160, 120)]


def _sub(text, pattern, repl, expect, what):
    new, n = re.subn(pattern, repl, text)
    if n != expect:
        raise SystemExit(f"patch {what}: expected {expect} sites, matched {n}")
    return new


def patch_common(src, n_stride_sites):
    # P1: compile-time screen size (raytracer.cpp:67-68, rasteriser.cpp:35-36)
    src = _sub(src, r"const int SCREEN_WIDTH = 500;", "const int SCREEN_WIDTH = REF_W;", 1, "P1w")
    src = _sub(src, r"const int SCREEN_HEIGHT = 500;", "const int SCREEN_HEIGHT = REF_H;", 1, "P1h")
    # P2: row stride y*SCREEN_HEIGHT -> y*SCREEN_WIDTH (no-op for square screens)
    src = _sub(src, r"([yz\)])\*SCREEN_HEIGHT", r"\1*SCREEN_WIDTH", n_stride_sites, "P2")
    return src


def patch_rt(src):
    src = patch_common(src, 9)
    # P7 (timing only): let the harness render a strided subset of rows so the CPU baseline can be a
    # bounded, representative sample of a 4K frame.  Defaults (0, SCREEN_HEIGHT, 1) = the reference loop.
    src = _sub(src, r"for \(int y = 0; y < SCREEN_HEIGHT; y\+\+\)",
               "for (int y = ref_y0; y < ref_y1; y += ref_ystep)", 1, "P7")
    return src


def patch_ras(src):
    src = patch_common(src, 7)
    # P3: define the entries Bresenham() skips (rasteriser.cpp:599,663,606)
    src = _sub(
        src,
        r"vector<Pixel> line \(pixels\);",
        "vector<Pixel> line (pixels); for(int q_=0;q_<pixels;++q_){line[q_].x=-1;line[q_].y=-1;line[q_].zinv=0.0f;}",
        1,
        "P3",
    )
    # P4: remember which triangle wrote the depth buffer (rasteriser.cpp:477,606-608)
    src = _sub(
        src,
        r"DrawPolygon\( vertices , triangles\[i\]\.color, triangles\[i\]\.normal\);",
        "ref_cur_tri = (int)i; DrawPolygon( vertices , triangles[i].color, triangles[i].normal);",
        1,
        "P4a",
    )
    src = _sub(
        src,
        r"depthBuffer\[line\[i\]\.y\]\[line\[i\]\.x\] = line\[i\]\.zinv;",
        "depthBuffer[line[i].y][line[i].x] = line[i].zinv; "
        "ref_winner[line[i].y*SCREEN_WIDTH + line[i].x] = ref_cur_tri; ++ref_depth_passes;",
        1,
        "P4b",
    )
    src = _sub(
        src,
        r"(\t\t// Ensure pixel is on the screen and is closer)",
        "\t\tif(line[i].y < SCREEN_HEIGHT && line[i].y >= 0 && line[i].x < SCREEN_WIDTH && line[i].x >= 0) ++ref_depth_tests;\n\\1",
        1,
        "P4c",
    )
    return src


def build(sizes=SIZES, verbose=True):
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not present: cannot (re)build oracle/_ref here")
    os.makedirs(OUT, exist_ok=True)
    rt_src = open(os.path.join(REF, "raytracer/Source/raytracer.cpp")).read()
    ras_src = open(os.path.join(REF, "rasteriser/Source/rasteriser.cpp")).read()
    glm = os.path.join(REF, "raytracer")  # <glm/glm.hpp> -> raytracer/glm (GLM 0.9.7.2)
    flags = ["g++", "-fopenmp", "-O3", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
             "-w", "-std=gnu++14", "-I" + os.path.join(HERE, "sdl_stub"), "-I" + glm]
    jobs = []
    with tempfile.TemporaryDirectory(prefix="b2r_ref_") as tmp:
        prt = os.path.join(tmp, "rt_patched.cpp")
        pras = os.path.join(tmp, "ras_patched.cpp")
        open(prt, "w").write(patch_rt(rt_src))
        open(pras, "w").write(patch_ras(ras_src))
        for (w, h) in sizes:
            for prog, patched, srcdir, harness in (
                ("rt", prt, "raytracer/Source", "ref_harness_rt.cpp"),
                ("ras", pras, "rasteriser/Source", "ref_harness_ras.cpp"),
            ):
                out = os.path.join(OUT, f"libref_{prog}_{w}x{h}.so")
                src = os.path.join(HERE, harness)
                if os.path.exists(out) and os.path.getmtime(out) > max(
                        os.path.getmtime(src), os.path.getmtime(__file__),
                        os.path.getmtime(os.path.join(HERE, "sdl_stub/SDL.h"))):
                    continue
                cmd = flags + [f"-DREF_W={w}", f"-DREF_H={h}", f'-DREF_PATCHED_SOURCE="{patched}"',
                               "-I" + os.path.join(REF, srcdir), src, "-o", out]
                jobs.append((out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        fail = False
        for out, p in jobs:
            log = p.communicate()[0].decode()
            if p.returncode != 0:
                fail = True
                sys.stderr.write(f"FAILED {out}\n{log}\n")
            elif verbose:
                print("built", os.path.relpath(out, HERE))
        if fail:
            raise SystemExit(1)


def build_integration(sizes=((500, 500), (160, 120)), verbose=True):
    """The reference programs THEMSELVES -- main(), Update(), LoadTestModel, SDLauxiliary.h -- with only the body of
    Draw() replaced by the calls INTEGRATION.md gives (patch P8: the reference definition is renamed
    Draw_reference and oracle/integration_draw_{rt,ras}.inc is appended).  Executables in oracle/_ref/, linked against
    libb2r.so; tests/test_integration.py runs them on the GPU box with B2R_STUB_FRAMES=1 and compares the
    screenshot.bmp their own main() writes with the oracle's frame."""
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not present")
    os.makedirs(OUT, exist_ok=True)
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "cpp-raytracer-rasterizer_b200", "lib")
    glm = os.path.join(REF, "raytracer")
    with tempfile.TemporaryDirectory(prefix="b2r_int_") as tmp:
        for prog, srcdir, patch in (("raytracer", "raytracer/Source", patch_common), ("rasteriser", "rasteriser/Source", patch_common)):
            src = open(os.path.join(REF, srcdir, prog + ".cpp")).read()
            src = patch_common(src, 9 if prog == "raytracer" else 7)  # P1, P2 only: no oracle instrumentation here
            src = _sub(src, r"\nvoid Draw\(\)\n\{", "\nvoid Draw_reference()\n{", 1, "P8")
            inc = open(os.path.join(HERE, "integration_draw_rt.inc" if prog == "raytracer" else "integration_draw_ras.inc")).read()
            patched = os.path.join(tmp, prog + "_b2r.cpp")
            open(patched, "w").write("#include <cstring>\n#include <cstdint>\n" + src + "\n" + inc)
            for (w, h) in sizes:
                out = os.path.join(OUT, f"{prog}_b2r_{w}x{h}")
                cmd = ["g++", "-fopenmp", "-O3", "-ffp-contract=off", "-w", "-std=gnu++14", f"-DREF_W={w}", f"-DREF_H={h}",
                       "-I" + os.path.join(HERE, "sdl_stub"), "-I" + glm, "-I" + os.path.join(REF, srcdir),
                       "-I" + os.path.join(root, "include"), patched, "-o", out, "-L" + libdir, "-lb2r",
                       "-Wl,-rpath,$ORIGIN/../../cpp-raytracer-rasterizer_b200/lib"]
                subprocess.check_call(cmd)
                if verbose:
                    print("built", os.path.relpath(out, HERE))


if __name__ == "__main__":
    build()
    if "--integration" in sys.argv:
        build_integration()
