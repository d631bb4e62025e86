#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref, compiled from /root/reference).

Run in the build container (needs /root/reference and `python oracle/build_ref.py`):
    python tests/golden/make_golden.py
The fixtures pin the CPU oracle (and through it the CUDA path) to outputs of the unmodified
reference arithmetic; they are small (<= 160x120 frames, a few hundred sub-stage vectors).
Sub-stage vectors come from the reference's own functions with the reference signatures
(ClosestIntersection, DirectLight, VertexShader, Interpolate, ComputePolygonRows, PixelShader).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.refbind import RefRaytracer, RefRasteriser, PIXEL_DTYPE, INTERSECTION_DTYPE  # noqa: E402
import __graft_entry__ as g  # noqa: E402
from util import random_soup, rot_y  # noqa: E402

pkg = g.load_package()
LIGHT = np.array([[0, -0.5, -0.7, 1, 1, 1, 14]], np.float32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rt_frames():
    for (w, h) in [(96, 64), (64, 96), (160, 120)]:
        rt = RefRaytracer(w, h)
        tris = rt.load_test_model()
        table = rt.add_light_reference(1, True, [0, -0.5, -0.7], [1, 1, 1], 14.0)
        for name, aa, soft, yaw, pos in [("default", 0, 0, 0.0, [0, 0, -2]), ("aa3", 3, 0, 0.0, [0, 0, -2]),
                                         ("soft16", 0, 1, 0.0, [0, 0, -2]), ("aa2_soft16_yaw", 2, 1, 0.6, [1.1, 0, -1.7])]:
            if (w, h) == (160, 120) and name != "default":
                continue
            rt.set_lights(LIGHT, table)
            rot = rt.set_camera_yaw(pos, yaw, h / 2.0)
            rt.set_flags(aa=aa > 0, aa_samples=max(aa, 1), soft=bool(soft), soft_samples=16)
            r = rt.draw()
            np.savez_compressed(os.path.join(HERE, f"rt_{w}x{h}_{name}.npz"), tris=tris, table=table, rot=rot,
                                pos=np.array(pos, np.float32), focal=np.float32(h / 2.0), aa=aa, soft=soft,
                                pixelColours=r["pixelColours"], focalDistances=r["focalDistances"],
                                closest=r["closest"], surface=r["surface"])
    # a random multi-light scene
    rng = np.random.default_rng(11)
    w, h = 96, 64
    rt = RefRaytracer(w, h)
    tris = random_soup(rng, 40)
    lights = np.concatenate([rng.uniform(-1, 1, (3, 3)), rng.uniform(0.2, 1, (3, 3)), rng.uniform(2, 20, (3, 1))], 1).astype(np.float32)
    table = rng.uniform(-1, 1, (256, 3)).astype(np.float32)
    rt.set_triangles(tris)
    rt.set_lights(lights, table)
    rot = rot_y(0.3)
    pos = np.array([0.2, -0.1, -2.5], np.float32)
    rt.set_camera(pos, rot, h / 2.0)
    rt.set_flags(aa=True, aa_samples=2, soft=True, soft_samples=4)
    r = rt.draw()
    np.savez_compressed(os.path.join(HERE, "rt_96x64_soup_3lights.npz"), tris=tris, table=table, rot=rot, pos=pos,
                        focal=np.float32(h / 2.0), aa=2, soft=1, soft_samples=4, lights=lights,
                        pixelColours=r["pixelColours"], focalDistances=r["focalDistances"], closest=r["closest"],
                        surface=r["surface"])


def ras_frames():
    for (w, h) in [(96, 64), (64, 96), (160, 120)]:
        ra = RefRasteriser(w, h)
        tris = ra.load_test_model()
        ra.set_lights(LIGHT)
        ra.set_flags()
        for name, yaw, pos in [("default", 0.0, [0, 0, -3]), ("yaw", 0.25, [0.7, 0, -2.9])]:
            rot, culled, _ = ra.update_yaw(pos, yaw, float(h))
            r = ra.draw()
            np.savez_compressed(os.path.join(HERE, f"ras_{w}x{h}_{name}.npz"), tris=tris, rot=rot, culled=culled,
                                pos=np.array(pos, np.float32), focal=np.float32(h), depthBuffer=r["depthBuffer"],
                                pixelColours=r["pixelColours"], focalDistances=r["focalDistances"],
                                winner=r["winner"], surface=r["surface"],
                                counts=np.array([r["depth_tests"], r["depth_passes"]], np.int64))


def substage_vectors():
    rng = np.random.default_rng(5)
    rt = RefRaytracer(96, 64)
    tris = rt.load_test_model()
    rt.set_lights(LIGHT)
    rt.set_flags()
    n = 400
    starts = np.tile(np.array([0, 0, -2], np.float32), (n, 1))
    starts[n // 2:] = rng.uniform(-0.9, 0.9, (n - n // 2, 3)).astype(np.float32)
    dirs = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    dirs[:n // 2, 2] = np.abs(dirs[:n // 2, 2]) + 0.5
    hits = np.zeros(n, np.uint8)
    clo = np.zeros(n, INTERSECTION_DTYPE)
    light = np.zeros((n, 3), np.float32)
    for i in range(n):
        hit, c = rt.closest_intersection(starts[i], dirs[i], is_light=(i % 2 == 1))
        hits[i], clo[i] = hit, c
        if hit:
            light[i] = rt.direct_light(c)
    np.savez_compressed(os.path.join(HERE, "sub_rt.npz"), tris=tris, starts=starts, dirs=dirs, hits=hits, closest=clo,
                        direct_light=light)

    ra = RefRasteriser(160, 120)
    rtris = ra.load_test_model()
    ra.set_lights(LIGHT)
    ra.set_flags()
    rot, _, _ = ra.update_yaw([0.3, 0.1, -3.0], 0.2, 120.0)
    verts = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    vp = np.zeros(300, PIXEL_DTYPE)
    for i in range(300):
        vp[i] = ra.vertex_shader(verts[i])
    lefts, rights, offs = [], [], [0]
    for i in range(100):
        l, r = ra.compute_polygon_rows(vp[3 * i:3 * i + 3])
        lefts.append(l)
        rights.append(r)
        offs.append(offs[-1] + len(l))
    interp = [ra.interpolate(vp[2 * i], vp[2 * i + 1], 1 + abs(int(vp[2 * i]["y"]) - int(vp[2 * i + 1]["y"]))) for i in range(50)]
    ioffs = np.cumsum([0] + [len(x) for x in interp])
    ps_in = np.zeros(200, PIXEL_DTYPE)
    ps_in["x"] = rng.integers(0, 160, 200)
    ps_in["y"] = rng.integers(0, 120, 200)
    ps_in["zinv"] = rng.uniform(0.2, 0.6, 200).astype(np.float32)
    ps_in["pos3d"] = np.concatenate([rng.uniform(-0.4, 0.4, (200, 2)), np.ones((200, 1))], 1).astype(np.float32)
    ps_col = rng.uniform(0.1, 0.9, (200, 3)).astype(np.float32)
    ps_nrm = rng.uniform(-1, 1, (200, 3)).astype(np.float32)
    ps_out = np.zeros((200, 3), np.float32)
    ps_foc = np.zeros(200, np.float32)
    for i in range(200):
        ps_out[i], ps_foc[i] = ra.pixel_shader(ps_in[i], ps_col[i], ps_nrm[i])
    np.savez_compressed(os.path.join(HERE, "sub_ras.npz"), rot=rot, pos=np.array([0.3, 0.1, -3.0], np.float32),
                        focal=np.float32(120.0), verts=verts, vertex_pixels=vp, left=np.concatenate(lefts),
                        right=np.concatenate(rights), row_offsets=np.array(offs), interp=np.concatenate(interp),
                        interp_offsets=ioffs, ps_in=ps_in, ps_col=ps_col, ps_nrm=ps_nrm, ps_out=ps_out, ps_foc=ps_foc)


def kat_500():
    """Digests + known answers of the two default 500x500 frames."""
    rt = RefRaytracer(500, 500)
    rt.load_test_model()
    rt.set_lights(LIGHT)
    rt.set_camera_yaw([0, 0, -2], 0.0, 250.0)
    rt.set_flags()
    r = rt.draw()
    idx = r["closest"]["triangleIndex"]
    k = {"rt": {"hit_histogram": {int(a): int(b) for a, b in zip(*np.unique(idx, return_counts=True))},
                "black_surface_pixels": int(((r["surface"] & 0xFFFFFF) == 0).sum()),
                "sha256": {n: sha(r[n]) for n in ("pixelColours", "focalDistances", "closest", "surface")}}}
    ra = RefRasteriser(500, 500)
    ra.load_test_model()
    ra.set_lights(LIGHT)
    ra.set_flags()
    _, culled, _ = ra.update_yaw([0, 0, -3], 0.0, 500.0)
    r = ra.draw()
    k["ras"] = {"culled": "".join(map(str, culled)), "covered": int((r["winner"] >= 0).sum()),
                "depth_tests": int(r["depth_tests"]), "depth_passes": int(r["depth_passes"]),
                "sha256": {n: sha(r[n]) for n in ("depthBuffer", "pixelColours", "focalDistances", "winner", "surface")}}
    k["scene_fnv1a32"] = "b715a8a2"
    json.dump(k, open(os.path.join(HERE, "kat_500x500.json"), "w"), indent=1, sort_keys=True)


def kat_4k():
    """Digests of the BASELINE configurations at the size bench.py quotes them (3840x2160), from the reference itself:
    config 3 (AA 4x4), config 3b (16 jittered light samples), three frames of the config-5 orbit, config 4 (1,004,670
    triangles).  The GPU test compares digests only, so no CPU run is needed on the GPU box."""
    w, h = 3840, 2160
    k = {}
    rt = RefRaytracer(w, h)
    rt.load_test_model()
    table = rt.add_light_reference(1, True, [0, -0.5, -0.7], [1, 1, 1], 14.0)
    names = ("pixelColours", "focalDistances", "closest", "surface")
    rt.set_lights(LIGHT, table)
    rt.set_camera_yaw([0, 0, -2], 0.0, h / 2.0)
    rt.set_flags(aa=True, aa_samples=4)
    r = rt.draw()
    k["config3_aa4x4"] = {"sha256": {n: sha(r[n]) for n in names},
                          "hit_pixels": int((r["closest"]["triangleIndex"] >= 0).sum())}
    rt.set_flags(soft=True, soft_samples=16)
    r = rt.draw()
    k["config3b_soft16"] = {"sha256": {n: sha(r[n]) for n in names}, "jitter_table_sha256": sha(table)}
    rt.set_flags()
    k["config5_orbit"] = {}
    for f in (45, 170, 300):
        pos, rot = pkg.orbit_camera(f, 360)
        rt.set_camera(pos, rot, h / 2.0)
        r = rt.draw()
        k["config5_orbit"][str(f)] = {"sha256": {n: sha(r[n]) for n in names}}
    del rt
    ra = RefRasteriser(w, h)
    big = pkg.tessellate(pkg.cornell_box(), 183)
    ra.set_triangles(big)
    ra.set_lights(LIGHT)
    ra.set_flags()
    _, culled, _ = ra.update_yaw([0, 0, -3], 0.0, float(h))
    r = ra.draw()
    k["config4_ras_1m"] = {"triangles": int(len(big)), "culled": int(culled.sum()), "culled_sha256": sha(culled.astype(np.uint8)),
                           "covered": int((r["winner"] >= 0).sum()), "depth_tests": int(r["depth_tests"]),
                           "sha256": {n: sha(r[n]) for n in ("depthBuffer", "pixelColours", "focalDistances", "winner", "surface")}}
    json.dump(k, open(os.path.join(HERE, "kat_3840x2160.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "kat4k":
        kat_4k()
        sys.exit(0)
    rt_frames()
    ras_frames()
    substage_vectors()
    kat_500()
    kat_4k()
    print("golden fixtures written to", HERE)
