"""ctypes bindings for oracle/liboracle.so (the CPU restatement) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

INTERSECTION_DTYPE = np.dtype([("position", np.float32, 3), ("distance", np.float32), ("triangleIndex", np.int32)])
PIXEL_DTYPE = np.dtype([("x", np.int32), ("y", np.int32), ("zinv", np.float32), ("pos3d", np.float32, 3)])

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.oracle_bmp_payload_bytes.restype = C.c_size_t
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _t15(tris):
    return np.ascontiguousarray(tris, np.float32).reshape(-1, 15)


def rt_draw(tris15, fp, w, h, y0=0, y1=None, threads=None, ystep=1):
    """oracle_rt_draw: returns dict like the reference harness."""
    t = _t15(tris15)
    y1 = h if y1 is None else y1
    col = np.zeros((h, w, 3), np.float32)
    clo = np.zeros((h, w), INTERSECTION_DTYPE)
    clo["distance"] = np.finfo(np.float32).max
    clo["triangleIndex"] = -1
    foc = np.zeros((h, w), np.float32)
    cnt = (C.c_ulonglong * 2)()
    threads = threads or (os.cpu_count() or 1)
    rc = lib().oracle_rt_draw(_p(t), len(t), C.byref(fp), w, h, y0, y1, ystep, _p(col), _p(clo), _p(foc), threads, cnt)
    assert rc == 0, rc
    return dict(pixelColours=col, closest=clo, focalDistances=foc, primary_rays=int(cnt[0]),
                shadow_rays=int(cnt[1]))


def rt_closest_intersection(tris15, start, direction, closest=None, is_light=False, dof_focal=0.0):
    t = _t15(tris15)
    c = np.zeros(1, INTERSECTION_DTYPE)
    if closest is None:
        c["distance"] = np.finfo(np.float32).max
        c["triangleIndex"] = -1
    else:
        c[0] = closest
    s = np.ascontiguousarray(start, np.float32)
    d = np.ascontiguousarray(direction, np.float32)
    foc = C.c_float(0)
    hit = lib().oracle_rt_closest_intersection(_p(s), _p(d), _p(t), len(t), _p(c), int(is_light),
                                               C.c_float(dof_focal), C.byref(foc))
    return bool(hit), c[0], foc.value


def rt_direct_light(tris15, fp, closest):
    t = _t15(tris15)
    c = np.atleast_1d(np.array(closest, INTERSECTION_DTYPE))
    out = np.zeros(3, np.float32)
    lib().oracle_rt_direct_light(_p(c), _p(t), len(t), C.byref(fp), _p(out))
    return out


def ras_draw(tris15, culled, fp, w, h):
    t = _t15(tris15)
    m = None if culled is None else np.ascontiguousarray(culled, np.uint8)
    dep = np.zeros((h, w), np.float32)
    col = np.zeros((h, w, 3), np.float32)
    foc = np.zeros((h, w), np.float32)
    win = np.zeros((h, w), np.int32)
    cnt = (C.c_ulonglong * 4)()
    rc = lib().oracle_ras_draw(_p(t), _p(m), len(t), C.byref(fp), w, h, _p(dep), _p(col), _p(foc), _p(win), cnt)
    assert rc == 0, rc
    return dict(depthBuffer=dep, pixelColours=col, focalDistances=foc, winner=win, triangles=int(cnt[0]),
                rows=int(cnt[1]), depth_tests=int(cnt[2]), depth_passes=int(cnt[3]))


def ras_cull(tris15, fp, w, h):
    t = _t15(tris15)
    out = np.zeros(len(t), np.uint8)
    rc = lib().oracle_ras_cull(_p(t), len(t), C.byref(fp), w, h, _p(out))
    assert rc == 0
    return out


def ras_vertex_shader(fp, w, h, v):
    v = np.ascontiguousarray(v, np.float32)
    out = np.zeros(1, PIXEL_DTYPE)
    lib().oracle_ras_vertex_shader(C.byref(fp), w, h, _p(v), _p(out))
    return out[0]


def ras_interpolate(a, b, n):
    a = np.atleast_1d(np.array(a, PIXEL_DTYPE))
    b = np.atleast_1d(np.array(b, PIXEL_DTYPE))
    out = np.zeros(n, PIXEL_DTYPE)
    lib().oracle_ras_interpolate(_p(a), _p(b), _p(out), n)
    return out


def ras_compute_polygon_rows(vertex_pixels, max_rows=8192):
    vp = np.ascontiguousarray(vertex_pixels, PIXEL_DTYPE)
    left = np.zeros(max_rows, PIXEL_DTYPE)
    right = np.zeros(max_rows, PIXEL_DTYPE)
    rows = lib().oracle_ras_compute_polygon_rows(_p(vp), _p(left), _p(right), max_rows)
    assert 0 < rows <= max_rows
    return left[:rows].copy(), right[:rows].copy()


def ras_pixel_shader(fp, w, h, pixel, color, normal):
    p = np.atleast_1d(np.array(pixel, PIXEL_DTYPE))
    c = np.ascontiguousarray(color, np.float32)
    nrm = np.ascontiguousarray(normal, np.float32)
    out = np.zeros(3, np.float32)
    foc = C.c_float(0)
    lib().oracle_ras_pixel_shader(C.byref(fp), w, h, _p(p), _p(c), _p(nrm), _p(out), C.byref(foc))
    return out, foc.value


def resolve_surface(pixel_colours, focal_distances, dof_enabled=False, dof_kernel=8):
    col = np.ascontiguousarray(pixel_colours, np.float32)
    h, w = col.shape[:2]
    foc = None if focal_distances is None else np.ascontiguousarray(focal_distances, np.float32)
    if dof_enabled:
        assert foc is not None
    out = np.zeros((h, w), np.uint32)
    rc = lib().oracle_resolve_surface(_p(col), _p(foc), w, h, int(dof_enabled), dof_kernel, _p(out))
    assert rc == 0
    return out


def surface_to_bgr8(surface):
    s = np.ascontiguousarray(surface, np.uint32)
    h, w = s.shape
    out = np.zeros(lib().oracle_bmp_payload_bytes(w, h), np.uint8)
    lib().oracle_surface_to_bgr8(_p(s), w, h, _p(out))
    return out
