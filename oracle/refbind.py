"""ctypes bindings for oracle/_ref/libref_{rt,ras}_WxH.so -- TEST INFRASTRUCTURE ONLY.

These objects are the reference's own raytracer.cpp / rasteriser.cpp compiled
by oracle/build_ref.py.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

INTERSECTION_DTYPE = np.dtype([("position", np.float32, 3), ("distance", np.float32), ("triangleIndex", np.int32)])
PIXEL_DTYPE = np.dtype([("x", np.int32), ("y", np.int32), ("zinv", np.float32), ("pos3d", np.float32, 3)])

_fp = C.POINTER(C.c_float)


def _p(a, ty=C.c_void_p):
    return None if a is None else a.ctypes.data_as(ty)


def available(prog, w, h):
    return os.path.exists(os.path.join(REF_DIR, f"libref_{prog}_{w}x{h}.so"))


class RefRaytracer:
    """The reference raytracer at one compile-time screen size."""

    def __init__(self, w, h):
        self.lib = C.CDLL(os.path.join(REF_DIR, f"libref_rt_{w}x{h}.so"))
        L = self.lib
        L.ref_rt_draw.restype = C.c_double
        L.ref_rt_set_camera.argtypes = [_fp, _fp, C.c_float]
        L.ref_rt_set_camera_yaw.argtypes = [_fp, C.c_float, C.c_float]
        L.ref_rt_set_flags.argtypes = [C.c_int] * 5 + [C.c_float, C.c_int]
        L.ref_rt_add_light_reference.argtypes = [C.c_uint, C.c_int, _fp, _fp, C.c_float]
        self.w, self.h = L.ref_rt_width(), L.ref_rt_height()
        assert (self.w, self.h) == (w, h)
        assert L.ref_rt_sizeof_triangle() == 60 and L.ref_rt_sizeof_intersection() == 20
        assert L.ref_rt_sizeof_light() == 28

    def load_test_model(self):
        n = self.lib.ref_rt_load_test_model()
        out = np.zeros((n, 15), np.float32)
        self.lib.ref_rt_get_triangles(_p(out))
        return out

    def set_triangles(self, t15):
        t15 = np.ascontiguousarray(t15, np.float32).reshape(-1, 15)
        self.lib.ref_rt_set_triangles(_p(t15), len(t15))

    def set_camera(self, pos, rot9_colmajor, focal):
        pos = np.ascontiguousarray(pos, np.float32)
        rot = np.ascontiguousarray(rot9_colmajor, np.float32).reshape(9)
        self.lib.ref_rt_set_camera(_p(pos, _fp), _p(rot, _fp), C.c_float(focal))

    def set_camera_yaw(self, pos, yaw, focal):
        pos = np.ascontiguousarray(pos, np.float32)
        self.lib.ref_rt_set_camera_yaw(_p(pos, _fp), C.c_float(yaw), C.c_float(focal))
        rot = np.zeros(9, np.float32)
        self.lib.ref_rt_get_camera_rot(_p(rot))
        return rot

    def set_lights(self, lights7, random768=None):
        lights7 = np.ascontiguousarray(lights7, np.float32).reshape(-1, 7)
        r = None if random768 is None else np.ascontiguousarray(random768, np.float32).reshape(768)
        self.lib.ref_rt_set_lights(len(lights7), _p(lights7), _p(r))

    def add_light_reference(self, seed, reset, pos, color, intensity):
        pos = np.ascontiguousarray(pos, np.float32)
        color = np.ascontiguousarray(color, np.float32)
        self.lib.ref_rt_add_light_reference(seed, int(reset), _p(pos, _fp), _p(color, _fp), C.c_float(intensity))
        out = np.zeros((256, 3), np.float32)
        self.lib.ref_rt_get_random_positions(_p(out))
        return out

    def set_flags(self, aa=False, aa_samples=3, soft=False, soft_samples=16, dof=False, dof_focal=1.3, threads=0):
        self.lib.ref_rt_set_flags(int(aa), aa_samples, int(soft), soft_samples, int(dof), C.c_float(dof_focal), threads)

    def draw(self, want_surface=True):
        n = self.w * self.h
        col = np.zeros((self.h, self.w, 3), np.float32)
        foc = np.zeros((self.h, self.w), np.float32)
        clo = np.zeros((self.h, self.w), INTERSECTION_DTYPE)
        surf = np.zeros((self.h, self.w), np.uint32) if want_surface else None
        secs = self.lib.ref_rt_draw(_p(col), _p(foc), _p(clo), _p(surf))
        return dict(pixelColours=col, focalDistances=foc, closest=clo, surface=surf, seconds=secs)

    def time_draw(self):
        return self.lib.ref_rt_draw(None, None, None, None)

    def set_rows(self, y0=0, y1=None, step=1):
        """P7: restrict Draw() to rows y0, y0+step, ... (timing samples only)."""
        self.lib.ref_rt_set_rows(y0, self.h if y1 is None else y1, step)

    def closest_intersection(self, start, direction, closest=None, is_light=False):
        c = np.zeros((), INTERSECTION_DTYPE)
        if closest is None:
            c["distance"] = np.finfo(np.float32).max
            c["triangleIndex"] = -1
        else:
            c[...] = closest
        s = np.ascontiguousarray(start, np.float32)
        d = np.ascontiguousarray(direction, np.float32)
        c = np.atleast_1d(c)
        hit = self.lib.ref_rt_closest_intersection(_p(s), _p(d), _p(c), int(is_light))
        return bool(hit), c[0]

    def direct_light(self, closest):
        c = np.atleast_1d(np.array(closest, INTERSECTION_DTYPE))
        out = np.zeros(3, np.float32)
        self.lib.ref_rt_direct_light(_p(c), _p(out))
        return out


class RefRasteriser:
    """The reference rasteriser at one compile-time screen size."""

    def __init__(self, w, h):
        self.lib = C.CDLL(os.path.join(REF_DIR, f"libref_ras_{w}x{h}.so"))
        L = self.lib
        L.ref_ras_draw.restype = C.c_double
        L.ref_ras_update_yaw.restype = C.c_double
        L.ref_ras_update_yaw.argtypes = [_fp, C.c_float, C.c_float]
        L.ref_ras_set_camera.argtypes = [_fp, _fp, C.c_float]
        L.ref_ras_set_flags.argtypes = [C.c_int] * 3 + [C.c_float]
        L.ref_ras_depth_tests.restype = C.c_longlong
        L.ref_ras_depth_passes.restype = C.c_longlong
        self.w, self.h = L.ref_ras_width(), L.ref_ras_height()
        assert (self.w, self.h) == (w, h)
        assert L.ref_ras_sizeof_triangle() == 64 and L.ref_ras_sizeof_pixel() == 24

    def load_test_model(self):
        n = self.lib.ref_ras_load_test_model()
        return self.get_triangles(n)

    def load_stl(self, reference_rasteriser_dir):
        cwd = os.getcwd()
        os.chdir(reference_rasteriser_dir)
        try:
            n = self.lib.ref_ras_load_stl_cwd()
        finally:
            os.chdir(cwd)
        return self.get_triangles(n)

    def get_triangles(self, n=None):
        n = self.lib.ref_ras_num_triangles() if n is None else n
        out = np.zeros((n, 15), np.float32)
        self.lib.ref_ras_get_triangles(_p(out))
        return out

    def set_triangles(self, t15):
        t15 = np.ascontiguousarray(t15, np.float32).reshape(-1, 15)
        self.lib.ref_ras_set_triangles(_p(t15), len(t15))

    def set_culled(self, mask):
        m = np.ascontiguousarray(mask, np.uint8)
        assert len(m) == self.lib.ref_ras_num_triangles()
        self.lib.ref_ras_set_culled(_p(m))

    def get_culled(self):
        m = np.zeros(self.lib.ref_ras_num_triangles(), np.uint8)
        self.lib.ref_ras_get_culled(_p(m))
        return m

    def set_camera(self, pos, rot9_colmajor, focal):
        pos = np.ascontiguousarray(pos, np.float32)
        rot = np.ascontiguousarray(rot9_colmajor, np.float32).reshape(9)
        self.lib.ref_ras_set_camera(_p(pos, _fp), _p(rot, _fp), C.c_float(focal))

    def update_yaw(self, pos, yaw, focal):
        """Reference Update(): cameraRot from yaw + isCulled.  Returns (rot9, culled, seconds)."""
        pos = np.ascontiguousarray(pos, np.float32)
        secs = self.lib.ref_ras_update_yaw(_p(pos, _fp), C.c_float(yaw), C.c_float(focal))
        rot = np.zeros(9, np.float32)
        self.lib.ref_ras_get_camera_rot(_p(rot))
        return rot, self.get_culled(), secs

    def set_lights(self, lights7):
        lights7 = np.ascontiguousarray(lights7, np.float32).reshape(-1, 7)
        self.lib.ref_ras_set_lights(len(lights7), _p(lights7))

    def set_flags(self, backface=True, frustum=True, dof=False, dof_focal=1.9):
        self.lib.ref_ras_set_flags(int(backface), int(frustum), int(dof), C.c_float(dof_focal))

    def draw(self, want_surface=True):
        dep = np.zeros((self.h, self.w), np.float32)
        col = np.zeros((self.h, self.w, 3), np.float32)
        foc = np.zeros((self.h, self.w), np.float32)
        win = np.zeros((self.h, self.w), np.int32)
        surf = np.zeros((self.h, self.w), np.uint32) if want_surface else None
        secs = self.lib.ref_ras_draw(_p(dep), _p(col), _p(foc), _p(win), _p(surf))
        return dict(depthBuffer=dep, pixelColours=col, focalDistances=foc, winner=win, surface=surf,
                    seconds=secs, depth_tests=self.lib.ref_ras_depth_tests(),
                    depth_passes=self.lib.ref_ras_depth_passes())

    def time_draw(self):
        return self.lib.ref_ras_draw(None, None, None, None, None)

    def vertex_shader(self, v):
        v = np.ascontiguousarray(v, np.float32)
        out = np.zeros(1, PIXEL_DTYPE)
        self.lib.ref_ras_vertex_shader(_p(v), _p(out))
        return out[0]

    def compute_polygon_rows(self, vertex_pixels, max_rows=8192):
        vp = np.ascontiguousarray(vertex_pixels, PIXEL_DTYPE)
        left = np.zeros(max_rows, PIXEL_DTYPE)
        right = np.zeros(max_rows, PIXEL_DTYPE)
        rows = self.lib.ref_ras_compute_polygon_rows(_p(vp), _p(left), _p(right), max_rows)
        assert rows <= max_rows
        return left[:rows].copy(), right[:rows].copy()

    def interpolate(self, a, b, n):
        a = np.atleast_1d(np.array(a, PIXEL_DTYPE))
        b = np.atleast_1d(np.array(b, PIXEL_DTYPE))
        out = np.zeros(n, PIXEL_DTYPE)
        self.lib.ref_ras_interpolate(_p(a), _p(b), _p(out), n)
        return out

    def pixel_shader(self, pixel, color, normal):
        p = np.atleast_1d(np.array(pixel, PIXEL_DTYPE))
        c = np.ascontiguousarray(color, np.float32)
        nrm = np.ascontiguousarray(normal, np.float32)
        out = np.zeros(3, np.float32)
        foc = C.c_float(0)
        self.lib.ref_ras_pixel_shader(_p(p), _p(c), _p(nrm), _p(out), C.byref(foc))
        return out, foc.value
