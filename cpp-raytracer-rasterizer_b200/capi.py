"""ctypes view of include/b2r.h (libb2r.so), used by tests/ and bench.py.

The product is the C ABI; this module is only glue so Python can call it with
numpy host buffers or raw device pointers.  It never falls back to a CPU
implementation: if libb2r.so is missing the import of the library raises.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2R_LIB") or os.path.join(PKG_DIR, "lib", "libb2r.so")  # B2R_LIB: e.g. the bounds-checking debug build

MAX_LIGHTS = 32
RANDOM_POSITIONS = 256

INTERSECTION_DTYPE = np.dtype(
    [("position", np.float32, 3), ("distance", np.float32), ("triangleIndex", np.int32)])
assert INTERSECTION_DTYPE.itemsize == 20
PIXEL_DTYPE = np.dtype([("x", np.int32), ("y", np.int32), ("zinv", np.float32), ("pos3d", np.float32, 3)])
assert PIXEL_DTYPE.itemsize == 24


class Light(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("color", C.c_float * 3), ("intensity", C.c_float)]


class FrameParams(C.Structure):
    """== b2r_frame_params (include/b2r.h)."""
    _fields_ = [
        ("cameraPos", C.c_float * 3),
        ("cameraRot", C.c_float * 9),
        ("focalLength", C.c_float),
        ("numLights", C.c_int32),
        ("lights", Light * MAX_LIGHTS),
        ("randomPositions", C.c_float * (RANDOM_POSITIONS * 3)),
        ("aaEnabled", C.c_int32),
        ("aaSamples", C.c_int32),
        ("softShadowsEnabled", C.c_int32),
        ("softShadowsSamples", C.c_int32),
        ("indirectLight", C.c_float * 3),
        ("dofFocalLength", C.c_float),
        ("currentReflectance", C.c_float * 3),
        ("dofEnabled", C.c_int32),
        ("dofKernelSize", C.c_int32),
        ("backfaceCulling", C.c_int32),
        ("frustumCulling", C.c_int32),
    ]

    def set_camera(self, pos, rot9_colmajor, focal):
        self.cameraPos[:] = [float(v) for v in pos]
        self.cameraRot[:] = [float(v) for v in np.asarray(rot9_colmajor, np.float32).reshape(9)]
        self.focalLength = float(focal)
        return self

    def set_lights(self, lights7):
        lights7 = np.asarray(lights7, np.float32).reshape(-1, 7)
        assert len(lights7) <= MAX_LIGHTS
        self.numLights = len(lights7)
        for i, l in enumerate(lights7):
            self.lights[i].position[:] = [float(v) for v in l[0:3]]
            self.lights[i].color[:] = [float(v) for v in l[3:6]]
            self.lights[i].intensity = float(l[6])
        return self

    def set_random_positions(self, table):
        t = np.ascontiguousarray(table, np.float32).reshape(-1)
        assert t.size == RANDOM_POSITIONS * 3
        C.memmove(self.randomPositions, t.ctypes.data, t.nbytes)
        return self

    def lights_array(self):
        out = np.zeros((self.numLights, 7), np.float32)
        for i in range(self.numLights):
            out[i, 0:3] = self.lights[i].position[:]
            out[i, 3:6] = self.lights[i].color[:]
            out[i, 6] = self.lights[i].intensity
        return out

    def copy(self):
        other = FrameParams()
        C.memmove(C.byref(other), C.byref(self), C.sizeof(FrameParams))
        return other


STAT_NAMES = ["primary_rays", "shadow_rays", "exact_tests", "ras_triangles", "ras_rows", "ras_depth_tests",
              "shadow_rays_evaluated", "reserved7"]

# Every symbol include/b2r.h declares (tests check the library exports each one).
SYMBOLS = [
    "b2r_abi_version", "b2r_create", "b2r_destroy", "b2r_last_error", "b2r_default_frame_params",
    "b2r_set_triangles", "b2r_set_culled", "b2r_set_frame", "b2r_rt_draw", "b2r_ras_draw", "b2r_ras_cull",
    "b2r_resolve_surface", "b2r_resolve_bgr8", "b2r_bmp_payload_bytes", "b2r_write_bmp", "b2r_rt_frame",
    "b2r_ras_frame", "b2r_set_stream", "b2r_get_stream", "b2r_synchronize", "b2r_rt_draw_device_async",
    "b2r_ras_draw_device_async", "b2r_resolve_surface_device_async", "b2r_rt_frame_device_async", "b2r_ras_frame_device_async", "b2r_rt_frame_split_device_async", "b2r_pin_host_buffer", "b2r_unpin_host_buffer", "b2r_launch_count", "b2r_get_stats",
    "b2r_enable_stats", "b2r_set_option", "b2r_measure_fp32_peak", "b2r_selftest_division", "b2r_scene_cornell_box",
    "b2r_scene_tessellate", "b2r_camera_rot_from_yaw", "b2r_orbit_camera", "b2r_jitter_table",
    "b2r_shared_alloc", "b2r_shared_free", "b2r_shared_open", "b2r_shared_close",
    "b2r_resolve_surface_multi_device_async", "b2r_copy_device_async", "b2r_scene_load_stl",
    "b2r_rt_closest_intersection_batch", "b2r_rt_direct_light_batch", "b2r_ras_vertex_shader_batch",
    "b2r_ras_interpolate", "b2r_ras_compute_polygon_rows", "b2r_ras_pixel_shader_batch",
    "b2r_rt_frame_gather_device_async", "b2r_stream_wait_value32", "b2r_rt_frame_part", "b2r_rt_frame_part_async",
    "b2r_rt_frame_bgr8_async", "b2r_ras_frame_part_async",
    "b2r_group_create", "b2r_group_destroy", "b2r_group_size", "b2r_group_ctx", "b2r_group_last_error",
    "b2r_group_set_triangles", "b2r_group_set_frame", "b2r_group_rt_frame", "b2r_group_ras_frame", "b2r_group_rt_frames",
]

_lib = None


def load_library():
    """Load libb2r.so (built by __graft_entry__.build()).  No fallback: raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is the product and there is no CPU fallback. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` first.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, fp = C.c_void_p, C.c_int, C.POINTER(C.c_float)
    lib.b2r_create.argtypes = [C.POINTER(vp), i32, i32, i32]
    lib.b2r_destroy.argtypes = [vp]
    lib.b2r_last_error.argtypes = [vp]
    lib.b2r_last_error.restype = C.c_char_p
    lib.b2r_default_frame_params.argtypes = [C.POINTER(FrameParams), i32, i32, i32]
    lib.b2r_set_triangles.argtypes = [vp, vp, i32, i32]
    lib.b2r_set_culled.argtypes = [vp, vp, i32]
    lib.b2r_set_frame.argtypes = [vp, C.POINTER(FrameParams)]
    lib.b2r_rt_draw.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.b2r_ras_draw.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    lib.b2r_ras_cull.argtypes = [vp, vp]
    lib.b2r_resolve_surface.argtypes = [vp, vp]
    lib.b2r_resolve_bgr8.argtypes = [vp, vp]
    lib.b2r_bmp_payload_bytes.argtypes = [i32, i32]
    lib.b2r_bmp_payload_bytes.restype = C.c_size_t
    lib.b2r_write_bmp.argtypes = [C.c_char_p, vp, i32, i32]
    lib.b2r_rt_frame.argtypes = [vp, vp, vp, vp, vp]
    lib.b2r_ras_frame.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.b2r_set_stream.argtypes = [vp, vp]
    lib.b2r_get_stream.argtypes = [vp]
    lib.b2r_get_stream.restype = vp
    lib.b2r_synchronize.argtypes = [vp]
    lib.b2r_rt_draw_device_async.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.b2r_ras_draw_device_async.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    lib.b2r_resolve_surface_device_async.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.b2r_rt_frame_device_async.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    lib.b2r_ras_frame_device_async.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.b2r_pin_host_buffer.argtypes = [vp, vp, C.c_size_t]
    lib.b2r_unpin_host_buffer.argtypes = [vp, vp]
    lib.b2r_rt_frame_split_device_async.argtypes = [vp, i32, i32, C.POINTER(vp), i32, vp, vp, vp]
    lib.b2r_shared_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), vp]
    lib.b2r_shared_free.argtypes = [vp, vp]
    lib.b2r_shared_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.b2r_shared_close.argtypes = [vp, vp]
    lib.b2r_copy_device_async.argtypes = [vp, vp, vp, C.c_size_t]
    lib.b2r_resolve_surface_multi_device_async.argtypes = [vp, i32, i32, vp, vp, C.POINTER(vp), i32]
    lib.b2r_rt_closest_intersection_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    lib.b2r_rt_direct_light_batch.argtypes = [vp, i32, vp, vp]
    lib.b2r_ras_vertex_shader_batch.argtypes = [vp, i32, vp, vp]
    lib.b2r_ras_interpolate.argtypes = [vp, vp, vp, i32, vp]
    lib.b2r_ras_compute_polygon_rows.argtypes = [vp, vp, vp, vp, i32, C.POINTER(i32)]
    lib.b2r_ras_pixel_shader_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.b2r_launch_count.argtypes = [vp]
    lib.b2r_launch_count.restype = C.c_ulonglong
    lib.b2r_get_stats.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    lib.b2r_enable_stats.argtypes = [vp, i32]
    lib.b2r_set_option.argtypes = [vp, i32, i32]
    lib.b2r_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.b2r_selftest_division.argtypes = [vp, C.c_ulonglong, C.c_uint, C.POINTER(C.c_ulonglong), fp]
    lib.b2r_scene_cornell_box.argtypes = [vp, i32, i32]
    lib.b2r_scene_tessellate.argtypes = [vp, i32, i32, i32, vp, i32]
    lib.b2r_scene_tessellate.restype = C.c_longlong
    lib.b2r_scene_load_stl.argtypes = [C.c_char_p, vp, C.c_longlong, i32]
    lib.b2r_scene_load_stl.restype = C.c_longlong
    lib.b2r_camera_rot_from_yaw.argtypes = [C.c_float, C.c_float, fp]
    lib.b2r_orbit_camera.argtypes = [i32, i32, C.c_float, fp, fp]
    lib.b2r_jitter_table.argtypes = [C.c_uint, fp, fp]
    lib.b2r_rt_frame_gather_device_async.argtypes = [vp, i32, i32, vp, vp]
    lib.b2r_stream_wait_value32.argtypes = [vp, vp, C.c_uint]
    lib.b2r_rt_frame_part.argtypes = [vp, i32, i32, vp]
    lib.b2r_rt_frame_part_async.argtypes = [vp, i32, i32, vp]
    lib.b2r_rt_frame_bgr8_async.argtypes = [vp, vp]
    lib.b2r_ras_frame_part_async.argtypes = [vp, i32, i32, vp]
    lib.b2r_group_create.argtypes = [C.POINTER(vp), C.POINTER(i32), i32, i32, i32]
    lib.b2r_group_destroy.argtypes = [vp]
    lib.b2r_group_size.argtypes = [vp]
    lib.b2r_group_ctx.argtypes = [vp, i32]
    lib.b2r_group_ctx.restype = vp
    lib.b2r_group_last_error.argtypes = [vp]
    lib.b2r_group_last_error.restype = C.c_char_p
    lib.b2r_group_set_triangles.argtypes = [vp, vp, i32, i32]
    lib.b2r_group_set_frame.argtypes = [vp, C.POINTER(FrameParams)]
    lib.b2r_group_rt_frame.argtypes = [vp, vp]
    lib.b2r_group_ras_frame.argtypes = [vp, vp]
    lib.b2r_group_rt_frames.argtypes = [vp, C.POINTER(FrameParams), i32, vp, C.c_char_p]
    assert lib.b2r_abi_version() == 1
    _lib = lib
    return lib


# option ids (b2r_set_option)
OPT_RT_FILTER = 0        # 1 (default): conservative FMA filter in front of the exact test; 0: exact test on every pair
OPT_RT_VARIANT = 1       # kernel variant selector
OPT_RAS_VARIANT = 2
OPT_DOF_VARIANT = 3


class B2RError(RuntimeError):
    pass


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One b2r_ctx (one GPU, one screen size)."""

    def __init__(self, width, height, device=0):
        self.lib = load_library()
        self.w, self.h = int(width), int(height)
        h = C.c_void_p()
        rc = self.lib.b2r_create(C.byref(h), device, self.w, self.h)
        if rc != 0:
            raise B2RError(f"b2r_create -> {rc}: {self.lib.b2r_last_error(None).decode()}")
        self.handle = h
        self.ntris = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b2r_destroy(self.handle)
            self.handle = None

    __del__ = close

    def _chk(self, rc):
        if rc != 0:
            raise B2RError(f"b2r error {rc}: {self.lib.b2r_last_error(self.handle).decode()}")

    # scene / frame --------------------------------------------------------------
    def set_triangles(self, tris, stride=None):
        """tris: (n,15) float32 (raytracer Triangle, stride 60) or raw bytes with stride 60/64."""
        if isinstance(tris, np.ndarray) and tris.dtype == np.float32:
            tris = np.ascontiguousarray(tris).reshape(-1, 15)
            n, stride = len(tris), 60
            buf = tris
        else:
            buf = np.frombuffer(tris, np.uint8)
            n = len(buf) // stride
        self._chk(self.lib.b2r_set_triangles(self.handle, _ptr(buf), n, stride))
        self.ntris = n

    def set_culled(self, mask):
        m = np.ascontiguousarray(mask, np.uint8)
        self._chk(self.lib.b2r_set_culled(self.handle, _ptr(m), len(m)))

    def set_frame(self, fp):
        self._chk(self.lib.b2r_set_frame(self.handle, C.byref(fp)))

    def set_option(self, opt, value):
        self._chk(self.lib.b2r_set_option(self.handle, opt, value))

    def enable_stats(self, on=True):
        self._chk(self.lib.b2r_enable_stats(self.handle, int(on)))

    def stats(self):
        out = (C.c_ulonglong * 8)()
        self._chk(self.lib.b2r_get_stats(self.handle, out))
        return dict(zip(STAT_NAMES, [int(v) for v in out]))

    def launch_count(self):
        return int(self.lib.b2r_launch_count(self.handle))

    def synchronize(self):
        self._chk(self.lib.b2r_synchronize(self.handle))

    # host-buffer entry points ---------------------------------------------------
    def rt_draw(self, y0=0, y1=None, colours=True, closest=True, focal=True):
        y1 = self.h if y1 is None else y1
        col = np.zeros((self.h, self.w, 3), np.float32) if colours else None
        clo = np.zeros((self.h, self.w), INTERSECTION_DTYPE) if closest else None
        foc = np.zeros((self.h, self.w), np.float32) if focal else None
        self._chk(self.lib.b2r_rt_draw(self.handle, y0, y1, _ptr(col), _ptr(clo), _ptr(foc)))
        return dict(pixelColours=col, closest=clo, focalDistances=foc)

    def ras_draw(self, y0=0, y1=None):
        y1 = self.h if y1 is None else y1
        dep = np.zeros((self.h, self.w), np.float32)
        col = np.zeros((self.h, self.w, 3), np.float32)
        foc = np.zeros((self.h, self.w), np.float32)
        win = np.full((self.h, self.w), -1, np.int32)
        self._chk(self.lib.b2r_ras_draw(self.handle, y0, y1, _ptr(dep), _ptr(col), _ptr(foc), _ptr(win)))
        return dict(depthBuffer=dep, pixelColours=col, focalDistances=foc, winner=win)

    def ras_cull(self):
        out = np.zeros(self.ntris, np.uint8)
        self._chk(self.lib.b2r_ras_cull(self.handle, _ptr(out)))
        return out

    def resolve_surface(self):
        s = np.zeros((self.h, self.w), np.uint32)
        self._chk(self.lib.b2r_resolve_surface(self.handle, _ptr(s)))
        return s

    def resolve_bgr8(self):
        b = np.zeros(self.lib.b2r_bmp_payload_bytes(self.w, self.h), np.uint8)
        self._chk(self.lib.b2r_resolve_bgr8(self.handle, _ptr(b)))
        return b

    def rt_frame(self, surface=None):
        """Draw() drop-in: trace + resolve, copying out only the 32-bit surface."""
        if surface is None:
            surface = np.zeros((self.h, self.w), np.uint32)
        self._chk(self.lib.b2r_rt_frame(self.handle, _ptr(surface), None, None, None))
        return surface

    def ras_frame(self, surface=None):
        if surface is None:
            surface = np.zeros((self.h, self.w), np.uint32)
        self._chk(self.lib.b2r_ras_frame(self.handle, _ptr(surface), None, None, None, None))
        return surface

    # device-pointer entry points (ints = CUDA device addresses) ------------------
    def set_stream(self, cuda_stream):
        self._chk(self.lib.b2r_set_stream(self.handle, C.c_void_p(cuda_stream)))

    def rt_draw_device_async(self, y0, y1, d_colours=0, d_closest=0, d_focal=0):
        self._chk(self.lib.b2r_rt_draw_device_async(self.handle, y0, y1, C.c_void_p(d_colours),
                                                    C.c_void_p(d_closest), C.c_void_p(d_focal)))

    def rt_frame_device_async(self, y0, y1, d_surface, d_colours=0, d_closest=0, d_focal=0):
        self._chk(self.lib.b2r_rt_frame_device_async(self.handle, y0, y1, C.c_void_p(d_surface), C.c_void_p(d_colours),
                                                     C.c_void_p(d_closest), C.c_void_p(d_focal)))

    def ras_draw_device_async(self, y0, y1, d_depth=0, d_colours=0, d_focal=0, d_winner=0):
        self._chk(self.lib.b2r_ras_draw_device_async(self.handle, y0, y1, C.c_void_p(d_depth),
                                                     C.c_void_p(d_colours), C.c_void_p(d_focal),
                                                     C.c_void_p(d_winner)))

    def pin_host_buffer(self, array):
        """Page-lock a numpy array that will receive frames (cudaHostRegister)."""
        self._chk(self.lib.b2r_pin_host_buffer(self.handle, _ptr(array), array.nbytes))

    def unpin_host_buffer(self, array):
        self._chk(self.lib.b2r_unpin_host_buffer(self.handle, _ptr(array)))

    def rt_frame_split_device_async(self, part, nparts, surfaces, d_colours=0, d_closest=0, d_focal=0):
        """Tile rows part, part+nparts, ... of the frame, every pixel stored into all `surfaces` (device addresses)."""
        arr = (C.c_void_p * len(surfaces))(*surfaces)
        self._chk(self.lib.b2r_rt_frame_split_device_async(self.handle, part, nparts, arr, len(surfaces),
                                                           C.c_void_p(d_colours), C.c_void_p(d_closest), C.c_void_p(d_focal)))

    def rt_frame_gather_device_async(self, part, nparts, d_root_surface, d_arrive=0):
        """Tile rows part, part+nparts, ... of the frame into ONE surface (the root's); the launch's last thread block
        adds 1 to *d_arrive when every pixel has landed."""
        self._chk(self.lib.b2r_rt_frame_gather_device_async(self.handle, part, nparts, C.c_void_p(d_root_surface),
                                                            C.c_void_p(d_arrive)))

    def stream_wait_value32(self, d_word, value):
        self._chk(self.lib.b2r_stream_wait_value32(self.handle, C.c_void_p(d_word), value))

    def rt_frame_part(self, part, nparts, surface):
        """Host side of a split frame: this GPU's tile rows into the caller's full-frame host surface (synchronous)."""
        self._chk(self.lib.b2r_rt_frame_part(self.handle, part, nparts, _ptr(surface)))

    def rt_frame_part_async(self, part, nparts, surface):
        self._chk(self.lib.b2r_rt_frame_part_async(self.handle, part, nparts, _ptr(surface)))

    def rt_frame_bgr8_async(self, bgr):
        self._chk(self.lib.b2r_rt_frame_bgr8_async(self.handle, _ptr(bgr)))

    def ras_frame_part_async(self, y0, y1, surface):
        self._chk(self.lib.b2r_ras_frame_part_async(self.handle, y0, y1, _ptr(surface)))

    def ras_frame_device_async(self, y0, y1, d_surface, d_depth=0, d_colours=0, d_focal=0, d_winner=0):
        self._chk(self.lib.b2r_ras_frame_device_async(self.handle, y0, y1, C.c_void_p(d_surface), C.c_void_p(d_depth),
                                                      C.c_void_p(d_colours), C.c_void_p(d_focal), C.c_void_p(d_winner)))

    def resolve_surface_device_async(self, y0, y1, d_colours, d_focal, d_surface):
        self._chk(self.lib.b2r_resolve_surface_device_async(self.handle, y0, y1, C.c_void_p(d_colours),
                                                            C.c_void_p(d_focal), C.c_void_p(d_surface)))

    # fused band exchange (one process per GPU, peer-mapped surfaces) -----------------------------
    def shared_alloc(self, nbytes):
        """cudaMalloc on this GPU; returns (device address, 64-byte IPC handle)."""
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        self._chk(self.lib.b2r_shared_alloc(self.handle, nbytes, C.byref(p), h))
        return p.value, bytes(h)

    def shared_free(self, addr):
        self._chk(self.lib.b2r_shared_free(self.handle, C.c_void_p(addr)))

    def shared_open(self, handle_bytes):
        p = C.c_void_p()
        h = (C.c_ubyte * 64).from_buffer_copy(handle_bytes)
        self._chk(self.lib.b2r_shared_open(self.handle, h, C.byref(p)))
        return p.value

    def shared_close(self, addr):
        self._chk(self.lib.b2r_shared_close(self.handle, C.c_void_p(addr)))

    def copy_device_async(self, d_dst, d_src, nbytes):
        self._chk(self.lib.b2r_copy_device_async(self.handle, C.c_void_p(d_dst), C.c_void_p(d_src), nbytes))

    def resolve_surface_multi_device_async(self, y0, y1, d_colours, d_focal, d_surfaces):
        arr = (C.c_void_p * len(d_surfaces))(*d_surfaces)
        self._chk(self.lib.b2r_resolve_surface_multi_device_async(self.handle, y0, y1, C.c_void_p(d_colours),
                                                                  C.c_void_p(d_focal), arr, len(d_surfaces)))

    # sub-stage entry points (the reference's callee functions, batched) ---------------------------
    def closest_intersection(self, starts, dirs, closest=None, is_light=None):
        starts = np.ascontiguousarray(starts, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(starts)
        io = np.zeros(n, INTERSECTION_DTYPE)
        if closest is None:
            io["distance"] = np.finfo(np.float32).max
            io["triangleIndex"] = -1
        else:
            io[:] = closest
        il = None if is_light is None else np.ascontiguousarray(is_light, np.int32)
        hit = np.zeros(n, np.int32)
        foc = np.zeros(n, np.float32)
        self._chk(self.lib.b2r_rt_closest_intersection_batch(self.handle, n, _ptr(starts), _ptr(dirs), _ptr(il), _ptr(io),
                                                             _ptr(hit), _ptr(foc)))
        return hit.astype(bool), io, foc

    def direct_light(self, hits):
        h = np.ascontiguousarray(hits, INTERSECTION_DTYPE).reshape(-1)
        out = np.zeros((len(h), 3), np.float32)
        self._chk(self.lib.b2r_rt_direct_light_batch(self.handle, len(h), _ptr(h), _ptr(out)))
        return out

    def vertex_shader(self, verts):
        v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        out = np.zeros(len(v), PIXEL_DTYPE)
        self._chk(self.lib.b2r_ras_vertex_shader_batch(self.handle, len(v), _ptr(v), _ptr(out)))
        return out

    def interpolate(self, a, b, n):
        a = np.atleast_1d(np.array(a, PIXEL_DTYPE))
        b = np.atleast_1d(np.array(b, PIXEL_DTYPE))
        out = np.zeros(n, PIXEL_DTYPE)
        self._chk(self.lib.b2r_ras_interpolate(self.handle, _ptr(a), _ptr(b), n, _ptr(out)))
        return out

    def compute_polygon_rows(self, vertex_pixels, max_rows=8192):
        vp = np.ascontiguousarray(vertex_pixels, PIXEL_DTYPE)
        left = np.zeros(max_rows, PIXEL_DTYPE)
        right = np.zeros(max_rows, PIXEL_DTYPE)
        rows = C.c_int(0)
        self._chk(self.lib.b2r_ras_compute_polygon_rows(self.handle, _ptr(vp), _ptr(left), _ptr(right), max_rows,
                                                        C.byref(rows)))
        return left[:rows.value].copy(), right[:rows.value].copy()

    def pixel_shader(self, pixels, colors, normals):
        p = np.ascontiguousarray(pixels, PIXEL_DTYPE).reshape(-1)
        col = np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        out = np.zeros((len(p), 3), np.float32)
        foc = np.zeros(len(p), np.float32)
        self._chk(self.lib.b2r_ras_pixel_shader_batch(self.handle, len(p), _ptr(p), _ptr(col), _ptr(nrm), _ptr(out), _ptr(foc)))
        return out, foc

    def selftest_division(self, n, seed=1):
        """(mismatches, [a, b, expected, got] of the first one) of the shared-reciprocal division against div.rn.f32."""
        bad = C.c_ulonglong()
        first = (C.c_float * 4)()
        self._chk(self.lib.b2r_selftest_division(self.handle, n, seed, C.byref(bad), first))
        return bad.value, list(first)

    def measure_fp32_peak(self):
        t, s = C.c_double(), C.c_double()
        self._chk(self.lib.b2r_measure_fp32_peak(self.handle, C.byref(t), C.byref(s)))
        return t.value, s.value


class Group:
    """b2r_group: several GPUs of one box behind one Draw(), in one process."""

    def __init__(self, width, height, devices):
        self.lib = load_library()
        self.w, self.h = int(width), int(height)
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.b2r_group_create(C.byref(h), devs, len(devices), self.w, self.h)
        if rc != 0:
            raise B2RError(f"b2r_group_create -> {rc}: {self.lib.b2r_group_last_error(None).decode()}")
        self.handle = h
        self.n = len(devices)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b2r_group_destroy(self.handle)
            self.handle = None

    __del__ = close

    def _chk(self, rc):
        if rc != 0:
            raise B2RError(f"b2r group error {rc}: {self.lib.b2r_group_last_error(self.handle).decode()}")

    def set_triangles(self, tris):
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 15)
        self._chk(self.lib.b2r_group_set_triangles(self.handle, _ptr(tris), len(tris), 60))
        self.ntris = len(tris)

    def set_frame(self, fp):
        self._chk(self.lib.b2r_group_set_frame(self.handle, C.byref(fp)))

    def member_set_option(self, i, opt, value):
        rc = self.lib.b2r_set_option(self.lib.b2r_group_ctx(self.handle, i), opt, value)
        if rc != 0:
            raise B2RError(f"b2r_set_option -> {rc}")

    def member_ras_cull(self, i):
        rc = self.lib.b2r_ras_cull(self.lib.b2r_group_ctx(self.handle, i), None)
        if rc != 0:
            raise B2RError(f"b2r_ras_cull -> {rc}")

    def launch_count(self):
        return sum(int(self.lib.b2r_launch_count(self.lib.b2r_group_ctx(self.handle, i))) for i in range(self.n))

    def rt_frame(self, surface=None):
        if surface is None:
            surface = np.zeros((self.h, self.w), np.uint32)
        self._chk(self.lib.b2r_group_rt_frame(self.handle, _ptr(surface)))
        return surface

    def ras_frame(self, surface=None):
        if surface is None:
            surface = np.zeros((self.h, self.w), np.uint32)
        self._chk(self.lib.b2r_group_ras_frame(self.handle, _ptr(surface)))
        return surface

    def rt_frames(self, frames, surfaces=None, bmp_pattern=None):
        """frames: list of FrameParams; surfaces: (n, h, w) uint32 array or None; bmp_pattern: e.g. '/tmp/f_%04d.bmp'."""
        arr = (FrameParams * len(frames))(*frames)
        self._chk(self.lib.b2r_group_rt_frames(self.handle, arr, len(frames), _ptr(surfaces),
                                               bmp_pattern.encode() if bmp_pattern else None))
        return surfaces


def write_bmp(path, bgr_payload, w, h):
    lib = load_library()
    b = np.ascontiguousarray(bgr_payload, np.uint8)
    rc = lib.b2r_write_bmp(path.encode(), _ptr(b), w, h)
    if rc != 0:
        raise B2RError(f"b2r_write_bmp -> {rc}")


# ---- host-side scene helpers (C++ in csrc/scenes.cpp; no GPU needed) ------------
def default_frame_params(which, w, h):
    """which: 0 raytracer defaults, 1 rasteriser defaults."""
    fp = FrameParams()
    rc = load_library().b2r_default_frame_params(C.byref(fp), which, w, h)
    if rc != 0:
        raise B2RError(f"b2r_default_frame_params -> {rc}")
    return fp


def cornell_box():
    """LoadTestModel (TestModel.h:51-192): (30,15) float32 raytracer-layout triangles."""
    out = np.zeros((30, 15), np.float32)
    n = load_library().b2r_scene_cornell_box(_ptr(out), 30, 60)
    assert n == 30
    return out


def tessellate(tris15, k):
    """split every triangle into k*k; parent order kept."""
    tris15 = np.ascontiguousarray(tris15, np.float32).reshape(-1, 15)
    n = len(tris15) * k * k
    out = np.zeros((n, 15), np.float32)
    got = load_library().b2r_scene_tessellate(_ptr(tris15), len(tris15), 60, k, _ptr(out), 60)
    assert got == n, (got, n)
    return out


def load_stl(path):
    """ASCII STL with the reference loader's semantics (LoadSTL.cpp:17-81): (n,15) float32 triangles."""
    lib = load_library()
    n = lib.b2r_scene_load_stl(path.encode(), None, 0, 60)
    if n < 0:
        raise B2RError(f"b2r_scene_load_stl -> {n}")
    out = np.zeros((n, 15), np.float32)
    got = lib.b2r_scene_load_stl(path.encode(), _ptr(out), n, 60)
    assert got == n
    return out


def camera_rot_from_yaw(yaw, rot11):
    """cameraRot as Update() builds it (raytracer.cpp:377-382); rot11 = 1.0 (raytracer) or 1.01 (rasteriser)."""
    out = np.zeros(9, np.float32)
    load_library().b2r_camera_rot_from_yaw(C.c_float(yaw), C.c_float(rot11), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def orbit_camera(frame, nframes, radius=2.0):
    """yaw = frame*2pi/nframes, camera on a circle looking at the origin."""
    pos = np.zeros(3, np.float32)
    rot = np.zeros(9, np.float32)
    fpp = C.POINTER(C.c_float)
    load_library().b2r_orbit_camera(frame, nframes, C.c_float(radius), pos.ctypes.data_as(fpp), rot.ctypes.data_as(fpp))
    return pos, rot


def jitter_table(seed, light_pos):
    """The soft-shadow jitter table AddLight() builds (raytracer.cpp:186-190) from glibc rand()."""
    lp = np.ascontiguousarray(light_pos, np.float32)
    out = np.zeros((256, 3), np.float32)
    fpp = C.POINTER(C.c_float)
    load_library().b2r_jitter_table(seed, lp.ctypes.data_as(fpp), out.ctypes.data_as(fpp))
    return out
