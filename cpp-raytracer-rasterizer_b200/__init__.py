"""B200-native render hot paths of ArchDD/CPP-Raytracer-Rasterizer.

The product is csrc/ (hand-written sm_100a CUDA behind the C ABI of
include/b2r.h) and host/ (C++ drop-ins for the reference's Draw()).  This
Python package is only the ctypes view used by tests/ and bench.py.  The
directory name is not a Python identifier, so import it through
`__graft_entry__.load_package()` (module name `cpp_raytracer_rasterizer_b200`).
"""
from . import capi  # noqa: F401
from .capi import (B2RError, Context, FrameParams, Group, camera_rot_from_yaw, cornell_box,  # noqa: F401
                   default_frame_params, jitter_table, load_library, load_stl, orbit_camera, tessellate, write_bmp)


def __getattr__(name):
    # `parallel` needs torch.distributed; import it only when asked for
    if name == "parallel":
        import importlib
        return importlib.import_module(__name__ + ".parallel")
    raise AttributeError(name)
