"""Short fixed command for ncu captures: a few launches of the hot kernels (no timing printed is a bench value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
W, H = 3840, 2160
which = sys.argv[1] if len(sys.argv) > 1 else "rt"
dev = torch.device("cuda:0")
tris = pkg.cornell_box()
if which == "rt":
    ctx = pkg.Context(W, H)
    ctx.set_triangles(tris)
    fp = pkg.default_frame_params(0, W, H)
    fp.aaEnabled, fp.aaSamples = 1, 4
    ctx.set_frame(fp)
    col = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    surf = torch.empty((H, W), dtype=torch.int32, device=dev)
    for _ in range(3):
        ctx.rt_frame_device_async(0, H, surf.data_ptr(), col.data_ptr())
    ctx.synchronize()
elif which == "rtsurf":  # the bench's step at N = 1: surface only
    ctx = pkg.Context(W, H)
    ctx.set_triangles(tris)
    fp = pkg.default_frame_params(0, W, H)
    fp.aaEnabled, fp.aaSamples = 1, 4
    ctx.set_frame(fp)
    surf = torch.empty((H, W), dtype=torch.int32, device=dev)
    for _ in range(3):
        ctx.rt_frame_device_async(0, H, surf.data_ptr())
    ctx.synchronize()
elif which == "rt1":  # config 5's frame: 1 sample per pixel, surface only
    ctx = pkg.Context(W, H)
    ctx.set_triangles(tris)
    ctx.set_frame(pkg.default_frame_params(0, W, H))
    surf = torch.empty((H, W), dtype=torch.int32, device=dev)
    for _ in range(3):
        ctx.rt_frame_device_async(0, H, surf.data_ptr())
    ctx.synchronize()
elif which == "rtsoft":  # config 3b: 1 sample per pixel, 16 jittered light samples
    ctx = pkg.Context(W, H)
    ctx.set_triangles(tris)
    fp = pkg.default_frame_params(0, W, H)
    fp.softShadowsEnabled = 1
    fp.set_random_positions(pkg.jitter_table(1, [0, -0.5, -0.7]))
    ctx.set_frame(fp)
    surf = torch.empty((H, W), dtype=torch.int32, device=dev)
    for _ in range(3):
        ctx.rt_frame_device_async(0, H, surf.data_ptr())
    ctx.synchronize()
elif which.startswith("rtpart"):  # one part of the frame split N ways (gather form, local surface): kernel time vs 1/N
    n = int(which[6:])
    ctx = pkg.Context(W, H)
    ctx.set_triangles(tris)
    fp = pkg.default_frame_params(0, W, H)
    fp.aaEnabled, fp.aaSamples = 1, 4
    ctx.set_frame(fp)
    surf = torch.zeros((H * W + 64,), dtype=torch.int32, device=dev)
    for part in (0, 0, 1, n - 1):
        ctx.rt_frame_gather_device_async(part, n, surf.data_ptr(), surf.data_ptr() + H * W * 4)
    ctx.synchronize()
elif which == "dof":
    ctx = pkg.Context(W, H)
    ctx.set_triangles(tris)
    fp = pkg.default_frame_params(0, W, H)
    fp.dofEnabled = 1
    ctx.set_frame(fp)
    col = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    foc = torch.empty((H, W), dtype=torch.float32, device=dev)
    surf = torch.empty((H, W), dtype=torch.int32, device=dev)
    ctx.rt_draw_device_async(0, H, col.data_ptr(), 0, foc.data_ptr())
    for _ in range(3):
        ctx.resolve_surface_device_async(0, H, col.data_ptr(), foc.data_ptr(), surf.data_ptr())
    ctx.synchronize()
elif which == "ras30":
    ctx = pkg.Context(W, H)
    ctx.set_triangles(tris)
    ctx.set_frame(pkg.default_frame_params(1, W, H))
    ctx.ras_cull()
    dep = torch.empty((H, W), dtype=torch.float32, device=dev)
    col = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    for _ in range(3):
        ctx.ras_draw_device_async(0, H, dep.data_ptr(), col.data_ptr())
    ctx.synchronize()
else:  # "ras" (sort-last, default) or "rastiles" (B2R_OPT_RAS_VARIANT = 2)
    big = pkg.tessellate(tris, 183)
    ctx = pkg.Context(W, H)
    ctx.set_option(pkg.capi.OPT_RAS_VARIANT, 2 if which == "rastiles" else 0)
    ctx.set_triangles(big)
    ctx.set_frame(pkg.default_frame_params(1, W, H))
    ctx.ras_cull()
    dep = torch.empty((H, W), dtype=torch.float32, device=dev)
    col = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    for _ in range(3):
        ctx.ras_draw_device_async(0, H, dep.data_ptr(), col.data_ptr())
    ctx.synchronize()
print("profile_run ok", which)
