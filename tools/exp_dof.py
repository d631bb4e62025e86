"""Timing of the depth-of-field resolve (CalculateDOF with DOF_ENABLED) at 4K; development aid."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
W, H = 3840, 2160
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
ctx = pkg.Context(W, H); ctx.set_stream(stream.cuda_stream); ctx.set_triangles(pkg.cornell_box())
col = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
foc = torch.empty((H, W), dtype=torch.float32, device=dev)
surf = torch.empty((H, W), dtype=torch.int32, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for _ in range(n): fn()
        ev[1].record(stream)
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n
for dof in (0, 1):
    fp = pkg.default_frame_params(0, W, H); fp.dofEnabled = dof; ctx.set_frame(fp)
    ctx.rt_draw_device_async(0, H, col.data_ptr(), 0, foc.data_ptr()); ctx.synchronize()
    print(f"dof={dof}: resolve_surface {timeit(lambda: ctx.resolve_surface_device_async(0, H, col.data_ptr(), foc.data_ptr(), surf.data_ptr())):.3f} ms")
