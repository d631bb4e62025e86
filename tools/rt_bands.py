"""Trace time of each sixteenth of the config-3 frame (CUDA events, device-resident): the cost profile over the rows,
which decides how much of the device-to-host copy of a host-buffer frame can hide behind the tracing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
W, H = 3840, 2160
ctx = pkg.Context(W, H)
ctx.set_stream(stream.cuda_stream)
ctx.set_triangles(pkg.cornell_box())
fp = pkg.default_frame_params(0, W, H)
fp.aaEnabled, fp.aaSamples = 1, 4
ctx.set_frame(fp)
surf = torch.empty((H, W), dtype=torch.int32, device=dev)
nb = 16
per = (H // 8 + nb - 1) // nb * 8
out = []
with torch.cuda.stream(stream):
    for y0 in range(0, H, per):
        y1 = min(H, y0 + per)
        for _ in range(3):
            ctx.rt_frame_device_async(y0, y1, surf.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            ctx.rt_frame_device_async(y0, y1, surf.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        out.append((y0, y1, e0.elapsed_time(e1) / 10))
for y0, y1, ms in out:
    print(f"rows {y0:4d}-{y1:4d}: {ms*1e3:7.1f} us")
print(f"sum {sum(o[2] for o in out)*1e3:.1f} us")
