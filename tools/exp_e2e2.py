"""Where the host-buffer frame's time goes (config 3): wall clock per call, 40 calls each."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
W, H = 3840, 2160
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
ctx = pkg.Context(W, H); ctx.set_stream(stream.cuda_stream); ctx.set_triangles(pkg.cornell_box())
fp = pkg.default_frame_params(0, W, H); fp.aaEnabled, fp.aaSamples = 1, 4; ctx.set_frame(fp)
surf = torch.empty((H, W), dtype=torch.int32, device=dev)
host = torch.empty((H, W), dtype=torch.int32).pin_memory(); hnp = host.numpy().view(np.uint32)

def timeit(fn, n=40):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def dev_sync():
    ctx.rt_frame_device_async(0, H, surf.data_ptr()); ctx.synchronize()
def copy_only():
    with torch.cuda.stream(stream):
        host.copy_(surf, non_blocking=True)
    stream.synchronize()
def dev_then_copy():
    ctx.rt_frame_device_async(0, H, surf.data_ptr())
    with torch.cuda.stream(stream):
        host.copy_(surf, non_blocking=True)
    stream.synchronize()
print("set_frame only           %.4f ms" % timeit(lambda: ctx.set_frame(fp)))
print("device frame + sync      %.4f ms" % timeit(dev_sync))
print("D2H 33 MB only           %.4f ms" % timeit(copy_only))
print("device frame, then copy  %.4f ms" % timeit(dev_then_copy))
print("rt_frame(host)           %.4f ms" % timeit(lambda: ctx.rt_frame(hnp)))
print("set_frame + rt_frame     %.4f ms" % timeit(lambda: (ctx.set_frame(fp), ctx.rt_frame(hnp))))
ctx.set_option(pkg.capi.OPT_RT_VARIANT, 4)
print("rt_frame(host) variant 4 %.4f ms" % timeit(lambda: ctx.rt_frame(hnp)))
