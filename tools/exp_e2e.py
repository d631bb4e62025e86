import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
W,H=3840,2160
dev=torch.device("cuda:0")
stream=torch.cuda.Stream()
ctx=pkg.Context(W,H); ctx.set_stream(stream.cuda_stream); ctx.set_triangles(pkg.cornell_box())
fp=pkg.default_frame_params(0,W,H); fp.aaEnabled, fp.aaSamples=1,4; ctx.set_frame(fp)
surf=torch.empty((H,W),dtype=torch.int32,device=dev)
def timeit(fn,n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/n*1e3
print("full device (B2R_EXP_BANDS=%s):" % os.environ.get("B2R_EXP_BANDS"), timeit(lambda: ctx.rt_frame_device_async(0,H,surf.data_ptr())))
host=torch.empty((H,W),dtype=torch.int32).pin_memory(); hnp=host.numpy().view(np.uint32)
for v in (0,4,0,4):
    ctx.set_option(pkg.capi.OPT_RT_VARIANT, v)
    print("host rt_frame variant",v, timeit(lambda: ctx.rt_frame(hnp)))
