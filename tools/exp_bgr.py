import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
W, H = 3840, 2160
ctx = pkg.Context(W, H); ctx.set_triangles(pkg.cornell_box())
ctx.set_frame(pkg.default_frame_params(0, W, H))
pitch = (3 * W + 3) // 4 * 4
host = torch.empty(pitch * H, dtype=torch.uint8).pin_memory(); hnp = host.numpy()
for _ in range(3):
    ctx.rt_frame_bgr8_async(hnp); ctx.synchronize()
print("ok")
