"""One part of the config-3 frame split N ways (gather form, local surface, no arrival word), timed alone on one GPU:
what a launch costs beyond 1/N of the frame."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
W, H = 3840, 2160
ctx = pkg.Context(W, H)
ctx.set_stream(stream.cuda_stream)
ctx.set_triangles(pkg.cornell_box())
fp = pkg.default_frame_params(0, W, H)
fp.aaEnabled, fp.aaSamples = 1, 4
ctx.set_frame(fp)
surf = torch.zeros((H, W), dtype=torch.int32, device=dev)
with torch.cuda.stream(stream):
    for n in (1, 2, 4, 8, 16):
        for part in sorted(set((0, n - 1))):
            for _ in range(3):
                ctx.rt_frame_gather_device_async(part, n, surf.data_ptr())
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(10):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.rt_frame_gather_device_async(part, n, surf.data_ptr())
                e1.record(stream)
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            print(f"order={os.environ.get('B2R_EXP_ORDER', '0')} part {part} of {n}: {tot / 10 * 1e3:7.1f} us  (frame/{n} = {772.0 / n:6.1f})", flush=True)
