"""Launch floor of the trace kernel: time of bands of 8 .. 2160 rows at the top of the config-3 frame, and at 1 spp."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
W, H = 3840, 2160
for aa in (4, 1):
    ctx = pkg.Context(W, H)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_triangles(pkg.cornell_box())
    fp = pkg.default_frame_params(0, W, H)
    fp.aaEnabled, fp.aaSamples = int(aa > 1), aa
    ctx.set_frame(fp)
    surf = torch.empty((H, W), dtype=torch.int32, device=dev)
    with torch.cuda.stream(stream):
        for rows in (8, 16, 32, 64, 136, 272, 544, 1080, 2160):
            for _ in range(3):
                ctx.rt_frame_device_async(0, rows, surf.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20):
                ctx.rt_frame_device_async(0, rows, surf.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            print(f"aa {aa} rows 0-{rows:4d}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us", flush=True)
    ctx.close()
