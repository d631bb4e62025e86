"""Device timings of the raytracer configs (CUDA events, L2 flushed between frames); development aid, not the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()


def main():
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tris = pkg.cornell_box()
    W, H = 3840, 2160
    for name, aa, soft in [("config3 AA4x4", 4, 0), ("config3b soft16", 0, 1), ("1spp", 0, 0), ("AA2 soft16", 2, 1)]:
        ctx = pkg.Context(W, H)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_triangles(tris)
        fp = pkg.default_frame_params(0, W, H)
        fp.aaEnabled, fp.aaSamples, fp.softShadowsEnabled = int(aa > 0), max(aa, 1), soft
        fp.set_random_positions(pkg.jitter_table(1, [0, -0.5, -0.7]))
        ctx.set_frame(fp)
        surf = torch.empty((H, W), dtype=torch.int32, device=dev)
        fn = lambda: ctx.rt_frame_device_async(0, H, surf.data_ptr())
        with torch.cuda.stream(stream):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            tot, n = 0.0, 10
            for _ in range(n):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
        print(f"rt {name}: {tot / n:.4f} ms", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
