"""Device timings of the rasteriser configs (CUDA events, L2 flushed between frames); development aid, not the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()


VARIANTS = tuple(int(v) for v in os.environ.get('RAS_VARIANTS', '0,4').split(','))
CASES = [tuple(int(v) for v in c.split('x')) for c in os.environ.get('RAS_CASES', '3840x2160x183').split(',')]  # WxHxtessellation


def main():
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tris = pkg.cornell_box()
    for (w, h, k) in CASES:
        t = pkg.tessellate(tris, k) if k > 1 else tris
        for variant in VARIANTS:
            ctx = pkg.Context(w, h)
            ctx.set_option(pkg.capi.OPT_RAS_VARIANT, variant)
            ctx.set_stream(stream.cuda_stream)
            ctx.set_triangles(t)
            ctx.set_frame(pkg.default_frame_params(1, w, h))
            ctx.ras_cull()
            dep = torch.empty((h, w), dtype=torch.float32, device=dev)
            col = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
            fn = lambda: ctx.ras_draw_device_async(0, h, dep.data_ptr(), col.data_ptr())
            with torch.cuda.stream(stream):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                tot, best = 0.0, 1e9
                n = 10
                for _ in range(n):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    fn()
                    e1.record(stream)
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                    tot += ms
                    best = min(best, ms)
            ms = tot / n
            ab = 64 * len(t) + 16 * w * h
            print(f"ras variant {variant} {w}x{h} tris={len(t)}: mean {ms:.4f} ms best {best:.4f} ms -> {1e3/ms:.1f} frames/s, "
                  f"algorithmic {ab/ms/1e6:.1f} GB/s = {ab/ms/1e6/6459:.3f} of 6459", flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
