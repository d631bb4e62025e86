"""Quick device timings (CUDA events through torch) for the main configs; development aid, not the bench."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()

def timeit(fn, stream, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for _ in range(n): fn()
        ev[1].record(stream)
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n

def main():
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    tris = pkg.cornell_box()
    print("fp32 peak TFLOP/s:", pkg.Context(8, 8).measure_fp32_peak())
    for (w, h, aa, soft) in [(500, 500, 0, 0), (3840, 2160, 0, 0), (3840, 2160, 4, 0), (3840, 2160, 0, 1)]:
        ctx = pkg.Context(w, h)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_triangles(tris)
        fp = pkg.default_frame_params(0, w, h)
        fp.aaEnabled, fp.aaSamples, fp.softShadowsEnabled = int(aa > 0), max(aa, 1), soft
        fp.set_random_positions(pkg.jitter_table(1, [0, -0.5, -0.7]))
        ctx.set_frame(fp)
        col = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
        for filt in (1, 0):
            ctx.set_option(pkg.capi.OPT_RT_FILTER, filt)
            ms = timeit(lambda: ctx.rt_draw_device_async(0, h, col.data_ptr()), stream, n=5 if aa else 10)
            ctx.enable_stats(True); ctx.rt_draw_device_async(0, h, col.data_ptr()); st = ctx.stats(); ctx.enable_stats(False)
            rays = st["primary_rays"] + st["shadow_rays"]
            print(f"rt {w}x{h} aa={aa} soft={soft} filter={filt}: {ms:.3f} ms  rays={rays} -> {rays/ms/1e3:.1f} Mrays/s  exact_tests={st['exact_tests']} ({st['exact_tests']/(rays*30):.3f} of pairs)")
        ctx.close()
    for (w, h, k) in [(500, 500, 1), (3840, 2160, 1), (3840, 2160, 183)]:
        t = pkg.tessellate(tris, k) if k > 1 else tris
        ctx = pkg.Context(w, h)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_triangles(t)
        ctx.set_frame(pkg.default_frame_params(1, w, h))
        ctx.ras_cull()
        dep = torch.empty((h, w), dtype=torch.float32, device=dev)
        col = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
        ms = timeit(lambda: ctx.ras_draw_device_async(0, h, dep.data_ptr(), col.data_ptr()), stream)
        print(f"ras {w}x{h} tris={len(t)}: {ms:.3f} ms -> {1e3/ms:.1f} frames/s, algorithmic {(64*len(t)+16*w*h)/ms/1e6:.1f} GB/s")
        ctx.close()

if __name__ == "__main__":
    main()
