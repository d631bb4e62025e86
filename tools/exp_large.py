"""Timing of the raytracer's large-scene path (constants in HBM, chunk lists per warp); development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for _ in range(n): fn()
        ev[1].record(stream)
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n
for (w, h, k, aa) in [(500, 500, 17, 0), (1920, 1080, 17, 0), (3840, 2160, 17, 0), (3840, 2160, 17, 4), (3840, 2160, 4, 4)]:
    tris = pkg.tessellate(pkg.cornell_box(), k)
    ctx = pkg.Context(w, h); ctx.set_stream(stream.cuda_stream); ctx.set_triangles(tris)
    fp = pkg.default_frame_params(0, w, h); fp.aaEnabled, fp.aaSamples = int(aa > 0), max(aa, 1); ctx.set_frame(fp)
    surf = torch.empty((h, w), dtype=torch.int32, device=dev)
    ms = timeit(lambda: ctx.rt_frame_device_async(0, h, surf.data_ptr()))
    ctx.enable_stats(True); ctx.rt_frame_device_async(0, h, surf.data_ptr()); st = ctx.stats(); ctx.enable_stats(False)
    rays = st["primary_rays"] + st["shadow_rays"]
    print(f"rt {w}x{h} tris={len(tris)} aa={aa}: {ms:.3f} ms, {rays/ms/1e3:.0f} Mrays/s, exact tests/ray {st['exact_tests']/rays:.2f}")
    ctx.close()
for (w, h, k) in [(500, 500, 17), (3840, 2160, 17), (3840, 2160, 60)]:
    tris = pkg.tessellate(pkg.cornell_box(), k)
    ctx = pkg.Context(w, h); ctx.set_stream(stream.cuda_stream); ctx.set_triangles(tris)
    ctx.set_frame(pkg.default_frame_params(1, w, h)); ctx.ras_cull()
    surf = torch.empty((h, w), dtype=torch.int32, device=dev)
    ms = timeit(lambda: ctx.ras_frame_device_async(0, h, surf.data_ptr()), n=20)
    print(f"ras {w}x{h} tris={len(tris)}: {ms:.3f} ms")
    ctx.close()
