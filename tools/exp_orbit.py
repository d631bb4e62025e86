"""Config 5 end to end on one GPU: 4K orbit frames -> 24-bit BMP files in /dev/shm through b2r_group_rt_frames."""
import sys, os, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
W, H = 3840, 2160
n = int(os.environ.get("ORBIT_FRAMES", "180"))
frames = []
for i in range(n):
    f = pkg.default_frame_params(0, W, H)
    pos, rot = pkg.orbit_camera(i, 360)
    f.set_camera(pos, rot, H / 2)
    frames.append(f)
grp = pkg.Group(W, H, [0])
grp.set_triangles(pkg.cornell_box())
pattern = "/dev/shm/b2r_exp_orbit_%04d.bmp"
grp.rt_frames(frames[:2], bmp_pattern=pattern)
t0 = time.perf_counter()
grp.rt_frames(frames, bmp_pattern=pattern)
s = time.perf_counter() - t0
nbytes = sum(os.path.getsize(p) for p in glob.glob(pattern.replace("%04d", "*")))
for p in glob.glob(pattern.replace("%04d", "*")):
    os.unlink(p)
print(f"ring={os.environ.get('B2R_EXP_RING')} writers={os.environ.get('B2R_EXP_WRITERS')}: {n / s:.1f} frames/s, {nbytes / s / 1e9:.2f} GB/s of files")
