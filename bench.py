#!/usr/bin/env python3
"""bench.py -- headline benchmark of the render hot paths (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json config 3, named in config.workload): the Cornell box raytraced at
3840x2160 with AA 4x4 = 16 sub-samples per pixel and one hard-shadow ray per hit sample.
A "step" is one Draw() (trace + shade + resolve to the 32-bit surface) of ONE frame -- at N GPUs the frame is
split: rank r traces tile rows r, r+N, ... and stores its pixels into the root GPU's surface over NVLink
(b2r_rt_frame_gather_device_async); the last thread block of every launch bumps an arrival word in the root's
memory and the root's stream waits for it on the GPU.  No collective, strong scaling.
metric = Mrays/s, a ray being one ClosestIntersection call of the reference's Draw() for this frame (primary rays +
one shadow ray per hit sub-sample and light sample; counted by the kernel's own counters in an untimed pass and equal to
the oracle's count).  The kernel answers every one of them with the reference's bits; extra.shadow_rays_evaluated says
how many shadow rays it had to trace to do so (a sub-sample whose hit does not replace the pixel's carried Intersection
shades the same point as the sub-sample before: DirectLight's value is reused).

  value         device-resident: inputs already in HBM, CUDA events around each step on the
                launching stream, L2 flushed between steps, max over ranks.
  e2e           the same metric through the host-buffer C ABI a reference user calls: b2r_set_frame +
                b2r_rt_frame (N = 1) / b2r_rt_frame_part (N > 1: every rank copies its own tile rows into ONE
                page-locked host frame shared by the ranks, each over its own PCIe link); wall clock on rank 0,
                a step ends when every rank's rows are in the frame.
  roofline      rt_trace_shade kernel alone: EXECUTED FP32 flops (fadd + fmul + 2 ffma lane-operations per launch,
                from the tracked ncu capture profiles/r02_rt_trace_exec.json) / the launch time measured here,
                against the FP32 FFMA peak measured in this run -- a fraction <= 1 by construction.  The
                brute-force count of SURVEY.md 8d (every ray x every triangle) is kept as
                algorithmic_speedup_vs_bruteforce: the conservative culling removes 98.5 % of those tests.
  roofline_rasteriser  BASELINE config 4 (1,004,670 triangles at 4K): algorithmic bytes / frame time against the
                measured HBM bandwidth, the north star's second roofline.
  cpu_baseline  the reference's own raytracer.cpp (oracle/_ref, compiled from /root/reference)
                timed on this host's cores on a bounded row sample of the same workload.
  extra         the other BASELINE configs, the weak-scaling (frames per rank) numbers, the other exchange forms.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
WORKLOAD = "raytracer Cornell box (30 tris) 3840x2160, AA 4x4 = 16 samples/pixel, primary + shadow ray"
FLOP_PER_TEST, FLOP_PER_HIT, FLOP_PER_SHADE = 19, 21, 70


def algorithmic_flops(primary, shadow, ntris, lights_x_samples):
    """19 per ray/triangle test, 21 per accepted hit (one per hitting ray), 70 per shaded sample."""
    shaded = shadow // max(lights_x_samples, 1)
    return FLOP_PER_TEST * ntris * (primary + shadow) + FLOP_PER_HIT * (shaded + shadow) + FLOP_PER_SHADE * shaded


def rt4k_params(pkg):
    fp = pkg.default_frame_params(0, W4K, H4K)
    fp.aaEnabled, fp.aaSamples = 1, 4
    return fp


def reference_rt4k_params(pkg):
    """The same frame params built in Python (raytracer.cpp:33-81,116,162), so that the reference arm never loads
    libb2r.so: FrameParams is a plain ctypes struct."""
    fp = pkg.FrameParams()
    fp.numLights = 1
    fp.lights[0].position[:] = [0.0, -0.5, -0.7]
    fp.lights[0].color[:] = [1.0, 1.0, 1.0]
    fp.lights[0].intensity = 14.0
    fp.aaEnabled, fp.aaSamples = 1, 4
    fp.softShadowsSamples = 16
    fp.dofKernelSize = 8
    fp.indirectLight[:] = [0.2, 0.2, 0.2]
    fp.currentReflectance[:] = [1.0, 1.0, 1.0]
    fp.cameraPos[:] = [0.0, 0.0, -2.0]
    fp.focalLength = H4K / 2.0
    fp.cameraRot[:] = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]  # yaw 0: cos 0 = 1, sin 0 = 0, [1][1] = 1
    fp.dofFocalLength = 1.3
    return fp


def reference_cornell_box():
    """The 30 triangles of LoadTestModel from the reference itself (oracle/_ref), without libb2r.so."""
    from oracle import refbind
    for (w, h) in ((500, 500), (W4K, H4K), (96, 64)):
        if refbind.available("ras", w, h):
            ra = refbind.RefRasteriser(w, h)
            ra.load_test_model()
            return ra.get_triangles()
    return None


def cpu_reference_sample(pkg, row_step, steps, warmup, standalone=False):
    """Times the reference CPU implementation on rows 0, row_step, 2*row_step, ... of the workload.
    standalone: build scene and params without libb2r.so (the reference arm loads only oracle/ objects)."""
    from oracle import refbind, portbind
    cores = os.cpu_count() or 1
    tris = reference_cornell_box() if standalone else None
    if tris is None:
        tris = pkg.cornell_box()
        fp = rt4k_params(pkg)
    else:
        fp = reference_rt4k_params(pkg)
    rows = len(range(0, H4K, row_step))
    # rays of the sample, from the port's counters (identical arithmetic, untimed)
    cnt = portbind.rt_draw(tris, fp, W4K, H4K, 0, H4K, threads=cores, ystep=row_step)
    rays = cnt["primary_rays"] + cnt["shadow_rays"]
    if refbind.available("rt", W4K, H4K):
        kind = "reference"
        rt = refbind.RefRaytracer(W4K, H4K)
        rt.load_test_model()
        rt.set_lights(fp.lights_array())
        rt.set_camera(fp.cameraPos[:], fp.cameraRot[:], fp.focalLength)
        rt.set_flags(aa=True, aa_samples=4, threads=cores)
        rt.set_rows(0, H4K, row_step)
        run = rt.time_draw
    else:
        kind = "port"
        def run():
            t0 = time.perf_counter()
            portbind.rt_draw(tris, fp, W4K, H4K, 0, H4K, threads=cores, ystep=row_step)
            return time.perf_counter() - t0
    for _ in range(warmup):
        run()
    secs = [run() for _ in range(steps)]
    mean = sum(secs) / len(secs)
    return dict(value=rays / mean / 1e6, unit="Mrays/s", cores=cores, kind=kind,
                sample=f"{rows} of {H4K} rows (every {row_step}th) of the 4K AA4x4 frame, {rays} rays per step, "
                       f"{steps} steps, OpenMP on all {cores} host threads",
                ms_per_step=mean * 1e3, rays_per_step=rays)


def run_reference_arm(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # Every 4th row of the frame (0.6 s per step on 16 threads) as long as the run stays near two minutes; a thinner
    # sample would leave the reference's OpenMP threads unevenly loaded and understate it.
    budget_steps = max(1, args.steps + args.warmup)
    row_step = max(4, -(-4 * budget_steps * 6 // 1200))  # ceil(4 * steps * 0.6 s / 120 s)
    base = cpu_reference_sample(pkg, row_step=row_step, steps=args.steps, warmup=args.warmup, standalone=True)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": base["value"], "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": base["sample"]},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()  # the exact PID we started
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smmax.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _load_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def run_gpu_arm(args, pkg):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) write to fd 1: park stdout on stderr until the JSON line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    par = pkg.parallel

    stream = torch.cuda.Stream(device=dev)
    tris = pkg.cornell_box()
    fp = rt4k_params(pkg)
    ctx = pkg.Context(W4K, H4K, device=local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    npx = W4K * H4K
    d_col = torch.empty((H4K, W4K, 3), dtype=torch.float32, device=dev)
    d_surf = torch.empty((H4K, W4K), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # untimed: ray counts of one frame from the kernel's counters
    ctx.enable_stats(True)
    ctx.rt_frame_device_async(0, H4K, d_surf.data_ptr())
    st = ctx.stats()
    ctx.enable_stats(False)
    rays = st["primary_rays"] + st["shadow_rays"]
    flops = algorithmic_flops(st["primary_rays"], st["shadow_rays"], len(tris), fp.numLights * 1)

    # ---- the step: ONE frame, split over the ranks, gathered on rank 0's GPU -------------------------------------
    # rank 0 owns the surface and, right behind it, the arrival word; the others map both (CUDA IPC, NVLink)
    root_bytes = npx * 4
    if rank == 0:
        root_mem, handle = ctx.shared_alloc(root_bytes + 256)
        ctx.copy_device_async(root_mem + root_bytes, torch.zeros(64, dtype=torch.int32, device=dev).data_ptr(), 256)
        ctx.synchronize()
    else:
        root_mem, handle = 0, None
    if world > 1:
        box = [handle]
        dist.broadcast_object_list(box, src=0)
        root = root_mem if rank == 0 else ctx.shared_open(box[0])
    else:
        root = root_mem
    arrive = root + root_bytes
    calls = [0]

    def frame_split():
        # one kernel per rank: traces its interleaved tile rows, stores every pixel into the root's surface; its
        # last thread block tells the root.  The root's stream then waits (on the GPU) for all N launches.
        ctx.rt_frame_gather_device_async(rank, world, root, arrive)
        calls[0] += 1
        if rank == 0:
            ctx.stream_wait_value32(arrive, world * calls[0])

    # untimed: the assembled frame must be bit-equal to a frame rank 0 rendered alone
    barrier()
    frame_split()
    barrier()
    split_verified = None
    if rank == 0:
        got = torch.empty((H4K, W4K), dtype=torch.int32, device=dev)
        ctx.copy_device_async(got.data_ptr(), root, root_bytes)
        ctx.rt_frame_device_async(0, H4K, d_surf.data_ptr())
        ctx.synchronize()
        split_verified = bool(torch.equal(got, d_surf))
        del got
    barrier()

    def timed_loop(fn, steps, warmup, do_flush=True):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                fn()
            barrier()
            evs = []
            for _ in range(steps):
                if do_flush:
                    flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                evs.append((e0, e1))
            barrier()
        return sum(a.elapsed_time(b) for a, b in evs)  # ms over all steps, device time

    # clocks are sampled from before the warm-up to after the timed loop (nvidia-smi needs ~100 ms to start;
    # the extra warm-up frames below keep the GPU under the same load while it does)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nwarm = min(4000, int(400 * world / 1.05))  # ~0.4 s; the same number of calls on every rank (the arrival count must agree)
    for _ in range(nwarm):
        frame_split()
    barrier()
    launches0 = ctx.launch_count()
    total_ms = allmax(timed_loop(frame_split, args.steps, args.warmup))
    launches = ctx.launch_count() - launches0 - args.warmup
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = rays * args.steps / (total_ms * 1e-3) / 1e6  # Mrays/s of the one frame all ranks work on

    # roofline: the trace kernel alone (one GPU, whole frame), executed flops from the tracked ncu capture
    kern_ms = timed_loop(lambda: ctx.rt_frame_device_async(0, H4K, d_surf.data_ptr()), args.steps, 1) / args.steps
    peak_tf, _ = ctx.measure_fp32_peak()
    execp = _load_json("r02_rt_trace_exec.json") or {}
    exec_flops = execp.get("executed_fp32_flops_per_launch")
    exec_tf = exec_flops / (kern_ms * 1e-3) / 1e12 if exec_flops else None
    algo_tf = flops / (kern_ms * 1e-3) / 1e12

    # ---- e2e: host-buffer C ABI, every step = frame constants H2D + this rank's rows D2H into the one host frame ----
    if world > 1:
        shm = par.SharedHostFrame(W4K, H4K, rank, world, tag=os.environ.get("MASTER_PORT", "0"))
        ctx.pin_host_buffer(shm.frame)
        host_frame = shm.frame
        estep = [0]

        def frame_e2e():
            ctx.set_frame(fp)
            ctx.rt_frame_part(rank, world, host_frame)
            estep[0] += 1
            shm.publish(estep[0])
            if rank == 0:
                shm.wait_all(estep[0])
        call = "b2r_set_frame + b2r_rt_frame_part (own tile rows into one page-locked host frame shared by the ranks)"
        d2h_bytes = npx * 4 // world
    else:
        surf_host = torch.empty((H4K, W4K), dtype=torch.int32).pin_memory()
        host_frame = surf_host.numpy().view(np.uint32)

        def frame_e2e():
            ctx.set_frame(fp)
            ctx.rt_frame(host_frame)
        call = "b2r_set_frame + b2r_rt_frame (pinned host surface)"
        d2h_bytes = npx * 4
    for _ in range(max(args.warmup, 1)):
        frame_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        mine = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.broadcast(mine, src=0)  # rank 0's clock: its steps end when every rank has delivered
        e2e_s = float(mine.item())
    e2e_value = rays * args.steps / e2e_s / 1e6
    h2d_bytes = 1728 + 16 * (1 + fp.numLights)  # DevFrame up to and including the used ray origins (b2r_set_frame)
    e2e_verified = None
    barrier()
    if rank == 0:
        ctx.rt_frame_device_async(0, H4K, d_surf.data_ptr())
        ctx.synchronize()
        e2e_verified = bool(np.array_equal(host_frame, d_surf.cpu().numpy().view(np.uint32)))
    barrier()
    if world > 1:
        ctx.unpin_host_buffer(shm.frame)
        host_frame = None
        shm.close()

    extra = {"frames_per_s": args.steps / (total_ms * 1e-3), "rays_per_frame": rays,
             "primary_rays": st["primary_rays"], "shadow_rays": st["shadow_rays"],
             "shadow_rays_evaluated": st["shadow_rays_evaluated"],
             "algorithmic_gflop_per_frame": flops / 1e9, "e2e_frames_per_s": args.steps / e2e_s,
             "split_verified": split_verified, "e2e_frame_verified": e2e_verified}

    # weak scaling for comparison: every rank draws its own frames (nothing exchanged), with pixelColours as in round 1
    def frame_own():
        ctx.rt_frame_device_async(0, H4K, d_surf.data_ptr(), d_col.data_ptr())
    own_ms = allmax(timed_loop(frame_own, max(args.steps // 2, 5), args.warmup)) / max(args.steps // 2, 5)
    extra["frames_per_rank_weak"] = {"ms_per_frame": own_ms, "value": world * rays / (own_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                                     "scaling": "weak", "outputs": "pixelColours + 32-bit surface in HBM, one frame per rank"}

    # BASELINE config 5: 360-frame camera orbit, frames partitioned across the ranks (no collective on the data path);
    # every frame re-uploads its camera (b2r_set_frame) and is traced + resolved at 4K, 1 spp + hard shadow
    fp_orbit = pkg.default_frame_params(0, W4K, H4K)
    my_frames = par.frames_for_rank(rank, world, 360)

    def orbit_pass():
        for fidx in my_frames:
            pos, rot = pkg.orbit_camera(fidx, 360)
            fp_orbit.set_camera(pos, rot, H4K / 2)
            ctx.set_frame(fp_orbit)
            ctx.rt_frame_device_async(0, H4K, d_surf.data_ptr())
    orbit_pass()  # warm-up
    barrier()
    t0 = time.perf_counter()
    orbit_pass()
    barrier()
    orbit_s = allmax(time.perf_counter() - t0)
    extra["orbit_360_frames_4k_1spp"] = {"seconds": orbit_s, "frames_per_s": 360 / orbit_s,
                                         "partition": f"frames f = rank (mod {world})", "timing": "wall clock incl. per-frame b2r_set_frame"}
    ctx.set_frame(fp)
    # ... and end to end: every frame through the host ABI to a 24-bit BMP file (what the reference's SDL_SaveBMP
    # leaves, raytracer.cpp:175): BGR conversion on the GPU, 3 bytes per pixel over PCIe, files written by host
    # threads (b2r_group_rt_frames, one group member per rank)
    try:
        import glob
        sv = os.statvfs("/dev/shm")
        free_gb = sv.f_bavail * sv.f_frsize / 1e9
        nfr = 360 if free_gb > 16 else 96
        frames = []
        for fidx in range(rank, nfr, world):
            f = pkg.default_frame_params(0, W4K, H4K)
            pos, rot = pkg.orbit_camera(fidx, 360)
            f.set_camera(pos, rot, H4K / 2)
            frames.append(f)
        grp = pkg.Group(W4K, H4K, [local])
        grp.set_triangles(tris)
        pattern = f"/dev/shm/b2r_orbit_{os.environ.get('MASTER_PORT', '0')}_{rank}_%04d.bmp"
        grp.rt_frames(frames[:2], bmp_pattern=pattern)  # warm-up
        barrier()
        t0 = time.perf_counter()
        grp.rt_frames(frames, bmp_pattern=pattern)
        barrier()
        bmp_s = allmax(time.perf_counter() - t0)
        nbytes = sum(os.path.getsize(p) for p in glob.glob(pattern.replace("%04d", "*")))
        for pth in glob.glob(pattern.replace("%04d", "*")):
            os.unlink(pth)
        grp.close()
        extra["orbit_360_frames_4k_e2e_with_bmp"] = {
            "frames": nfr, "seconds": bmp_s, "frames_per_s": nfr / bmp_s, "bmp_bytes_written_this_rank": nbytes,
            "path": "/dev/shm (files removed afterwards)",
            "call": "b2r_group_rt_frames(bmp_pattern): b2r_set_frame + trace + BGR24 on the GPU + D2H + b2r_write_bmp on host threads"}
    except Exception as ex:  # a full /dev/shm must not cost the headline
        extra["orbit_360_frames_4k_e2e_with_bmp"] = {"error": repr(ex)}
    barrier()

    if world > 1 and H4K % world == 0:
        # the other exchange forms, for comparison (strong scaling, same frame)
        band = H4K // world
        y0, y1 = rank * band, (rank + 1) * band

        def frame_band():
            ctx.rt_frame_device_async(y0, y1, d_surf.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(d_surf.view(-1), d_surf.view(-1)[y0 * W4K:y1 * W4K])
        bms = allmax(timed_loop(frame_band, args.steps, args.warmup)) / args.steps
        extra["band_split_nccl_allgather"] = {"ms_per_frame": bms, "value": rays / (bms * 1e-3) / 1e6, "unit": "Mrays/s",
                                              "scaling": "strong", "collective": "contiguous row bands + nccl all_gather of 32-bit surface bands"}
        # all-gather inside the trace kernel: every rank stores its pixels into every rank's surface (round 1's form)
        mine, handle = ctx.shared_alloc(npx * 4)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        ptrs = [mine if r == rank else ctx.shared_open(handles[r]) for r in range(world)]
        tick = torch.zeros(1, dtype=torch.int32, device=dev)
        order = [mine] + [p for r, p in enumerate(ptrs) if r != rank]

        def frame_band_fused():
            ctx.rt_frame_split_device_async(rank, world, order)
            with torch.cuda.stream(stream):
                dist.all_reduce(tick)
        fms = allmax(timed_loop(frame_band_fused, args.steps, args.warmup)) / args.steps
        extra["split_allgather_in_kernel"] = {"ms_per_frame": fms, "value": rays / (fms * 1e-3) / 1e6, "unit": "Mrays/s",
                                              "scaling": "strong",
                                              "exchange": "trace kernel stores into every rank's peer-mapped surface + a 4-byte nccl all-reduce per frame"}
        # rasteriser config 4, sort-first: the 1,004,670 triangles are replicated, every rank rasterises and shades
        # its row band and the shade kernel's surface rows are gathered with NCCL.  Per-triangle work (vertex
        # shading, classification) is not divided by the band, so this split is bounded by it.
        rctx = pkg.Context(W4K, H4K, device=local)
        rctx.set_stream(stream.cuda_stream)
        rctx.set_triangles(pkg.tessellate(tris, 183))
        rctx.set_frame(pkg.default_frame_params(1, W4K, H4K))
        rctx.ras_cull()

        def ras_band():
            rctx.ras_frame_device_async(y0, y1, d_surf.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(d_surf.view(-1), d_surf.view(-1)[y0 * W4K:y1 * W4K])
        nr = max(args.steps // 4, 5)
        rms = allmax(timed_loop(ras_band, nr, args.warmup)) / nr
        extra["ras_band_split_4k_1m_tris"] = {"ms_per_frame": rms, "frames_per_s": 1e3 / rms, "scaling": "strong",
                                              "collective": "nccl all_gather of 32-bit surface bands"}
        rctx.close()
        barrier()
        for r in range(world):
            if r != rank:
                ctx.shared_close(ptrs[r])
        barrier()
        ctx.shared_free(mine)

    # the same frame through the single-process group ABI (b2r_group_rt_frame: what host/raytracer_dropin.cpp's Draw()
    # calls with several devices): rank 0 drives every GPU of the job while the other ranks wait
    barrier()
    done_flag = f"/dev/shm/b2r_group_done_{os.environ.get('MASTER_PORT', '0')}"
    if rank == 0 and os.path.exists(done_flag):
        os.unlink(done_flag)
    barrier()
    if rank == 0 and world > 1 and torch.cuda.device_count() >= world:
        try:
            grp = pkg.Group(W4K, H4K, list(range(world)))
            grp.set_triangles(tris)
            gsurf = torch.empty((H4K, W4K), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
            for _ in range(3):
                grp.set_frame(fp)
                grp.rt_frame(gsurf)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                grp.set_frame(fp)
                grp.rt_frame(gsurf)
            gs = (time.perf_counter() - t0) / args.steps
            ctx.rt_frame_device_async(0, H4K, d_surf.data_ptr())
            ctx.synchronize()
            extra["group_rt_frame_e2e"] = {"ms_per_frame": gs * 1e3, "value": rays / gs / 1e6, "unit": "Mrays/s",
                                           "verified": bool(np.array_equal(gsurf, d_surf.cpu().numpy().view(np.uint32))),
                                           "call": f"b2r_group_set_frame + b2r_group_rt_frame, one process driving {world} GPUs"}
            grp.close()
        except Exception as ex:
            extra["group_rt_frame_e2e"] = {"error": repr(ex)}
    # the other ranks wait on the host (a flag file), not in an NCCL barrier whose kernel would share their GPU with
    # rank 0's group members
    if rank == 0:
        open(done_flag, "w").close()
    else:
        torch.cuda.synchronize(dev)
        while not os.path.exists(done_flag):
            time.sleep(0.005)
    barrier()
    if rank == 0 and os.path.exists(done_flag):
        os.unlink(done_flag)
    if world > 1 and rank != 0:
        ctx.shared_close(root)
    barrier()
    if rank == 0:
        ctx.shared_free(root_mem)

    if rank == 0:
        others = other_configs(pkg, torch, dev, stream, flush, local, cpu=(world == 1 and not args.no_cpu))
        ras_roof = others.pop("roofline_rasteriser", None)
        extra.update(others)
        cpu = cpu_reference_sample(pkg, row_step=4, steps=3, warmup=1) if world == 1 and not args.no_cpu else None
        roofline = {
            "bound": "fp32", "achieved": exec_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (exec_tf / peak_tf) if (exec_tf and peak_tf) else None,
            "traffic": execp.get("dram_bytes_per_launch"),
            "kernel": "rt_trace_shade_kernel", "kernel_ms": kern_ms,
            "executed_fp32_flops_per_launch": exec_flops,
            "issue_slot_frac": execp.get("issue_slot_frac"), "fma_pipe_frac": execp.get("fma_pipe_frac"),
            "source": "profiles/r02_rt_trace_exec.json (ncu: smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on, "
                      "ffma = 2 flops; per launch of this exact workload) / kernel time measured in this run",
            "algorithmic_speedup_vs_bruteforce": algo_tf / peak_tf if peak_tf else None,
            "algorithmic_tflops_bruteforce_model": algo_tf,
            "exact_tests_frac": st["exact_tests"] / float(rays * len(tris)),
            "note": "frac = executed FP32 flops / time / measured FFMA peak, <= 1 by construction (exact-order mul+add code "
                    "cannot use FMA, so 0.5 would already saturate the pipe). algorithmic_speedup_vs_bruteforce is round 1's "
                    "number: SURVEY 8d flops of the brute-force formulation / time / peak; the conservative culling leaves "
                    "exact_tests_frac of those tests to execute.",
            "peak_source": "FFMA microbenchmark measured in this run (b2r_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry"}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": 1,
                       "parallelism": f"one frame split over {world} GPU(s): interleaved 8-row tile rows, gathered on rank 0 over NVLink"
                                      if world > 1 else "one frame on one GPU",
                       "l2": "flushed between steps (256 MB fill)", "outputs": "32-bit surface in HBM (on rank 0's GPU)",
                       "split_verified": split_verified},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_s / args.steps * 1e3, "call": call,
                    "frame_verified": e2e_verified},
            "gpu_launches": launches,
            "roofline": roofline,
            "clocks": clocks,
            "extra": extra,
        }
        if ras_roof:
            line["roofline_rasteriser"] = ras_roof
        if cpu:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def other_configs(pkg, torch, dev, stream, flush, local, cpu=False):
    """The remaining BASELINE configs, device-timed (reported under extra, not the headline)."""
    out = {}
    tris = pkg.cornell_box()

    def avg_ms(fn, n=10, warm=3):
        with torch.cuda.stream(stream):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize(dev)
            tot = 0.0
            for _ in range(n):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                torch.cuda.synchronize(dev)
                tot += e0.elapsed_time(e1)
        return tot / n

    # config 1: raytracer 500x500
    ctx = pkg.Context(500, 500, device=local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_triangles(tris)
    ctx.set_frame(pkg.default_frame_params(0, 500, 500))
    col = torch.empty((500, 500, 3), dtype=torch.float32, device=dev)
    surf = torch.empty((500, 500), dtype=torch.int32, device=dev)
    def f1():
        ctx.rt_frame_device_async(0, 500, surf.data_ptr(), col.data_ptr())
    ms = avg_ms(f1)
    out["rt_500x500"] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms, "Mrays_per_s": 500000 / ms / 1e3}
    # the same frame through the host ABI all the way to a BMP on disk (what the reference does on Esc, raytracer.cpp:175)
    import tempfile as _tf
    bmp = os.path.join(_tf.gettempdir(), f"b2r_bench_{os.getpid()}.bmp")
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.rt_frame()
        pkg.write_bmp(bmp, ctx.resolve_bgr8(), 500, 500)
    out["rt_500x500"]["host_abi_with_bmp_write_frames_per_s"] = 20 / (time.perf_counter() - t0)
    os.unlink(bmp)
    # config 2: rasteriser 500x500
    ctx.set_frame(pkg.default_frame_params(1, 500, 500))
    ctx.ras_cull()
    dep = torch.empty((500, 500), dtype=torch.float32, device=dev)
    def f2():
        ctx.ras_frame_device_async(0, 500, surf.data_ptr(), dep.data_ptr(), col.data_ptr())
    ms = avg_ms(f2)
    out["ras_500x500"] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms}
    ctx.close()
    # config 3, second reading (SURVEY 8d "3b"): 16 jittered light samples per pixel (soft shadows), 1 primary ray
    ctx = pkg.Context(W4K, H4K, device=local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_triangles(tris)
    fps = pkg.default_frame_params(0, W4K, H4K)
    fps.softShadowsEnabled, fps.softShadowsSamples = 1, 16
    fps.set_random_positions(pkg.jitter_table(1, [0, -0.5, -0.7]))
    ctx.set_frame(fps)
    surf4 = torch.empty((H4K, W4K), dtype=torch.int32, device=dev)
    ms = avg_ms(lambda: ctx.rt_frame_device_async(0, H4K, surf4.data_ptr()))
    ctx.enable_stats(True)
    ctx.rt_frame_device_async(0, H4K, surf4.data_ptr())
    st = ctx.stats()
    ctx.enable_stats(False)
    out["rt_4k_soft_shadows_16"] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                                    "Mrays_per_s": (st["primary_rays"] + st["shadow_rays"]) / ms / 1e3}
    ctx.close()
    # config 4: rasteriser 4K, 1,004,670 triangles
    big = pkg.tessellate(tris, 183)
    ctx = pkg.Context(W4K, H4K, device=local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_triangles(big)
    ctx.set_frame(pkg.default_frame_params(1, W4K, H4K))
    ctx.ras_cull()
    dep = torch.empty((H4K, W4K), dtype=torch.float32, device=dev)
    col = torch.empty((H4K, W4K, 3), dtype=torch.float32, device=dev)
    ms = avg_ms(lambda: ctx.ras_draw_device_async(0, H4K, dep.data_ptr(), col.data_ptr()))
    abytes = 64 * len(big) + 16 * W4K * H4K
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    src = "MEASURED_PEAKS.json hbm_gbs" if peak else "fallback 6650 GB/s (B200_PROFILING.md)"
    peak = peak or 6650.0
    prof = _load_json("r02_ras_traffic.json") or {}
    out["roofline_rasteriser"] = {"bound": "hbm", "achieved": abytes / ms / 1e6, "peak": peak, "unit": "GB/s",
                                  "frac": abytes / ms / 1e6 / peak, "traffic": prof.get("sortlast_dram_bytes_per_frame"),
                                  "algorithmic_bytes": abytes, "ms_per_frame": ms, "triangles": len(big),
                                  "workload": "rasteriser, tessellated Cornell box (1,004,670 triangles) 3840x2160, depth + colour out",
                                  "pipeline": "sort-last (default): ras_small + ras_shade",
                                  "peak_source": src, "scope": "whole Draw() pipeline, all kernels",
                                  "traffic_source": "profiles/r02_ras_traffic.json (ncu dram__bytes_read+write, summed over the frame's kernels)"}
    out["ras_4k_1m_tris"] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms, "triangles": len(big)}
    # the screen-tile pipeline on the same frame (B2R_OPT_RAS_VARIANT = 2)
    ctx.set_option(pkg.capi.OPT_RAS_VARIANT, 2)
    ms2 = avg_ms(lambda: ctx.ras_draw_device_async(0, H4K, dep.data_ptr(), col.data_ptr()))
    out["ras_4k_1m_tris_tiled"] = {"ms_per_frame": ms2, "frames_per_s": 1e3 / ms2,
                                   "roofline": {"bound": "hbm", "achieved": abytes / ms2 / 1e6, "peak": peak, "unit": "GB/s",
                                                "frac": abytes / ms2 / 1e6 / peak, "traffic": prof.get("tiles_dram_bytes_per_frame")},
                                   "pipeline": "screen tiles: ras_setup + ras_bin + ras_tile (keys and spans in shared memory)"}
    ctx.set_option(pkg.capi.OPT_RAS_VARIANT, 0)
    ctx.close()
    if cpu:
        out["cpu_reference"] = cpu_reference_other_configs(pkg, big)
    return out


def cpu_reference_other_configs(pkg, big):
    """Draw() of the reference itself (oracle/_ref) for the other configs, on this host: raytracer 500^2 with OpenMP on
    every thread, rasteriser single-threaded (the reference default, rasteriser.cpp:22; its OpenMP mode races)."""
    from oracle import refbind
    res = {"cores": os.cpu_count()}
    if refbind.available("rt", 500, 500):
        rt = refbind.RefRaytracer(500, 500)
        rt.load_test_model()
        rt.set_lights([[0, -0.5, -0.7, 1, 1, 1, 14]])
        rt.set_camera_yaw([0, 0, -2], 0.0, 250.0)
        rt.set_flags(threads=os.cpu_count())
        rt.time_draw()
        res["rt_500x500_ms"] = min(rt.time_draw() for _ in range(5)) * 1e3
    if refbind.available("ras", 500, 500):
        ra = refbind.RefRasteriser(500, 500)
        ra.load_test_model()
        ra.set_lights([[0, -0.5, -0.7, 1, 1, 1, 14]])
        ra.set_flags()
        ra.update_yaw([0, 0, -3], 0.0, 500.0)
        ra.time_draw()
        res["ras_500x500_ms"] = min(ra.time_draw() for _ in range(5)) * 1e3
    if refbind.available("ras", W4K, H4K):
        ra = refbind.RefRasteriser(W4K, H4K)
        ra.set_triangles(big)
        ra.set_lights([[0, -0.5, -0.7, 1, 1, 1, 14]])
        ra.set_flags()
        ra.update_yaw([0, 0, -3], 0.0, float(H4K))
        res["ras_4k_1m_tris_ms"] = min(ra.time_draw() for _ in range(2)) * 1e3
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2r", choices=["b2r", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (development)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2r" else args.warmup
    import __graft_entry__ as g
    pkg = g.load_package()
    if args.impl == "reference":
        run_reference_arm(args, pkg)
    else:
        run_gpu_arm(args, pkg)


if __name__ == "__main__":
    main()
