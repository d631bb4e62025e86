"""Experiment: concurrent D2H bandwidth of all ranks, with and without binding each rank to its GPU's NUMA node."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local)) if world > 1 else None
def cpulist_of_gpu(i):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    if isinstance(bus, bytes): bus = bus.decode()
    bus = bus.lower()
    if len(bus.split(":")[0]) == 8: bus = bus[4:]
    p = f"/sys/bus/pci/devices/{bus}/local_cpulist"
    node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
    return open(p).read().strip(), node
def parse(cl):
    out = []
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part: out.append(int(part))
    return out
def bw(tag):
    dev = torch.device("cuda", local)
    src = torch.empty(33177600, dtype=torch.uint8, device=dev)
    dst = torch.empty(33177600, dtype=torch.uint8).pin_memory()
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"[{tag}] rank {rank}: {50 * 33.1776e-3 / dt:.1f} GB/s  affinity={sorted(os.sched_getaffinity(0))[:4]}..({len(os.sched_getaffinity(0))})", flush=True)
    if world > 1: dist.barrier()
try:
    cl, node = cpulist_of_gpu(local)
except Exception as e:
    cl, node = "", f"err {e}"
print(f"rank {rank}: gpu {local} local_cpulist={cl} numa_node={node}", flush=True)
bw("default")
allowed = os.sched_getaffinity(0)
want = set(parse(cl)) & allowed if cl else set()
if want:
    os.sched_setaffinity(0, want)
bw("bound" if want else "no-binding-possible")
