"""Turn `ncu --metrics ... --csv --log-file X.csv` captures into the small tracked JSON files bench.py reads.

    python tools/ncu_metrics.py rt_exec  gpurun_out/r2_rt_exec.csv   > profiles/r02_rt_trace_exec.json
    python tools/ncu_metrics.py ras_traffic gpurun_out/r2_ras_sl.csv gpurun_out/r2_ras_tiles.csv > profiles/r02_ras_traffic.json
"""
import collections
import csv
import json
import sys


def read(path):
    """-> list of launches: {"kernel": name, metric: float, ...} in launch order."""
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, mi, vi, ui, idi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("ID")
    launches = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[idi], {"kernel": r[ki]})
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "us": 1e-6, "nsecond": 1e-9,
                 "ns": 1e-9, "msecond": 1e-3, "ms": 1e-3, "second": 1.0}.get(unit, 1.0)
        d[r[mi]] = v * scale
    return list(launches.values())


def rt_exec(path):
    ls = [l for l in read(path) if "rt_trace_shade_kernel" in l["kernel"]]
    l = ls[-1]
    g = lambda k: l.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k, 0.0)
    flops = g("fadd") + g("fmul") + 2 * g("ffma") + 2 * g("fadd2") + 2 * g("fmul2") + 4 * g("ffma2")
    out = {
        "note": "ncu --clock-control none, tools/profile_run.py rtsurf (3840x2160, AA 4x4, 30 triangles, surface only): "
                "one launch of rt_trace_shade_kernel. flops = fadd + fmul + 2 ffma + 2 fadd2 + 2 fmul2 + 4 ffma2 "
                "(thread-level, predicated-on).",
        "kernel": l["kernel"].split("(")[0],
        "thread_inst": {k: g(k) for k in ("fadd", "fmul", "ffma", "fadd2", "fmul2", "ffma2")},
        "executed_fp32_flops_per_launch": flops,
        "warp_inst_per_launch": l.get("smsp__inst_executed.sum"),
        "issue_slot_frac": l.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) / 100.0,
        "fma_pipe_frac": l.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 0.0) / 100.0,
        "xu_pipe_frac": l.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 0.0) / 100.0,
        "alu_pipe_frac": l.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 0.0) / 100.0,
        "dram_bytes_per_launch": l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0),
        "time_under_ncu_s": l.get("gpu__time_duration.sum"),
    }
    print(json.dumps(out, indent=1))


def frame_traffic(path, last_kernel):
    """DRAM bytes and time of the last complete frame in the capture (kernels up to and including last_kernel)."""
    ls = read(path)
    ends = [i for i, l in enumerate(ls) if last_kernel in l["kernel"]]
    end = ends[-1]
    start = ends[-2] + 1 if len(ends) > 1 else 0
    frame = ls[start:end + 1]
    per = [{"kernel": l["kernel"].split("(")[0], "dram_read": l.get("dram__bytes_read.sum", 0.0),
            "dram_write": l.get("dram__bytes_write.sum", 0.0), "time_under_ncu_s": l.get("gpu__time_duration.sum")} for l in frame]
    return per, sum(p["dram_read"] + p["dram_write"] for p in per)


def ras_traffic(sl_path, tiles_path):
    sl, slb = frame_traffic(sl_path, "ras_shade_kernel")
    tl, tlb = frame_traffic(tiles_path, "ras_tile_kernel")
    print(json.dumps({
        "note": "ncu --clock-control none, tools/profile_run.py ras / rastiles (3840x2160, 1,004,670 triangles, depth + colour "
                "out): dram__bytes_read.sum + dram__bytes_write.sum of every kernel of one frame. Algorithmic bytes "
                "(SURVEY 8d): 64 T + 16 W H = 197.0 MB.",
        "algorithmic_bytes": 64 * 1004670 + 16 * 3840 * 2160,
        "sortlast_dram_bytes_per_frame": slb, "sortlast_kernels": sl,
        "tiles_dram_bytes_per_frame": tlb, "tiles_kernels": tl}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "rt_exec":
        rt_exec(sys.argv[2])
    else:
        ras_traffic(sys.argv[2], sys.argv[3])
