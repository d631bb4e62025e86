"""Summarise an `ncu --set full --import-source on` report into the small JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_rt_v9.ncu-rep "note text" > profiles/r01_rt_trace_shade_v9_ncu_full.json

Reads the report with `ncu -i ... --page raw|source --csv` (works without a GPU).  Per kernel in the report:
selected raw metrics, the SASS opcode mix weighted by executed instructions, the warp-stall mix, and the
executed-instruction histogram by "how often was this instruction executed" (which separates per-tile, per-sample
and per-candidate code).
"""
import collections
import csv
import json
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    kernels = []
    for row in raw[2:]:
        k = {"kernel": row[hdr.index("Kernel Name")], "metrics": {}}
        for m in RAW:
            if m in hdr:
                k["metrics"][m] = [row[hdr.index(m)], units[hdr.index(m)]]
        kernels.append(k)
    import re
    for k in kernels:
        # the source page holds one kernel at a time: select it by (escaped) base name
        base = re.sub(r"^(void )?(\w+::)*", "", k["kernel"]).split("(")[0].split("<")[0]
        src = ncu_csv(rep, "source", ("-k", "regex:" + base))
        b = {"hdr": None, "rows": []}
        seen = 0
        for r in src:
            if r and r[0] == "Kernel Name":
                seen += 1
                if seen > 1:  # further launches of the same kernel: the first one is enough
                    break
                continue
            if b["hdr"] is None:
                b["hdr"] = r
            else:
                b["rows"].append(r)
        if not b["hdr"]:
            continue
        h = b["hdr"]
        ia, isrc = h.index("Instructions Executed"), h.index("Source")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        ops, stalls, hist = collections.Counter(), collections.Counter(), collections.Counter()
        total = 0
        for r in b["rows"]:
            if len(r) <= ia:
                continue
            n = int(r[ia])
            total += n
            toks = r[isrc].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            ops[op.split(".")[0]] += n
            hist[n] += n
            for i, c in stall_cols:
                stalls[c] += int(r[i] or 0)
        st = sum(stalls.values()) or 1
        k["sass_instructions_static"] = len(b["rows"])
        k["opcode_share_pct"] = {o: round(100.0 * n / total, 2) for o, n in ops.most_common(24)}
        k["warp_stall_share_pct"] = {c: round(100.0 * n / st, 2) for c, n in stalls.most_common(8)}
        k["executed_by_frequency"] = [
            {"times_executed": t, "share_pct": round(100.0 * n / total, 2)} for t, n in hist.most_common(12)]
    print(json.dumps({"note": note, "report": rep.split("/")[-1], "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
