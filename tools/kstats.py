"""Per-kernel table from an `ncu --metrics ... --csv` capture: last launch of each kernel, the metrics side by side."""
import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from ncu_metrics import read

SHORT = {
    "gpu__time_duration.sum": "time_us", "smsp__inst_executed.sum": "winst_M",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes", "dram__bytes_read.sum": "rd_MB", "dram__bytes_write.sum": "wr_MB",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "st_long",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "st_short",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "st_wait",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "st_lg",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "st_mio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "st_br",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "st_bar",
    "l1tex__t_sector_hit_rate.pct": "l1hit%", "lts__t_sector_hit_rate.pct": "l2hit%",
}
METRICS = ",".join(SHORT)

if __name__ == "__main__":
    if sys.argv[1] == "--list":
        print(METRICS)
        sys.exit(0)
    last = {}
    for l in read(sys.argv[1]):
        last[l["kernel"].split("(")[0]] = l
    for k, l in last.items():
        out = []
        for m, s in SHORT.items():
            if m in l:
                v = l[m]
                if s == "time_us": v *= 1e6
                if s in ("winst_M", "rd_MB", "wr_MB"): v /= 1e6
                out.append(f"{s}={v:.2f}")
        print(k[-40:], " ".join(out))
