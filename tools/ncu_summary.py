"""Condense ncu outputs into the text summaries committed under profiles/.

  python tools/ncu_summary.py full  <report.ncu-rep> <out.txt>     (from `ncu --set full`)
  python tools/ncu_summary.py list  <launches.csv>   <out.txt>     (from `ncu --metrics gpu__time_duration.sum --csv`)
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    ("gpu__time_duration.sum", "duration_us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep} (per launch; cold caches, serialised replays)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[idx['Kernel Name']]}\n")
            for k, short in KEYS:
                if k in idx:
                    f.write(f"  {short:22s} {r[idx[k]]} {units[idx[k]]}\n")


def launches(path, out):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = next(r for r in rows if "Kernel Name" in r)
    i_name, i_val, i_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) == len(hdr) and r is not hdr and r[i_metric] == "gpu__time_duration.sum":
            name = r[i_name].split("(")[0]
            agg[name][0] += 1
            agg[name][1] += float(r[i_val].replace(",", ""))
    tot = sum(v[1] for v in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write(f"# launch list summary of {path}: kernel, launches, total device time (ns as reported), share\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name[:90]:90s} {c:5d} {t:14.0f} {100 * t / tot:6.2f}%\n")


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
