#!/usr/bin/env python
"""Turns ncu output brought back from the GPU box (gpurun_out/) into the tracked summaries in profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/r01_launches.csv
    python profiles/summarize.py kernel   gpurun_out/prof_X.ncu-rep  profiles/r01_kernel_X.txt

`launches`: per-kernel aggregate of the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised:
compare shares, not absolutes).  `kernel`: the key `--set full` metrics of one capture.
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "gpc__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        if not name.startswith("rn_"):
            continue  # torch's kernels that generate the synthetic inputs
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg.setdefault((name, row.get("Grid Size", ""), row.get("Block Size", "")), []).append(v)
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "launches", "mean_us", "min_us", "max_us", "total_us"])
        for (name, grid, block), v in agg.items():
            w.writerow([name, grid, block, len(v), "%.2f" % (sum(v) / len(v)), "%.2f" % min(v), "%.2f" % max(v), "%.1f" % sum(v)])
    print(open(dst).read())


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write("kernel: %s\n" % name)
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("  %-85s %-14s %s\n" % (k, units[i], vals[i]))
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
