#!/usr/bin/env python
"""Turns the .ncu-rep files brought back from the GPU box (gpurun_out/) into the small text
summaries committed here.  Usage:  python profiles/summarize.py <report.ncu-rep> <out_prefix>"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(head) if h in KEYS]
    with open(prefix + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in data:
            w.writerow([r[i] for i in cols])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    cur, fn, agg = None, None, {}
    for r in csv.reader(src.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) == 2 and r[0] == "Function Name":
            fn = r[1][:60]
        elif len(r) > 7 and r[0].isdigit():
            try:
                agg.setdefault(fn, []).append((int(r[4] or 0), int(r[7] or 0), cur, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    with open(prefix + "_hot_lines.txt", "w") as f:
        for fn, a in agg.items():
            tot = sum(x[0] for x in a) or 1
            f.write(f"== {fn}  (warp stall samples: {tot})\n")
            for x in sorted(a, reverse=True)[:25]:
                f.write(f"{100 * x[0] / tot:5.1f}%  inst={x[1]:>10d}  {x[2]}:{x[3]}  {x[4]}\n")


if __name__ == "__main__":
    main()
