#!/usr/bin/env python
"""Copies the round's artefacts from gpurun_out/ into profiles/ and prints the numbers the README
quotes: python profiles/collect.py  (after the gpurun call listed in profiles/README.md)."""
import collections
import csv
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "gpurun_out")


def jline(path):
    for l in open(path):
        if l.startswith("{"):
            return json.loads(l)
    return None


def main():
    subprocess.run(["python", os.path.join(HERE, "summarize.py"), os.path.join(OUT, "r1_top_kernels.ncu-rep"),
                    os.path.join(HERE, "r1_top_kernels")], check=True)
    shutil.copy(os.path.join(OUT, "r1_launches.csv"), os.path.join(HERE, "r1_launches.csv"))
    rows = list(csv.reader(open(os.path.join(HERE, "r1_top_kernels_metrics.csv"))))
    h = rows[0]
    traffic = {}
    print("kernel | time us | DRAM MB | issue % | warp instr M")
    for r in rows[2:]:
        d = dict(zip(h, r))
        nm = d["Kernel Name"]
        b = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e9
        ent = ("ipb_region_stats" if "region_stats" in nm else "ipb_hist_select" if "pq_" in nm
               else "ipb_fret_pixels" if "fret_pixels" in nm else "ipb_fa_segment")
        traffic[ent] = traffic.get(ent, 0) + b
        print(nm[:44], "|", round(float(d["gpu__time_duration.sum"]), 1), "|", round(b / 1e6), "|",
              round(float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]), 1), "|",
              round(float(d["smsp__inst_executed.sum"]) / 1e6))
    traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per step (64 frames) of the entry point's kernels, "
                        "from profiles/r1_top_kernels_metrics.csv (ncu --set full, one step of python bench.py --steps 1 "
                        "--warmup 3 --no-cpu-baseline)")
    json.dump(traffic, open(os.path.join(HERE, "r1_traffic.json"), "w"), indent=1)
    for name in ("r1_bench_n1", "r1_bench_default", "r1_bench_reference"):
        p = os.path.join(OUT, name + ".log")
        if os.path.exists(p) and jline(p):
            d = jline(p)
            if "roofline" in d:
                d["roofline"]["traffic"] = traffic.get(d["roofline"]["kernel"])
            open(os.path.join(HERE, name + ".jsonl"), "w").write(json.dumps(d) + "\n")
            print(name, round(d["value"], 2), "Mpix/s", round(d["ms_per_step"], 3), "ms/step e2e", round(d["e2e"]["value"], 1),
                  d.get("roofline", {}).get("frac"), d.get("pipeline_roofline", {}).get("frac_of_peak"),
                  d.get("cpu_baseline", {}).get("value"))
            if "kernels" in d:
                tot = d["ms_per_step_serialized"] * d["steps"]
                print({k: (round(v["ms"] / d["steps"], 4), round(v["ms"] / tot, 3)) for k, v in d["kernels"].items()})
    rws = list(csv.reader(l for l in open(os.path.join(HERE, "r1_launches.csv")) if l.startswith('"')))
    ki, vi = rws[0].index("Kernel Name"), rws[0].index("Metric Value")
    seq = [(r[ki][:48], float(r[vi].replace(",", "")) / 1000) for r in rws[1:]]
    idx = [i for i, (k, _) in enumerate(seq) if k.startswith("void ipb_k_raster<0>")]
    last = seq[idx[-2]:idx[-1]]
    tot = sum(v for _, v in last)
    agg = collections.OrderedDict()
    for k, v in last:
        agg[k] = agg.get(k, 0) + v
    print("launches in the last complete step:", len(last), "total us", round(tot, 1))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"| `{k.split('(')[0].replace('void ', '')}` | {v:.0f} | {100 * v / tot:.1f} % |")


if __name__ == "__main__":
    main()
