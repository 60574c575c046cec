#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small committed text files.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv
    python profiles/summarize.py rep gpurun_out/prof_policy_r1.ncu-rep [...]
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        t = float(r[mv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    unit = rows[1][hdr.index("Metric Unit")]
    print("kernel | launches | total %s | share | avg" % unit)
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s | %4d | %12.1f | %5.1f%% | %10.1f" % (k[:60], n, t, 100 * t / tot, t / n))


def rep(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-70s %18s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        for p in sys.argv[2:]:
            print("==", p)
            rep(p)
