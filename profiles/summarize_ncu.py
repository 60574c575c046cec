"""Summarise `ncu --set full` reports into one table + the DRAM-traffic json bench.py reads.
    python profiles/summarize_ncu.py gpurun_out/r2_ncu_index.ncu-rep [more.ncu-rep ...] > profiles/r2_ncu_full.txt
Needs the `ncu` CLI (reads the report with `--page raw --csv`)."""
import csv
import io
import json
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time_us", 1e-3), ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
        ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
        ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%act", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_%el", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1),
        ("lts__t_sectors_op_read.sum", "l2_rd_Msec", 1e-6), ("lts__t_sectors_op_write.sum", "l2_wr_Msec", 1e-6)]


UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]          # header, units, one row per captured launch


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")


def main():
    traffic = {}
    print("%-44s %9s %10s %10s %7s %11s %10s %6s %6s %5s %6s" % ("kernel (report)", "time_us", "dram_rd_MB", "dram_wr_MB",
                                                                "dram_%", "tensor_%act", "tensor_%el", "sm_%", "occ_%",
                                                                "regs", "grid"))
    for path in sys.argv[1:]:
        head, units, rows = rows_of(path)
        ix = {c: head.index(c) for c, _, _ in COLS if c in head}
        kn = head.index("Kernel Name")
        for r in rows:
            if len(r) <= kn:
                continue
            name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            v = {}
            for c, lab, sc in COLS:
                x = num(r[ix[c]]) if c in ix else float("nan")
                if c in ix and units[ix[c]] in UNIT:          # normalise to us / bytes first
                    x *= UNIT[units[ix[c]]]
                    sc = 1.0 if lab == "time_us" else (1e-6 if lab.endswith("_MB") else sc)
                v[lab] = x * sc
            tag = "%s (%s)" % (name[:30], path.split("/")[-1].replace(".ncu-rep", "").replace("r2_ncu_", ""))
            print("%-44s %9.1f %10.1f %10.1f %7.1f %11.1f %10.1f %6.1f %6.1f %5d %6d" % (
                tag[:44], v["time_us"], v["dram_rd_MB"], v["dram_wr_MB"], v["dram_%"], v["tensor_%act"], v["tensor_%el"],
                v["sm_%"], v["occ_%"], int(v["regs"]), int(v["grid"])))
            traffic.setdefault(tag, dict(dram_bytes=int((v["dram_rd_MB"] + v["dram_wr_MB"]) * 1e6), time_us=v["time_us"],
                                         tensor_pct_active=v["tensor_%act"]))
    json.dump(traffic, open("/tmp/ncu_summary.json", "w"), indent=1)


if __name__ == "__main__":
    main()
