"""Summarise `ncu --set full` captures (.ncu-rep) into a markdown table and the per-launch DRAM traffic file bench.py reads.

    python scripts/ncu_summary.py --md profiles/r02_ncu_summary.md --traffic profiles/ncu_traffic.json \
        c2|256|fp16=gpurun_out/r02_ncu_conv_c2.ncu-rep  adm256|64|fp16=gpurun_out/r02_ncu_conv_c5.ncu-rep  gn=...

Each positional argument is KEY=REPORT.  For keys of the form workload|batch|precision the launch with the longest duration
whose kernel name matches --kernel (default conv_) becomes that key's entry in the traffic file: dram__bytes_read.sum +
dram__bytes_write.sum of ONE launch (bench.py: roofline.traffic).  Every launch of every report goes into the table."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2->SM read sectors"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor inst %"),
    ("sm__cycles_active.avg", "SM cycles"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
]
SCALE = {"": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[start], rows[start + 1]
    return hdr, units, rows[start + 2:]


def value(row, hdr, units, name):
    if name not in hdr:
        return None
    i = hdr.index(name)
    try:
        v = float(row[i].replace(",", ""))
    except ValueError:
        return None
    return v * SCALE.get(units[i], 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", required=True)
    ap.add_argument("--traffic", default=None)
    ap.add_argument("--kernel", default="conv_")
    ap.add_argument("reports", nargs="+")
    args = ap.parse_args()
    traffic = {}
    if args.traffic and os.path.exists(args.traffic):
        with open(args.traffic) as f:
            traffic = json.load(f)
    lines = ["# ncu --set full --clock-control none summaries (scripts/ncu_summary.py)", ""]
    for spec in args.reports:
        key, rep = spec.split("=", 1)
        hdr, units, rows = raw_page(rep)
        lines += ["## %s  (`%s`)" % (key, os.path.basename(rep)), "",
                  "| kernel | " + " | ".join(lbl for _, lbl in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
        best = None
        for r in rows:
            name = r[hdr.index("Kernel Name")]
            vals = [value(r, hdr, units, m) for m, _ in METRICS]
            cells = []
            for (m, _), v in zip(METRICS, vals):
                if v is None:
                    cells.append("-")
                elif m.startswith("gpu__time"):
                    cells.append("%.1f us" % (v * 1e6))
                elif "bytes" in m:
                    cells.append("%.1f MB" % (v / 1e6))
                elif "sectors" in m:
                    cells.append("%.1f MB" % (v * 32 / 1e6))
                elif "pct" in m:
                    cells.append("%.1f" % v)
                else:
                    cells.append("%.0f" % v)
            lines.append("| `%s` | %s |" % (name[:70], " | ".join(cells)))
            if args.kernel in name and vals[0] is not None and (best is None or vals[0] > best[0]):
                best = (vals[0], name, vals)
        lines.append("")
        if best is not None and key.count("|") == 2 and best[2][1] is not None and best[2][2] is not None:
            traffic[key] = {"dram_bytes": best[2][1] + best[2][2], "dram_read": best[2][1], "dram_write": best[2][2],
                            "kernel": best[1][:100], "duration_us_under_ncu": best[0] * 1e6,
                            "source": "ncu --set full --clock-control none, %s (longest %s launch of the capture)" % (
                                os.path.basename(rep), args.kernel)}
    with open(args.md, "w") as f:
        f.write("\n".join(lines) + "\n")
    if args.traffic:
        with open(args.traffic, "w") as f:
            json.dump(traffic, f, indent=1, sort_keys=True)
    print("wrote", args.md, args.traffic or "")


if __name__ == "__main__":
    sys.exit(main())
