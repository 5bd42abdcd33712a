"""Per-kernel shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <command>`):
    python scripts/launch_summary.py gpurun_out/X.csv "title" > profiles/X_summary.md
Per-launch times under ncu are cold-cache and serialised; the SHARES are what compares with bench.py's live CUDA-event
figures (roofline.conv_share_of_step)."""
import collections
import csv
import re
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
im = hdr.index("Metric Name")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("nlc::", "")
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
tot = sum(v[1] for v in agg.values())
print("# %s\n" % title)
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.0f | %.1f %% |" % (k[:80], n, us, 100 * us / tot))
conv = sum(v[1] for k, v in agg.items() if k.startswith("conv_tc") or k.startswith("conv_slab"))
gn = sum(v[1] for k, v in agg.items() if k.startswith("gn_"))
print("\nconv_tc + conv_slab together: %.1f %% of the listed time; GroupNorm kernels: %.1f %%; %d launches, %.1f ms listed."
      % (100 * conv / tot, 100 * gn / tot, sum(v[0] for v in agg.values()), tot / 1e3))
