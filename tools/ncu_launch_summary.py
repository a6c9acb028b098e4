"""Per-kernel shares from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file LIST.csv ...).
usage: python tools/ncu_launch_summary.py LIST.csv OUT.csv "command that was profiled" """
import csv
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
agg = {}
for r in rows[1:]:
    ms = float(r[vi].replace(",", "")) * scale[r[ui]]
    n, t = agg.get(r[ki], (0, 0.0))
    agg[r[ki]] = (n + 1, t + ms)
total = sum(t for _, t in agg.values())
with open(dst, "w") as f:
    f.write("# %s\n# per-launch times are cold-cache and serialised: compare SHARES\n" % cmd)
    f.write("kernel,launches,total_ms,share_pct,avg_ms\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.3f,%.2f,%.4f\n' % (k[:100], n, t, 100 * t / total, t / n))
print(open(dst).read())
