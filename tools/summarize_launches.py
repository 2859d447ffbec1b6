"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total ms, share)."""
import collections
import csv
import sys

src, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
lines = [l for l in open(src) if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3, "s": 1e3}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * scale[r[ui]]
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# ncu launch list summary (gpu__time_duration.sum, --clock-control none): {title}")
print("# per-launch times are cold-cache and serialised; compare shares")
for k, (c, t) in agg.items():
    print(f"{k:64s} {c:4d} launches {t:10.3f} ms {100 * t / tot:5.1f}%")
