"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
if len(sys.argv) > 2:            # only the first N launches (e.g. the device-resident steps before the e2e chunks)
    rows = rows[:1 + int(sys.argv[2])]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    ms = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    name = r[ki].split("(")[0][:60]
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + ms)
total = sum(t for _, t in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:60s} {n:8d} {t:10.3f} {t / n:9.4f} {100 * t / total:6.1f}%")
