"""Aggregate an ncu source page (cuda,sass view) per source line: instruction and stall-sample shares."""
import csv
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, agg, hdr = None, [], None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0] != "":
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:110], int(r[6] or 0), int(r[7] or 0)))
        except ValueError:
            pass
ts, ti = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print(f"kernel {kernel}: {ti} warp instructions, {ts} stall samples")
for a in sorted(agg, key=lambda x: -x[4])[:top]:
    print(f"{a[0]:18s} {a[1]:4d} inst {100 * a[4] / ti:5.1f}% samp {100 * a[3] / ts:5.1f}%  {a[2]}")
