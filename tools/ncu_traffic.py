"""DRAM bytes and duration per kernel launch of an .ncu-rep (ncu --set full) as the JSON bench.py reads."""
import csv
import json
import subprocess
import sys

rep, bytes_per_gpu, capture = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1}
    return v * scale.get(u, 1)


kernels = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").strip()
    key = name if name not in kernels else name + "#2"
    kernels[key] = {"dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                    "duration_s": val(r, "gpu__time_duration.sum")}
json.dump({"capture": capture, "bytes_per_gpu": bytes_per_gpu, "kernels": kernels}, sys.stdout, indent=1)
