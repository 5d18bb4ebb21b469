"""DRAM traffic of the convolution kernels over one adaptation iteration, from the ncutraffic step of scripts/gpu_round.sh
(gpurun_out/conv_traffic<TAG>.csv): per-kernel sums -> profiles/<tag>_conv_traffic.md and the `traffic` entries bench.py
copies into its roofline objects (profiles/traffic.json).   python scripts/summarize_traffic.py <csv> <tag>"""
import csv
import io
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, tag = sys.argv[1], sys.argv[2]
lines = [l for l in open(path, newline="") if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
per = defaultdict(lambda: defaultdict(float))          # (id) -> metric -> value
names = {}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
         "ms": 1e3, "msecond": 1e3}
for r in rows:
    v = float(r["Metric Value"].replace(",", "")) * scale.get(r["Metric Unit"], 1.0)
    per[r["ID"]][r["Metric Name"]] = v
    names[r["ID"]] = "conv_wgrad_kernel" if "conv_wgrad" in r["Kernel Name"] else "conv_tc_kernel"
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, m in per.items():
    a = agg[names[i]]
    a[0] += 1
    a[1] += m.get("dram__bytes_read.sum", 0.0)
    a[2] += m.get("dram__bytes_write.sum", 0.0)
    a[3] += m.get("gpu__time_duration.sum", 0.0)
with open(os.path.join(ROOT, "profiles", f"{tag}_conv_traffic.md"), "w") as f:
    f.write(f"# {tag}: DRAM traffic of the convolution kernels over ONE adaptation iteration (ncu, one metric pass, eager launch "
            f"of the bench workload)\n\n| kernel | launches | DRAM read MB | DRAM write MB | total MB | sum of durations us |\n"
            f"|---|---:|---:|---:|---:|---:|\n")
    for k, (n, rd, wr, us) in agg.items():
        f.write(f"| `{k}` | {n} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / 1e6:.1f} | {us:.1f} |\n")
tj = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(tj)) if os.path.isfile(tj) else {}
for k, (n, rd, wr, us) in agg.items():
    t[k] = rd + wr
json.dump(t, open(tj, "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", f"{tag}_conv_traffic.md")).read())
