"""BASELINE configs[3] op sweep on its own (bench.py's extra.op_sweep), printed with the fraction of the measured HBM peak.
    python scripts/sweep_ops.py [--json out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
peak = bench.load_peaks()["hbm_gbs"]
res = bench.op_sweep(dev)
for k, v in res.items():
    print(f"{k:38s} {v:8.1f} GB/s  {v / peak:5.2f}" if not k.endswith("_us") else f"{k:38s} {v:8.2f} us")
if "--json" in sys.argv:
    json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
