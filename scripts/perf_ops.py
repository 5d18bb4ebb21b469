"""Quick op-level bandwidth sweep (run on the GPU box): python scripts/perf_ops.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
print(json.dumps({"roofline_upfirdn2d": bench.roofline_upfirdn2d(dev), "op_sweep": bench.op_sweep(dev)}, indent=1))
