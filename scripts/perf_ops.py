"""Quick op-level bandwidth sweep (run on the GPU box): python scripts/perf_ops.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
print(json.dumps({"roofline_upfirdn2d": bench.roofline_upfirdn2d(dev), "op_sweep": bench.op_sweep(dev)}, indent=1))

# reference points for the write-heavy up=2 case: pure write, pure copy, 1:4 read:write expand (library kernels)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
big = torch.empty(32, 512, 256, 256, device=dev)
src = torch.randn(32, 512, 128, 128, device=dev)
ms = bench._time_kernel(lambda: big.fill_(1.0), flush, iters=5)
print("memset 4.3 GB:", round(big.numel() * 4 / ms / 1e6, 1), "GB/s")
dst = torch.empty_like(big)
ms = bench._time_kernel(lambda: dst.copy_(big), flush, iters=5)
print("copy 4.3 GB -> 4.3 GB:", round(2 * big.numel() * 4 / ms / 1e6, 1), "GB/s")
v = big.view(32, 512, 128, 2, 128, 2)
ms = bench._time_kernel(lambda: v.copy_(src[:, :, :, None, :, None]), flush, iters=5)
print("nearest 2x expand (1 read : 4 writes, ATen copy):", round(5 * src.numel() * 4 / ms / 1e6, 1), "GB/s")
