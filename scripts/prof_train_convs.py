"""The convolution launches that dominate one adaptation iteration, on their TRAINING shapes (batch 4 joint D pass,
batch 2 G pass), a few calls each -- run under ``ncu --set full`` (scripts/gpu_round.sh ncuops with
NCU_SCRIPT=scripts/prof_train_convs.py) to see what bounds them: tensor-pipe activity, L2 hit rate, DRAM traffic.

Order of launches per layer: fprop, dgrad, wgrad (each followed by its split-K fold when the plan has one).
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rick_b200 import conv

# (name, B, H, W, Cin, Cout, k, stride, pad, transposed)
LAYERS = [("D b4 conv1 256", 4, 256, 256, 128, 128, 3, 1, 1, False),
          ("D b4 conv2 256>128", 4, 257, 257, 128, 256, 3, 2, 0, False),
          ("D b4 skip 256>128", 4, 255, 255, 128, 256, 1, 2, 0, False),
          ("D b4 conv1 128", 4, 128, 128, 256, 256, 3, 1, 1, False),
          ("D b4 conv1 64", 4, 64, 64, 512, 512, 3, 1, 1, False),
          ("G b2 conv 256", 2, 256, 256, 128, 128, 3, 1, 1, False),
          ("G b2 up 128>256", 2, 128, 128, 256, 128, 3, 2, 0, True)]

only = os.environ.get("PROF_LAYERS")
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, b, h, w, cin, cout, k, stride, pad, tr in LAYERS:
    if only and not any(s in name for s in only.split(",")):
        continue
    cfg = (stride, pad, tr)
    x = cl(torch.randn(b, cin, h, w, device="cuda"))
    wt = cl(torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k))
    y = conv._fprop(x, wt, cfg)
    g = cl(torch.randn_like(y))
    for rep in range(int(os.environ.get('PROF_REPS', '1'))):
        flush.zero_()
        conv._fprop(x, wt, cfg)
        flush.zero_()
        conv._dgrad(g, wt, x, cfg)
        flush.zero_()
        conv._wgrad(g, x, wt, cfg)
    torch.cuda.synchronize()
    print("done", name, flush=True)
