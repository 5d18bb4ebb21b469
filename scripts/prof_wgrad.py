"""ncu target: the weight-gradient kernel and the small-map forward conv on adaptation-loop shapes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rick_b200 import conv_tc as ct

dev = "cuda"
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
for (b, h, cin, cout) in [(4, 256, 128, 128), (4, 64, 512, 512), (4, 128, 256, 256)]:
    x = torch.randn(b, h, h, cin, device=dev)
    g = torch.randn(b, h, h, cout, device=dev)
    like = cl(torch.empty(cout, cin, 3, 3, device=dev))
    geom = ct.geom_wgrad(b, h, h, cin, cout, 3, 1, 1)
    for _ in range(2):
        ct.conv_wgrad_tc(g, x, geom, like)
    torch.cuda.synchronize()
for (b, h, cin, cout) in [(4, 8, 512, 512), (4, 16, 512, 512)]:
    x = torch.randn(b, h, h, cin, device=dev)
    w = cl(torch.randn(cout, cin, 3, 3, device=dev))
    geom = ct.geom_conv(b, h, h, cin, cout, 3, 1, 1)
    for _ in range(2):
        ct.conv_tc_nhwc(x, w, geom)
    torch.cuda.synchronize()
