import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rick_b200 import conv_tc as ct
dev = "cuda"
for (b, h, cin, cout) in [(16, 128, 256, 128), (16, 64, 512, 256)]:
    x = torch.randn(b, h, h, cin, device=dev)
    wt = torch.randn(9, cout, cin, device=dev) / math.sqrt(cin * 9)
    geom = ct.geom_conv_transpose_s2(b, h, h, cin, cout)
    for _ in range(3):
        ct.conv_tc_nhwc(x, wt, geom)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        ct.conv_tc_nhwc(x, wt, geom)
    e.record(); e.synchronize()
    ms = s.elapsed_time(e) / 5
    fl = 2 * b * h * h * cin * cout * 9
    print(f"convT b{b} {h}->{2*h+1} {cin}->{cout}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s; MMA-bound {fl/820e12*1e3:.3f} ms; write-bound {b*(2*h+1)**2*cout*4/6e12*1e3:.3f} ms")
