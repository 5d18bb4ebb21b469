import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rick_b200 import conv_tc as ct
dev = "cuda"
b, h, cin, cout = 16, 256, 128, 128
x = torch.randn(b, h, h, cin, device=dev)
wt = torch.randn(9, cout, cin, device=dev) / math.sqrt(cin * 9)
demod = torch.rand(b, cout, device=dev) + 0.5
noise = torch.randn(b, h, h, device=dev)
nw = torch.randn(1, device=dev)
bias = torch.randn(cout, device=dev)
sn = torch.randn(b, cout, device=dev)
geom = ct.geom_conv(b, h, h, cin, cout, 3, 1, 1)
for _ in range(3):
    ct.conv_tc_nhwc(x, wt, geom, demod=demod, noise=noise, noise_weight=nw, bias=bias, act=True, s_next=sn, want_out2=True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ct.conv_tc_nhwc(x, wt, geom, demod=demod, noise=noise, noise_weight=nw, bias=bias, act=True, s_next=sn, want_out2=True)
e.record(); e.synchronize()
print("dual-output styled epilogue:", s.elapsed_time(e) / 5, "ms")
s.record()
for _ in range(5):
    ct.conv_tc_nhwc(x, wt, geom)
e.record(); e.synchronize()
print("plain single output:", s.elapsed_time(e) / 5, "ms   (MMA-bound would be", 2 * b * h * h * cin * cout * 9 / 820e12 * 1e3, "ms)")
