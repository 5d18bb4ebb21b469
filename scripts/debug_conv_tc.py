"""Layout diagnostics for rick_conv_tc (run on the GPU box).  Each probe is small and prints enough to tell a
descriptor / swizzle / TMEM-mapping mistake from a pipeline mistake."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.nn import functional as F
from rick_b200 import conv_tc as ct

torch.backends.cudnn.allow_tf32 = False
dev = "cuda"


def run(name, b, h, w, cin, cout, k, stride, pad, x, wgt):
    geom = ct.geom_conv(b, h, w, cin, cout, k, stride, pad)
    got = ct.conv_tc_nhwc(x.permute(0, 2, 3, 1).contiguous(), ct.pack_weight(wgt), geom).permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    want = F.conv2d(x, wgt, stride=stride, padding=pad)
    err = (got - want).abs().max().item() / max(want.abs().max().item(), 1e-30)
    print(f"{name}: rel err {err:.3e}  (want max {want.abs().max().item():.3f}, got max {got.abs().max().item():.3f})", flush=True)
    return got, want, err


# probe 1: identity weights on the first 32 output channels -> out[p, co] = x[p, co]
b, h, w, cin, cout = 1, 16, 16, 32, 128
x = (torch.arange(h * w, device=dev).float().view(1, 1, h, w) + torch.arange(cin, device=dev).float().view(1, cin, 1, 1) / 100)
wgt = torch.zeros(cout, cin, 1, 1, device=dev)
for c in range(cin):
    wgt[c, c] = 1.0
got, want, err = run("identity 1x1 K=32", b, h, w, cin, cout, 1, 1, 0, x, wgt)
if err > 1e-3:
    g = got[0].permute(1, 2, 0).reshape(h * w, cout)
    wn = want[0].permute(1, 2, 0).reshape(h * w, cout)
    for p in (0, 1, 2, 7, 8, 9, 17, 255):
        print(" pixel", p, "got", [round(v, 2) for v in g[p, :12].tolist()], "want", [round(v, 2) for v in wn[p, :12].tolist()])
    print(" channels 32..40 of pixel 3 (should be 0):", g[3, 32:40].tolist())

# probe 2: random, one k-block
x = torch.randn(1, 32, 16, 16, device=dev)
wgt = torch.randn(128, 32, 1, 1, device=dev) / math.sqrt(32)
run("random 1x1 K=32", 1, 16, 16, 32, 128, 1, 1, 0, x, wgt)
# probe 3: K loop
x = torch.randn(1, 256, 16, 16, device=dev)
wgt = torch.randn(128, 256, 1, 1, device=dev) / 16
run("random 1x1 K=256", 1, 16, 16, 256, 128, 1, 1, 0, x, wgt)
# probe 4: several tiles / cout tiles / batch
x = torch.randn(2, 128, 32, 32, device=dev)
wgt = torch.randn(256, 128, 1, 1, device=dev) / math.sqrt(128)
run("random 1x1 multi-tile", 2, 32, 32, 128, 256, 1, 1, 0, x, wgt)
# probe 5: 3x3 taps + padding
x = torch.randn(2, 64, 16, 16, device=dev)
wgt = torch.randn(128, 64, 3, 3, device=dev) / math.sqrt(64 * 9)
run("random 3x3", 2, 16, 16, 64, 128, 3, 1, 1, x, wgt)
# probe 6: stride 2
x = torch.randn(2, 64, 17, 17, device=dev)
run("random 3x3 stride 2", 2, 17, 17, 64, 128, 3, 2, 0, x, wgt)

# timing at a generator-like shape (batch 16, 64x64, 512 -> 512)
b, h, cin, cout = 16, 64, 512, 512
x = torch.randn(b, h, h, cin, device=dev)
wt = torch.randn(9, cout, cin, device=dev) / math.sqrt(cin * 9)
geom = ct.geom_conv(b, h, h, cin, cout, 3, 1, 1)
for _ in range(3):
    ct.conv_tc_nhwc(x, wt, geom)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    ct.conv_tc_nhwc(x, wt, geom)
e.record(); e.synchronize()
ms = s.elapsed_time(e) / 10
fl = 2 * b * h * h * cin * cout * 9
print(f"conv_tc 3x3 b{b} {h}x{h} {cin}->{cout}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s (tf32)")
xc = x.permute(0, 3, 1, 2).contiguous()
wc = wt.view(3, 3, cout, cin).permute(2, 3, 0, 1).contiguous()
torch.backends.cudnn.allow_tf32 = True
for _ in range(3):
    F.conv2d(xc, wc, padding=1)
s.record()
for _ in range(10):
    F.conv2d(xc, wc, padding=1)
e.record(); e.synchronize()
ms = s.elapsed_time(e) / 10
print(f"cudnn tf32 same shape: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s")
