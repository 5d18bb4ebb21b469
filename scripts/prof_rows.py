"""ncu target: the three up=1 shapes of the op sweep (bulk-staged rows kernel) + the NHWC blur."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rick_b200 import op, conv_tc as ct
t = torch.tensor([1., 3., 3., 1.], device="cuda")
taps = torch.outer(t, t) / 64
n, c = 32, 512
xb = torch.randn(n, c, 129, 129, device="cuda")
x = torch.randn(n, c, 128, 128, device="cuda")
xn = torch.randn(n, 129, 129, c, device="cuda")
for _ in range(2):
    op.upfirdn2d(xb, taps * 4, pad=(1, 1))
    op.upfirdn2d(x, taps, pad=(2, 2))
    op.upfirdn2d(x, taps, down=2, pad=(1, 1))
    ct.blur_nhwc(xn, taps * 4, (1, 1))
torch.cuda.synchronize()
