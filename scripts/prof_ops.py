"""Tiny driver for ncu captures of the memory-bound kernels at the BASELINE op-sweep shapes."""
import sys
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from rick_b200 import op

dev = "cuda"
t = torch.tensor([1., 3., 3., 1.], device=dev)
taps4, taps1 = torch.outer(t, t) / 16, torch.outer(t, t) / 64
x = torch.randn(32, 512, 128, 128, device=dev)
b = torch.randn(512, device=dev)
for _ in range(3):
    y = op.upfirdn2d(x, taps4, up=2, pad=(2, 1))          # (32,512,256,256)
    z = op.upfirdn2d(x, taps1, pad=(2, 2))                # D blur
    d = op.upfirdn2d(x, taps1, down=2, pad=(1, 1))
    xr = x.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    a = op.fused_leaky_relu(xr, br)
    a.backward(torch.ones_like(a))
xb = torch.randn(32, 512, 129, 129, device=dev)
for _ in range(3):
    w = op.upfirdn2d(xb, taps4, pad=(1, 1))
torch.cuda.synchronize()
print("ok")
