"""Sample-generation throughput of the fused (tcgen05) generator at batch 64 + per-kernel breakdown (torch.profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import defaultdict
from torch.autograd import DeviceType
from torch.profiler import profile, ProfilerActivity
import bench
from rick_b200.fused import FusedGenerator

dev = torch.device("cuda", 0)
G, D, Ge, De = bench.build_networks(256, dev)
print("samples/s fused:", round(bench.g_samples_per_s(Ge, dev, fused=True), 1), " module path:", round(bench.g_samples_per_s(Ge, dev, fused=False), 1))
FG = FusedGenerator(Ge)
z = torch.randn(64, 512, device=dev)
for _ in range(2):
    FG([z])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        FG([z])
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == DeviceType.CUDA:
        agg[ev.name][0] += 1
        agg[ev.name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"kernel time per batch of 64: {tot / 3e3:.2f} ms")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{100 * t / tot:6.2f} %  {t / 3e3:8.3f} ms  x{c / 3:5.1f}  {k[:110]}")
# per-layer conv_tc timings
evs = [ev for ev in prof.events() if ev.device_type == DeviceType.CUDA and "conv_tc" in ev.name]
per = len(evs) // 3
print("conv_tc launches in forward order (us):", [round(e.device_time) for e in evs[:per]])
