"""Kernel-time breakdown of the path-length sub-step (G forward at batch 1, grad w.r.t. the latents with create_graph, second
backward, optimiser step) -- eager, torch.profiler.  Guidance only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import defaultdict
from torch.autograd import DeviceType
from torch.profiler import profile, ProfilerActivity
import bench
from rick_b200.adapt import AdaptConfig
from rick_b200.graphs import GraphedRickAdapter

dev = torch.device("cuda", 0)
cfg = AdaptConfig(size=256, batch=2, warmup_iter=0)
G, D, Ge, De = bench.build_networks(256, dev)
A = GraphedRickAdapter(cfg, G, D, Ge, De, fused_generator=True)
which = sys.argv[1] if len(sys.argv) > 1 else "path"
body = getattr(A, "_body_" + which)
A._real.copy_(bench.synthetic_shots(10, 256).to(dev)[:2])
for _ in range(3):
    body()
torch.cuda.synchronize()
n = 4
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for _ in range(n):
        body()
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == DeviceType.CUDA:
        agg[ev.name][0] += 1
        agg[ev.name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"{which}: total kernel time {tot / 1e3 / n:.2f} ms over {sum(v[0] for v in agg.values()) // n} kernels")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{100 * t / tot:6.2f} %  {t / n:8.1f} us  x{c / n:6.1f}  {k[:120]}")
ops = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.name.startswith("aten::") and ev.self_device_time_total > 0:
        ops[(ev.name, str(ev.input_shapes)[:80])][0] += 1
        ops[(ev.name, str(ev.input_shapes)[:80])][1] += ev.self_device_time_total
print("--- ATen ops by self device time")
for (name, shp), (c, t) in sorted(ops.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t / n:8.1f} us x{c / n:5.1f}  {name:26s} {shp}")
