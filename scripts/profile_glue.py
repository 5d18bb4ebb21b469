"""Which module-level lines launch the remaining ATen kernels of one adaptation iteration: torch.profiler with shapes and
Python stacks, eager executor, grouped by (op, input shapes, innermost rick_b200 frame).  Guidance only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import defaultdict
from torch.profiler import profile, ProfilerActivity
import bench
from rick_b200.adapt import AdaptConfig, DrawStream, RickAdapter

dev = torch.device("cuda", 0)
cfg = AdaptConfig(size=256, batch=2, warmup_iter=0)
G, D, Ge, De = bench.build_networks(256, dev)
A = RickAdapter(cfg, G, D, Ge, De, fused_generator=True)
shots = bench.synthetic_shots(10, 256).to(dev)
draws = DrawStream(1, dev, cpu_seeded=False)
A.fisher_round(torch.randn(5, 512, device=dev), shots[:5])
for i in range(1, 4):
    A.step(i, shots[:2], draws)
torch.cuda.synchronize()
steps = [5, 6, 7, 9]
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    for i in steps:
        A.step(i, shots[:2], draws)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if not ev.name.startswith("aten::") or ev.self_device_time_total <= 0:
        continue
    frame = ""
    for fr in (ev.stack or []):
        if "rick_b200" in fr and "_lib" not in fr:
            frame = fr.split("rick_b200/")[-1]
            break
    shapes = str(ev.input_shapes)[:70]
    agg[(ev.name, shapes, frame[:60])][0] += 1
    agg[(ev.name, shapes, frame[:60])][1] += ev.self_device_time_total
n = len(steps)
tot = sum(v[1] for v in agg.values())
print(f"ATen self device time {tot / 1e3 / n:.3f} ms / iteration")
for (name, shapes, frame), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{t / n:8.1f} us/it x{c / n:5.1f}  {name:28s} {shapes:70s} {frame}")
