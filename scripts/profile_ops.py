"""Which ATen operators (with input shapes) the adaptation iteration still spends device time in: torch.profiler with
record_shapes, grouped by (op, shapes).  Guidance for glue fusion; the judged evidence is the ncu data under profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from rick_b200.adapt import AdaptConfig, DrawStream, RickAdapter

dev = torch.device("cuda", 0)
cfg = AdaptConfig(size=256, batch=2, warmup_iter=0)
G, D, Ge, De = bench.build_networks(256, dev)
A = RickAdapter(cfg, G, D, Ge, De, fused_generator=True)
shots = bench.synthetic_shots(10, 256).to(dev)
lat = torch.randn(5, 512, device=dev)
draws = DrawStream(1, dev, cpu_seeded=False)
A.fisher_round(lat, shots[:5])
for i in range(1, 4):
    A.step(i, shots[:2], draws)
torch.cuda.synchronize()
steps = [5, 6, 7, 9]          # plain iterations
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for i in steps:
        A.step(i, shots[:2], draws)
    torch.cuda.synchronize()
n = len(steps)
rows = prof.key_averages(group_by_input_shape=True)
sel = [r for r in rows if r.self_device_time_total > 0]
tot = sum(r.self_device_time_total for r in sel)
print(f"total self device time {tot / 1e3 / n:.2f} ms / iteration")
for r in sorted(sel, key=lambda r: -r.self_device_time_total)[:90]:
    print(f"{100 * r.self_device_time_total / tot:6.2f} %  {r.self_device_time_total / 1e3 / n:7.3f} ms/it  x{r.count / n:6.1f}  "
          f"{r.key[:48]:48s} {str(r.input_shapes)[:110]}")
