"""Which ATen ops (with input shapes) own the device time of a path-length / R1 iteration -- guidance only."""
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
draws = DrawStream(1, dev, cpu_seeded=False)
A.fisher_round(torch.randn(5, 512, device=dev), shots[:5])
for i in (4, 16, 4):
    A.step(i, shots[:2], draws)
torch.cuda.synchronize()
which = int(sys.argv[1]) if len(sys.argv) > 1 else 4          # 4: path-length iteration, 16: R1 + path, 5: plain
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    A.step(which, shots[:2], draws)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45,
                                                         max_name_column_width=60, max_shapes_column_width=110))
