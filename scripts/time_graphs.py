"""Replay time of each captured sub-step graph of the adaptation iteration (D / R1 / G / path / EMA / Fisher image)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rick_b200.adapt import AdaptConfig
from rick_b200.graphs import GraphedRickAdapter

dev = torch.device("cuda", 0)
cfg = AdaptConfig(size=256, batch=2, warmup_iter=0)
G, D, Ge, De = bench.build_networks(256, dev)
A = GraphedRickAdapter(cfg, G, D, Ge, De, fused_generator=True)
shots = bench.synthetic_shots(10, 256).to(dev)
A._real.copy_(shots[:2])
A.prepare()
for key in ("d", "r1", "g", "path", "ema", "fisher"):
    for _ in range(3):
        A._graphs[key].replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        A._graphs[key].replay()
    e.record()
    e.synchronize()
    print(f"{key:7s} {s.elapsed_time(e) / 10:7.3f} ms   ({A._graph_launches[key]} rick_b200 kernels)")
