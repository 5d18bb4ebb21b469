"""Per-layer timing of the three convolution primitives (fprop / dgrad / wgrad) on the shapes of the 256 px adaptation
iteration: the tcgen05 kernels of this package against the cuDNN TF32 kernels torch dispatches to for the same call
(the bar SURVEY section 2a sets for the convolution).  CUDA events on the launching stream, L2 flushed between launches.

    python scripts/bench_convs.py [--json out.json] [--quick]
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rick_b200 import conv

# (name, B, H, W, Cin, Cout, k, stride, pad, transposed)
D_LAYERS = [("D conv1 256", 4, 256, 256, 128, 128, 3, 1, 1, False), ("D conv2 256>128", 4, 257, 257, 128, 256, 3, 2, 0, False),
            ("D skip 256>128", 4, 255, 255, 128, 256, 1, 2, 0, False),
            ("D conv1 128", 4, 128, 128, 256, 256, 3, 1, 1, False), ("D conv2 128>64", 4, 129, 129, 256, 512, 3, 2, 0, False),
            ("D conv1 64", 4, 64, 64, 512, 512, 3, 1, 1, False), ("D conv2 64>32", 4, 65, 65, 512, 512, 3, 2, 0, False),
            ("D conv1 32", 4, 32, 32, 512, 512, 3, 1, 1, False), ("D conv2 32>16", 4, 33, 33, 512, 512, 3, 2, 0, False),
            ("D conv1 16", 4, 16, 16, 512, 512, 3, 1, 1, False), ("D conv1 8", 4, 8, 8, 512, 512, 3, 1, 1, False),
            ("D final 4", 4, 4, 4, 544, 512, 3, 1, 1, False)]
G_LAYERS = [("G conv 4", 2, 4, 4, 512, 512, 3, 1, 1, False), ("G up 4>8", 2, 4, 4, 512, 512, 3, 2, 0, True),
            ("G conv 16", 2, 16, 16, 512, 512, 3, 1, 1, False), ("G up 16>32", 2, 16, 16, 512, 512, 3, 2, 0, True),
            ("G conv 64", 2, 64, 64, 512, 512, 3, 1, 1, False), ("G up 64>128", 2, 64, 64, 512, 256, 3, 2, 0, True),
            ("G conv 128", 2, 128, 128, 256, 256, 3, 1, 1, False), ("G up 128>256", 2, 128, 128, 256, 128, 3, 2, 0, True),
            ("G conv 256", 2, 256, 256, 128, 128, 3, 1, 1, False)]


def timed(fn, flush, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sum(ts) / len(ts)


def run(layers, flush):
    rows = []
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    for name, b, h, w, cin, cout, k, stride, pad, tr in layers:
        cfg = (stride, pad, tr)
        x = cl(torch.randn(b, cin, h, w, device="cuda"))
        wt = cl(torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k))
        y = conv._fprop(x, wt, cfg)
        g = cl(torch.randn_like(y))
        flops = 2 * b * cin * cout * k * k * (h * w if tr else y.shape[2] * y.shape[3])
        row = {"layer": name, "gflop": flops / 1e9}
        for backend in ("tc", "cudnn"):
            conv._FORCE = "cudnn" if backend == "cudnn" else ""
            row[f"fprop_{backend}_us"] = timed(lambda: conv._fprop(x, wt, cfg), flush)
            row[f"dgrad_{backend}_us"] = timed(lambda: conv._dgrad(g, wt, x, cfg), flush)
            row[f"wgrad_{backend}_us"] = timed(lambda: conv._wgrad(g, x, wt, cfg), flush)
        conv._FORCE = ""
        rows.append(row)
        print(f"{name:16s} {row['gflop']:7.1f} GF | " + " | ".join(
            f"{op} {row[op + '_tc_us']:7.1f} us ({row['gflop'] / row[op + '_tc_us'] * 1e3:5.0f} TF/s) cudnn {row[op + '_cudnn_us']:7.1f}"
            for op in ("fprop", "dgrad", "wgrad")), flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    layers = D_LAYERS + G_LAYERS
    if args.quick:
        layers = [layers[i] for i in (0, 3, 5, 9, 11, 12, 16, 19)]
    rows = run(layers, flush)
    tot = {k: sum(r[k] for r in rows) for k in rows[0] if k.endswith("_us")}
    print("totals (us):", {k: round(v, 1) for k, v in tot.items()})
    if args.json:
        json.dump({"rows": rows, "totals_us": tot}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
