#!/usr/bin/env python
"""bench.py -- RICK adaptation throughput on B200 (BASELINE.json metric: adapt iters/s, 256 px, batch 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (configs[1] of BASELINE.json, named in config.workload): the RICK FFHQ->Babies-shaped loop -- StyleGAN2
256 px G + D (random init, FFHQ-256 architecture), batch 2 per GPU, 10 synthetic shots, Fisher round every 50
iterations on 5 images, fisher_quantile 40, prune_quantile 0.1, R1 every 16, path-length every 4, mixing 0.9.
A "step" is one adaptation iteration (D step [+R1], G step [+path-length], masks, Adam, EMA).  Iteration indices run
0..W-1 untimed and W..W+K-1 timed with warmup_iter = 0, so a Fisher round falls on iterations 0, 50, ...; the JSON
line says how many Fisher / R1 / path-length iterations were inside the timed region.

  value   whole-job iters/s with the real images already resident in HBM and no loss read-back
  e2e     the same loop through the public API with HOST inputs: every step copies its real images from pinned host
          memory and reads the step's losses back to the host
  N > 1   one process per GPU (torchrun), DDP-style gradient all-reduce over NCCL, per-GPU batch 2 (weak scaling);
          value = N x (optimiser iterations/s): every rank runs its own batch-2 iteration, so this is batch-2
          iterations processed per second over the whole job (``optimizer_iterations_per_s`` is the un-multiplied rate)
  roofline  the dominant kernel of the TIMED step, conv_tc_kernel (forward + data-gradient convolutions: 29 % of the
            iteration's kernel time in profiles/r02*_launches.md): the algorithmic TF32 flops of all its launches in one
            plain adaptation iteration / the sum of their durations, each launch timed live with CUDA events on the
            launching stream (L2 flushed between launches), against the TF32 matmul peak MEASURED IN THIS RUN (cuBLAS
            8192^3, burst).  ``roofline.step`` is the whole iteration: algorithmic conv flops per iteration / ms_per_step.
            ``rooflines`` adds the weight-gradient kernel, the sample-generation shape (b64 64x64 512->512) and the
            memory-bound headline kernel (upfirdn2d, BASELINE configs[3] shape) against hbm_gbs; extra.conv_vs_cudnn
            times every layer shape on cuDNN TF32 as well (SURVEY 2a's bar); extra.op_sweep is all of configs[3]
  cpu_baseline / --impl reference   the oracle port of the same iteration on the box's host cores (the reference's
          CPU path: upfirdn2d_native + native leaky-ReLU semantics), bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "rick_adapt_ffhq256_b2_10shot_fisher50x5_q40_p0.1"
METRIC = "adapt_iters_per_s"
UNIT = "iters/s"


# ------------------------------------------------------------------------------------------------ helpers

class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def measure_tf32_peak(device, n=8192, reps=10):
    """Dense TF32 matmul throughput of THIS GPU, measured like MEASURED_PEAKS.json's bf16 entry (torch.matmul n^3, best
    of ``reps``, CUDA events): the denominator of every tensor-bound fraction below."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=device)
        b = torch.randn(n, n, device=device)
        for _ in range(3):
            a @ b
        best = float("inf")
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            a @ b
            e.record()
            e.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2.0 * n ** 3 / (best / 1e3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def build_networks(size, device, seed=1):
    """Random init, FFHQ-256 architecture (no checkpoint is available offline); G and G_ema start identical, as after
    loading one source checkpoint into both (train:876-879)."""
    from rick_b200 import stylegan2 as sg
    torch.manual_seed(seed)
    G = sg.Generator(size, 512, 8)
    D = sg.Discriminator(size)
    Ge = sg.Generator(size, 512, 8)
    De = sg.Discriminator(size)
    Ge.load_state_dict(G.state_dict())
    De.load_state_dict(D.state_dict())
    return G.to(device), D.to(device), Ge.to(device), De.to(device)


def synthetic_shots(n, size, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.clamp(torch.randn(n, 3, size, size, generator=g) * 0.5, -1, 1)


# ------------------------------------------------------------------------------------------------ our arm

def run_ours(args):
    from rick_b200 import _lib
    from rick_b200 import dist as rdist
    from rick_b200.adapt import AdaptConfig, DrawStream, RickAdapter, generate_samples

    rank, world, local_rank = rdist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in rick_b200)"
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    lib = _lib.lib()
    # default: BASELINE configs[1] (the metric's configuration).  --workload ffhq1024: configs[4], the 1024 px architecture
    # at per-GPU batch 8 under DDP; no Fisher round in it (the reference hard-codes 256 px there, train:237, 281, 336)
    big = args.workload == "ffhq1024"
    size, batch = (1024, 8) if big else (256, 2)
    n_shots = 16 if big else 10
    workload = "stylegan2_ffhq1024_adapt_b8_per_gpu_ddp_no_fisher" if big else WORKLOAD
    if big:
        args.quick = True
    cfg = AdaptConfig(size=size, batch=batch, warmup_iter=0, fisher_freq=10 ** 9 if big else 50, num_fisher_img=5, fisher_quantile=40.0,
                      prune_quantile=0.1, d_reg_every=16, g_reg_every=4, mixing=0.9, lr=0.002)
    G, D, Ge, De = build_networks(size, device, seed=1)       # identical weights on every rank (DDP)
    mode = args.mode
    if mode == "auto":
        mode = "graphs"
    if mode == "graphs":
        from rick_b200.graphs import GraphedRickAdapter
        adapter = GraphedRickAdapter(cfg, G, D, Ge, De, fused_generator=True)
    else:
        adapter = RickAdapter(cfg, G, D, Ge, De, fused_generator=True)
    shots_host = synthetic_shots(n_shots, size, seed=100 + rank).pin_memory()
    shots_dev = shots_host.to(device)
    fisher_lat = torch.randn(cfg.num_fisher_img, 512, generator=torch.Generator().manual_seed(7)).to(device)
    draws = DrawStream(1234 + rank, device, cpu_seeded=False)

    def iteration(i, e2e):
        if i % cfg.fisher_freq == 0 and not big:
            adapter.fisher_round(fisher_lat, shots_dev[:cfg.num_fisher_img])
        j = (i * batch) % n_shots
        if e2e:
            real = shots_host[j:j + batch].to(device, non_blocking=True)          # H2D from pinned memory
        else:
            real = shots_dev[j:j + batch]
        out = adapter.step(i, real, draws)
        if e2e:
            host = torch.stack([out["d"], out["g"]]).to("cpu")                      # D2H read of the step's losses
            return host
        return None

    def timed(first, k, e2e):
        rdist.barrier()
        torch.cuda.synchronize()
        n0 = lib.rick_launch_count() + getattr(adapter, "replayed_launches", 0)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.cuda_profiler and not e2e:
            torch.cuda.profiler.start()           # ncu --profile-from-start off: capture exactly the timed steps
        start.record()
        for i in range(first, first + k):
            iteration(i, e2e)
        end.record()
        torch.cuda.synchronize()
        if args.cuda_profiler and not e2e:
            torch.cuda.profiler.stop()
        rdist.barrier()
        ms = start.elapsed_time(end)
        if world > 1:
            t = torch.tensor([ms], device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        # kernels of this package executed in the region: direct launches + kernel nodes of replayed CUDA graphs
        return ms, lib.rick_launch_count() + getattr(adapter, "replayed_launches", 0) - n0

    W, K = args.warmup, args.steps
    try:
        if mode == "graphs":                      # capture every graph before iteration 0
            adapter._real.copy_(shots_dev[:batch])
            adapter.prepare(fisher=not big)       # capture leaves weights / optimiser state / RNG untouched
        for i in range(W):
            iteration(i, False)
    except Exception as exc:                       # graph capture refused: fall back to the eager executor, say so
        if mode != "graphs" or args.mode == "graphs":
            raise
        sys.stderr.write(f"bench: CUDA-graph capture failed ({type(exc).__name__}: {exc}); falling back to eager\n")
        torch.cuda.synchronize()
        mode = "eager (graph capture failed)"
        G, D, Ge, De = build_networks(size, device, seed=1)
        adapter = RickAdapter(cfg, G, D, Ge, De, fused_generator=True)
        for i in range(W):
            iteration(i, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(W, K, False)
    ms_e2e, _ = timed(W + K, K, True)
    clocks = sampler.stop() if rank == 0 else None

    def count(first, k, every):
        return sum(1 for i in range(first, first + k) if i % every == 0)

    value = world * K / (ms / 1e3)
    e2e_value = world * K / (ms_e2e / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "optimizer_iterations_per_s": K / (ms / 1e3),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload, "size": size, "batch_per_gpu": batch, "global_batch": batch * world,
                   "parallelism": f"dp{world}", "conv_math": "tf32 (fp32 storage, fp32 accumulate)",
                   "value_is": "rank-iterations per second: N ranks x optimiser iterations/s, each rank on its own "
                               f"batch-{batch} shots (weak scaling)",
                   "exec_mode": mode,
                   "timing": "inputs regenerated on device every step (fresh latents/noise); activations + weights "
                             "(~1.5 GB/iter) exceed L2",
                   "fisher_rounds_in_timed": count(W, K, cfg.fisher_freq), "r1_iters_in_timed": count(W, K, 16),
                   "path_iters_in_timed": count(W, K, 4)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": batch * 3 * size * size * 4,
                "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / K},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }

    extra = {}
    # Fisher round (5 images sharded over ranks + one all-reduce of the grad^2 accumulators + masks), every N
    rdist.barrier()
    if not big:
        fisher_ms = time_fisher_round(adapter, fisher_lat, shots_dev)
        if world > 1:
            t = torch.tensor([fisher_ms], device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            fisher_ms = float(t.item())
        extra["fisher_round_ms"] = fisher_ms
    extra["samples_per_s"] = value * batch
    if not args.quick or args.samplegen:
        # BASELINE config 3 as specified: 5000 samples at batch 64, sharded over the ranks (gan_training/eval.py:31-46),
        # every image copied to host memory, feature statistics all-reduced once at the end
        from rick_b200 import sample as rsample
        extra["samplegen_5000"] = rsample.run(Ge, 5000, 64, rank, world, seed=1000, to_host=True, stats=True)
        extra["samplegen_5000_device_only"] = rsample.run(Ge, 5000, 64, rank, world, seed=1000, to_host=False,
                                                          stats=False)["samples_per_s"]
    if rank == 0 and not args.quick:
        # the micro-benchmarks behind the roofline objects and the op sweep: each guarded, so that a failure in one of them
        # is reported in the line instead of costing the line (the timed numbers above are already final)
        def guarded(name, fn):
            try:
                return fn()
            except Exception as exc:                      # noqa: BLE001
                extra.setdefault("errors", {})[name] = f"{type(exc).__name__}: {exc}"[:300]
                torch.cuda.synchronize()
                return None
        tf32_peak = guarded("tf32_peak", lambda: measure_tf32_peak(device)) or load_peaks()["bf16_tflops"] / 2
        prof = guarded("conv_step_profile", lambda: conv_step_profile(device, tf32_peak, ms / K, with_cudnn=True))
        samplegen_roof = guarded("roofline_conv_tc", lambda: roofline_conv_tc(device, tf32_peak))
        up_roof = guarded("roofline_upfirdn2d", lambda: roofline_upfirdn2d(device))
        if prof is not None:
            line["roofline"] = prof["roofline"]
            extra["conv_vs_cudnn"] = prof["table"]
        else:                                             # keep the contract's key: the next tensor-bound / memory-bound entry
            line["roofline"] = samplegen_roof or up_roof
        line["rooflines"] = [r for r in ((prof or {}).get("roofline"), (prof or {}).get("wgrad"), samplegen_roof, up_roof)
                             if r is not None]
        extra["tf32_peak_measured_tflops"] = tf32_peak
        extra["op_sweep"] = guarded("op_sweep", lambda: op_sweep(device))
        extra["g_samples_per_s_b64_per_gpu"] = guarded("g_samples", lambda: g_samples_per_s(Ge, device, fused=True))
        extra["g_samples_per_s_b64_per_gpu_library_conv_module_path"] = guarded(
            "g_samples_library", lambda: g_samples_per_s(Ge, device, fused=False, library=True))
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = guarded("cpu_baseline", lambda: cpu_baseline(max_seconds=40.0))
    if world > 1:
        rdist.barrier()
    line["extra"] = extra
    if rank == 0:
        emit(line)
    if world > 1:
        # Tearing down a NCCL process group while CUDA graphs that recorded its collectives are still alive can block
        # (observed in round 1: the JSON line was out, the process never exited).  Everything is flushed: leave hard.
        torch.cuda.synchronize()
        rdist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


_JSON_OUT = None


def emit(line: dict):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _traffic(kernel):
    """DRAM bytes per launch of ``kernel`` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
    except Exception:
        return None


def _flush_l2(buf):
    buf.zero_()          # 256 MB write > 126 MB L2


def _time_kernel(fn, flush_buf, iters=10, warm=3, graph=True):
    """average device time of fn() in ms: CUDA events on the launching stream, L2 flushed before every launch.  The call
    is replayed from a CUDA graph (as the timed iteration replays it), so the Python / ctypes / autograd dispatch of
    ``fn`` -- 20-50 us, more than a small-map kernel takes -- is not inside the events (``graph=False``: eager launch,
    for calls that cannot be captured)."""
    for _ in range(warm):
        fn()
    g = None
    if graph:
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()                            # noqa: F841  (outputs live in the graph's pool until g is dropped)
        g.replay()
    times = []
    for _ in range(iters):
        _flush_l2(flush_buf)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay() if g is not None else fn()
        e.record()
        e.synchronize()
        times.append(s.elapsed_time(e))
    del g
    return sum(times) / len(times)


def roofline_upfirdn2d(device):
    """upfirdn2d (up=2, 4x4 blur), (32,512,128,128) -> (32,512,256,256) fp32: algorithmic bytes = 4*N*C*(HinWin+HoutWout)."""
    from rick_b200 import op
    peaks = load_peaks()
    n, c, r = 32, 512, 128
    x = torch.randn(n, c, r, r, device=device)
    taps = torch.tensor([1., 3., 3., 1.], device=device)
    taps = torch.outer(taps, taps) / 64 * 4
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    ms = _time_kernel(lambda: op.upfirdn2d(x, taps, up=2, pad=(2, 1)), flush)
    bytes_alg = 4 * n * c * (r * r + 4 * r * r)
    achieved = bytes_alg / (ms / 1e3) / 1e9
    return {"kernel": "upfirdn2d_direct<up=2>", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": _traffic("upfirdn2d_direct<up=2>"),
            "peak_source": peaks["source"],
            "ms_per_launch": ms, "algorithmic_bytes": bytes_alg}


def op_sweep(device):
    """BASELINE configs[3], all of it: achieved ALGORITHMIC GB/s of the memory-bound ops on (32, 512, r, r) tensors --
    upfirdn2d up = 2 (forward, and its gradient = the decimating adjoint), the blur after the transposed convolution,
    D's two blurs, and fused_leaky_relu forward / backward (backward includes grad_bias) for r = 4 ... 128 (R up to 256),
    fp32 and bf16 storage -- plus the 12 x 12 augmentation taps, the optimiser step and the channels-last variants the
    adaptation loop runs.  Each call is replayed from a CUDA graph between the events, L2 flushed before it."""
    from rick_b200 import op
    from rick_b200.op.fused_act import FusedLeakyReLUFunctionBackward as _Bwd
    from rick_b200.op.upfirdn2d import upfirdn2d_adjoint
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    taps = torch.tensor([1., 3., 3., 1.], device=device)
    taps4 = torch.outer(taps, taps) / 64 * 4
    taps1 = torch.outer(taps, taps) / 64
    n, c = 32, 512

    def rate(fn, nbytes, iters=5):
        return nbytes / _time_kernel(fn, flush, iters=iters) / 1e6

    # what this timing method reads for a launch that does (almost) nothing: events + one graph launch around a 16-byte
    # fill.  A 5 MB op (r = 4) cannot show more than 5 MB / floor, whatever its kernel does.
    tiny = torch.zeros(4, device=device)
    out["timing_floor_us"] = _time_kernel(lambda: tiny.add_(1.0), flush, iters=10) * 1e3

    for dtype, tag, esz in ((torch.float32, "", 4), (torch.bfloat16, "bf16_", 2)):
        for r in (4, 8, 16, 32, 64, 128):
            x = torch.randn(n, c, r, r, device=device, dtype=dtype)
            out[f"{tag}upfirdn2d_up2_{r}->{2 * r}"] = rate(lambda: op.upfirdn2d(x, taps4, up=2, pad=(2, 1)),
                                                           esz * n * c * 5 * r * r)
            go = torch.randn(n, c, 2 * r, 2 * r, device=device, dtype=dtype)
            out[f"{tag}upfirdn2d_up2_bwd_{2 * r}->{r}"] = rate(
                lambda: upfirdn2d_adjoint(go, taps4, up=2, pad=(2, 1), in_size=(r, r)), esz * n * c * 5 * r * r)
            del go
            xb = torch.randn(n, c, 2 * r + 1, 2 * r + 1, device=device, dtype=dtype)
            out[f"{tag}blur_{2 * r + 1}->{2 * r}"] = rate(lambda: op.upfirdn2d(xb, taps4, pad=(1, 1)),
                                                         esz * n * c * ((2 * r + 1) ** 2 + 4 * r * r))
            del xb
            out[f"{tag}d_blur_pad22_{r}->{r + 1}"] = rate(lambda: op.upfirdn2d(x, taps1, pad=(2, 2)),
                                                         esz * n * c * (r * r + (r + 1) ** 2))
            out[f"{tag}d_blur_pad11_{r}->{r - 1}"] = rate(lambda: op.upfirdn2d(x, taps1, pad=(1, 1)),
                                                         esz * n * c * (r * r + (r - 1) ** 2))
            del x
        x = torch.randn(n, c, 128, 128, device=device, dtype=dtype)
        out[f"{tag}down2_128->64"] = rate(lambda: op.upfirdn2d(x, taps1, down=2, pad=(1, 1)), esz * n * c * (128 * 128 + 64 * 64))
        del x
        bias = torch.randn(c, device=device, dtype=dtype)
        for r in (4, 8, 16, 32, 64, 128, 256):
            xa = torch.randn(n, c, r, r, device=device, dtype=dtype)
            out[f"{tag}bias_act_fwd_{r}"] = rate(lambda: op.fused_leaky_relu(xa, bias), esz * 2 * xa.numel())
            y = op.fused_leaky_relu(xa, bias)
            go = torch.randn_like(y)
            out[f"{tag}bias_act_bwd_{r}"] = rate(lambda: _Bwd.apply(go, y, 0.2, 2 ** 0.5), esz * 3 * xa.numel())
            del xa, y, go
    # the 12x12 antialiasing filter of non_leaking.py:338, 359 (generic kernel; small, latency-bound tensors)
    sym6 = torch.tensor([0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633,
                         0.4910559419267466, 0.787641141030194, 0.3379294217276218, -0.07263752278646252,
                         -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148],
                        device=device)
    k12 = torch.outer(sym6, sym6)
    xa = torch.randn(8, 3, 282, 282, device=device)
    ms = _time_kernel(lambda: op.upfirdn2d(xa, k12, up=2), flush, iters=5)
    ya = op.upfirdn2d(xa, k12, up=2)
    out["aug_sym6_up2_282"] = 4 * (xa.numel() + ya.numel()) / ms / 1e6
    ms = _time_kernel(lambda: op.upfirdn2d(ya, k12, down=2), flush, iters=5)
    out["aug_sym6_down2_553"] = 4 * (ya.numel() + ya.numel() // 4) / ms / 1e6
    del xa, ya
    # fused masks + Adam + EMA over 33.5 M parameters (G + D trainables are 59 M): 36 B per parameter
    from rick_b200.optim import FusedMaskedAdam
    pw = torch.nn.Parameter(torch.randn(64, 512, 1024, device=device))
    pe = torch.nn.Parameter(pw.detach().clone())
    pw.grad = torch.randn_like(pw)
    fopt = FusedMaskedAdam({"w": pw}, [pw], lr=2e-3, betas=(0.0, 0.99), ema_named={"w": pe}, ema_decay=0.998)
    ms = _time_kernel(lambda: fopt.step(ema=True), flush, iters=5)
    out["adam_mask_ema_33M"] = 36 * pw.numel() / ms / 1e6
    del pw, pe, fopt
    from rick_b200 import conv_tc as ct
    xn = torch.randn(n, 129, 129, c, device=device)                       # NHWC blur after the transposed conv
    ms = _time_kernel(lambda: ct.blur_nhwc(xn, taps4, (1, 1)), flush, iters=5)
    out["blur_nhwc_129->128"] = 4 * n * c * (129 * 129 + 128 * 128) / ms / 1e6
    del xn
    xcl = torch.randn(n, c, 128, 128, device=device).to(memory_format=torch.channels_last)
    bcl = torch.randn(c, device=device)
    ms = _time_kernel(lambda: op.fused_leaky_relu(xcl, bcl), flush, iters=5)
    out["bias_act_fwd_nhwc_128"] = 4 * 2 * xcl.numel() / ms / 1e6
    y = op.fused_leaky_relu(xcl, bcl)
    go = torch.randn_like(y)
    ms = _time_kernel(lambda: _Bwd.apply(go, y, 0.2, 2 ** 0.5), flush, iters=5)
    out["bias_act_bwd_nhwc_128"] = 4 * 3 * xcl.numel() / ms / 1e6
    del xcl, y, go
    return {k: round(v, 1) for k, v in out.items()}


def time_fisher_round(adapter, fisher_lat, shots_dev):
    torch.cuda.synchronize()
    t = time.perf_counter()
    adapter.fisher_round(fisher_lat, shots_dev[:adapter.cfg.num_fisher_img])
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3


def roofline_conv_tc(device, tf32_peak):
    """rick_conv_tc on the sample-generation shape of the 64x64 layers (batch 64, 512 -> 512, 3x3): TF32 tensor-core
    work 2*B*H*W*Cin*Cout*9 flop per launch, CUDA events on the launching stream, L2 flushed between launches."""
    from rick_b200 import conv_tc as ct
    b, h, cin, cout = 64, 64, 512, 512
    x = torch.randn(b, h, h, cin, device=device)
    wt = torch.randn(9, cout, cin, device=device) / (cin * 9) ** 0.5
    geom = ct.geom_conv(b, h, h, cin, cout, 3, 1, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    ms = _time_kernel(lambda: ct.conv_tc_nhwc(x, wt, geom), flush)
    flops = 2 * b * h * h * cin * cout * 9
    achieved = flops / (ms / 1e3) / 1e12
    return {"kernel": "conv_tc_kernel (3x3, b64 64x64 512->512: the sample-generation shape)", "bound": "tensor",
            "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
            "traffic": _traffic("conv_tc_kernel (3x3, b64 64x64 512->512)"),
            "peak_source": "measured in this run: torch.matmul TF32 8192^3, best of 10", "ms_per_launch": ms,
            "algorithmic_flops": flops}


# Convolution launches of ONE plain adaptation iteration (256 px, batch 2): (name, B, H, W, Cin, Cout, k, stride, pad,
# transposed, primitives).  D step: G forward on the tcgen05 executor (no grad), D on the joint fake+real batch of 4 with
# forward / data gradient (not into the image) / weight gradient.  G step: G forward, D forward, data gradients through D
# and G, weight gradients of G's trained layers.
def _iteration_convs():
    d = [("conv1 256", 256, 256, 128, 128, 3, 1, 1), ("conv2 256>128", 257, 257, 128, 256, 3, 2, 0),
         ("skip 256>128", 255, 255, 128, 256, 1, 2, 0), ("conv1 128", 128, 128, 256, 256, 3, 1, 1),
         ("conv2 128>64", 129, 129, 256, 512, 3, 2, 0), ("skip 128>64", 127, 127, 256, 512, 1, 2, 0),
         ("conv1 64", 64, 64, 512, 512, 3, 1, 1), ("conv2 64>32", 65, 65, 512, 512, 3, 2, 0),
         ("skip 64>32", 63, 63, 512, 512, 1, 2, 0), ("conv1 32", 32, 32, 512, 512, 3, 1, 1),
         ("conv2 32>16", 33, 33, 512, 512, 3, 2, 0), ("skip 32>16", 31, 31, 512, 512, 1, 2, 0),
         ("conv1 16", 16, 16, 512, 512, 3, 1, 1), ("conv2 16>8", 17, 17, 512, 512, 3, 2, 0),
         ("skip 16>8", 15, 15, 512, 512, 1, 2, 0), ("conv1 8", 8, 8, 512, 512, 3, 1, 1),
         ("conv2 8>4", 9, 9, 512, 512, 3, 2, 0), ("skip 8>4", 7, 7, 512, 512, 1, 2, 0), ("final 4", 4, 4, 544, 512, 3, 1, 1)]
    g = [("conv 4", 4, 4, 512, 512, False)]
    for r, ci, co in ((4, 512, 512), (8, 512, 512), (16, 512, 512), (32, 512, 512), (64, 512, 256), (128, 256, 128)):
        g += [(f"up {r}>{2 * r}", r, r, ci, co, True), (f"conv {2 * r}", 2 * r, 2 * r, co, co, False)]
    out = []
    for name, h, w, ci, co, k, s_, p_ in d:
        out.append((f"D b4 {name}", 4, h, w, ci, co, k, s_, p_, False, ("fprop", "dgrad", "wgrad")))      # D step
        out.append((f"D b2 {name}", 2, h, w, ci, co, k, s_, p_, False, ("fprop", "dgrad")))               # G step
    for name, h, w, ci, co, tr in g:
        cfg = (3, 2, 0) if tr else (3, 1, 1)
        out.append((f"G b2 {name}", 2, h, w, ci, co, *cfg, tr, ("fprop", "fprop", "dgrad", "wgrad")))     # D step + G step
    return out


def conv_step_profile(device, tf32_peak, ms_per_step, with_cudnn=True):
    """Every convolution launch of one plain adaptation iteration, timed live (CUDA events on the launching stream, L2
    flushed before each launch) on the tcgen05 kernels and -- the bar SURVEY 2a sets -- on cuDNN's TF32 kernels for the
    same call.  Aggregates per kernel: algorithmic flops / summed duration."""
    import math
    from rick_b200 import conv
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    agg = {"fprop": [0.0, 0.0, 0.0, 0, 0.0], "dgrad": [0.0, 0.0, 0.0, 0, 0.0],
           "wgrad": [0.0, 0.0, 0.0, 0, 0.0]}                   # flops, us tc, us cudnn, launches, algorithmic DRAM bytes
    table = []
    prev_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        for name, b, h, w, cin, cout, k, stride, pad, tr, prims in _iteration_convs():
            cfg = (stride, pad, tr)
            x = cl(torch.randn(b, cin, h, w, device=device))
            wt = cl(torch.randn(cout, cin, k, k, device=device) / math.sqrt(cin * k * k))
            y = conv._fprop(x, wt, cfg)
            g = cl(torch.randn_like(y))
            flops = 2.0 * b * cin * cout * k * k * (h * w if tr else y.shape[2] * y.shape[3])
            fns = {"fprop": lambda: conv._fprop(x, wt, cfg), "dgrad": lambda: conv._dgrad(g, wt, x, cfg),
                   "wgrad": lambda: conv._wgrad(g, x, wt, cfg)}
            row = {"layer": name, "gflop": round(flops / 1e9, 2)}
            for prim in dict.fromkeys(prims):
                conv._FORCE = ""
                us = _time_kernel(fns[prim], flush, iters=4, warm=2) * 1e3
                us_lib = None
                if with_cudnn:
                    conv._FORCE = "cudnn"
                    us_lib = _time_kernel(fns[prim], flush, iters=4, warm=2) * 1e3
                    conv._FORCE = ""
                n = prims.count(prim)
                a = agg[prim]
                a[0] += flops * n
                a[4] += 4.0 * (x.numel() + y.numel() + wt.numel()) * n     # each operand read / result written once
                a[1] += us * n
                a[2] += (us_lib or 0.0) * n
                a[3] += n
                row[prim] = [round(us, 1), None if us_lib is None else round(us_lib, 1)]
            table.append(row)
            del x, wt, y, g
    finally:
        conv._FORCE = ""
        torch.backends.cudnn.allow_tf32 = prev_tf32

    def entry(kernel, prims, note):
        fl = sum(agg[p][0] for p in prims)
        us = sum(agg[p][1] for p in prims)
        us_lib = sum(agg[p][2] for p in prims)
        ach = fl / (us * 1e-6) / 1e12
        return {"kernel": kernel, "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": ach / tf32_peak, "traffic": _traffic(kernel),
                "peak_source": "measured in this run: torch.matmul TF32 8192^3, best of 10 (MEASURED_PEAKS.json has "
                               "bf16 only: bf16_tflops / 2 = %.1f)" % (load_peaks()["bf16_tflops"] / 2),
                "traffic_note": "summed DRAM bytes of this kernel's launches in one iteration (profiles/r02*_conv_traffic.md)",
                "algorithmic_bytes_per_iteration": sum(agg[p][4] for p in prims),
                "launches_per_iteration": sum(agg[p][3] for p in prims), "sum_launch_us": us,
                "share_of_step": us * 1e-3 / ms_per_step, "algorithmic_flops_per_iteration": fl,
                "cudnn_tf32_same_launches_us": us_lib if with_cudnn else None, "note": note}
    total_flops = sum(a[0] for a in agg.values())
    roof = entry("conv_tc_kernel", ("fprop", "dgrad"),
                 "all forward + data-gradient convolution launches of one plain iteration (D on the joint batch of 4, "
                 "G and D on batch 2; includes the split-K fold of the small maps); each launch replayed from a CUDA "
                 "graph between the events, as the timed iteration replays it")
    roof["step"] = {"algorithmic_tflop_per_iteration": total_flops / 1e12,
                    "achieved_tflops": total_flops / 1e12 / (ms_per_step / 1e3),
                    "frac": total_flops / 1e12 / (ms_per_step / 1e3) / tf32_peak,
                    "note": "convolution flops of a PLAIN iteration over the measured ms_per_step (the timed steps also "
                            "contain R1 / path-length iterations whose extra work is not counted: conservative)"}
    wg = entry("conv_wgrad_kernel", ("wgrad",), "all weight-gradient launches of one plain iteration, fold kernel included")
    return {"roofline": roof, "wgrad": wg, "table": table}


@torch.no_grad()
def g_samples_per_s(G, device, batches=6, batch=64, fused=False, library=False):
    """``library=True``: the module path with its convolutions forced onto the library (cuDNN) for comparison."""
    from rick_b200 import conv
    from rick_b200.adapt import generate_samples
    G.eval()
    conv._FORCE = "cudnn" if library else ""
    it = generate_samples(G, (batches + 2) * batch, batch, seed=1000, fused=fused)
    next(it), next(it)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    n = 0
    for _, img in it:
        n += img.shape[0]
    e.record()
    e.synchronize()
    conv._FORCE = ""
    return n / (s.elapsed_time(e) / 1e3)


# ------------------------------------------------------------------------------------------------ CPU arm

def _oracle_adapter(threads):
    from oracle import adapt_oracle as ao
    from oracle import synth
    from rick_b200.adapt import AdaptConfig, DrawStream
    torch.set_num_threads(threads)
    size = int(os.environ.get("RICK_BENCH_REF_SIZE", "256"))     # tests shrink the CPU arm; the bench always runs 256
    cfg = AdaptConfig(size=size, batch=2, warmup_iter=0, fisher_freq=50, num_fisher_img=5, fisher_quantile=40.0,
                      prune_quantile=0.1)
    gp, dp = synth.g_state(size, 1), synth.d_state(size, 2)
    A = ao.OracleAdapter(cfg, gp, dp, {k: v.clone() for k, v in gp.items()}, {k: v.clone() for k, v in dp.items()})
    return A, DrawStream(5, "cpu"), synth.shots(10, size, 0)


def cpu_baseline(max_seconds=40.0):
    """Oracle port of ONE plain adaptation iteration (D step + G step, no regulariser, no Fisher round) on the host
    cores of this box -- the reference's CPU path (upfirdn2d_native + native leaky-ReLU)."""
    cores = len(os.sched_getaffinity(0))
    A, draws, shots = _oracle_adapter(cores)
    t = time.perf_counter()
    A.step(1, shots[:2], draws, explicit_layer_noise=False)          # i = 1: no R1, no path-length
    dt = time.perf_counter() - t
    return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 plain adaptation iteration (D step + G step, 256 px, batch 2), {dt:.1f} s, torch CPU "
                      f"{torch.__version__}, {cores} threads"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the reference itself is
    Python + JIT CUDA and cannot travel; its CPU branch is upfirdn2d_native, restated and pinned in oracle/)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    A, draws, shots = _oracle_adapter(cores)
    W, K = args.warmup, args.steps
    budget = 240.0
    t0 = time.perf_counter()
    A.step(1, shots[:2], draws, explicit_layer_noise=False)
    first = time.perf_counter() - t0
    # bounded sample: each step is one plain iteration; keep the whole run within a few minutes
    k_run = max(1, min(K, int((budget - first) / max(first, 1e-3))))
    w_run = 0 if k_run < K else max(0, min(W - 1, int((budget - first * (1 + k_run)) / max(first, 1e-3))))
    for i in range(w_run):
        A.step(1, shots[:2], draws, explicit_layer_noise=False)
    t = time.perf_counter()
    for i in range(k_run):
        j = (2 * i) % 10
        A.step(1 + 4 * i + 1, shots[j:j + 2], draws, explicit_layer_noise=False)   # indices that skip R1 / path-length
    dt = time.perf_counter() - t
    value = k_run / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "steps_run": k_run, "warmup": W, "warmup_run": w_run + 1, "ms_per_step": dt / k_run * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "size": 256, "batch_per_gpu": 2, "global_batch": 2,
                       "parallelism": "cpu", "note": "plain iterations (D step + G step) of the oracle port on host "
                                                     "cores; bounded sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{k_run} plain adaptation iterations, 256 px, batch 2, {cores} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "graphs", "eager"],
                    help="iteration executor: CUDA graphs (single GPU) or eager; auto = graphs when N == 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-profiler", action="store_true", help="bracket the timed steps with cudaProfilerStart/Stop")
    ap.add_argument("--workload", default="rick256", choices=["rick256", "ffhq1024"],
                    help="rick256: BASELINE configs[1] (default, the metric's configuration); ffhq1024: configs[4], the "
                         "1024 px architecture at per-GPU batch 8 under DDP")
    ap.add_argument("--samplegen", action="store_true", help="with --quick: still run the config-3 sample-generation loop")
    ap.add_argument("--quick", action="store_true", help="skip the op sweep / roofline micro-benchmarks (profiler runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    # stdout carries exactly ONE line, the JSON: anything a library writes to file descriptor 1 on the way (NCCL's
    # version banner, cuDNN notes) is sent to stderr instead; the JSON line goes out through the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
