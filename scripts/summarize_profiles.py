"""Turn the raw ncu artefacts a gpurun call brought back (gpurun_out/) into the small tracked summaries under profiles/.

  python scripts/summarize_profiles.py r01            # writes profiles/r01_launches.md, profiles/r01_<report>.md
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

OURS = ("rick::", "upfirdn2d", "bias_act", "bias_grad", "fisher_multi", "filter_fim", "percentile_kernel", "decide_kernel",
        "mask_apply", "conv_tc", "blur_nhwc", "to_rgb_nhwc", "adam_mask_ema", "colsum_bwd", "modulate_kernel",
        "styled_epilogue")


def short(name):
    name = name.replace("void ", "")
    cut = name.find("(")
    return name[:cut] if cut > 0 else name


def launches(tag, fname="launches.csv"):
    path = os.path.join(OUT, fname)
    if not os.path.isfile(path):
        return
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        val_us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        rows.append((short(r["Kernel Name"]), val_us))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    ours = sum(v for k, v in rows if any(o in k for o in OURS))
    with open(os.path.join(PROF, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write(f"command: `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none "
                f"python bench.py --steps 1 --warmup 3 --quick --cuda-profiler --mode eager` -- the timed iteration of the "
                f"bench workload, launched eagerly (the default graphs mode replays the same kernels); {len(rows)} launches; "
                f"cold-cache, serialised -- compare SHARES, not absolutes\n\n")
        f.write(f"total {tot / 1e3:.2f} ms over {len(rows)} launches; rick_b200 kernels {ours / 1e3:.2f} ms "
                f"({100 * ours / max(tot, 1e-9):.1f} %)\n\n| kernel | launches | total us | share % | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            mark = " **(ours)**" if any(o in k for o in OURS) else ""
            f.write(f"| `{k[:110]}`{mark} | {n} | {v:.1f} | {100 * v / tot:.2f} | {v / n:.2f} |\n")
    print("wrote", f"profiles/{tag}_launches.md", f"({len(rows)} launches, ours {100 * ours / max(tot, 1e-9):.1f} %)")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]


def report(tag, rep):
    path = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.isfile(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    if len(rd) < 3:
        print("empty report", rep)
        return
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(PROF, f"{tag}_{rep}.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none ({rep}.ncu-rep), per launch\n\n")
        for row in rd[2:]:
            name = short(row[col["Kernel Name"]])
            f.write(f"## `{name[:140]}`  grid {row[col.get('Grid Size', 0)]} block {row[col.get('Block Size', 0)]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in col:
                    f.write(f"| {k} | {row[col[k]]} | {units[col[k]]} |\n")
            try:
                rdb = float(row[col["dram__bytes_read.sum"]].replace(",", ""))
                wrb = float(row[col["dram__bytes_write.sum"]].replace(",", ""))
                dur = float(row[col["gpu__time_duration.sum"]].replace(",", ""))
                f.write(f"| **traffic (read+write)** | {rdb + wrb:.4g} | {units[col['dram__bytes_read.sum']]} |\n")
                f.write(f"| **duration** | {dur:.4g} | {units[col['gpu__time_duration.sum']]} |\n")
            except Exception:
                pass
            f.write("\n")
    print("wrote", f"profiles/{tag}_{rep}.md")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    for rep in sorted(f[:-8] for f in os.listdir(OUT) if f.endswith(".ncu-rep")):
        report(tag, rep)
