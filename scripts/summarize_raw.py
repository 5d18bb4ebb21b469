"""Compact per-launch table from an exported ncu raw page (``ncu -i X.ncu-rep --page raw --csv``, written by
scripts/gpu_round.sh as gpurun_out/prof_ops<TAG>_raw.csv).   python scripts/summarize_raw.py <csv> [out.md] [title]"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active", "umma %"),
        ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("lts__t_bytes.sum", "L2 MB"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "xbar->SM MB"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"), ("launch__grid_size", "grid")]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    have = [(k, t) for k, t in COLS if k in col]
    lines = ["| # | kernel | " + " | ".join(t for _, t in have) + " |", "|---|---|" + "---:|" * len(have)]
    for i, r in enumerate(body):
        name = r[col["Kernel Name"]].replace("void ", "").split("(")[0].replace("rick::<unnamed>::", "")
        vals = []
        for k, t in have:
            v, u = num(r[col[k]]), units[col[k]]
            if v is None:
                vals.append(r[col[k]])
                continue
            if "MB" in t:
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            if t == "us":
                v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
            vals.append(f"{v:.1f}" if t != "grid" else f"{int(v)}")
        lines.append(f"| {i} | `{name[:48]}` | " + " | ".join(vals) + " |")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        title = sys.argv[3] if len(sys.argv) > 3 else path
        open(sys.argv[2], "w").write(f"# {title}\n\n" + text)
    print(text)


if __name__ == "__main__":
    main()
