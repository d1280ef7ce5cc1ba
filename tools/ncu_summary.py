#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports into the markdown tables kept under profiles/ and refresh
profiles/traffic.json (DRAM bytes per launch of each hot kernel, read by bench.py for roofline.traffic).

    python tools/ncu_summary.py --config C2 gpurun_out/r01_f_c2_raw.csv [--config C1 gpurun_out/r01_f_c1_raw.csv]
                                [--launches gpurun_out/r01_f_launches_c2.csv] [--update-traffic]
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import re
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX (smem + L1) busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe busy %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
]

# SASS/demangled kernel name -> bench.py kernel key
KEYS = [("k_fft_rows_fwd", "fft_rows_fwd"), ("k_fft_cols", "fft_cols"), ("k_fft_rows_inv", "fft_rows_inv"),
        ("k_conv2d_sym", "mtf"), ("k_conv2d", "mtf"), ("k_grain_finish", "grain"), ("k_pointwise", "pointwise"),
        ("k_expose", "expose"), ("k_finish", "finish")]


def short(name: str) -> str:
    name = re.sub(r"\(.*\)$", "", name)
    name = name.replace("void ", "").replace("r2f::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()


def key_of(name: str):
    for pat, key in KEYS:
        if pat in name:
            return key
    return None


def to_bytes(val: str, unit: str) -> float:
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(val) * scale


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    out = OrderedDict()
    for r in rows[2:]:
        out.setdefault(short(r[ki]), (hdr, units, r))  # first launch of each kernel
    return out


def table(kernels):
    names = list(kernels)
    lines = ["| metric | " + " | ".join(f"`{n}`" for n in names) + " | unit |", "|---|" + "---|" * (len(names) + 1)]
    for metric, label in METRICS:
        cells, unit = [], ""
        for n in names:
            hdr, units, r = kernels[n]
            if metric in hdr:
                i = hdr.index(metric)
                v = r[i]
                try:
                    f = float(v)
                    v = f"{f:.3f}".rstrip("0").rstrip(".") if abs(f) < 1e6 else f"{f:.4g}"
                except ValueError:
                    pass
                cells.append(v)
                unit = units[i]
            else:
                cells.append("-")
        lines.append(f"| {label} (`{metric}`) | " + " | ".join(cells) + f" | {unit} |")
    return "\n".join(lines)


def launches_table(path):
    rows = list(csv.reader(line for line in open(path) if not line.startswith("==")))
    hdr = rows[0]
    ni, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(short(r[ni]), [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    lines = ["| kernel | launches | avg ms | share |", "|---|---|---|---|"]
    for n, (cnt, ms) in agg.items():
        lines.append(f"| `{n}` | {cnt} | {ms / cnt:.3f} | {100 * ms / total:.1f}% |")
    return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", action="append", nargs=2, metavar=("NAME", "RAW_CSV"), default=[])
    ap.add_argument("--launches", default=None)
    ap.add_argument("--update-traffic", action="store_true")
    args = ap.parse_args()
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for name, path in args.config:
        kernels = read_raw(path)
        print(f"## config {name}\n")
        print(table(kernels))
        print()
        for n, (hdr, units, r) in kernels.items():
            key = key_of(n)
            if key is None:
                continue
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic.setdefault(name, {})[key] = int(to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]))
    if args.launches:
        print("## launch list (cold-cache, serialised under ncu: shares, not absolutes)\n")
        print(launches_table(args.launches))
        print()
    if args.update_traffic:
        with open(traffic_path, "w") as f:
            json.dump(traffic, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
