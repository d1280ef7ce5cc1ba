#!/usr/bin/env python3
"""Per-source-line view of an `ncu --set full --import-source on` capture.

ncu's CSV source page is per SASS instruction and carries no line numbers; nvdisasm -g prints the
`//## File "...", line N` markers of the same instruction stream.  This tool joins the two by instruction
order and sums stall samples, instructions, shared-memory wavefronts and global tag requests per source line.

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX CUBIN MANGLED_SUBSTRING [top]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def sass_lines(cubin, mangled):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    out, on, cur, inl = [], False, ("?", 0), ""
    for ln in txt:
        if ln.startswith(".text."):
            on = mangled in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            inl = m.group(3)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append((cur, ln.split("*/", 1)[1].strip()))
    return out


def main():
    rep, kregex, cubin, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    hdr, data = rows[1], []
    for r in rows[2:]:          # ncu prints the table once per view: keep the first copy
        if r and r[0] == "Kernel Name":
            break
        if len(r) >= len(hdr):
            data.append(r)
    ix = {n: i for i, n in enumerate(hdr)}
    sl = sass_lines(cubin, mangled)
    if len(sl) != len(data):
        print(f"warning: {len(sl)} disassembled instructions vs {len(data)} profiled", file=sys.stderr)
    cols = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Tag Requests Global", "stall_long_sb",
            "stall_short_sb", "stall_mio", "stall_math", "stall_barrier", "stall_wait", "stall_lg"]
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for (line, _), r in zip(sl, data):
        for c in cols:
            try:
                agg[line][c] += float(r[ix[c]])
            except (ValueError, KeyError):
                pass
    tot = {c: sum(a[c] for a in agg.values()) or 1.0 for c in cols}
    print("line                         samp%  inst%  smemWF%  gtag%   long  short  mio  math  bar  wait")
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        s = a["# Samples"] or 1.0
        print(f"{line[0][:20]:20s}:{line[1]:<5d} {100 * a['# Samples'] / tot['# Samples']:6.1f} "
              f"{100 * a['Instructions Executed'] / tot['Instructions Executed']:6.1f} "
              f"{100 * a['L1 Wavefronts Shared'] / tot['L1 Wavefronts Shared']:7.1f} "
              f"{100 * a['L1 Tag Requests Global'] / tot['L1 Tag Requests Global']:6.1f}  "
              + " ".join(f"{100 * a[c] / s:5.0f}" for c in cols[4:10]))


if __name__ == "__main__":
    main()
