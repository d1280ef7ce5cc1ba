#!/usr/bin/env python
"""SASS evidence for profiles/: per hot kernel of libr2f_b200.so, the mnemonic counts that show what the code
is (UTMALDG = TMA bulk tensor load, FFMA2 / FADD2 / FMUL2 = packed float32x2, LDGSTS = cp.async, SYNCS = mbarrier,
LDCU = uniform-datapath constant load (the correlation weights: FFMA2 takes the uniform register as an operand),
MUFU, no UTC*MMA / LDTM: no tensor-core contraction on this path) and a short excerpt of the hottest loop.

    python tools/sass_excerpt.py > profiles/r02_sass.md        (needs cuobjdump; runs on the build machine)
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "raw2film_b200", "libr2f_b200.so")
KERNELS = [  # (substring of the mangled name, title, regex of a mnemonic that marks the hot loop)
    ("k_pointwise_fastILi0ELi1024ELb1E", "k_pointwise_fast<F32x3, 1024 threads, packed pairs>", r"FADD2"),
    ("k_conv2d_symILi17ELi16E", "k_conv2d_sym<17, 16> (MTF)", r"FFMA2"),
    ("k_grain_finish_symILi7ELi8ELb1ELb1E", "k_grain_finish_sym<7, 8, GEN, FASTCURVE>", r"FFMA2"),
    ("k_fft_rows_fwdILi1ELi1ELi1E", "k_fft_rows_fwd<SRC=F32x3, ROWS=1, PLAN=1>", r"LDGSTS"),
    ("k_fft_cols_ipILi1ELi2E", "k_fft_cols_ip<1, 2> (n = 4096, two columns per CTA)", r"FFMA"),
    ("k_fft_rows_invILi0ELi1ELi1ELi1E", "k_fft_rows_inv<0, 1, 1, 1>", r"MUFU"),
    ("k_resize_areaIfE", "k_resize_area<float>", r"FMUL"),
]
WATCH = ["UTMALDG", "UTMAPF", "SYNCS", "LDGSTS", "LDCU", "CCTL", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "DFMA", "MUFU",
         "F2I", "I2FP", "FRND", "LDS", "LDG", "STG", "STS", "BAR", "UTCHMMA", "UTCQMMA", "LDTM", "HMMA"]


def symbols():
    out = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
    return sorted(set(re.findall(r"_ZN3r2f[A-Za-z0-9_]+", out)))


def sass(sym):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    t = text.split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]


def main():
    syms = symbols()
    print("# SASS evidence (round 2)\n")
    print("`python tools/sass_excerpt.py` on `raw2film_b200/libr2f_b200.so` (nvcc 12.9, `-gencode "
          "arch=compute_100a,code=sm_100a`).  Counts are static instruction counts of the whole kernel.\n")
    total = collections.Counter()
    for sub, title, hot in KERNELS:
        cands = [s for s in syms if sub in s and not s.endswith("_param_0")]
        if not cands:
            print(f"## {title}\n\nnot found\n")
            continue
        ins = sass(cands[0])
        cnt = collections.Counter(opcode(t) for _, t in ins)
        total.update(cnt)
        print(f"## {title}\n")
        print(f"`{cands[0]}`: {len(ins)} instructions\n")
        print("| " + " | ".join(k for k in WATCH if cnt.get(k)) + " |")
        print("|" + "---|" * sum(1 for k in WATCH if cnt.get(k)))
        print("| " + " | ".join(str(cnt[k]) for k in WATCH if cnt.get(k)) + " |\n")
        # excerpt: the densest 24-instruction window of the marker mnemonic
        marks = [i for i, (_, t) in enumerate(ins) if re.search(hot, t)]
        if marks:
            best, best_i = -1, 0
            for i in range(0, max(1, len(ins) - 24)):
                c = sum(1 for j in marks if i <= j < i + 24)
                if c > best:
                    best, best_i = c, i
            print("```")
            for a, t in ins[best_i:best_i + 24]:
                print(f"/*{a:05x}*/  {t}")
            print("```\n")
    print("## all listed kernels\n")
    print("tensor-core / TMEM mnemonics (UTCHMMA, UTCQMMA, LDTM, HMMA): "
          + str(sum(total[k] for k in ("UTCHMMA", "UTCQMMA", "LDTM", "HMMA")))
          + " -- correct: nothing on this path is a dense contraction.")


if __name__ == "__main__":
    sys.exit(main())
