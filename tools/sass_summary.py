#!/usr/bin/env python
"""tools/sass_summary.py -- opcode census + hot-loop excerpts of the built libvrt_cuda.so (cuobjdump -sass), for profiles/.

    python tools/sass_summary.py > profiles/r02_sass_summary.md

The .so is git-ignored, so this file is the committed evidence of what the compiler emitted: packed FFMA2/FMUL2 in the
inner term, MUFU.RCP/EX2, the TMA bulk copy (UBLKCP) + mbarrier (SYNCS) staging, 128-bit framebuffer stores, warp
votes / reductions, and the absence of tensor-core instructions (this path is not a contraction)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "simd-gaussian-ray-tracing_b200", "csrc", "libvrt_cuda.so")
KERNELS = ["k2_bandILi0ELi4", "k2_band_longILi0ELi4", "k2_band_longILi0ELi3", "k2_renderILi0ELi8ELb1ELi3ELb0ELb0", "k2_renderILi0ELi8ELb1ELi1ELb1ELb0", "k1_leafILi0", "k1_binILb0", "k1_binILb1", "k0_prepare", "k3_combine"]
WATCH = ["FFMA2", "FMUL2", "FFMA", "FMUL", "FADD", "MUFU.RCP", "MUFU.EX2", "LOP3", "LDS.128", "LDS", "STS", "LDG", "STG.E.128", "STG", "UBLKCP", "SYNCS", "VOTE", "REDUX", "SHFL",
         "ATOM", "RED", "BRA", "UTCMMA", "HMMA", "LDTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(1).strip())
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    print(f"# SASS census of `libvrt_cuda.so` (sm_100a, nvcc 12.9; built from the tree at / after commit {head})\n")
    print("`cuobjdump -sass` of the in-tree library; counts are static instructions per kernel (not executed counts).\n")
    print("| kernel | instructions | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    picked = {}
    for key in KERNELS:
        for name, ins in funcs.items():
            if key in name:
                picked[key] = ins
                ops = [i.split()[1] if i.startswith("@") else i.split()[0] for i in ins]
                row = []
                for w in WATCH:
                    if w in ("FFMA", "FMUL", "LDS", "STG"):
                        row.append(sum(1 for o in ops if o == w or (o.startswith(w + ".") and not o.startswith(w + "2") and not (w == "LDS" and o.startswith("LDS.128")) and not (w == "STG" and "128" in o))))
                    else:
                        row.append(sum(1 for o in ops if o.startswith(w)))
                print(f"| `{key}` | {len(ins)} | " + " | ".join(str(r) for r in row) + " |")
                break
    total_tc = sum(1 for ins in funcs.values() for i in ins if re.match(r"(@\S+\s+)?(UTC|HMMA|IMMA|LDTM|STTM)", i))
    print(f"\nTensor-core / TMEM instructions in the whole library: **{total_tc}** (none expected: every (pixel, sample, occluder) triple needs its own erf).\n")
    # hot-loop excerpt: the longest run of packed math in the banded kernel and in the strict kernel
    for key, title in (("k2_bandILi0ELi4", "k2_band<A&S>: sign-uniform body of one pair group (10 terms: 5 x [FFMA2 t, 4 FFMA2 Horner, 2 FMUL2, 2 MUFU.RCP, FFMA2 accumulate])"),
                       ("k2_renderILi0ELi8ELb1ELi3ELb0ELb0", "k2_render<A&S, Q=8, packed>: the strict kernel's inner term")):
        ins = picked.get(key, [])
        best, start, run, s0 = 0, 0, 0, 0
        for k, i in enumerate(ins):
            if re.match(r"(FFMA2|FMUL2|MUFU\.RCP)", i):
                if run == 0:
                    s0 = k
                run += 1
                if run > best:
                    best, start = run, s0
            elif not re.match(r"(MOV|IMAD\.MOV|HFMA2|FSEL)", i):
                run = 0
        print(f"### {title}\n\n```")
        for i in ins[start : start + min(best, 64)]:
            print("    " + i)
        print("```\n")
    # the TMA staging of the contiguous-list variant
    ins = picked.get("k2_renderILi0ELi8ELb1ELi1ELb1ELb0", [])
    tma = [i for i in ins if re.match(r"(@\S+\s+)?(UBLKCP|SYNCS|FENCE)", i)]
    print("### TMA bulk copy + mbarrier staging (k2_render, contiguous lists)\n\n```")
    for i in tma[:12]:
        print("    " + i)
    print("```")


if __name__ == "__main__":
    main()
