#!/usr/bin/env python
"""Opcode histogram of libb2h.so per kernel (cuobjdump -sass): the mnemonics that prove which hardware path a kernel
uses -- UTCHMMA/UTCQMMA... (tcgen05.mma), UTMALDG (TMA loads), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), LDGMC
(multimem), FFMA (CUDA-core fp32), HMMA (legacy mma.sync) -- written as a markdown table.

    python tools/sass_histogram.py [path/to/libb2h.so] > profiles/sass_rNN.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "multimodal-hand-pose-enhancement-for-sign-language_b200", "libb2h.so")
WATCH = ["UTCHMMA", "UTCMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS",
         "LDGMC", "HMMA", "FFMA", "DFMA", "ATOMG", "RED", "STG", "LDG", "STS", "LDS", "SHFL", "BAR", "ACQBULK", "CCTL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            cur[op] += 1
            cur["_total"] += 1
            if op in ("UTCHMMA", "UTCMMA", "UTCQMMA", "UTMALDG", "LDGMC", "LDTM"):
                cur[op + mods] += 1
    names = demangle(list(per))
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# SASS opcode histogram of `{os.path.relpath(LIB, ROOT)}` ({', '.join(sorted(set(arch)))})\n")
    print("Produced by `tools/sass_histogram.py` (`cuobjdump -sass`); counts are static instruction counts per kernel. "
          "`UTCHMMA` = `tcgen05.mma` (kind::f16 / kind::tf32), `UTMALDG` = TMA tensor load, `LDTM` = `tcgen05.ld`, "
          "`UTCBAR` = `tcgen05.commit`, `LDGMC` = `multimem.ld_reduce`.\n")
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("## Whole library\n")
    print("| opcode | count |\n|---|---:|")
    for k in WATCH:
        if tot[k]:
            print(f"| {k} | {tot[k]} |")
    print(f"| (all instructions) | {tot['_total']} |\n")
    variants = sorted(k for k in tot if "." in k)
    if variants:
        print("Variants: " + ", ".join(f"`{k}` x{tot[k]}" for k in variants) + "\n")
    print("## Per kernel\n")
    cols = [k for k in WATCH if tot[k]]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for fn, c in per.items():
        name = names.get(fn, fn)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*$", "", name)
        print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")


if __name__ == "__main__":
    main()
