#!/usr/bin/env python
"""Opcode histogram of every kernel in libb200grbm.so from `cuobjdump -sass` (no GPU needed).

    python tools/sass_histogram.py > profiles/r2_sass_opcode_histogram.txt

Lists, per kernel, the instruction count and the mnemonics that prove the claimed hardware paths: tcgen05 MMA
(UTCIMMA / UTCHMMA / UTCOMMA = int8 / bf16 / block-scaled e2m1, .2CTA for cta_group::2), TMEM loads (LDTM), TMA tensor loads (UTMALDG), bulk copies (UBLKCP),
R2P predicate moves, MUFU.EX2, shared-memory atomics (ATOMS), and the top opcodes by count."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image-generation_b200", "csrc", "libb200grbm.so")
MARK = ("UTCIMMA", "UTCHMMA", "UTCOMMA", "STTM", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "R2P", "MUFU.EX2", "ATOMS", "ATOMG", "RED",
        "IMAD.WIDE", "FFMA", "FADD", "LDS", "STS", "SHFL", "BAR.SYNC", "REDUX", "VOTE", "POPC", "DADD", "DFMA")


def main():
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*", "", name)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    print("# cuobjdump -sass", os.path.relpath(LIB, ROOT), "-- opcode histogram per kernel (static instruction counts)")
    for k, c in kernels.items():
        total = sum(c.values())
        marks = {}
        for op, n in c.items():
            for mk in MARK:
                if op == mk or op.startswith(mk + ".") or (mk in ("MUFU.EX2", "IMAD.WIDE", "BAR.SYNC") and op.startswith(mk)):
                    marks[mk] = marks.get(mk, 0) + n
        two_cta = sum(n for op, n in c.items() if "2CTA" in op)
        top = ", ".join(f"{op} {n}" for op, n in c.most_common(8))
        print(f"\n## {k}\n   instructions {total}" + (f"   (.2CTA forms: {two_cta})" if two_cta else ""))
        print("   marks: " + ", ".join(f"{m} {n}" for m, n in sorted(marks.items(), key=lambda x: -x[1])))
        print("   top:   " + top)


if __name__ == "__main__":
    sys.exit(main())
