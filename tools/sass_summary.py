#!/usr/bin/env python
"""Instruction-class summary of every kernel in libuwu_b200.so from `cuobjdump -sass` (runs in the build container, no GPU):
counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel — UTCHMMA / UTCQMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA), UTCBAR (tcgen05.commit), SYNCS (mbarrier) — against the
legacy tensor path HMMA (mma.sync) and LDGSTS (cp.async).  See /opt/skills/guides/B200_PROFILING.md for the mnemonics.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uwudiff_b200", "libuwu_b200.so")
CLASSES = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "MUFU",
           "RED", "ATOM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for c in CLASSES:
                if op.startswith(c):
                    kernels[cur][c] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({len(kernels)} kernels)")
    print("# " + " ".join(f"{c:>8s}" for c in ["instrs"] + CLASSES) + "  kernel")
    for (name, cnt), dn in zip(kernels.items(), demangle):
        depth, cut = 0, len(dn)
        for i, ch in enumerate(dn):  # cut the parameter list: first '(' outside template angle brackets
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0 and not dn[i:].startswith("(anonymous"):
                cut = i
                break
        short = dn[:cut]
        if len(short) > 110:
            short = short[:107] + "..."
        native = any(cnt[c] for c in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG"))
        legacy = cnt["HMMA"] > 0
        tag = "tcgen05/TMA" if native else ("mma.sync" if legacy else "simt")
        print("  " + " ".join(f"{cnt[c]:8d}" for c in ["_total"] + CLASSES) + f"  [{tag}] {short}")


if __name__ == "__main__":
    sys.exit(main())
