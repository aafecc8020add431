#!/usr/bin/env python3
"""Instruction-class summary per kernel of libkoreb200.so (cuobjdump -sass), for profiles/:
which Blackwell paths each kernel uses (bulk/TMA copies UBLKCP, mbarrier SYNCS, cp.async LDGSTS,
256-bit LDG/STG, REDUX, DFMA vs DMMA, tcgen05 UTC*MMA / TMEM LDTM -- none for complex128: tcgen05 has
no FP64 kind).   python tools/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "kore_b200", "libkoreb200.so")
CLASSES = [
    ("DFMA", r"^DFMA"), ("DMUL/DADD", r"^(DMUL|DADD)"), ("DMMA", r"^DMMA"), ("UTC*MMA (tcgen05)", r"^UTC.*MMA"),
    ("LDTM/STTM (TMEM)", r"^(LDTM|STTM)"), ("UTMALDG/UTMASTG (TMA tensor)", r"^UTMA"),
    ("UBLKCP (bulk copy)", r"^UBLKCP"), ("UBLKPF (bulk prefetch)", r"^UBLKPF"), ("SYNCS (mbarrier)", r"^SYNCS"),
    ("LDGSTS (cp.async)", r"^LDGSTS"), ("LDG/STG .256", r"^(LDG|STG).*\.256"), ("LDG/STG .128", r"^(LDG|STG).*\.128"),
    ("LDS/STS", r"^(LDS|STS)"), ("SHFL", r"^SHFL"), ("CREDUX (warp reduce)", r"^C?REDUX"), ("BAR", r"^BAR"), ("ATOMG/REDG/ATOMS", r"^(ATOMG|REDG|ATOMS|ATOM|RED)\b"),
    ("spill LDL/STL", r"^(LDL|STL)"),
]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(name.split("(")[0], collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(2)
            cur["total"] += 1
            for cname, pat in CLASSES:
                if re.search(pat, op):
                    cur[cname] += 1
    names = [c for c, _ in CLASSES]
    print("SASS instruction classes per kernel, kore_b200/libkoreb200.so (sm_100a), static counts")
    print("%-46s %7s  %s" % ("kernel", "total", "non-zero classes"))
    tot = collections.Counter()
    for k, c in kernels.items():
        if c["total"] == 0:
            continue
        tot.update(c)
        print("%-46s %7d  %s" % (k[:46], c["total"], ", ".join("%s %d" % (n, c[n]) for n in names if c[n])))
    print("%-46s %7d  %s" % ("ALL", tot["total"], ", ".join("%s %d" % (n, tot[n]) for n in names)))


if __name__ == "__main__":
    main()
