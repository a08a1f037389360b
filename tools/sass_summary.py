#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA2 (packed fp32 FMA), LDCU (constant bank).

    python tools/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pfs-neural-net_b200", "csrc", "libpfs_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "FFMA", "LDCU", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, order = {}, None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for o in OPS:
                if op == o:
                    counts[cur][o] += 1
    names = demangle(order)
    short = lambda s: re.sub(r"\(.*", "", re.sub(r"^void ", "", s)).replace("pfs::", "")
    print("SASS mnemonic counts per kernel of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-52s %7s " % ("kernel", "instr") + " ".join("%7s" % o for o in OPS))
    tot = collections.Counter()
    for k in sorted(order, key=lambda k: short(names[k])):
        c = counts[k]
        tot.update(c)
        print("%-52s %7d " % (short(names[k])[:52], c["total"]) + " ".join("%7d" % c[o] for o in OPS))
    print("%-52s %7d " % ("TOTAL (%d kernels)" % len(order), tot["total"]) + " ".join("%7d" % tot[o] for o in OPS))
    tc = sorted({short(names[k]) for k in order if counts[k]["UTCHMMA"] or counts[k]["UTCQMMA"]})
    print("\nkernels issuing tcgen05.mma (UTCHMMA): " + ", ".join(tc))
    tma = sorted({short(names[k]) for k in order if counts[k]["UTMALDG"] or counts[k]["UBLKCP"]})
    print("kernels using TMA (UTMALDG / UBLKCP): " + ", ".join(tma))


if __name__ == "__main__":
    main()
