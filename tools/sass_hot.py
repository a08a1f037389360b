#!/usr/bin/env python
"""Static instruction mix of one kernel's hot region (first to last FFMA2 / UTCHMMA) from cuobjdump -sass.
The per-edge kernels are fully unrolled, so the static mix of that region is the dynamic mix per tile.

    python tools/sass_hot.py <kernel substring> [marker mnemonic] [--dump N]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pfs-neural-net_b200", "csrc", "libpfs_b200.so")


def functions():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)(\S*)\s*(.*?);", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2), m.group(3), m.group(4)))
    return out


def main():
    want = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "FFMA2"
    dump = int(sys.argv[sys.argv.index("--dump") + 1]) if "--dump" in sys.argv else 0
    fns = functions()
    names = subprocess.run(["c++filt"], input="\n".join(fns), capture_output=True, text=True).stdout.splitlines()
    for mangled, name in zip(fns, names):
        if want not in name:
            continue
        ops = fns[mangled]
        idx = [i for i, o in enumerate(ops) if o[1] == marker]
        if not idx:
            print(name, ": no", marker)
            continue
        a, b = idx[0], idx[-1]
        c = collections.Counter(o[1] for o in ops[a:b + 1])
        n = b - a + 1
        print("%s\n  %d instructions total, hot region %d (first..last %s)" % (name[:110], len(ops), n, marker))
        print("  " + "  ".join("%s %d (%.0f%%)" % (k, v, 100.0 * v / n) for k, v in c.most_common(16)))
        for o in ops[a:a + dump]:
            print("    %05x %s%s %s" % (o[0], o[1], o[2], o[3]))


if __name__ == "__main__":
    main()
