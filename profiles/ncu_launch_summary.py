"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/ncu_launch_summary.py launches.csv "<command line>" > summary.txt"""
import csv, re, sys
from collections import OrderedDict
def main(path, cmd=""):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]]); name = re.sub(r"\(.*$", "", name).replace("pfs::", "")
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(cmd); print("(durations are cold-cache and serialised under ncu: compare SHARES)")
    print("%-64s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-64s %8d %12.1f %6.1f%%" % (k[:64], a[0], a[1], 100 * a[1] / tot))
    print("%-64s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
