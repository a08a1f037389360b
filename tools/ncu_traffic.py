#!/usr/bin/env python
"""DRAM bytes per launch of every kernel in an `ncu --set full` report -> profiles/ncu_traffic.json, the file
bench.py reads `roofline.traffic` from (stamped with the commit the capture was taken at).

    python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep [commit]
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep = sys.argv[1]
    commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(
        ["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"<.*", "", re.sub(r"^(void )?(pfs::)?", "", r[ix["Kernel Name"]]).split("(")[0])
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[m]].replace(",", "")) * scale.get(units[ix[m]], 1.0)
        acc.setdefault(name, []).append(tot)
    out = {"file": os.path.relpath(rep, ROOT), "commit": commit,
           "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches), C3 default workload",
           "kernels": {k: sum(v) / len(v) for k, v in sorted(acc.items())}}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
