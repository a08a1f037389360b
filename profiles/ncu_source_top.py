"""Top source lines / SASS instructions by warp-stall samples from an ncu report (--import-source on).
usage: python profiles/ncu_source_top.py report.ncu-rep [N]"""
import csv, subprocess, sys
def main(path, n=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hi], rows[hi + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    s_i, smp, ex = ix["Source"], ix["# Samples"], ix["Instructions Executed"]
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[smp] or 0) for r in data if len(r) > smp)
    print("total samples", tot)
    data = [r for r in data if len(r) > smp]
    for r in sorted(data, key=lambda r: -int(r[smp] or 0))[:n]:
        top = sorted(((int(r[ix[k]] or 0), k) for k in stall), reverse=True)[:3]
        print("%6d %5.1f%% exec %9s | %-70s | %s" % (int(r[smp] or 0), 100.0 * int(r[smp] or 0) / max(tot, 1), r[ex], r[s_i][:70],
              ", ".join("%s %d" % (k[6:], v) for v, k in top if v)))
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
