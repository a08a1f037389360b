import csv, subprocess, sys, collections
path, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); sm = collections.Counter()
tot_ex = tot_sm = 0
for r in data:
    if len(r) <= ix["# Samples"] or not r[0].startswith("0x"): continue
    ins = r[ix["Source"]].strip()
    if ins.startswith("@"): ins = ins.split(None, 1)[1]
    op = ins.split()[0].split(".")[0] if ins else "?"
    if op in ("LDS", "STS", "LDG", "STG", "FFMA2", "LDC", "LDCU"):
        op = ins.split()[0]
    e = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ex[op] += e; sm[op] += s; tot_ex += e; tot_sm += s
print("total executed", tot_ex, "samples", tot_sm)
for op, e in ex.most_common(28):
    print("%-16s exec %10d %5.1f%%   samples %6d %5.1f%%" % (op, e, 100.0 * e / tot_ex, sm[op], 100.0 * sm[op] / max(tot_sm, 1)))
