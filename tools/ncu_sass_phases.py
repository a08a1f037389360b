import csv, subprocess, sys, collections
path, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
seg_ex = seg_sm = 0; seg_start = 0; n = 0
segs = []
marks = ("BAR.SYNC", "UTCHMMA", "UTCBAR", "SYNCS", "TCGEN", "LDTM", "UTCQMMA", "UTCMMA")
ops = collections.Counter()
for r in data:
    if len(r) <= ix["# Samples"] or not r[0].startswith("0x"): continue
    ins = r[ix["Source"]].strip()
    e = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    seg_ex += e; seg_sm += s; n += 1
    op = ins.split(None, 1)[1] if ins.startswith("@") else ins
    ops[op.split()[0].split(".")[0]] += e
    if any(m in ins for m in marks):
        segs.append((n, seg_ex, seg_sm, ins[:60], ops.most_common(4)))
        seg_ex = seg_sm = 0; ops = collections.Counter()
segs.append((n, seg_ex, seg_sm, "END", ops.most_common(4)))
tot = sum(s[1] for s in segs); tots = sum(s[2] for s in segs)
for s in segs:
    if s[1] > tot * 0.004 or s[2] > tots * 0.004:
        print("upto#%5d exec %10d %5.1f%% samples %6d %5.1f%% | %-50s | %s" % (s[0], s[1], 100.0 * s[1] / tot, s[2], 100.0 * s[2] / tots, s[3], s[4]))
