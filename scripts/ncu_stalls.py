"""Stall-reason totals and the hottest instructions of one kernel from an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_stalls.py <rep> <kernel-regex> [top-n]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # which matching launch (0 = first)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][inst]
hdr = rows[hi]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = []
for r in rows[hi + 1:]:
    if "Instructions Executed" in r:
        break
    if len(r) > iE and r[iE].strip().isdigit():
        body.append(r)
tot = sum(int(r[iSamp] or 0) for r in body)
agg = collections.Counter()
for r in body:
    for i, h in stall_cols:
        agg[h] += int(r[i] or 0)
print("kernel %s: %d instructions, %d samples" % (kre, len(body), tot))
print("stall reasons: " + ", ".join("%s %.1f%%" % (h.replace("stall_", ""), 100.0 * n / max(1, sum(agg.values()))) for h, n in agg.most_common(10)))
for idx, r in sorted(sorted(enumerate(body), key=lambda x: -int(x[1][iSamp] or 0))[:topn]):
    st = sorted(((int(r[i] or 0), h.replace("stall_", "")) for i, h in stall_cols), reverse=True)[:2]
    print("%5d %7d %5.1f%%  exec %10s  %-64s %s" % (idx, int(r[iSamp] or 0), 100.0 * int(r[iSamp] or 0) / max(1, tot), r[iE], r[iS][:64], st))
