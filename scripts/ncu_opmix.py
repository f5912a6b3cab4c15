"""Dynamic instruction mix of one kernel from an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_opmix.py <rep> <kernel-regex> <units> [unit-name]   (units = divisor, e.g. frames or tokens)"""
import collections, csv, io, subprocess, sys

rep, kre, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
uname = sys.argv[4] if len(sys.argv) > 4 else "unit"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
iS, iE, iSamp, iW, iWx = (hdr.index(k) for k in ("Source", "Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive"))
byop, samp, wav, wavx = (collections.Counter() for _ in range(4))
tot = 0
first = True
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[iE].strip().isdigit():
        if "Instructions Executed" in r:   # a second kernel instance follows: stop at the first
            break
        continue
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    n = int(r[iE]); byop[op] += n; tot += n
    samp[op] += int(r[iSamp] or 0); wav[op] += int(r[iW] or 0); wavx[op] += int(r[iWx] or 0)
print("total warp-instructions %d = %.1f per %s; stall samples %d" % (tot, tot / units, uname, sum(samp.values())))
for op, n in byop.most_common(30):
    print("%-12s %8.2f /%s  samples %6d  smem wavefronts %.2f (excess %.2f)" % (op, n / units, uname, samp[op], wav[op] / units, wavx[op] / units))
