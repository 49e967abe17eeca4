"""Summarise an `ncu --page source --csv` export: opcode mix by executed warp instructions + samples."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot, samp, data = collections.Counter(), collections.Counter(), []
for r in rows[2:]:
    if len(r) <= iE:
        continue
    op = r[iS].strip().split()
    o = op[1] if op[0].startswith('@') else op[0]
    n = int(r[iE])
    tot[o] += n
    samp[o] += int(r[iSamp])
    data.append((r[0], r[iS].strip(), n, int(r[iSamp])))
T = sum(tot.values())
S = sum(samp.values())
print("total warp instructions", T, "samples", S)
for k, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{k:28s} {v:12d} {100*v/T:5.1f}%  samples {samp[k]:6d} {100*samp[k]/max(S,1):5.1f}%")
cnt = collections.Counter(d[2] for d in data)
print()
for c, n in cnt.most_common(8):
    print("exec count", c, "x", n, "instrs")
