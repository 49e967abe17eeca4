"""Per-source-line warp instructions (per unit of work) and stall-sample shares of one kernel in an ncu report.
    python tools/ncu_lines.py report.ncu-rep kernel_regex units [min_pct]"""
import csv, io, subprocess, sys
rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.6
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(txt)))
hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hdr]
ie, src, smp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
out = []
for r in rows[hdr + 1:]:
    if len(r) > ie and r[0].strip().isdigit():
        try:
            out.append((int(r[0]), int(r[ie]), int(r[smp] or 0), r[src].strip()[:110]))
        except ValueError:
            pass
tot, ts = sum(o[1] for o in out), sum(o[2] for o in out)
for l, v, s, t in out:
    if v > minp / 100 * tot or s > 2 * minp / 100 * ts:
        print(f"L{l:4d} {v / units:7.1f}/unit {100 * v / tot:5.1f}% smp {100 * s / ts:5.1f}%  {t}")
print(f"total {tot / units:.1f} warp instructions per unit, {ts} samples")
