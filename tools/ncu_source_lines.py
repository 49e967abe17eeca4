"""Per-source-line instruction / stall-sample totals from an ncu report (needs -lineinfo and --import-source on).
    python tools/ncu_source_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    h = rows[hdr]
    ie, src, smp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    out, tot, tots = [], 0, 0
    for r in rows[hdr + 1:]:
        if len(r) <= ie or not r[0].strip().isdigit():   # keep the per-line totals, skip the SASS rows under them
            continue
        try:
            v, s = int(r[ie]), int(r[smp] or 0)
        except ValueError:
            continue
        tot += v
        tots += s
        out.append((v, s, r[0], r[src].strip()[:120]))
    out.sort(reverse=True)
    print(f"{kern}: {tot} warp instructions, {tots} stall samples")
    for v, s, line, text in out[:top]:
        print(f"{v:10d} {100 * v / max(tot, 1):5.1f}%  samples {100 * s / max(tots, 1):5.1f}%  L{line}: {text}")


if __name__ == "__main__":
    main()
