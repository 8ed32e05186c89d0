"""Per-source-line instruction / stall-sample counts from an ncu source-page CSV.

  ncu -i rep.ncu-rep --page source --print-source cuda --csv --kernel-name regex:nfst_fwd > src.csv
  python tools/ncu_lines.py src.csv [top_n]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for i, r in enumerate(rows[:10]):
    if "Source" in r:
        hdr, start = r, i + 1
        break
iS, iI, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data, tot, tots = [], 0, 0
for r in rows[start:]:
    try:
        n = int(r[iI])
    except ValueError:
        continue
    w = int(r[iW] or 0)
    tot += n
    tots += w
    data.append((n, w, r[0], r[iS]))
print("total warp-instructions", tot, "total stall samples", tots)
for n, w, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{n:>12d} {100 * n / max(tot, 1):5.1f}%  samples {100 * w / max(tots, 1):5.1f}%  L{ln}: {src.strip()[:105]}")
