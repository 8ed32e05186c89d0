"""Hot SASS regions of an ncu source-page CSV (SASS view).
  ncu -i rep --page source --csv --kernel-name regex:X > src.csv ; python tools/ncu_sass.py src.csv [min_frac]
Prints, for the FIRST launch in the file, every instruction whose executed count or stall
samples exceed min_frac of the total, with running region totals."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]
iS, iI, iW, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
data = []
first_addr = None
for r in rows[2:]:
    try:
        n = int(r[iI])
    except (ValueError, IndexError):
        continue
    if first_addr is None:
        first_addr = r[0]
    elif r[0] == first_addr:
        break  # second launch starts
    data.append((n, int(r[iW] or 0), r[iS], r[iT]))
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print(len(data), "SASS instructions; executed", tot, "stall samples", tots)
for k, (n, w, s, t) in enumerate(data):
    if n > tot * thr or w > tots * thr * 2:
        print(f"{k:4d} {n / 1e6:8.2f}M {100 * n / tot:5.2f}%  smp {100 * w / max(tots, 1):5.1f}%  thr={t:>5s}  {s[:80]}")
