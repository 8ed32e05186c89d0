"""Hottest SASS instructions (stall samples) of a kernel from an ncu source-page CSV:
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv
  python tools/ncu_sass_hot.py src.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iL, iB = hdr.index("stall_long_sb"), hdr.index("stall_barrier")
data = []
tot_s = tot_i = 0
for k, r in enumerate(rows[2:]):
    try:
        n, i = int(r[iN]), int(r[iI])
    except (ValueError, IndexError):
        continue
    tot_s += n
    tot_i += i
    data.append((n, i, k, r[iS].strip(), int(r[iL] or 0), int(r[iB] or 0)))
print("instructions", len(data), "executed warp-instr", tot_i, "samples", tot_s)
for n, i, k, src, lsb, bar in sorted(data, reverse=True)[:top]:
    print(f"#{k:5d} samples {100 * n / tot_s:5.1f}% (long_sb {lsb:6d} barrier {bar:6d}) exec {i:>10d}  {src[:90]}")
