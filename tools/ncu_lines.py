"""Per-source-line executed warp-instructions and stall samples from an .ncu-rep (needs -lineinfo + --import-source on):
python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
i_line, i_src, i_exec, i_smp = 0, 1, hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
per = defaultdict(lambda: [0, 0, "", defaultdict(int)])
tot = smp = 0
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[i_line] or not r[i_exec].isdigit():
        continue  # SASS rows under a source line, elisions
    n, s = int(r[i_exec]), int(r[i_smp] or 0)
    e = per[r[i_line]]
    e[0] += n
    e[1] += s
    e[2] = r[i_src]
    for c in stall_cols:
        if r[c] and r[c] != "0":
            e[3][hdr[c]] += int(r[c])
    tot += n
    smp += s
print(f"total warp-instructions {tot}, stall samples {smp}")
for line, (n, s, src, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top_n]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * n / tot:5.1f}% instr {100 * s / max(smp, 1):5.1f}% samples  L{line:>4s}: {src.strip()[:100]:100s} | {top}")
