"""Turn the raw captures in gpurun_out/ into the committed summaries under profiles/.

    python tools/make_profiles.py TAG [--rep gpurun_out/prof_bench.ncu-rep] [--launches gpurun_out/launches.csv]
                                      [--bench gpurun_out/bench_line.json] [--ref gpurun_out/bench_ref_line.json]

Writes profiles/TAG_{ncu_full_summary.txt, traffic.json, launch_list_summary.txt, launches.csv, bench_line.json,
bench_reference_line.json}.  bench.py reads the newest *_traffic.json whose workload matches for roofline.traffic.
"""
import argparse
import collections
import csv
import io
import json
import re
import shutil
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--rep", default="gpurun_out/prof_bench.ncu-rep")
ap.add_argument("--launches", default="gpurun_out/launches.csv")
ap.add_argument("--bench", default="gpurun_out/bench_line.json")
ap.add_argument("--ref", default="gpurun_out/bench_ref_line.json")
a = ap.parse_args()
tag = a.tag


def short(name: str) -> str:
    """'void <unnamed>::sell_pull_kernel<0, 1, ...>(...)' -> 'sell_pull_kernel'"""
    m = re.search(r"(\w+)\s*(<|\()", name.replace("void ", "").replace("<unnamed>::", ""))
    return m.group(1) if m else name[:40]


with open(f"profiles/{tag}_ncu_full_summary.txt", "w") as f:
    f.write(subprocess.run([sys.executable, "tools/ncu_summary.py", a.rep], capture_output=True, text=True).stdout)

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def val(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].split("/")[0]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)


bench = json.loads(open(a.bench).read().strip().splitlines()[-1])
out = {"source": f"ncu --set full --clock-control none --import-source on -k regex:tile_pull|tile_flow -s 6 -c 2 python bench.py "
                 f"--steps 3 --warmup 3 --no-e2e --no-cpu --no-sweep ({tag})",
       "workload_arcs_per_gpu": bench["config"]["arcs_per_gpu"], "kernels": {}}
for r in rows[2:]:
    key = short(r[hdr.index("Kernel Name")])
    out["kernels"][key] = {
        "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
        "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
        "gpu_time_us": val(r, "gpu__time_duration.sum"), "grid": int(val(r, "launch__grid_size")),
        "block": int(val(r, "launch__block_size")), "registers": int(val(r, "launch__registers_per_thread")),
        "smem_dynamic_bytes": round(val(r, "launch__shared_mem_per_block_dynamic")),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
        "warp_instructions_per_arc": val(r, "smsp__inst_executed.sum") / bench["config"]["arcs_per_gpu"],
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    }
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)

lrows = list(csv.reader(open(a.launches)))
for i, r in enumerate(lrows):
    if "Kernel Name" in r:
        h, st = r, i + 1
        break
agg = collections.OrderedDict()
for r in lrows[st:]:
    if len(r) < len(h):
        continue
    n = short(r[h.index("Kernel Name")])
    v = float(r[h.index("Metric Value")].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[h.index("Metric Unit")], 1)
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
L = ["ncu launch list: ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tile_|sell_|nfst_ -c 12 python bench.py "
     "--steps 3 --warmup 3 --no-e2e --no-cpu --no-sweep",
     "(a bench step launches exactly these kernels; per-launch times under ncu are cold-cache and serialised: compare SHARES)", ""]
for k, v in agg.items():
    L.append(f"{k:24s} launches {len(v):3d}  mean {sum(v) / len(v):8.1f} us  share of step {100 * sum(v) / tot:5.1f}%")
r = bench["roofline"]
o = r["other_kernel"]
L += ["", f"bench.py (CUDA events, same build): {r['kernel'].split(' ')[0]} {r['kernel_ms'] * 1e3:.1f} us, "
          f"{o['kernel'].split(' ')[0]} {o['kernel_ms'] * 1e3:.1f} us -> dominant kernel's share "
          f"{100 * r['kernel_ms'] / (r['kernel_ms'] + o['kernel_ms']):.1f}%"]
open(f"profiles/{tag}_launch_list_summary.txt", "w").write("\n".join(L) + "\n")
shutil.copy(a.launches, f"profiles/{tag}_launches.csv")
json.dump(bench, open(f"profiles/{tag}_bench_line.json", "w"))
try:
    json.dump(json.loads(open(a.ref).read().strip().splitlines()[-1]), open(f"profiles/{tag}_bench_reference_line.json", "w"))
except Exception as e:  # noqa: BLE001
    print("no reference line:", e)
print(open(f"profiles/{tag}_launch_list_summary.txt").read())
print(json.dumps(out["kernels"], indent=1))
