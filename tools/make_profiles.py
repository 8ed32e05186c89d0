"""Turn the raw captures in gpurun_out/ into the committed summaries under profiles/.

    python tools/make_profiles.py r1      # expects gpurun_out/{prof_r1.ncu-rep, launches_r1.csv, bench_r1*.json}
"""
import collections
import csv
import io
import json
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rep = f"gpurun_out/prof_{tag}.ncu-rep"

with open(f"profiles/{tag}_ncu_full_summary.txt", "w") as f:
    f.write(subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def val(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].split("/")[0]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)


bench = json.load(open(f"gpurun_out/bench_{tag}.json"))
out = {"source": f"ncu --set full --clock-control none -k regex:nfst_ -s 6 -c 2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu ({tag})",
       "workload_arcs_per_gpu": bench["config"]["arcs_per_gpu"], "kernels": {}}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    key = "nfst_fwd_kernel" if "nfst_fwd" in name else "nfst_bwd_kernel"
    out["kernels"][key] = {
        "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
        "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
        "gpu_time_us": val(r, "gpu__time_duration.sum"), "grid": int(val(r, "launch__grid_size")),
        "block": int(val(r, "launch__block_size")), "registers": int(val(r, "launch__registers_per_thread")),
        "smem_dynamic_bytes": round(val(r, "launch__shared_mem_per_block_dynamic")),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
        "warp_instructions_per_arc": val(r, "smsp__inst_executed.sum") / bench["config"]["arcs_per_gpu"],
    }
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)

lrows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
for i, r in enumerate(lrows):
    if "Kernel Name" in r:
        h, st = r, i + 1
        break
agg = collections.OrderedDict()
for r in lrows[st:]:
    if len(r) < len(h):
        continue
    n = r[h.index("Kernel Name")]
    n = "nfst_fwd_kernel" if "nfst_fwd" in n else ("nfst_bwd_kernel" if "nfst_bwd" in n else n[:40])
    v = float(r[h.index("Metric Value")].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[h.index("Metric Unit")], 1)
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
L = ["ncu launch list: ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nfst_ -c 12 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu",
     "(a bench step launches exactly these two kernels; per-launch times under ncu are cold-cache and serialised: compare SHARES)", ""]
for k, v in agg.items():
    L.append(f"{k:20s} launches {len(v):3d}  mean {sum(v) / len(v):8.1f} us  share of step {100 * sum(v) / tot:5.1f}%")
r = bench["roofline"]
L += ["", f"bench.py (CUDA events, same build): fwd {r['fwd_kernel']['kernel_ms'] * 1e3:.1f} us, bwd {r['kernel_ms'] * 1e3:.1f} us "
          f"-> bwd share {100 * r['kernel_ms'] / (r['kernel_ms'] + r['fwd_kernel']['kernel_ms']):.1f}%"]
open(f"profiles/{tag}_launch_list_summary.txt", "w").write("\n".join(L) + "\n")
shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launches.csv")
shutil.copy(f"gpurun_out/bench_{tag}.json", f"profiles/{tag}_bench_line.json")
shutil.copy(f"gpurun_out/bench_{tag}_ref.json", f"profiles/{tag}_bench_reference_line.json")
shutil.copy(f"gpurun_out/configs_{tag}.txt", f"profiles/{tag}_configs.txt")
print(open(f"profiles/{tag}_launch_list_summary.txt").read())
print(json.dumps(out["kernels"], indent=1))
