"""Print the key numbers of bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    r = d["roofline"]
    o = r["other_kernel"]
    e = d.get("e2e") or {}
    print(f"{path}: {d['value']:.3e} arcs/s  step {d['ms_per_step']:.3f} ms | {r['kernel'].split(' ')[0]} {r['kernel_ms']:.3f} ms frac {r['frac']:.3f} | "
          f"{o['kernel'].split(' ')[0]} {o['kernel_ms']:.3f} ms frac {o['frac']:.3f} | step frac {r['step']['frac']:.3f} | e2e {e.get('value', 0):.3e} "
          f"({e.get('ms_per_step', 0):.2f} ms, {e.get('h2d_bytes_per_step', 0) / 1e6:.0f} MB) | {d['config'].get('execution')}")
    for s in d.get("sweep", []):
        print("   sweep", s)
    for c in d.get("configs", []):
        print(f"   {c['config']:42s} {c['execution']:22s} f+b {c['fwd_bwd_ms']:.3f} ms {c['fwd_bwd_arcs_per_s']:.3e} arcs/s | Viterbi {c['viterbi_ms']:.3f} ms {c['viterbi_arcs_per_s']:.3e}")
