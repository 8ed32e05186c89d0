N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu > gpurun_out/b$N.json 2> gpurun_out/b$N.err; echo "n$N rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/b$N.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("theta_step",{}).get("arcs_per_s"))
for r in d["ranks"]: print("   ", r["rank"], round(r["ms_per_step"],4), round(r["first_kernel_ms"],4), round(r["second_kernel_ms"],4), round(r["between_kernels_ms"],4), round(r["host_enqueue_ms"],4), r["sm_mhz"])
PY
