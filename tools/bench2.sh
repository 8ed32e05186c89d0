N=${1:-2}
python bench.py --no-sweep --no-cpu > gpurun_out/b1.json 2> gpurun_out/b1.err; echo "n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu > gpurun_out/b$N.json 2> gpurun_out/b$N.err; echo "n$N rc=$?"
tail -3 gpurun_out/b1.err gpurun_out/b$N.err
python - <<PY
import json
for f in ("gpurun_out/b1.json","gpurun_out/b$N.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("theta_step"))
        for r in d["ranks"]: print("   ", r)
    except Exception as e: print(f, "ERR", e)
PY
