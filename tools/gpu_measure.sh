#!/bin/bash
# One measurement pass on the GPU box: tests, bench line (with sweep), reference line, ncu launch list, ncu full capture.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_measure.sh TAG'
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench_err.log; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref_line.json 2>> gpurun_out/${TAG}_bench_err.log; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tile_|sell_|nfst_' -c 12 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-sweep > gpurun_out/${TAG}_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:tile_pull|tile_flow' -s 6 -c 2 -f -o gpurun_out/${TAG}_prof_bench \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-sweep > gpurun_out/${TAG}_ncu_f.log 2>&1; echo "ncu full rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench_line.json
