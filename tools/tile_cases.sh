for cfg in "100000 64 4 -1 pullcond" "100000 64 4 0 pullcond" "100000 64 4 1 pullcond" "100000 64 4 2 pullcond" "100000 64 4 3 pullcond" "100000 64 2 -1 fb" "20000 16 2 -1 fb"; do
  set -- $cfg
  echo "=== arcs=$1 levels=$2 B=$3 which=$4 mode=$5"
  CUDA_LAUNCH_BLOCKING=1 timeout 120 python tests/report/tile_debug2.py $1 $2 $3 $4 $5 2>&1 | grep -E "groups|ok|Error|error" | head -6
done
