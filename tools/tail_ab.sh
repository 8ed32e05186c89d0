#!/bin/bash
# tools/tile_tune.py over "arcs:tailmax[:warps]" cases (NFST_TILE_TAILMAX = out-degree above which a state is a slice of its own)
mkdir -p gpurun_out
for c in "$@"; do
  IFS=: read a t w <<< "$c"
  echo "== arcs=$a tailmax=$t warps=${w:-0}"
  NFST_TILE_TAILMAX=$t NFST_TILE_WARPS=${w:-0} timeout 300 python tools/tile_tune.py $a 2>&1 | grep -v "^$" | tail -4
done 2>&1 | tee gpurun_out/tail_ab.txt
