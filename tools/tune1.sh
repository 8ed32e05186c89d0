python tools/tile_tune.py 10000
NFST_TILE_WARPS=1 python tools/tile_tune.py 10000
NFST_TILE_WARPS=4 python tools/tile_tune.py 10000
python tools/tile_tune.py 30000
python tools/tile_tune.py 100000
python tools/tile_tune.py 1000000 1024
NFST_TILE_BLOCK_ARCS=2048 python tools/tile_tune.py 1000000 1024
NFST_TILE_BLOCK_ARCS=3072 python tools/tile_tune.py 1000000 1024
