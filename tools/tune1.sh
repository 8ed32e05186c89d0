NFST_TILE_WARPS=32 python tools/tile_tune.py 1000000 1024
NFST_TILE_WARPS=8 python tools/tile_tune.py 1000000 1024
