python tools/tile_tune.py 100000
NFST_FLOW_BITS=64 python tools/tile_tune.py 100000
NFST_FLOW_BITS=64 python tools/tile_tune.py 300000
NFST_FLOW_BITS=64 python tools/tile_tune.py 1000000
NFST_FLOW_BITS=64 python tools/tile_tune.py 10000
NFST_FLOW_BITS=64 timeout 600 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
