export NBE_CONVT_ONE_PHASE=1 NBE_FLAT_RES_MIN_ABUF=2
python -m pytest tests/test_conv_flat_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/one_flat.py convT 256 64 128
NBE_FLAT_ROUND_ROBIN=1 python tools/one_flat.py convT 256 64 128
for g in 37 33 30; do NBE_FLAT_ROUND_ROBIN=1 NBE_FLAT_GRID_PAIRS=$g python tools/one_flat.py convT 256 64 128; done
unset NBE_CONVT_ONE_PHASE NBE_FLAT_RES_MIN_ABUF
python tools/one_flat.py convT 256 64 128
