set -x
python tools/one_flat.py convT 256 64 128
NBE_FLAT_NO_RESIDENT=1 python tools/one_flat.py convT 256 64 128
NBE_FLAT_NO_RESIDENT=1 NBE_FLAT_ROUND_ROBIN=1 python tools/one_flat.py convT 256 64 128
for g in 20 27 37 50; do NBE_FLAT_NO_RESIDENT=1 NBE_FLAT_ROUND_ROBIN=1 NBE_FLAT_GRID_PAIRS=$g python tools/one_flat.py convT 256 64 128; done
for g in 20 27 37; do NBE_FLAT_GRID_PAIRS=$g python tools/one_flat.py convT 256 64 128; done
python tools/one_fir.py 256 128
for g in 216 188 148 96; do NBE_FIR_GRID=$g python tools/one_fir.py 256 128; done
python tools/one_flat.py convT 256 32 384
for g in 37 50; do NBE_FLAT_GRID_PAIRS=$g python tools/one_flat.py convT 256 32 384; done
python tools/one_fir.py 256 64
for g in 148 96; do NBE_FIR_GRID=$g python tools/one_fir.py 256 64; done
