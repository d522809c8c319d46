#!/bin/bash
# same-box A/B of the FIR pass's tile loop: warps released per buffer by an mbarrier (default) vs one block-wide barrier per tile
# (NBE_FIR_DEBUG=3, the loop before this change)
python -m pytest tests/test_conv_flat_gpu.py tests/test_up_fused_gpu.py -q -m gpu -x -k "fir or up_layer or up" 2>&1 | tail -2
for i in 1 2 3; do
for sz in "256 128" "256 64" "256 32"; do
echo -n "barrier  "; NBE_FIR_DEBUG=3 python tools/one_fir.py $sz 20
echo -n "free-run "; python tools/one_fir.py $sz 20
done
done
