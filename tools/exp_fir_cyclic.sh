#!/bin/bash
# same-box A/B of the FIR pass's strip order at 128^2: strips handed out round-robin (default) vs contiguous tile runs (NBE_FIR_CYCLIC=0)
python -m pytest tests/test_conv_flat_gpu.py tests/test_up_fused_gpu.py -q -m gpu -x -k "fir or up_layer or up" 2>&1 | tail -2
for i in 1 2 3; do
echo -n "runs    "; NBE_FIR_CYCLIC=0 python tools/one_fir.py 256 128 20
echo -n "cyclic  "; python tools/one_fir.py 256 128 20
done
echo -n "runs  64 "; NBE_FIR_CYCLIC=0 python tools/one_fir.py 256 64 20
echo -n "cyclic 64 "; python tools/one_fir.py 256 64 20
