#!/bin/bash
# same-box A/B of the last torch kernels taken out of the step: b4 input written by the styles launch (NBE_NO_STYLES_IN4=1 = the
# torch expression, 3 kernels) and (colors + 1) / 2 as one nbe_bias_act launch (NBE_TORCH_COLORS=1 = two torch kernels + a clone)
for i in 1 2 3; do
NBE_NO_STYLES_IN4=1 NBE_TORCH_COLORS=1 python bench.py --steps 50 --warmup 5 --no-incumbent --no-canvas --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('torch ', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'])"
python bench.py --steps 50 --warmup 5 --no-incumbent --no-canvas --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fused ', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'])"
done
