#!/bin/bash
# compute-sanitizer passes over the kernels with hand-rolled shared-memory pipelines:
#   persistent upfirdn2d (bulk copies, two buffers), the TMA-tiled FIR pass with noise tiles, the generator's flat path.
# Logs go to gpurun_out/ (copied into profiles/ by hand).  Each pass is bounded by `timeout`.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool tests...
  local name=$1 tool=$2; shift 2
  echo "=== $name ($tool): $*" > gpurun_out/sanitize_$name.log
  timeout 700 $SAN --tool $tool --print-limit 20 python -m pytest -x -q -m gpu "$@" >> gpurun_out/sanitize_$name.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitize_$name.log
  tail -n 6 gpurun_out/sanitize_$name.log
}
run memcheck_ops memcheck tests/test_ops_gpu.py -k "upfirdn or bias_act"
run racecheck_ops racecheck tests/test_ops_gpu.py -k "upfirdn2d_generator_and_upsample_shapes or upfirdn2d_golden or packed_path"
run memcheck_gen memcheck tests/test_generator_gpu.py -k "bf16 or mapping"
run racecheck_fir racecheck tests/test_conv_flat_gpu.py -k "up_layer or fir"
