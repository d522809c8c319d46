#!/bin/bash
# compute-sanitizer passes over the kernels rewritten at the end of round 2: row128 with eight epilogue warps, the templated flat
# kernel (compile-time tap programs, ld.shared staging read-back), the FIR pass in strip order (+ its out-of-line general path),
# the per-tap kernel with two tiles per CTA, the window-blend kernel.  Logs: gpurun_out/sanitize_r02c_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool tests...
  local name=$1 tool=$2; shift 2
  echo "=== $name ($tool): $*" > gpurun_out/sanitize_r02c_$name.log
  timeout 900 $SAN --tool $tool --print-limit 20 python -m pytest -x -q -m gpu "$@" >> gpurun_out/sanitize_r02c_$name.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitize_r02c_$name.log
  tail -n 5 gpurun_out/sanitize_r02c_$name.log
}
run memcheck_tc memcheck tests/test_conv_tc_gpu.py tests/test_conv_flat_gpu.py tests/test_blend_window_gpu.py
run memcheck_gen memcheck tests/test_generator_gpu.py -k "bf16 or encoder"
run racecheck_fir racecheck tests/test_conv_flat_gpu.py -k "up_layer or fir"
run racecheck_tc racecheck tests/test_conv_tc_gpu.py -k "row or torgb or 128"
