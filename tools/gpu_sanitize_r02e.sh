#!/bin/bash
# compute-sanitizer passes over what sessions 6-7 changed after the r02d run: the pack / unpack kernels of the operator-surface
# modulated_conv2d (pointers advancing by constant strides, gap column zeroed by the last image tile), enc7_pad_kernel with 8 columns
# per thread, the kernels' global reads moved below griddepcontrol.wait, and the single styles launch with its table of row ranges
# read through stride-0 views.  Logs: gpurun_out/sanitize_r02e_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool tests...
  local name=$1 tool=$2; shift 2
  echo "=== $name ($tool): $*" > gpurun_out/sanitize_r02e_$name.log
  timeout 240 $SAN --tool $tool --print-limit 20 python -m pytest -x -q -m gpu "$@" >> gpurun_out/sanitize_r02e_$name.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitize_r02e_$name.log
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|exit code" gpurun_out/sanitize_r02e_$name.log | tail -n 4
}
run memcheck_modconv memcheck tests/test_ops_gpu.py -k "modconv"
run memcheck_styles memcheck tests/test_parity_holes_gpu.py tests/test_generator_gpu.py -k "styles_table or bf16"
run memcheck_enc7 memcheck tests/test_enc7_toeplitz_gpu.py
run racecheck_modconv racecheck tests/test_ops_gpu.py -k "modconv_tensor_core_golden"
