#!/bin/bash
# compute-sanitizer passes over what session 5 changed: the up = 2 path of upfirdn2d_staged_kernel (zero-row / masked window loads,
# funnel-shifted 32-bit reads, persistent grid for 16-bit types), programmatic dependent launch on the tensor-core / FIR / encoder
# kernels, the --neg_slope encoder layout (affine kernels, transposed conv + one-tap FIR).  Logs: gpurun_out/sanitize_r02d_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool tests...
  local name=$1 tool=$2; shift 2
  echo "=== $name ($tool): $*" > gpurun_out/sanitize_r02d_$name.log
  timeout 900 $SAN --tool $tool --print-limit 20 python -m pytest -x -q -m gpu "$@" >> gpurun_out/sanitize_r02d_$name.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitize_r02d_$name.log
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|exit code" gpurun_out/sanitize_r02d_$name.log | tail -n 4
}
run memcheck_ops memcheck tests/test_ops_gpu.py -k "upfirdn or upsample"
run racecheck_ops racecheck tests/test_ops_gpu.py -k "upfirdn or upsample"
run memcheck_gen memcheck tests/test_generator_gpu.py tests/test_parity_holes_gpu.py -k "bf16 or encoder or neg_slope"
run racecheck_v2 racecheck tests/test_parity_holes_gpu.py -k "neg_slope_variant_bf16"
