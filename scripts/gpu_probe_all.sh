#!/bin/bash
# Runs every probe case in its own process with a timeout; never aborts on a failing case.
mkdir -p gpurun_out
N=${1:-70}
for c in chain blend_fp32 lbs_fma lbs_dense full_fp32_fma regress blend_bf16 blend_tf32 blend_bf16x3 lbs_tc full_bf16x3_tc full_fp32_tc_dense; do
  echo "=== $c"
  timeout 180 python scripts/gpu_probe.py $c $N 2>&1 | tail -12
  echo "exit=$?"
done
