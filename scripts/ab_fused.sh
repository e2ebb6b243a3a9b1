#!/bin/bash
# A/B timing of library builds (ab_libs/<name>.so, scripts/build_variant.sh) on ONE box: box-to-box variance is
# +-5 %, larger than most of the effects being measured.   usage: bash scripts/ab_fused.sh <n> <name> <name> ...
N=$1; shift
for r in 1 2 3; do for v in "$@"; do echo -n "$v: "; SMPLB200_LIB=/root/repo/ab_libs/$v.so python scripts/run_fused_only.py $N 30 2>&1 | grep -E "k_fused|Error" | head -1; done; done
for v in "$@"; do echo -n "check $v: "; SMPLB200_LIB=/root/repo/ab_libs/$v.so timeout 300 python -m pytest tests/test_gpu_fused.py -x -q -k "f16_full_batch or shard or other_model" 2>&1 | tail -1; done
