#!/bin/bash
# One-call GPU profiling recipe (B200_PROFILING.md): plain run first, then the ncu launch list and
# one `--set full` capture of the hot kernels.  Outputs land in gpurun_out/ (scratch); summarise
# them into profiles/ with scripts/summarize_ncu.py.
#   usage: bash scripts/profile_round.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --kernel-iters 5"
$B > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_list.log 2>&1
# the default path (k2 + the fused kernel) ...
ncu --set full --clock-control none --import-source on \
    -k regex:"k_fused_tc|k_pose_chain" -s 8 -c 2 -o gpurun_out/${TAG}_full $B > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
# ... and the unfused tensor-core kernels (timed by the bench's per-kernel section)
ncu --set full --clock-control none --import-source on \
    -k regex:"k_blend_tc|k_lbs_tc" -s 2 -c 2 -o gpurun_out/${TAG}_full_unfused $B > gpurun_out/${TAG}_ncu_full_unfused.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full_unfused.log
