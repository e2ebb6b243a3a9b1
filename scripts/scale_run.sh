#!/bin/bash
# Weak (4096 bodies per GPU) and strong (65,536 bodies in total) scaling at 1/2/4/8 GPUs of one box.
#   usage (on an 8-GPU box): bash scripts/scale_run.sh <tag> ["weak strong"] ["1 2 4 8"] [auto|dma|peer|collective|off]
TAG=${1:-r02}
MODES=${2:-"weak strong"}
NS=${3:-"1 2 4 8"}
XCHG=${4:-auto}
mkdir -p gpurun_out
P=29600
for N in $NS; do
  for MODE in $MODES; do
    EXTRA="--no-extras --no-cpu-baseline --steps 100 --warmup 10 --exchange $XCHG"
    [ $MODE = strong ] && EXTRA="$EXTRA --total-bodies 65536"
    OUT=gpurun_out/${TAG}_scale_${MODE}_${N}.json
    if [ $N = 1 ]; then
      python bench.py --gpus 1 $EXTRA > $OUT 2> gpurun_out/${TAG}_scale_${MODE}_${N}.err
    else
      P=$((P+1))
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
        bench.py --gpus $N $EXTRA > $OUT 2> gpurun_out/${TAG}_scale_${MODE}_${N}.err
    fi
    echo "$MODE N=$N rc=$? $(python -c "import json,sys; d=json.load(open('$OUT')); print(round(d['value']/1e6,2),'M bodies/s', round(d['ms_per_step'],4),'ms', 'e2e', round(d['e2e']['value']/1e6,2), 'small', round(d['e2e_small_outputs']['value']/1e6,2), d['impl_config']['exchange'] and d['impl_config']['exchange']['transport'])" 2>&1)"
  done
done
