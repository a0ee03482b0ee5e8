#!/bin/bash
# Multi-GPU evidence run: tools/scale_run.sh N  -> ring/head-shard check, then bench lines for the sharded configs
# and the default weak-scaling line at N GPUs.  Output: gpurun_out/scale_N.jsonl
N=${1:-2}
OUT=gpurun_out/scale_${N}.jsonl
: > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" -gt 1 ]; then
  timeout 300 $TR tools/ring_check.py 2>&1 | grep -E "^ring|^head" | tee -a gpurun_out/scale_${N}_check.log
  RUN="$TR bench.py --gpus $N"
else
  RUN="python bench.py --gpus 1"
fi
for w in c4s c5 c5dyn c5f8; do
  S=5; [ "$w" = c4s ] && S=30   # a 1.5 ms step: 5 steps would time host jitter, not the GPUs
  timeout 400 $RUN --workload $w --steps $S --warmup 3 2>&1 | grep '^{"metric"' | tee -a $OUT
done
timeout 400 $RUN --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{"metric"' | tee -a $OUT
