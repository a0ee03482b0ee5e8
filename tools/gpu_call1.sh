#!/bin/bash
# round-2 GPU call 1: first run of the 128-key-step kernel + peaks + reference-kernel goldens
OUT=gpurun_out/r2_call1; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== smoke subset" ; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention_matches_reference_kernel_golden or api_vs_oracle_and_sdpa or tail_masking or packed_int4 or int4_api" > $OUT/subset.log 2>&1; echo "subset rc=$?"; tail -5 $OUT/subset.log
echo "== timing"
for pf in 0 1 2 3; do LOWBIT_ATTN_PF=$pf timeout 120 python tools/time_attn.py c2 c2c c4:k4f16 c2:k4f16 >> $OUT/time.log 2>&1; done
LOWBIT_ATTN_WIDE=0 timeout 120 python tools/time_attn.py c2 c2c c4:k4f16 c2:k4f16 d128 d128c8k >> $OUT/time.log 2>&1
LOWBIT_ATTN_WIDE=2 timeout 120 python tools/time_attn.py c2:i8f8 c2:k4f8 >> $OUT/time.log 2>&1
LOWBIT_ATTN_WIDE=0 timeout 120 python tools/time_attn.py c2:i8f8 c2:k4f8 >> $OUT/time.log 2>&1
cat $OUT/time.log
echo "== ubench"; timeout 120 tools/ubench_wide > $OUT/ubench_wide.log 2>&1; cat $OUT/ubench_wide.log
echo "== peaks"; timeout 300 python tools/measure_peaks.py $OUT/peaks_int8_fp8.json > $OUT/peaks.log 2>&1; tail -6 $OUT/peaks.log
echo "== fused goldens"; timeout 300 python tools/make_golden_fused.py --variant fast --out $OUT/golden_fused > $OUT/golden_fused.log 2>&1; tail -8 $OUT/golden_fused.log
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests.log 2>&1; echo "suite rc=$?"; tail -8 $OUT/gpu_tests.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.json
