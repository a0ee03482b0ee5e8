OUT=gpurun_out/r2_call21; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp16_class or multi_precision or api_vs_oracle or golden or int4 or causal or lse" 2>&1 | tail -15
for pf in 2 3 1; do LOWBIT_ATTN_PF=$pf timeout 300 python tools/time_attn.py c2 c2c c2:k4f16 c4:k4f16 2>&1 | tee -a $OUT/time.log; done
