OUT=gpurun_out/r2_call29; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
python tools/time_attn.py c2 c2c c2:k4f16 c4:k4f16 2>&1 | tee $OUT/time.log
