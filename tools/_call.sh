OUT=gpurun_out/r2_call39; mkdir -p $OUT
timeout 600 python -m pytest tests/test_qattn_golden.py -q -s -m gpu > $OUT/test.log 2>&1; grep -E "npz f8|npz f16|passed|failed" $OUT/test.log
