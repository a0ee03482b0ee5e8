OUT=gpurun_out/r2_call53; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "one_call" 2>&1 | tail -5 | tee $OUT/test.log
timeout 300 python tools/time_small.py 2>&1 | tail -5 | tee $OUT/small.log
LOWBIT_ONE_CALL=0 timeout 300 python tools/time_small.py 2>&1 | tail -5 | sed 's/$/ (five-call path)/' | tee -a $OUT/small.log
timeout 300 python tools/time_prep.py 2>&1 | head -2 | tee -a $OUT/small.log
timeout 300 python -m pytest tests/test_tensor_handoff.py -m gpu -q 2>&1 | tail -2
