OUT=gpurun_out/r2_call49; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kv_cache.py tests/test_qattn_golden.py -m gpu -q 2>&1 | tail -4 | tee $OUT/test.log
LOWBIT_KV_MAGIC=1 timeout 600 python -m pytest tests/test_kv_cache.py -m gpu -q 2>&1 | tail -2 | tee -a $OUT/test.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kv_attn_partial -s 4 -c 1 -o $OUT/kv_v6 -f python tools/time_kv.py 4 32 16384 128 4 > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kv_attn -s 6 -c 4 --csv --log-file $OUT/r2_kv_attn_ncu.csv python tools/time_kv.py 4 32 16384 128 4 > /dev/null 2>&1
timeout 600 python -m pytest tests/test_qattn_golden.py -q -s -m gpu 2>&1 | grep -E "npz f8|npz f16" | sed 's/^[.F]*//' > $OUT/qattn_cmp.txt
