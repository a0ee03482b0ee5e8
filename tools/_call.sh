OUT=gpurun_out/r2_call59; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kv_attn_partial -s 4 -c 1 -o $OUT/kv_v7 -f python tools/time_kv.py 4 32 16384 128 4 > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kv_attn -s 6 -c 4 --csv --log-file $OUT/r2_kv_attn_ncu.csv python tools/time_kv.py 4 32 16384 128 4 > /dev/null 2>&1
for m in 2 4 8; do
LOWBIT_KV_DEBUG=$((m+1)) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kv_attn_partial -s 6 -c 2 --csv --log-file $OUT/l_$m.csv python tools/time_kv.py 4 32 16384 128 4 > /dev/null 2>&1
echo "tma no-arith, without stream $m: $(grep -E 'partial' $OUT/l_$m.csv | awk -F'","' '{print $(NF)}' | tr -d '"' | tr '\n' ' ')" | tee -a $OUT/dbg.log
done
