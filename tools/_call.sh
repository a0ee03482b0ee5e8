OUT=gpurun_out/r2_call18; mkdir -p $OUT
python bench.py --workload c2 --steps 3 --warmup 3 > $OUT/plain.json 2> $OUT/plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
python tools/time_attn.py d128 > $OUT/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 10 -c 1 -o $OUT/d128 python tools/time_attn.py d128 > $OUT/ncu2.log 2>&1; tail -2 $OUT/ncu2.log
