OUT=gpurun_out/r2_call51; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; tail -c 1500 $OUT/bench_c2.json | head -c 1200; echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
