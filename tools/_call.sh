OUT=gpurun_out/r2_call54; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2_call54/bench_c2.json') if l.startswith('{')][0])
print('value',d['value'],'ms',d['ms_per_step'],'attn',d['attn_only'],'e2e',d['e2e']['ms'],d['e2e']['value'])
P
