OUT=gpurun_out/r2_call50; mkdir -p $OUT
nvidia-smi -L | head -8 > $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench8.json 2> $OUT/bench8.err; echo "bench8 rc=$?"; tail -c 400 $OUT/bench8.json
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -5 | tee $OUT/test_multi.log
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532"
timeout 400 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/bench4.json 2> $OUT/bench4.err; echo "bench4 rc=$?"
timeout 300 $TR bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > $OUT/bench8_ref.json 2> $OUT/bench8_ref.err; echo "ref8 rc=$?"; tail -c 300 $OUT/bench8_ref.json
