# round 2, call 20 (2 GPUs): regression check of the multi-rank path with the TMA epilogue - NCCL scatter test, the new DBNet
# own-map polygon test + DocumentAnalyzer tests, skewed 2-GPU bench (10 steps)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_dbnet.py tests/test_gpu_api.py -q --durations=8 2>&1 | tail -30 ) > gpurun_out/t_c20.log; cat gpurun_out/t_c20.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_c20_n2.json 2> gpurun_out/bench_c20_n2.err; echo "exit $?" >> gpurun_out/bench_c20_n2.err )
grep -h '^{' gpurun_out/bench_c20_n2.json | cut -c1-400; tail -3 gpurun_out/bench_c20_n2.err
