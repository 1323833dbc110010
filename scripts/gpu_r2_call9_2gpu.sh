# round 2, call 9 (2 GPUs): three-phase crop scatter - NCCL test + skewed / unskewed 2-GPU bench (20 steps)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_kernels.py -q -k "multirank or scatter or attention" > gpurun_out/t_multirank.log 2>&1; echo "exit $?" >> gpurun_out/t_multirank.log )
tail -4 gpurun_out/t_multirank.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r02_n2_v2.json 2> gpurun_out/bench_r02_n2_v2.err; echo "exit $?" >> gpurun_out/bench_r02_n2_v2.err )
grep -h '^{' gpurun_out/bench_r02_n2_v2.json | cut -c1-300; tail -4 gpurun_out/bench_r02_n2_v2.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 --skew 0 > gpurun_out/bench_r02_n2_v2_noskew.json 2> gpurun_out/bench_r02_n2_v2_noskew.err )
grep -h '^{' gpurun_out/bench_r02_n2_v2_noskew.json | cut -c1-300
( timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_f.json 2> gpurun_out/bench_r02_f.err )
cut -c1-300 gpurun_out/bench_r02_f.json
