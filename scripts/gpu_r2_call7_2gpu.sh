# round 2, call 7 (2 GPUs): NCCL crop scatter test + skewed 2-GPU bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi2.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_multirank.py -q -s > gpurun_out/t_multirank.log 2>&1; echo "exit $?" >> gpurun_out/t_multirank.log )
tail -6 gpurun_out/t_multirank.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "exit $?" >> gpurun_out/bench_r02_n2.err )
grep -h '^{' gpurun_out/bench_r02_n2.json | cut -c1-500; tail -4 gpurun_out/bench_r02_n2.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --skew 0 > gpurun_out/bench_r02_n2_noskew.json 2> gpurun_out/bench_r02_n2_noskew.err )
grep -h '^{' gpurun_out/bench_r02_n2_noskew.json | cut -c1-300
