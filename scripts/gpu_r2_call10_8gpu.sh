# round 2, call 10 (8 GPUs): the default (skewed) multi-GPU bench as the driver launches it
mkdir -p gpurun_out
nproc > gpurun_out/nproc8.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err; echo "exit $?" >> gpurun_out/bench_r02_n8.err )
grep -h '^{' gpurun_out/bench_r02_n8.json | cut -c1-400; tail -3 gpurun_out/bench_r02_n8.err
