mkdir -p gpurun_out
( YTK_DEVICE_CROPS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu > gpurun_out/bench_n2_dev1.json 2> gpurun_out/bench_n2_dev1.err; echo "exit $?" >> gpurun_out/bench_n2_dev1.err )
( YTK_DEVICE_CROPS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu > gpurun_out/bench_n2_dev0.json 2> gpurun_out/bench_n2_dev0.err; echo "exit $?" >> gpurun_out/bench_n2_dev0.err )
grep -h '^{' gpurun_out/bench_n2_dev1.json gpurun_out/bench_n2_dev0.json | cut -c1-300
tail -3 gpurun_out/bench_n2_dev1.err
