mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_crops.py -x -q > gpurun_out/t_crops2.log 2>&1; echo "exit $?" >> gpurun_out/t_crops2.log )
( timeout 200 python scripts/gpu_diag_crops.py > gpurun_out/diag_crops.log 2>&1; echo "exit $?" >> gpurun_out/diag_crops.log )
( YTK_DEVICE_CROPS=1 timeout 300 python bench.py --no-cpu > gpurun_out/bench_dev1_v2.json 2> gpurun_out/bench_dev1_v2.err; echo "exit $?" >> gpurun_out/bench_dev1_v2.err )
tail -2 gpurun_out/t_crops2.log; tail -12 gpurun_out/diag_crops.log
