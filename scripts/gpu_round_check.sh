mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
( timeout 300 python -m pytest tests/test_gpu_crops.py -x -q > gpurun_out/t_crops.log 2>&1; echo "exit $?" >> gpurun_out/t_crops.log )
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log )
( timeout 120 python scripts/run_crops_once.py 16 5 > gpurun_out/crops_once.json 2> gpurun_out/crops_once.err )
( YTK_DEVICE_CROPS=1 timeout 400 python bench.py > gpurun_out/bench_dev1.json 2> gpurun_out/bench_dev1.err; echo "exit $?" >> gpurun_out/bench_dev1.err )
( timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:crop_ --csv --log-file gpurun_out/r01_crops_launches.csv python scripts/run_crops_once.py 16 2 > gpurun_out/ncu_crops.log 2>&1 )
( timeout 150 ncu --set full --clock-control none --import-source on -k regex:crop_ -c 2 -f -o gpurun_out/r01_crops_full python scripts/run_crops_once.py 16 1 > gpurun_out/ncu_crops_full.log 2>&1 )
( YTK_DEVICE_CROPS=0 timeout 300 python bench.py --no-cpu > gpurun_out/bench_dev0.json 2> gpurun_out/bench_dev0.err; echo "exit $?" >> gpurun_out/bench_dev0.err )
( timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "exit $?" >> gpurun_out/t_all.log )
tail -3 gpurun_out/t_crops.log gpurun_out/smoke.log gpurun_out/t_all.log
