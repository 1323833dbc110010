# round 2, call 11 (1 GPU): device post-processing front half: tests, kernel times, full suite, bench
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dbpost.py -x -q 2>&1 | tail -15 ) > gpurun_out/t_dbpost.log
cat gpurun_out/t_dbpost.log
( timeout 120 python scripts/run_dbpost_once.py 8 5 ) > gpurun_out/dbpost_time.log 2>&1; cat gpurun_out/dbpost_time.log
( timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_dbpost_launches.csv python scripts/run_dbpost_once.py 8 2 ) > gpurun_out/dbpost_ncu.log 2>&1
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/t_all2.log; cat gpurun_out/t_all2.log
( timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_g.json 2> gpurun_out/bench_r02_g.err; echo "exit $?" >> gpurun_out/bench_r02_g.err )
grep -h '^{' gpurun_out/bench_r02_g.json | cut -c1-600; tail -3 gpurun_out/bench_r02_g.err
( YTK_DEVICE_POST=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_g_hostpost.json 2> gpurun_out/bench_r02_g_hostpost.err )
grep -h '^{' gpurun_out/bench_r02_g_hostpost.json | cut -c1-300
