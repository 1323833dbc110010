# round 2, call 12 (1 GPU): chunk-run CCL: tests, kernel times, bench A/B device post vs host post, GEMM shape dump
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dbpost.py -x -q 2>&1 | tail -15 ) > gpurun_out/t_dbpost.log
cat gpurun_out/t_dbpost.log
( timeout 120 python scripts/run_dbpost_once.py 8 5 ) > gpurun_out/dbpost_time.log 2>&1; cat gpurun_out/dbpost_time.log
( timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_dbpost_launches_v2.csv python scripts/run_dbpost_once.py 8 2 ) > gpurun_out/dbpost_ncu.log 2>&1
( timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_h.json 2> gpurun_out/bench_r02_h.err; echo "exit $?" >> gpurun_out/bench_r02_h.err )
grep -h '^{' gpurun_out/bench_r02_h.json | cut -c1-300; tail -3 gpurun_out/bench_r02_h.err
( YTK_DEVICE_POST=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_h_hostpost.json 2> gpurun_out/bench_r02_h_hostpost.err )
grep -h '^{' gpurun_out/bench_r02_h_hostpost.json | cut -c1-300
rm -f gpurun_out/gemm_dump.csv
( YTK_GEMM_DUMP=gpurun_out/gemm_dump.csv timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/bench_dump.json 2> gpurun_out/bench_dump.err )
python scripts/gemm_shape_table.py gpurun_out/gemm_dump.csv gpurun_out/r02_gemm_shapes.json | head -60
( timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_dbnet.py -x -q 2>&1 | tail -5 ) > gpurun_out/t_api.log; cat gpurun_out/t_api.log
