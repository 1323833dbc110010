# round 2, call 13 (1 GPU): new GELU + deterministic ASF + staged post front: tests, GEMM shape dump, bench A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dbpost.py tests/test_gpu_dbnet.py tests/test_gpu_kernels.py tests/test_gpu_parseq.py -x -q 2>&1 | tail -15 ) > gpurun_out/t_c13.log
cat gpurun_out/t_c13.log
rm -f gpurun_out/gemm_dump.csv
( YTK_GEMM_DUMP=gpurun_out/gemm_dump.csv timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/bench_dump.json 2> gpurun_out/bench_dump.err )
python scripts/gemm_shape_table.py gpurun_out/gemm_dump.csv gpurun_out/r02_gemm_shapes_v2.json | head -12
( timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_i.json 2> gpurun_out/bench_r02_i.err; echo "exit $?" >> gpurun_out/bench_r02_i.err )
grep -h '^{' gpurun_out/bench_r02_i.json | cut -c1-300; tail -3 gpurun_out/bench_r02_i.err
( YTK_DEVICE_POST=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_i_hostpost.json 2> gpurun_out/bench_r02_i_hostpost.err )
grep -h '^{' gpurun_out/bench_r02_i_hostpost.json | cut -c1-300
( timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_i2.json 2> gpurun_out/bench_r02_i2.err )
grep -h '^{' gpurun_out/bench_r02_i2.json | cut -c1-300
( timeout 900 python -m pytest tests/test_gpu_parseq_identity.py -x -q 2>&1 | tail -5 ) > gpurun_out/t_ident.log; cat gpurun_out/t_ident.log
