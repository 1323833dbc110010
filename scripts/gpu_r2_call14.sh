# round 2, call 14 (1 GPU): wave-quantisation-aware N tile for small GEMMs: parity tests, GEMM shape dump, bench
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parseq.py -x -q 2>&1 | tail -8 ) > gpurun_out/t_c14.log
cat gpurun_out/t_c14.log
rm -f gpurun_out/gemm_dump.csv
( YTK_GEMM_DUMP=gpurun_out/gemm_dump.csv timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/bench_dump.json 2> gpurun_out/bench_dump.err )
python scripts/gemm_shape_table.py gpurun_out/gemm_dump.csv gpurun_out/r02_gemm_shapes_v3.json > gpurun_out/gemm_table_v3.txt; head -22 gpurun_out/gemm_table_v3.txt
( timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_j.json 2> gpurun_out/bench_r02_j.err; echo "exit $?" >> gpurun_out/bench_r02_j.err )
grep -h '^{' gpurun_out/bench_r02_j.json | cut -c1-300; tail -3 gpurun_out/bench_r02_j.err
( YTK_NO_WAVE_MODEL=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra --no-e2e > gpurun_out/bench_r02_j_nowave.json 2> gpurun_out/bench_r02_j_nowave.err )
grep -h '^{' gpurun_out/bench_r02_j_nowave.json | cut -c1-300
