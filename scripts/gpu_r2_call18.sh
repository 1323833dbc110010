# round 2, call 18 (1 GPU): TMA epilogue (UTMASTG / residual boxes by UTMALDG) - kernel tests, whole GPU suite, A/B bench vs the
# staged epilogue with the per-shape GEMM table, e2e trace, smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_c18.txt 2>&1
( timeout 400 python -m pytest tests/test_gpu_kernels.py -q 2>&1 | tail -25 ) > gpurun_out/t_c18_kernels.log; cat gpurun_out/t_c18_kernels.log
if grep -q "failed\|error\|Error\|Timeout" gpurun_out/t_c18_kernels.log; then
  echo "=== kernel tests FAILED with the TMA epilogue: diagnosing"
  ( YTK_EPI_SWZ=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "linear or conv" 2>&1 | tail -15 ) > gpurun_out/t_c18_kernels_noswz.log; cat gpurun_out/t_c18_kernels_noswz.log
  ( YTK_EPI=staged timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "linear or conv" 2>&1 | tail -8 ) > gpurun_out/t_c18_kernels_staged.log; cat gpurun_out/t_c18_kernels_staged.log
  export YTK_EPI=staged
fi
( timeout 1200 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_kernels.py 2>&1 | tail -25 ) > gpurun_out/t_c18_all.log; cat gpurun_out/t_c18_all.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 ) > gpurun_out/smoke_c18.log; cat gpurun_out/smoke_c18.log
rm -f gpurun_out/gemm_dump_tma.csv gpurun_out/gemm_dump_staged.csv
( YTK_GEMM_DUMP=gpurun_out/gemm_dump_tma.csv timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-e2e > gpurun_out/bench_c18_tma.json 2> gpurun_out/bench_c18_tma.err; echo "exit $?" >> gpurun_out/bench_c18_tma.err )
( YTK_EPI=staged YTK_GEMM_DUMP=gpurun_out/gemm_dump_staged.csv timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-e2e > gpurun_out/bench_c18_staged.json 2> gpurun_out/bench_c18_staged.err; echo "exit $?" >> gpurun_out/bench_c18_staged.err )
grep -h '^{' gpurun_out/bench_c18_tma.json gpurun_out/bench_c18_staged.json | cut -c1-260; tail -2 gpurun_out/bench_c18_tma.err gpurun_out/bench_c18_staged.err
python scripts/gemm_shape_table.py gpurun_out/gemm_dump_tma.csv gpurun_out/r02_gemm_shapes_v4_tma.json > gpurun_out/gemm_table_v4_tma.txt 2>&1; head -30 gpurun_out/gemm_table_v4_tma.txt
python scripts/gemm_shape_table.py gpurun_out/gemm_dump_staged.csv gpurun_out/r02_gemm_shapes_v4_staged.json > gpurun_out/gemm_table_v4_staged.txt 2>&1; head -30 gpurun_out/gemm_table_v4_staged.txt
( timeout 200 python scripts/gpu_trace_e2e.py 10 16 2 > gpurun_out/trace_c18.log 2>&1 ); head -40 gpurun_out/trace_c18.log
