# round 2, call 19 (1 GPU): TMA epilogue v2 (16-byte extent rule, whole-tile L2 prefetch of the residual, L2 promotion 256 B):
# kernel tests, A/B/C bench with the per-shape GEMM table, then the GPU suite with durations (identity test on fixtures)
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_kernels.py -q 2>&1 | tail -12 ) > gpurun_out/t_c19_kernels.log; cat gpurun_out/t_c19_kernels.log
for v in default nopf nopf_l2p128; do
  case $v in
    default) EXTRA="";;
    nopf) EXTRA="YTK_EPI_PF=0";;
    nopf_l2p128) EXTRA="YTK_EPI_PF=0 YTK_EPI_L2P=1";;
  esac
  rm -f gpurun_out/gemm_dump_$v.csv
  ( env $EXTRA YTK_GEMM_DUMP=gpurun_out/gemm_dump_$v.csv timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --no-e2e > gpurun_out/bench_c19_$v.json 2> gpurun_out/bench_c19_$v.err; echo "exit $?" >> gpurun_out/bench_c19_$v.err )
  grep -h '^{' gpurun_out/bench_c19_$v.json | cut -c1-200; tail -1 gpurun_out/bench_c19_$v.err
  python scripts/gemm_shape_table.py gpurun_out/gemm_dump_$v.csv gpurun_out/r02_gemm_shapes_v5_$v.json > gpurun_out/gemm_table_v5_$v.txt 2>&1; head -22 gpurun_out/gemm_table_v5_$v.txt
done
( timeout 900 python -m pytest tests -m gpu -q --durations=25 --ignore=tests/test_gpu_kernels.py 2>&1 | tail -45 ) > gpurun_out/t_c19_all.log; cat gpurun_out/t_c19_all.log
