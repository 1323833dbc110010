# round 2, call 5: PDL with late triggers for multi-wave kernels; A/B; bench
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parseq.py tests/test_gpu_kernels.py -q -x > gpurun_out/t_parseq.log 2>&1; echo "exit $?" >> gpurun_out/t_parseq.log )
tail -3 gpurun_out/t_parseq.log
for v in default nopdl nofused_nopdl; do
  case $v in
    default) envs="";;
    nopdl) envs="YTK_NO_PDL=1";;
    nofused_nopdl) envs="YTK_NO_FUSED_HEAD=1 YTK_NO_PDL=1";;
  esac
  ( env $envs timeout 200 python scripts/run_parseq_once.py 3200 264 4 0 > gpurun_out/parseq_once_$v.log 2>&1 )
  echo "== $v"; tail -2 gpurun_out/parseq_once_$v.log | cut -c1-200
done
( timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err; echo "exit $?" >> gpurun_out/bench_r02_c.err )
cut -c1-300 gpurun_out/bench_r02_c.json; tail -3 gpurun_out/bench_r02_c.err
( YTK_NO_PDL=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_c_nopdl.json 2> gpurun_out/bench_r02_c_nopdl.err )
cut -c1-300 gpurun_out/bench_r02_c_nopdl.json
