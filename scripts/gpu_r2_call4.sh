# round 2, call 4: fused head epilogue + PDL + attention prefetch - tests, A/B phase times, bench with 20 steps
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "exit $?" >> gpurun_out/t_all.log )
tail -6 gpurun_out/t_all.log
for v in default nopdl nofused legacyattn; do
  case $v in
    default) envs="";;
    nopdl) envs="YTK_NO_PDL=1";;
    nofused) envs="YTK_NO_FUSED_HEAD=1";;
    legacyattn) envs="YTK_ATTN=legacy";;
  esac
  ( env $envs timeout 200 python scripts/run_parseq_once.py 3200 264 4 0 > gpurun_out/parseq_once_$v.log 2>&1 )
  echo "== $v"; tail -2 gpurun_out/parseq_once_$v.log | cut -c1-200
done
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_parseq_launches_3200x264_v2.csv python scripts/run_parseq_once.py 3200 264 1 0 > gpurun_out/ncu_parseq.log 2>&1 )
( timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo "exit $?" >> gpurun_out/bench_r02_b.err )
cut -c1-400 gpurun_out/bench_r02_b.json; tail -3 gpurun_out/bench_r02_b.err
