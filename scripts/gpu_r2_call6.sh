# round 2, call 6: P-in-TMEM attention, 4-stage stream pipeline - full tests, A/B, bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "exit $?" >> gpurun_out/t_all.log )
tail -8 gpurun_out/t_all.log
for v in default attnsmem pdl; do
  case $v in
    default) envs="";;
    attnsmem) envs="YTK_ATTN=smem";;
    pdl) envs="YTK_PDL=1";;
  esac
  ( env $envs timeout 200 python scripts/run_parseq_once.py 3200 264 4 0 > gpurun_out/parseq_once_$v.log 2>&1 )
  echo "== $v"; tail -2 gpurun_out/parseq_once_$v.log | cut -c1-200
done
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_parseq_launches_3200x264_v3.csv python scripts/run_parseq_once.py 3200 264 1 0 > gpurun_out/ncu_parseq.log 2>&1 )
( timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err; echo "exit $?" >> gpurun_out/bench_r02_d.err )
cut -c1-300 gpurun_out/bench_r02_d.json; tail -3 gpurun_out/bench_r02_d.err
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log ); tail -3 gpurun_out/smoke.log
