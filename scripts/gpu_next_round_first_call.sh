# First gpurun call of the next round (1 GPU, ~6 min): re-validate HEAD on a fresh box and collect the A/B numbers the
# open decisions in DESIGN.md section 8 need.  Usage: gpurun --timeout 900 -- 'bash scripts/gpu_next_round_first_call.sh'
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "exit $?" >> gpurun_out/t_all.log )
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log )
( timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err )
( YTK_DEVICE_CROPS=1 timeout 300 python bench.py --no-cpu > gpurun_out/bench_device_crops.json 2> gpurun_out/bench_device_crops.err )
# per-launch list of one bench step (kernel shares; the timed step = the last ~1600 launches)
( timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-window > gpurun_out/ncu_bench.log 2>&1 )
tail -3 gpurun_out/t_all.log gpurun_out/smoke.log
grep -h '^{' gpurun_out/bench_default.json gpurun_out/bench_device_crops.json | cut -c1-260
