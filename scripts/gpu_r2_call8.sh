# round 2, call 8: attention with units dealt to both slots; quick tests + A/B + launch list
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parseq.py -q -x > gpurun_out/t_parseq.log 2>&1; echo "exit $?" >> gpurun_out/t_parseq.log )
tail -3 gpurun_out/t_parseq.log
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_parseq_launches_3200x264_v4.csv python scripts/run_parseq_once.py 3200 264 1 0 > gpurun_out/ncu_parseq.log 2>&1 )
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_attn_tc_full_v4 python scripts/run_parseq_once.py 3200 264 1 0 > gpurun_out/ncu_attn_full.log 2>&1 )
( timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r02_e.json 2> gpurun_out/bench_r02_e.err; echo "exit $?" >> gpurun_out/bench_r02_e.err )
cut -c1-300 gpurun_out/bench_r02_e.json; tail -3 gpurun_out/bench_r02_e.err
