# (the call itself ran with 1606 launches per step in the last line; a step has 1614 since the device post-processing
# kernels - the summary in profiles/ was recomputed with 1614 from the same launch list)
# round 2, call 23 (1 GPU): final state - whole GPU suite, smoke, the default bench (driver's shape: 20 steps), one --set full
# capture of the TMA-epilogue GEMMs of an encoder layer, launch list + DRAM bytes of one bench step
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -22 ) > gpurun_out/t_c23_all.log; cat gpurun_out/t_c23_all.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 ) > gpurun_out/smoke_c23.log; cat gpurun_out/smoke_c23.log
( timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c23.json 2> gpurun_out/bench_c23.err; echo "exit $?" >> gpurun_out/bench_c23.err )
grep -h '^{' gpurun_out/bench_c23.json | cut -c1-300; tail -2 gpurun_out/bench_c23.err
# encoder layer 0 of a 3200 x 264 px forward: launches 1..4 of gemm_tc_kernel = qkv, proj, fc1, fc2 (0 = patch embedding)
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 4 -f -o gpurun_out/r02_gemm_tma_full python scripts/run_parseq_once.py 3200 264 1 0 > gpurun_out/ncu_gemm_full.log 2>&1 )
( ncu -i gpurun_out/r02_gemm_tma_full.ncu-rep --page raw --csv > gpurun_out/r02_gemm_tma_full_raw.csv 2>/dev/null; ls -la gpurun_out/r02_gemm_tma_full* )
( timeout 420 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 4500 -c 4000 --csv --log-file gpurun_out/r02_bench_launches_n1_tma.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-window --no-extra > gpurun_out/ncu_bench_c23.log 2>&1 )
python scripts/ncu_traffic.py gpurun_out/r02_bench_launches_n1_tma.csv 1614 gpurun_out/r02_bench_step_traffic_tma.json > gpurun_out/traffic_c23.txt 2>&1; head -24 gpurun_out/traffic_c23.txt
