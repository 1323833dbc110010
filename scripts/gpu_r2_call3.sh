# round 2, call 3: full GPU suite with the tcgen05 attention default, launch lists, attention capture, new bench
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "exit $?" >> gpurun_out/t_all.log )
tail -4 gpurun_out/t_all.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log )
tail -4 gpurun_out/smoke.log
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_parseq_launches_3200_tc.csv python scripts/run_parseq_once.py 3200 184 1 0 > gpurun_out/ncu_parseq.log 2>&1 )
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_attn_tc_full python scripts/run_parseq_once.py 3200 184 1 0 > gpurun_out/ncu_attn_full.log 2>&1 )
( timeout 600 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; echo "exit $?" >> gpurun_out/bench_r02_a.err )
cut -c1-1500 gpurun_out/bench_r02_a.json; tail -5 gpurun_out/bench_r02_a.err
( timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_bench_launches_n1.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-window --no-extra > gpurun_out/ncu_bench.log 2>&1 )
ls -la gpurun_out | tail -20
