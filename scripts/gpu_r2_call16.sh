# round 2, call 16 (1 GPU): RT-DETR tests + API tests + launch list of one RT-DETR forward + bench with the new line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_rtdetr.py tests/test_gpu_api.py -x -q 2>&1 | tail -15 ) > gpurun_out/t_c16.log; cat gpurun_out/t_c16.log
( RT_ONCE=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_rtdetr_launches_b8.csv python scripts/gpu_probe_rtdetr.py layout 8 ) > gpurun_out/rtdetr_ncu.log 2>&1
LPF=$(grep launches_per_forward gpurun_out/rtdetr_ncu.log | awk '{print $2}'); echo "launches per forward: $LPF"
python scripts/ncu_traffic.py gpurun_out/r02_rtdetr_launches_b8.csv $((LPF + 1)) gpurun_out/r02_rtdetr_step_traffic_b8.json > gpurun_out/rtdetr_traffic.txt 2>&1; head -30 gpurun_out/rtdetr_traffic.txt
( timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_k.json 2> gpurun_out/bench_r02_k.err; echo "exit $?" >> gpurun_out/bench_r02_k.err )
grep -h '^{' gpurun_out/bench_r02_k.json | cut -c1-300; tail -3 gpurun_out/bench_r02_k.err
