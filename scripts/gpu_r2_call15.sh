# round 2, call 15 (1 GPU): first run of the RT-DETRv2 engine: stage errors vs the oracle, timing
mkdir -p gpurun_out
( timeout 600 python scripts/gpu_probe_rtdetr.py layout 1 ) > gpurun_out/rtdetr_probe_layout.log 2>&1; tail -25 gpurun_out/rtdetr_probe_layout.log
( timeout 600 python scripts/gpu_probe_rtdetr.py table 2 ) > gpurun_out/rtdetr_probe_table2.log 2>&1; tail -25 gpurun_out/rtdetr_probe_table2.log
( timeout 600 python scripts/gpu_probe_rtdetr.py layout 8 ) > gpurun_out/rtdetr_probe_layout8.log 2>&1; tail -6 gpurun_out/rtdetr_probe_layout8.log
( timeout 300 compute-sanitizer --tool memcheck python scripts/gpu_probe_rtdetr.py table 1 ) > gpurun_out/rtdetr_sanitizer.log 2>&1; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/rtdetr_sanitizer.log; grep "ERROR SUMMARY" gpurun_out/rtdetr_sanitizer.log
