# round 2, call 17 (1 GPU): RT-DETR + API + dbpost + dbnet tests after the test fix
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_rtdetr.py tests/test_gpu_api.py tests/test_gpu_dbpost.py tests/test_gpu_dbnet.py -q 2>&1 | tail -15 ) > gpurun_out/t_c17.log; cat gpurun_out/t_c17.log
