# round 2, call 22 (2 GPUs): handles on two GPUs in one process (device guard, per-device function attributes), API tests
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_gpu_api.py tests/test_gpu_multirank.py -q --durations=5 2>&1 | head -150 ) > gpurun_out/t_c22.log; tail -40 gpurun_out/t_c22.log
