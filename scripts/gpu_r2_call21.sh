# round 2, call 21 (1 GPU): the two failures of call 20 in full + compute-sanitizer over the DocumentAnalyzer test
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_dbnet.py -q -x -k "own_map" 2>&1 | tail -60 ) > gpurun_out/t_c21_dbnet.log; cat gpurun_out/t_c21_dbnet.log
( timeout 300 python -m pytest tests/test_gpu_api.py -q -x -k "batched_pages" 2>&1 | head -150 ) > gpurun_out/t_c21_api.log; head -90 gpurun_out/t_c21_api.log
( timeout 500 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_api.py -q -x -k "batched_pages" 2>&1 | grep -v "^$" | head -120 ) > gpurun_out/t_c21_sanitizer.log; head -100 gpurun_out/t_c21_sanitizer.log
