# round 2, call 2: tcgen05 attention kernel - op-level parity (both V-descriptor conventions), model tests, phase times
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -s -k attention > gpurun_out/t_attn.log 2>&1; echo "exit $?" >> gpurun_out/t_attn.log )
( YTK_ATTN=vswap timeout 300 python -m pytest tests/test_gpu_kernels.py -q -s -k "attention_tc" > gpurun_out/t_attn_vswap.log 2>&1; echo "exit $?" >> gpurun_out/t_attn_vswap.log )
tail -15 gpurun_out/t_attn.log; tail -8 gpurun_out/t_attn_vswap.log
( timeout 600 python -m pytest tests/test_gpu_parseq.py tests/test_gpu_api.py -q -x > gpurun_out/t_parseq.log 2>&1; echo "exit $?" >> gpurun_out/t_parseq.log )
tail -5 gpurun_out/t_parseq.log
( timeout 200 python scripts/run_parseq_once.py 3200 184 3 0 > gpurun_out/parseq_once_tc.log 2>&1 )
( YTK_ATTN=legacy timeout 200 python scripts/run_parseq_once.py 3200 184 3 0 > gpurun_out/parseq_once_legacy.log 2>&1 )
tail -3 gpurun_out/parseq_once_tc.log gpurun_out/parseq_once_legacy.log
