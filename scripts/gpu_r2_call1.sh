# round 2, call 1: fp16 operands - full GPU test-suite (incl. the 2 x 2048-crop identity test), smoke, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
( timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/t_all.log 2>&1; echo "exit $?" >> gpurun_out/t_all.log )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log )
( timeout 400 python bench.py --no-cpu > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err; echo "exit $?" >> gpurun_out/bench_f16.err )
grep -h "identity\|dbnet\]\|passed\|failed\|Error\|assert" gpurun_out/t_all.log | cut -c1-400 | tail -40
tail -5 gpurun_out/smoke.log
cut -c1-600 gpurun_out/bench_f16.json
