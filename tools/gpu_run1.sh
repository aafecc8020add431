mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
( time timeout 400 python -m pytest tests/test_gpu_robustness.py -m gpu -x -q ) > gpurun_out/r2a_pytest_robust.log 2>&1
tail -25 gpurun_out/r2a_pytest_robust.log
( time timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_robustness.py ) > gpurun_out/r2a_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_1gpu.json 2> gpurun_out/r2a_bench_1gpu.err
tail -c 2500 gpurun_out/r2a_bench_1gpu.json; tail -5 gpurun_out/r2a_bench_1gpu.err
timeout 200 python tools/microbench/fp64_peak.py > gpurun_out/r2a_fp64_peak.json 2>&1
head -30 gpurun_out/r2a_fp64_peak.json
