mkdir -p gpurun_out
timeout 120 python tools/dev_expired_wait.py 40 600 > gpurun_out/r2b_expired_plain.log 2>&1
tail -4 gpurun_out/r2b_expired_plain.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_1gpu.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','protocol_fallbacks')}, d['roofline']['ms_per_sweep'], d['e2e']['value'])"
tail -3 gpurun_out/r2b_bench_1gpu.err
( time timeout 400 python -m pytest tests/test_gpu_robustness.py -m gpu -x -q ) > gpurun_out/r2b_pytest_robust.log 2>&1
tail -8 gpurun_out/r2b_pytest_robust.log
