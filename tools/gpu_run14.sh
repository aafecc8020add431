mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide_nodes" ) > gpurun_out/r2o_pytest_wide.log 2>&1
tail -4 gpurun_out/r2o_pytest_wide.log
timeout 300 python bench.py --P 672 --b 676 --steps 4 --warmup 2 --no-cpu --e2e-steps 1 > gpurun_out/r2o_bench_kore_rule_E1e-8_P672_b676.json 2> gpurun_out/r2o_bench_kore_rule.err
python -c "
import json; d=json.load(open('gpurun_out/r2o_bench_kore_rule_E1e-8_P672_b676.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','op_applies_per_step','max_residual','protocol_fallbacks')}, d['roofline']['ms_per_sweep'], d['e2e']['value'])"; tail -3 gpurun_out/r2o_bench_kore_rule.err
