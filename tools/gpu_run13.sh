mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2n_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2n_pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_bench_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_1gpu.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','protocol_fallbacks')}, d['roofline']['ms_per_sweep'], d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2n_bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2n_launches.csv > gpurun_out/r2n_launches.summary.txt 2>&1; head -14 gpurun_out/r2n_launches.summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_sweep_fold -s 10 -c 1 -o gpurun_out/r2n_sweep_fold -f python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2n_ncu_full_sweep.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_fold_solution -s 10 -c 1 -o gpurun_out/r2n_fold_solution -f python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2n_ncu_full_sol.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_chain_factor -s 1 -c 1 -o gpurun_out/r2n_chain_factor -f python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2n_ncu_full_fac.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:kb_zgemm_batch -s 60 -c 1 -o gpurun_out/r2n_zgemm -f python tools/dev_zgemm.py > gpurun_out/r2n_ncu_full_zgemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
