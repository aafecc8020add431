mkdir -p gpurun_out
for v in r1 nobp cur; do
  if [ $v = cur ]; then export KB_LIB_PATH=$PWD/kore_b200/libkoreb200.so; else export KB_LIB_PATH=$PWD/kore_b200/libkoreb200_$v.so; fi
  timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2c_ab_$v.json 2> gpurun_out/r2c_ab_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r2c_ab_$v.json')); print('$v', {k:d[k] for k in ('value','ms_per_step','factor_ms')}, d['roofline']['ms_per_sweep'])"
done
