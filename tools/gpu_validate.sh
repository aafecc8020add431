#!/bin/bash
# What a round's GPU validation runs (under gpurun, from the repo root):
#   bash tools/gpu_validate.sh 1 <tag>     one GPU : pytest -m gpu, bench.py (with the CPU baseline leg)
#   bash tools/gpu_validate.sh N <tag>     N GPUs  : l-sharded parity checks on the committed fixtures (both paths),
#                                                    bench.py --mode lshard, bench.py (one shift per GPU + attached lshard leg)
# Results go to gpurun_out/<tag>_*.
N=${1:-1}
TAG=${2:-val}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$N" = "1" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
  tail -4 gpurun_out/${TAG}_pytest_gpu.log
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_1gpu.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','protocol_fallbacks','gpu_launches')}, d['roofline']['ms_per_sweep'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
else
  ( time timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest_sharded.log 2>&1
  tail -3 gpurun_out/${TAG}_pytest_sharded.log
  ( time timeout 200 $TR --nproc-per-node $N --master-port 29561 tools/run_sharded_check.py spinover magnetic_small dormy ) > gpurun_out/${TAG}_shard${N}_check.log 2>&1
  echo "sharded check N=$N rc=$?"; grep -cE " OK" gpurun_out/${TAG}_shard${N}_check.log
  ( time KB_SHARD_TIMING=1 timeout 200 $TR --nproc-per-node $N --master-port 29571 bench.py --gpus $N --steps 8 --warmup 3 --mode lshard --e2e-steps 2 ) > gpurun_out/${TAG}_n${N}_lshard.json 2> gpurun_out/${TAG}_n${N}_lshard.err
  echo "lshard N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_n${N}_lshard.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','max_residual')}, d['roofline']['ms_per_sweep'], d['e2e']['value'])"
  ( time timeout 300 $TR --nproc-per-node $N --master-port 29581 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/${TAG}_n${N}_shifts.json 2> gpurun_out/${TAG}_n${N}_shifts.err
  echo "shifts N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_n${N}_shifts.json')); print({k:d[k] for k in ('value','ms_per_step','protocol_fallbacks')}, d['e2e']['value'], d['lshard']['ms_per_step'], d['lshard']['speedup_vs_one_gpu_unit'])"
fi
