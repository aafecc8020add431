mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/r2j_pytest_sharded.log 2>&1
tail -5 gpurun_out/r2j_pytest_sharded.log
( time KB_SHARD_TIMING=1 timeout 300 $TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 2 --mode lshard --e2e-steps 1 ) > gpurun_out/r2j_n2_lshard.json 2> gpurun_out/r2j_n2_lshard.err
echo "lshard rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2j_n2_lshard.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','op_applies_per_step','max_residual')}, d['roofline']['ms_per_sweep'])"; grep "shard timing rank 0" gpurun_out/r2j_n2_lshard.err | tail -4
