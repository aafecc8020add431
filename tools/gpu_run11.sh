mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/r2l_pytest_sharded.log 2>&1
tail -4 gpurun_out/r2l_pytest_sharded.log
( time timeout 200 $TR --nproc-per-node 8 --master-port 29561 tools/run_sharded_check.py spinover magnetic_small dormy ) > gpurun_out/r2l_shard8_check.log 2>&1
echo "check8 rc=$?"; grep -E "rank 0|rror|WARN" gpurun_out/r2l_shard8_check.log | tail -5
for N in 8 4 2; do
( time KB_SHARD_TIMING=1 timeout 200 $TR --nproc-per-node $N --master-port 2957$N bench.py --gpus $N --steps 5 --warmup 3 --mode lshard --e2e-steps 2 ) > gpurun_out/r2l_n${N}_lshard.json 2> gpurun_out/r2l_n${N}_lshard.err
echo "lshard N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2l_n${N}_lshard.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','op_applies_per_step','max_residual')}, d['roofline']['ms_per_sweep'], d['e2e']['value'])"; grep "shard timing rank 0:" gpurun_out/r2l_n${N}_lshard.err | tail -3
done
( time timeout 200 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r2l_n8_shifts.json 2> gpurun_out/r2l_n8_shifts.err
echo "shifts N=8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2l_n8_shifts.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms')}, d['lshard'])"
