mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=8
( time timeout 120 $TR --nproc-per-node $N --master-port 29561 tools/run_sharded_check.py spinover magnetic_small dormy ) > gpurun_out/r2t_shard${N}_check.log 2>&1
echo "check$N rc=$?"; grep -E "rank 0|rror:|WARN" gpurun_out/r2t_shard${N}_check.log | tail -5 | cut -c1-200
( time KB_SHARD_TIMING=1 timeout 120 $TR --nproc-per-node $N --master-port 2957$N bench.py --gpus $N --steps 8 --warmup 3 --mode lshard --e2e-steps 2 ) > gpurun_out/r2t_n${N}_lshard.json 2> gpurun_out/r2t_n${N}_lshard.err
echo "lshard N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2t_n${N}_lshard.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','op_applies_per_step','max_residual')}, d['roofline']['ms_per_sweep'], d['e2e']['value'])"; grep "shard timing rank 0:" gpurun_out/r2t_n${N}_lshard.err | tail -3 | cut -c1-250; grep "rror:" gpurun_out/r2t_n${N}_lshard.err | head -3
( time KB_SHARD_NCCL_ONLY=1 timeout 120 $TR --nproc-per-node $N --master-port 2958$N bench.py --gpus $N --steps 8 --warmup 3 --mode lshard --e2e-steps 1 ) > gpurun_out/r2t_n${N}_lshard_nccl.json 2> gpurun_out/r2t_n${N}_lshard_nccl.err
echo "lshard nccl-only N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2t_n${N}_lshard_nccl.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms')}, d['roofline']['ms_per_sweep'])"
( time timeout 150 $TR --nproc-per-node 8 --master-port 29591 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r2t_n8_shifts.json 2> gpurun_out/r2t_n8_shifts.err
echo "shifts N=8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2t_n8_shifts.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','protocol_fallbacks')}, d['e2e']['value'], d['lshard'])"
