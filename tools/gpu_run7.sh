mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 $TR --nproc-per-node 2 --master-port 29541 tools/run_sharded_check.py spinover magnetic_small dormy ) > gpurun_out/r2g_shard_fast.log 2>&1
echo "fast rc=$?"; grep -E "rank|rror|WARN" gpurun_out/r2g_shard_fast.log | tail -12
( time KB_SHARD_GENERAL=1 timeout 300 $TR --nproc-per-node 2 --master-port 29542 tools/run_sharded_check.py spinover magnetic_small dormy ) > gpurun_out/r2g_shard_general.log 2>&1
echo "general rc=$?"; grep -E "rank|rror|WARN" gpurun_out/r2g_shard_general.log | tail -8
( time timeout 300 $TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 2 --mode lshard --e2e-steps 1 ) > gpurun_out/r2g_n2_lshard.json 2> gpurun_out/r2g_n2_lshard.err
echo "lshard rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2g_n2_lshard.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','op_applies_per_step','max_residual')}, d['roofline']['ms_per_sweep'])"; tail -5 gpurun_out/r2g_n2_lshard.err
