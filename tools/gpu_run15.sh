mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 150 $TR --nproc-per-node 4 --master-port 29561 tools/run_sharded_check.py spinover magnetic_small dormy ) > gpurun_out/r2r_shard4_check.log 2>&1
echo "check4 rc=$?"; grep -E "rank 0|rror:|WARN" gpurun_out/r2r_shard4_check.log | tail -5 | cut -c1-230
( time timeout 150 $TR --nproc-per-node 3 --master-port 29562 tools/run_sharded_check.py spinover dormy ) > gpurun_out/r2r_shard3_check.log 2>&1
echo "check3 rc=$?"; grep -E "rank 1|rror:|WARN" gpurun_out/r2r_shard3_check.log | tail -4 | cut -c1-230
( time KB_SHARD_GENERAL=1 timeout 150 $TR --nproc-per-node 4 --master-port 29563 tools/run_sharded_check.py spinover dormy ) > gpurun_out/r2r_shard4_general.log 2>&1
echo "general4 rc=$?"; grep -E "rank 0|rror:|WARN" gpurun_out/r2r_shard4_general.log | tail -4 | cut -c1-230
for N in 4 2; do
( time KB_SHARD_TIMING=1 timeout 150 $TR --nproc-per-node $N --master-port 2957$N bench.py --gpus $N --steps 5 --warmup 3 --mode lshard --e2e-steps 2 ) > gpurun_out/r2r_n${N}_lshard.json 2> gpurun_out/r2r_n${N}_lshard.err
echo "lshard N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2r_n${N}_lshard.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','op_applies_per_step','max_residual')}, d['roofline']['ms_per_sweep'], d['e2e']['value'])"; grep "shard timing rank 0:" gpurun_out/r2r_n${N}_lshard.err | tail -3 | cut -c1-250; grep "rror:" gpurun_out/r2r_n${N}_lshard.err | head -3
done
