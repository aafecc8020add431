mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
( time KB_SHARD_TIMING=1 timeout 200 $TR --nproc-per-node $N --master-port 2955$N bench.py --gpus $N --steps 3 --warmup 2 --mode lshard --e2e-steps 1 ) > gpurun_out/r2i_n${N}_lshard.json 2> gpurun_out/r2i_n${N}_lshard.err
echo "lshard N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2i_n${N}_lshard.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','op_applies_per_step','max_residual')}, d['roofline']['ms_per_sweep'])"; grep "shard timing rank [01]:" gpurun_out/r2i_n${N}_lshard.err | tail -6
done
