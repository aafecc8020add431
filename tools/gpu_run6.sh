mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 $TR --nproc-per-node 1 --master-port 29511 bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2f_n1_torchrun.json 2> gpurun_out/r2f_n1_torchrun.err
echo "n1 rc=$?"; tail -c 400 gpurun_out/r2f_n1_torchrun.json; tail -6 gpurun_out/r2f_n1_torchrun.err
( time timeout 300 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r2f_n8_torchrun.json 2> gpurun_out/r2f_n8_torchrun.err
echo "n8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2f_n8_torchrun.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','protocol_fallbacks','n_gpus')}, d['e2e']['value'])"; tail -6 gpurun_out/r2f_n8_torchrun.err
( time timeout 200 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r2f_n4_torchrun.json 2> gpurun_out/r2f_n4_torchrun.err
echo "n4 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2f_n4_torchrun.json')); print({k:d[k] for k in ('value','ms_per_step','factor_ms','protocol_fallbacks','n_gpus')}, d['e2e']['value'])"; tail -3 gpurun_out/r2f_n4_torchrun.err
