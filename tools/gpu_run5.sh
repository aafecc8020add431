mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 400 $TR --nproc-per-node 1 --master-port 29511 bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2e_n1_torchrun.json 2> gpurun_out/r2e_n1_torchrun.err
echo "n1 rc=$?"; tail -c 600 gpurun_out/r2e_n1_torchrun.json; tail -4 gpurun_out/r2e_n1_torchrun.err
( time timeout 400 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2e_n2_torchrun.json 2> gpurun_out/r2e_n2_torchrun.err
echo "n2 rc=$?"; tail -c 600 gpurun_out/r2e_n2_torchrun.json; tail -4 gpurun_out/r2e_n2_torchrun.err
( time timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/r2e_pytest_sharded.log 2>&1
tail -5 gpurun_out/r2e_pytest_sharded.log
( time timeout 300 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --mode lshard --e2e-steps 1 ) > gpurun_out/r2e_n2_lshard.json 2> gpurun_out/r2e_n2_lshard.err
echo "lshard rc=$?"; tail -c 600 gpurun_out/r2e_n2_lshard.json; tail -4 gpurun_out/r2e_n2_lshard.err
CUDA_VISIBLE_DEVICES=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2e_gpu1.json 2> gpurun_out/r2e_gpu1.err
echo "gpu1 rc=$?"; tail -c 300 gpurun_out/r2e_gpu1.json
