"""GPU, >= 2 devices: the l-sharded path (NCCL) against the committed oracle values.
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", ["fast", "general"])
def test_two_rank_lsharded_parity(path):
    """fast: the rank's interior on the strip factorisation + folded sweep (the default);
    general: the per-node kernels with spike recurrences (KB_SHARD_GENERAL=1, wide nodes)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533" if path == "fast" else "29534",
           os.path.join(ROOT, "tools", "run_sharded_check.py"), "spinover", "magnetic_small", "dormy"]
    env = dict(os.environ, KB_EXPECT_SHARD_PATH=path)
    if path == "general":
        env["KB_SHARD_GENERAL"] = "1"
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    print(r.stdout[-3000:])
    assert r.returncode == 0
    assert r.stdout.count(" OK") >= 6 and "FAIL" not in r.stdout
