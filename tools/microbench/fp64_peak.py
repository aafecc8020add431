"""FP64 throughput record for the factorisation's roofline denominator (MEASURED_PEAKS.json has no
FP64 entry).  Two comparators on the GPU box: cuBLAS DGEMM / ZGEMM through torch.matmul (what a
contraction-bound FP64 kernel could reach, tensor-core DMMA included if cuBLAS picks it) and the
dependent / independent DFMA issue microbenchmark (tools/microbench/fp64_latency.cu, the
SIMT-pipe peak the strip kernel's DFMAs are bound by).  Usage: python tools/microbench/fp64_peak.py"""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))


def gemm_tflops(dtype, n, flops_per_mac, reps=10):
    a = torch.randn(n, n, dtype=dtype, device="cuda")
    b = torch.randn(n, n, dtype=dtype, device="cuda")
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return flops_per_mac * n ** 3 / (best * 1e-3) / 1e12, best


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    for n in (4096, 8192):
        tf, ms = gemm_tflops(torch.float64, n, 2)
        out["dgemm_%d_tflops" % n] = tf
        tf, ms = gemm_tflops(torch.complex128, n, 8)
        out["zgemm_%d_tflops" % n] = tf
    src = os.path.join(ROOT, "tools", "microbench", "fp64_latency.cu")
    exe = "/tmp/fp64_latency"
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, src],
                       capture_output=True, text=True)
    if r.returncode == 0:
        r = subprocess.run([exe], capture_output=True, text=True)
        out["dfma_microbench"] = r.stdout.strip().splitlines()
    else:
        out["dfma_microbench"] = "nvcc failed: " + r.stderr[-500:]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.exit(main())
