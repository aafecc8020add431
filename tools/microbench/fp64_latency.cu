// Microbenchmark: dependent-chain latency and issue rate of FP64 DFMA / DADD / SHFL on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/microbench/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x + k;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void shfl_chain(double* out, long long* cyc, int iters) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) v += __shfl_xor_sync(0xffffffffu, v, 1 + (i & 15));
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 1024);
  long long h[4];
  const int iters = 4096;
  auto run = [&](auto kernel, int threads, const char* name, int ilp) {
    kernel<<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    kernel<<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s threads %4d: %.2f cycles per DFMA-step (per warp-instr: %.2f)\n", name, threads,
           (double)h[0] / iters, (double)h[0] / iters / ilp);
  };
  run(dfma_chain<1>, 32, "DFMA dependent chain ILP=1", 1);
  run(dfma_chain<2>, 32, "DFMA ILP=2", 2);
  run(dfma_chain<4>, 32, "DFMA ILP=4", 4);
  run(dfma_chain<8>, 32, "DFMA ILP=8", 8);
  run(dfma_chain<8>, 128, "DFMA ILP=8 4 warps", 8);
  run(dfma_chain<8>, 256, "DFMA ILP=8 8 warps", 8);
  run(dfma_chain<8>, 640, "DFMA ILP=8 20 warps", 8);
  shfl_chain<<<1, 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("SHFL+DADD dependent chain: %.2f cycles per step\n", (double)h[0] / iters);
  return 0;
}
