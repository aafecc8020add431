// FP64 tensor-core (mma.sync m8n8k4 f64) issue rate on sm_100a against the DFMA rate: is a
// contraction-bound complex128 product better off on DMMA than on the SIMT FP64 pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_rate tools/microbench/dmma_rate.cu && /tmp/dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = 0.0;
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(a, b, c[i]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 4 * 256);
  const int iters = 20000;
  for (int cpsm = 1; cpsm <= 4; cpsm *= 2) {
    const int grid = sms * cpsm;
    double ms = time_ms([&] { k_dmma<16><<<grid, 256>>>(out, iters, 1.0, 2.0); });
    // one m8n8k4 = 8*8*4 FMAs = 512 flops per warp instruction
    double fl = (double)grid * 8 /*warps*/ * iters * 16.0 * 512.0;
    printf("DMMA m8n8k4, 16 accumulators, %d x 256 threads per SM: %.2f ms  %.1f TFLOP/s\n", cpsm, ms, fl / ms / 1e9);
    ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0, 2.0); });
    fl = (double)grid * 256 * iters * 16.0 * 2.0;
    printf("DFMA, 16 accumulators,        %d x 256 threads per SM: %.2f ms  %.1f TFLOP/s\n", cpsm, ms, fl / ms / 1e9);
  }
  return 0;
}
