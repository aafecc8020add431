// Microbenchmark: cross-SM signalling latency through L2 on sm_100a (one-way "hop" in SM cycles).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hop_latency tools/microbench/hop_latency.cu
//  A) ping-pong between two CTAs on a 16-byte datum that is its own flag (plain / strong store, relaxed poll)
//  B) ping-pong with st.release / ld.acquire on a 4-byte flag
//  C) ping-pong where the sender first writes NB bytes of payload, then release-stores the flag, and the
//     receiver reads the payload after the acquire (flag-then-data)
//  D) all-to-all: G CTAs, every CTA writes 16-byte entries for every other CTA and polls the ones meant for it
//  E) plain load latency of a line last written by another SM (L2 hit), and membar.gl cost
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 ld_relaxed(const double2* p) {
  double2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(double2* p, double2 v) {
  asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_plain(double2* p, double2 v) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// mode 0: plain store of the datum, 1: strong store
__global__ void pingpong_datum(double2* a, double2* b, int iters, int mode, long long* cyc) {
  if (threadIdx.x != 0) return;
  const int me = blockIdx.x;
  if (me > 1) return;
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    double2 v = make_double2((double)i, (double)i);
    if (me == 0) {
      if (mode) st_relaxed(a, v); else st_plain(a, v);
      while (ld_relaxed(b).x != (double)i) {}
    } else {
      while (ld_relaxed(a).x != (double)i) {}
      if (mode) st_relaxed(b, v); else st_plain(b, v);
    }
  }
  if (me == 0) cyc[0] = clock64() - t0;
}

__global__ void pingpong_flag(unsigned* a, unsigned* b, int iters, int mode, long long* cyc) {
  if (threadIdx.x != 0) return;
  const int me = blockIdx.x;
  if (me > 1) return;
  long long t0 = clock64();
  for (unsigned i = 1; i <= (unsigned)iters; ++i) {
    if (me == 0) {
      if (mode == 0) st_release(a, i); else st_relaxed_u(a, i);
      if (mode == 0) { while (ld_acquire(b) != i) {} } else { while (ld_relaxed_u(b) != i) {} }
    } else {
      if (mode == 0) { while (ld_acquire(a) != i) {} } else { while (ld_relaxed_u(a) != i) {} }
      if (mode == 0) st_release(b, i); else st_relaxed_u(b, i);
    }
  }
  if (me == 0) cyc[0] = clock64() - t0;
}

// flag-then-data with a whole CTA: 256 threads write nb bytes, barrier, thread 0 releases; the receiver's
// thread 0 acquires, barrier, all threads read the payload (ld.cg) and reduce
__global__ void pingpong_payload(double2* pa, double2* pb, unsigned* fa, unsigned* fb, int n16, int iters,
                                 long long* cyc, double* sink) {
  const int me = blockIdx.x;
  if (me > 1) return;
  double acc = 0;
  long long t0 = clock64();
  for (unsigned i = 1; i <= (unsigned)iters; ++i) {
    double2* outp = me == 0 ? pa : pb;
    const double2* inp = me == 0 ? pb : pa;
    unsigned* fo = me == 0 ? fa : fb;
    unsigned* fi = me == 0 ? fb : fa;
    if (me == 1) {
      if (threadIdx.x == 0) while (ld_acquire(fi) != i) {}
      __syncthreads();
      for (int k = threadIdx.x; k < n16; k += blockDim.x) acc += __ldcg(&inp[k]).x;
    }
    for (int k = threadIdx.x; k < n16; k += blockDim.x) outp[k] = make_double2((double)i, acc);
    __syncthreads();
    if (threadIdx.x == 0) st_release(fo, i);
    if (me == 0) {
      if (threadIdx.x == 0) while (ld_acquire(fi) != i) {}
      __syncthreads();
      for (int k = threadIdx.x; k < n16; k += blockDim.x) acc += __ldcg(&inp[k]).x;
    }
  }
  if (threadIdx.x == 0 && me == 0) cyc[0] = clock64() - t0;
  if (acc == 12345.678) sink[0] = acc;
}

// all-to-all with G CTAs: step i: CTA c writes buf[i&1][c][0..G) (entry d is "for" CTA d), then polls
// buf[i&1][c'][me] for all c' (datum-as-flag, value = i).  One warp per CTA polls (lane <-> producers).
__global__ void all_to_all(double2* buf, int G, int iters, int strong, long long* cyc) {
  const int me = blockIdx.x;
  const int lane = threadIdx.x;
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    double2* slot = buf + (size_t)(i & 1) * G * G;
    for (int d = lane; d < G; d += 32) {
      double2 v = make_double2((double)i, 0.0);
      if (strong) st_relaxed(&slot[(size_t)me * G + d], v); else slot[(size_t)me * G + d] = v;
    }
    for (int c = lane; c < G; c += 32)
      while (ld_relaxed(&slot[(size_t)c * G + me]).x != (double)i) {}
    __syncwarp();
  }
  if (lane == 0) cyc[me] = clock64() - t0;
}

// counter barrier among G CTAs (red.release + ld.acquire poll), one thread per CTA
__global__ void counter_barrier(unsigned* ctr, int G, int iters, long long* cyc) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  for (unsigned i = 1; i <= (unsigned)iters; ++i) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    while (ld_acquire(ctr) < i * (unsigned)G) {}
  }
  cyc[blockIdx.x] = clock64() - t0;
}
__global__ void counter_barrier_relaxed(unsigned* ctr, int G, int iters, long long* cyc) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  for (unsigned i = 1; i <= (unsigned)iters; ++i) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    while (ld_relaxed_u(ctr) < i * (unsigned)G) {}
  }
  cyc[blockIdx.x] = clock64() - t0;
}

__global__ void fence_cost(double2* buf, int iters, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    buf[(size_t)blockIdx.x * 1024 + threadIdx.x] = make_double2(i, i);
    __threadfence();
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
__global__ void load_latency(const double2* buf, int iters, long long* cyc, double* sink) {
  // dependent chain of L2 loads (ld.cg) over lines written by the previous kernel
  size_t idx = threadIdx.x;
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double2 v = __ldcg(&buf[idx]);
    acc += v.x;
    idx = (idx + 64 + (size_t)v.y) & 65535;
  }
  if (threadIdx.x == 0) cyc[0] = clock64() - t0;
  if (acc == 1.2345) sink[0] = acc;
}

int main() {
  const int iters = 2000;
  double2 *a, *b, *buf;
  unsigned *fa, *fb;
  long long* cyc;
  double* sink;
  cudaMalloc(&a, 1 << 20);
  cudaMalloc(&b, 1 << 20);
  cudaMalloc(&buf, 64 << 20);
  cudaMalloc(&fa, 4096);
  cudaMalloc(&fb, 4096);
  cudaMalloc(&cyc, 8 * 1024);
  cudaMalloc(&sink, 64);
  long long h[256];
  auto report = [&](const char* name, int n = 1, double div = 4000.0) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, cyc, n * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < n; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-64s %8.0f cycles\n", name, (double)mx / div);
  };
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(a, 0, 64); cudaMemset(b, 0, 64);
    pingpong_datum<<<2, 32>>>(a, b, iters, mode, cyc);
    report(mode ? "A  datum-as-flag, strong store, relaxed poll: one-way hop" : "A  datum-as-flag, plain store, relaxed poll: one-way hop");
  }
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(fa, 0, 64); cudaMemset(fb, 0, 64);
    pingpong_flag<<<2, 32>>>(fa, fb, iters, mode, cyc);
    report(mode ? "B  4-byte flag, relaxed store / relaxed poll: one-way hop" : "B  4-byte flag, st.release / ld.acquire: one-way hop");
  }
  for (int n16 : {0, 64, 600, 4096}) {
    cudaMemset(fa, 0, 64); cudaMemset(fb, 0, 64);
    pingpong_payload<<<2, 256>>>(a, b, fa, fb, n16, iters, cyc, sink);
    char nm[128];
    snprintf(nm, sizeof nm, "C  payload %6d B + release flag + acquire + read: one-way", n16 * 16);
    report(nm);
  }
  for (int G : {2, 8, 74, 148}) {
    for (int strong = 0; strong < 2; ++strong) {
      cudaMemset(buf, 0, (size_t)2 * G * G * 16);
      all_to_all<<<G, 32>>>(buf, G, iters, strong, cyc);
      char nm[128];
      snprintf(nm, sizeof nm, "D  all-to-all G=%3d, 16 B per pair, %s store: per step", G, strong ? "strong" : "plain");
      report(nm, G, (double)iters);
    }
  }
  for (int G : {2, 74, 148}) {
    cudaMemset(fa, 0, 64);
    counter_barrier<<<G, 32>>>(fa, G, iters, cyc);
    char nm[128];
    snprintf(nm, sizeof nm, "D' counter barrier G=%3d (red.release + ld.acquire): per round", G);
    report(nm, G, (double)iters);
    cudaMemset(fa, 0, 64);
    counter_barrier_relaxed<<<G, 32>>>(fa, G, iters, cyc);
    snprintf(nm, sizeof nm, "D' counter barrier G=%3d (relaxed): per round", G);
    report(nm, G, (double)iters);
  }
  fence_cost<<<1, 32>>>(buf, iters, cyc);
  report("E  store + __threadfence, 1 warp: per iteration", 1, (double)iters);
  fence_cost<<<148, 256>>>(buf, iters, cyc);
  report("E  store + __threadfence, 148 x 256 threads: per iteration", 148, (double)iters);
  cudaMemset(buf, 0, 64 << 20);
  load_latency<<<1, 32>>>(buf, iters, cyc, sink);
  report("E  dependent ld.cg chain (L2 hit): per load", 1, (double)iters);
  return 0;
}
