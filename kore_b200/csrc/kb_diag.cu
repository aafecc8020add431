// Per-degree energy / dissipation / power integrals of a solution.
//
// Replaces the multiprocessing pool of /root/reference/bin/utils4pp.py:794-858 (`diagnose`: one
// Python worker per spherical-harmonic degree, utils4pp.py:426-488 flow_worker, :536-561
// thermal_worker) behind bin/spin_doctor.py:119-147, for hydrodynamic and Boussinesq thermal
// solutions: the kinetic energy, the kinetic and internal viscous dissipation and the buoyancy power
// of every degree of the flow, the thermal energy, dissipation and advection of every poloidal degree
// -- the terms of the power balance the reference checks its solutions with
// (spin_doctor.py:227-242).
//
// One CTA per (degree, solution).  The Chebyshev coefficients of the degree's radial function
// (poloidal P, toroidal T, and the temperature of the same degree when it enters) sit in shared
// memory; every thread owns Chebyshev-Gauss nodes and evaluates the series and its first three
// radial derivatives there with the forward three-term recurrences of T_j, T_j', T_j'', T_j'''
// (the reference differentiates in coefficient space and evaluates by Clenshaw; both are exact in
// exact arithmetic, tests hold the difference below 1e-9 of the largest degree), forms the
// integrands of utils4pp.py:221-296, 363-412 and the CTA sums them with the quadrature weights in a
// fixed order (deterministic).  FP64-pipe bound: 7.6e10 flops of recurrences and complex
// accumulations at N = lmax = 600 for eleven solutions in 3.19 ms = 24 TFLOP/s, 0.64 of the measured
// 37 TFLOP/s DFMA peak (profiles/r2B_ncu_launches_assembly_diagnostics.csv); the reference's pool
// takes minutes.
#include "kb_internal.cuh"

namespace {

struct KdParams {
  int N, N1, nb, nll, lfirst, pol_first, thermal, heating, full_sphere, parP, parT;
  int64_t n, sizmat;
  double scale;        // d/dr = scale d/dx
  const double* nodes; // [3][N]: x in the solution's Chebyshev domain, r, quadrature weight
  const double2* x;    // [nsol][sizmat]
  double* flow;        // [nsol][nll][6]
  double* therm;       // [nsol][nb][3]
};

struct KdSeries {
  double2 f[4];
};

// f, f', f'', f''' at x of sum_j c_j T_j (coefficient j lives at c[(j - par) / 2] when only one
// parity is stored, `step` = 2, else at c[j]); derivatives with respect to r
__device__ __forceinline__ KdSeries kd_series(const double2* c, int ncoef, int step, int par, int N, double x,
                                              double scale) {
  double t[4][2];  // [derivative][previous, current]
  t[0][0] = 1.0; t[0][1] = x;
  t[1][0] = 0.0; t[1][1] = 1.0;
  t[2][0] = t[2][1] = t[3][0] = t[3][1] = 0.0;
  KdSeries s;
#pragma unroll
  for (int p = 0; p < 4; ++p) s.f[p] = zmake(0.0, 0.0);
  auto coef = [&](int j, double2& cj) -> bool {
    if (step == 1) {
      if (j >= ncoef) return false;
      cj = c[j];
      return true;
    }
    if ((j & 1) != par) return false;
    const int jj = j >> 1;
    if (jj >= ncoef) return false;
    cj = c[jj];
    return true;
  };
  double2 cj;
  if (coef(0, cj)) s.f[0] = cj;
  if (N > 1 && coef(1, cj)) {
    s.f[0].x += cj.x * x;
    s.f[0].y += cj.y * x;
    s.f[1] = cj;
  }
  for (int j = 1; j < N - 1; ++j) {
    double nx[4];
    nx[3] = 2.0 * x * t[3][1] - t[3][0] + 6.0 * t[2][1];
    nx[2] = 2.0 * x * t[2][1] - t[2][0] + 4.0 * t[1][1];
    nx[1] = 2.0 * x * t[1][1] - t[1][0] + 2.0 * t[0][1];
    nx[0] = 2.0 * x * t[0][1] - t[0][0];
    if (coef(j + 1, cj)) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        s.f[p].x += cj.x * nx[p];
        s.f[p].y += cj.y * nx[p];
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      t[p][0] = t[p][1];
      t[p][1] = nx[p];
    }
  }
  double sc = scale;
#pragma unroll
  for (int p = 1; p < 4; ++p) {
    s.f[p].x *= sc;
    s.f[p].y *= sc;
    sc *= scale;
  }
  return s;
}

__device__ __forceinline__ double kd_abs2(double2 a) { return a.x * a.x + a.y * a.y; }
// Re(conj(a) b)
__device__ __forceinline__ double kd_redot(double2 a, double2 b) { return a.x * b.x + a.y * b.y; }

#define KD_THREADS 256
#define KD_NQ 7  // kinetic energy, kinetic dissipation, internal dissipation, buoyancy power, thermal energy, dissipation, advection

__global__ void __launch_bounds__(KD_THREADS) kd_degree(KdParams q) {
  extern __shared__ __align__(16) unsigned char kd_smem[];
  double2* cA = (double2*)kd_smem;  // P or T of this degree
  double2* cH = cA + q.N1;          // temperature of this degree (poloidal degrees of thermal runs)
  __shared__ double red[KD_NQ][KD_THREADS];
  const int a = blockIdx.x, sol = blockIdx.y;
  const int l = q.lfirst + a;
  const bool pol = ((a & 1) == q.pol_first);
  const int idx = a >> 1;  // block of the degree inside its section
  const double2* xs = q.x + (size_t)sol * q.sizmat;
  const double2* srcA = xs + (pol ? 0 : q.n) + (size_t)idx * q.N1;
  const bool withH = pol && q.thermal;
  const double2* srcH = xs + 2 * q.n + (size_t)idx * q.N1;
  for (int j = threadIdx.x; j < q.N1; j += KD_THREADS) {
    cA[j] = srcA[j];
    if (withH) cH[j] = srcH[j];
  }
  __syncthreads();
  const int step = q.full_sphere ? 2 : 1;
  const double L = (double)l * (double)(l + 1);
  const double f0 = 4.0 * 3.14159265358979323846 / (2.0 * l + 1.0);
  double acc[KD_NQ];
#pragma unroll
  for (int k = 0; k < KD_NQ; ++k) acc[k] = 0.0;
  for (int k = threadIdx.x; k < q.N; k += KD_THREADS) {
    const double x = q.nodes[k], r = q.nodes[q.N + k], w = q.nodes[2 * q.N + k];
    const double r2 = r * r;
    if (pol) {
      const KdSeries P = kd_series(cA, q.N1, step, q.parP, q.N, x, q.scale);
      // radial (q) and consoidal (s) components and their radial derivatives (utils4pp.py:163-201)
      double2 q0, s0, q1, s1, q2, s2;
      q0 = zmake(L * P.f[0].x / r, L * P.f[0].y / r);
      s0 = zmake(P.f[1].x + P.f[0].x / r, P.f[1].y + P.f[0].y / r);
      q1 = zmake((L * P.f[1].x - q0.x) / r, (L * P.f[1].y - q0.y) / r);
      s1 = zmake(P.f[2].x + q1.x / L, P.f[2].y + q1.y / L);
      q2 = zmake((L * P.f[2].x - 2.0 * q1.x) / r, (L * P.f[2].y - 2.0 * q1.y) / r);
      s2 = zmake(P.f[3].x + q2.x / L, P.f[3].y + q2.y / L);
      acc[0] += w * f0 * (r2 * kd_abs2(q0) + r2 * L * kd_abs2(s0));
      const double dk = L * r2 * kd_redot(s0, s2) + 2.0 * r * L * kd_redot(s0, s1) - L * L * kd_abs2(s0) -
                        ((double)l * l + l + 2.0) * kd_abs2(q0) + 2.0 * r * kd_redot(q0, q1) + r2 * kd_redot(q0, q2) +
                        4.0 * L * kd_redot(q0, s0);
      acc[1] += w * 2.0 * f0 * dk;
      const double2 e = zmake(q0.x + r * s1.x - s0.x, q0.y + r * s1.y - s0.y);
      acc[2] += w * 2.0 * f0 * (L * kd_abs2(e) + 3.0 * r2 * kd_abs2(q1) + L * (l - 1.0) * (l + 2.0) * kd_abs2(s0));
      if (withH) {
        const KdSeries H = kd_series(cH, q.N1, step, q.parP, q.N, x, q.scale);
        const double ph = 2.0 * kd_redot(P.f[0], H.f[0]);
        acc[3] += w * f0 * r2 * L * ph;
        acc[4] += w * f0 * r2 * kd_abs2(H.f[0]);
        acc[5] += w * f0 * (4.0 * r * kd_redot(H.f[1], H.f[0]) + 2.0 * r2 * kd_redot(H.f[2], H.f[0]) - 2.0 * L * kd_abs2(H.f[0]));
        const double fr = q.heating == 0 ? 1.0 / r : r2;
        acc[6] += w * f0 * fr * L * ph;
      }
    } else {
      const KdSeries T = kd_series(cA, q.N1, step, q.parT, q.N, x, q.scale);
      acc[0] += w * f0 * r2 * L * kd_abs2(T.f[0]);
      acc[1] += w * 2.0 * f0 * (L * r2 * kd_redot(T.f[0], T.f[2]) + 2.0 * r * L * kd_redot(T.f[0], T.f[1]) - L * L * kd_abs2(T.f[0]));
      const double2 e = zmake(r * T.f[1].x - T.f[0].x, r * T.f[1].y - T.f[0].y);
      acc[2] += w * 2.0 * f0 * (L * kd_abs2(e) + L * (l - 1.0) * (l + 2.0) * kd_abs2(T.f[0]));
    }
  }
#pragma unroll
  for (int k = 0; k < KD_NQ; ++k) red[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int s = KD_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < KD_NQ; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* f = q.flow + ((size_t)sol * q.nll + a) * 6;
    f[0] = red[0][0];
    f[1] = red[1][0];
    f[2] = red[2][0];
    f[3] = 0.0;  // Lorentz power: magnetic runs are not covered
    f[4] = withH ? red[3][0] : 0.0;
    f[5] = 0.0;  // compositional buoyancy: the host runs this kernel a second time on [u | v | c] (diagnostics.py)
    if (withH) {
      double* t = q.therm + ((size_t)sol * q.nb + idx) * 3;
      t[0] = red[4][0];
      t[1] = red[5][0];
      t[2] = red[6][0];
    }
  }
}

}  // namespace

extern "C" int kb_diagnose(kb_handle h, const kb_diag_params* p, const double* nodes, const double* x, int nsol,
                           double* flow, double* thermal) {
  if (!h) return KB_EINVAL;
  if (!p || !nodes || !x || !flow || nsol < 1) return kb_fail(h, KB_EINVAL, "kb_diagnose: null argument");
  if (p->N < 2 || p->N1 < 1 || p->nb < 1 || p->N1 > p->N || p->m < 0)
    return kb_fail(h, KB_EINVAL, "kb_diagnose: bad sizes");
  if (p->thermal && !thermal) return kb_fail(h, KB_EINVAL, "kb_diagnose: thermal output missing");
  if (2 * p->nb != p->lmax - p->m + 1) return kb_fail(h, KB_EINVAL, "kb_diagnose: nb does not match lmax and m");
  if (p->heating != 0 && p->heating != 1) return kb_fail(h, KB_EINVAL, "kb_diagnose: heating must be 0 or 1");
  if (p->symm != 1 && p->symm != -1) return kb_fail(h, KB_EINVAL, "kb_diagnose: symm must be +1 or -1");
  if (p->ricb < 0 || p->ricb >= p->rcmb) return kb_fail(h, KB_EINVAL, "kb_diagnose: bad radii");
  const bool full = p->ricb == 0.0;
  if (full ? (p->N1 != p->N / 2) : (p->N1 != p->N)) return kb_fail(h, KB_EINVAL, "kb_diagnose: N1 does not match N and ricb");
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  KdParams q;
  q.N = p->N;
  q.N1 = p->N1;
  q.nb = p->nb;
  q.nll = 2 * p->nb;
  q.lfirst = p->m == 0 ? 1 : p->m;
  // the first family of degrees (section u: poloidal) starts with the first degree exactly when
  // m > 0 and symm == 1 disagree (utils.py:174-183)
  q.pol_first = ((p->m == 0 ? 0 : 1) + (p->symm > 0 ? 1 : 0)) % 2;
  q.thermal = p->thermal != 0;
  q.heating = p->heating;
  q.full_sphere = full;
  const int sy = (p->symm + 1) / 2;
  q.parP = (p->m + 1 - sy) % 2;
  q.parT = (p->m + sy) % 2;
  q.n = (int64_t)p->N1 * p->nb;
  q.sizmat = q.n * (2 + (p->thermal ? 1 : 0));
  q.scale = 2.0 / (p->rcmb - (full ? -p->rcmb : p->ricb));
  DevBuf<double> d_nodes, d_flow, d_therm;
  DevBuf<double2> d_x;
  KB_CUDA(h, d_nodes.alloc(3 * (size_t)q.N));
  KB_CUDA(h, d_x.alloc((size_t)nsol * q.sizmat));
  KB_CUDA(h, d_flow.alloc((size_t)nsol * q.nll * 6));
  KB_CUDA(h, d_therm.alloc((size_t)nsol * q.nb * 3));
  KB_CUDA(h, cudaMemcpyAsync(d_nodes.p, nodes, 3 * (size_t)q.N * sizeof(double), cudaMemcpyHostToDevice, s));
  KB_CUDA(h, cudaMemcpyAsync(d_x.p, x, (size_t)nsol * q.sizmat * sizeof(double2), cudaMemcpyHostToDevice, s));
  KB_CUDA(h, cudaMemsetAsync(d_therm.p, 0, (size_t)nsol * q.nb * 3 * sizeof(double), s));
  q.nodes = d_nodes.p;
  q.x = d_x.p;
  q.flow = d_flow.p;
  q.therm = d_therm.p;
  const size_t smem = 2 * (size_t)q.N1 * sizeof(double2);
  if (smem > 48 * 1024)
    KB_CUDA(h, cudaFuncSetAttribute((const void*)kd_degree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kd_degree<<<dim3((unsigned)q.nll, (unsigned)nsol), KD_THREADS, smem, s>>>(q);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  KB_CUDA(h, cudaMemcpyAsync(flow, d_flow.p, (size_t)nsol * q.nll * 6 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (p->thermal)
    KB_CUDA(h, cudaMemcpyAsync(thermal, d_therm.p, (size_t)nsol * q.nb * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaStreamSynchronize(s));
  return KB_OK;
}
