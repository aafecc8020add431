// Krylov-Schur eigensolver on the shift-and-invert operator (A - sigma B)^{-1} B.
//
// Replaces SLEPc EPS(krylovschur) + ST(sinvert) + BV + DS, i.e. everything
// inside E.solve() (/root/reference/bin/solve.py:123) plus getConverged /
// getEigenpair (solve.py:131-149).  Semantics restated from SLEPc's documented
// defaults, since SLEPc itself is not vendored:
//   * GNHEP => standard inner product; ncv = max(2 nev, nev+15); restart keeps
//     half of the non-converged basis; converged Schur vectors are locked;
//   * Ritz values theta are compared through lambda = sigma + 1/theta by `which`;
//   * convergence: estimated residual beta |s_m^T z_i| <= tol |theta_i|
//     (EPS_CONV_REL) or, with true_residual, ||A x - lambda B x|| <= tol |lambda| ||x||;
//   * eigenvectors are purified (x <- OP x, EPSSetPurify) and have unit 2-norm.
// Basis vectors live on the device in chain order; Gram-Schmidt (classical, two
// passes), the restart update V <- V Q and all SpMVs are HBM-bound kernels;
// the ncv x ncv projected problem is solved on the host (kb_host_dense.hpp).
#include <math.h>

#include <random>

#include "kb_host_dense.hpp"
#include "kb_internal.cuh"

// rows per CTA of the multi-dot kernel: n / 512 = 700 CTAs at the E = 1e-8 size (a chunk of 2048 left
// one CTA per SM and the kernel at a third of the HBM rate: 47 us for 104 MB)
#define KB_DOT_CHUNK 512
#define KB_DOT_THREADS 256

// partial[c * nchunks + blockIdx] = sum_{i in chunk} conj(V[i,c]) w[i]
__global__ void __launch_bounds__(KB_DOT_THREADS)
kb_multidot_partial(int n, int ncols, const double2* __restrict__ V, int64_t ldv,
                    const double2* __restrict__ w, double2* __restrict__ partial) {
  __shared__ double2 ws[KB_DOT_CHUNK];
  const int i0 = blockIdx.x * KB_DOT_CHUNK;
  const int len = min(KB_DOT_CHUNK, n - i0);
  for (int i = threadIdx.x; i < len; i += blockDim.x) ws[i] = w[i0 + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = wid; c < ncols; c += nw) {
    const double2* v = V + (size_t)c * ldv + i0;
    double2 acc = zmake(0.0, 0.0), acc1 = zmake(0.0, 0.0);
    int i = lane;
    for (; i + 96 < len; i += 128) {  // four loads in flight per lane
      const double2 v0 = v[i], v1 = v[i + 32], v2 = v[i + 64], v3 = v[i + 96];
      zfmac(acc, v0, ws[i]);
      zfmac(acc1, v1, ws[i + 32]);
      zfmac(acc, v2, ws[i + 64]);
      zfmac(acc1, v3, ws[i + 96]);
    }
    for (; i < len; i += 32) zfmac(acc, v[i], ws[i]);
    acc = zadd(acc, acc1);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
    }
    if (lane == 0) partial[(size_t)c * gridDim.x + blockIdx.x] = acc;
  }
}

// h[c] (+)= sum_b partial[c, b]    (one warp per column, fixed order => deterministic)
__global__ void kb_reduce_cols(int ncols, int nchunks, const double2* __restrict__ partial,
                               double2* __restrict__ h, double2* __restrict__ hsum, int accumulate) {
  int c = blockIdx.x;
  int lane = threadIdx.x;
  double2 acc = zmake(0.0, 0.0);
  for (int b = lane; b < nchunks; b += 32) acc = zadd(acc, partial[(size_t)c * nchunks + b]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
  }
  if (lane == 0) {
    h[c] = acc;
    hsum[c] = accumulate ? zadd(hsum[c], acc) : acc;
  }
}

// w[i] -= sum_c V[i,c] h[c];  normpart[blockIdx] = sum |w_new|^2 over the block
__global__ void __launch_bounds__(256)
kb_multiaxpy(int n, int ncols, const double2* __restrict__ V, int64_t ldv, const double2* __restrict__ h,
             double2* __restrict__ w, double* __restrict__ normpart) {
  __shared__ double2 hs[128];
  __shared__ double red[8];
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) hs[c] = h[c];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double nn = 0.0;
  if (i < n) {
    double2 acc = w[i];
    for (int c = 0; c < ncols; ++c) zfms(acc, V[(size_t)c * ldv + i], hs[c]);
    w[i] = acc;
    nn = zabs2(acc);
  }
  for (int s = 16; s > 0; s >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nn;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    normpart[blockIdx.x] = t;
  }
}

// beta = sqrt(sum normpart); out[0] = beta; dst = w / beta
__global__ void kb_norm_finish(int nparts, const double* __restrict__ normpart, double* __restrict__ beta) {
  int lane = threadIdx.x;
  double acc = 0.0;
  for (int b = lane; b < nparts; b += 32) acc += normpart[b];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) beta[0] = sqrt(acc);
}
__global__ void kb_scale_by_recip(int n, const double2* __restrict__ w, const double* __restrict__ beta,
                                  double2* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double b = beta[0];
  double s = b > 0.0 ? 1.0 / b : 0.0;
  dst[i] = zscale(w[i], s);
}

// out[i] = sum_c V[i,c] z[c]
__global__ void kb_lincomb(int n, int ncols, const double2* __restrict__ V, int64_t ldv,
                           const double2* __restrict__ z, double2* __restrict__ out) {
  __shared__ double2 zs[128];
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) zs[c] = z[c];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 acc = zmake(0.0, 0.0);
  for (int c = 0; c < ncols; ++c) zfma(acc, V[(size_t)c * ldv + i], zs[c]);
  out[i] = acc;
}

// In-place V[:, 0:q] <- V[:, 0:mq] Q  (Q mq x q column-major), rows in chunks of 64
#define KB_RS_ROWS 64
__global__ void __launch_bounds__(256)
kb_restart_update(int n, int mq, int q, double2* __restrict__ V, int64_t ldv, const double2* __restrict__ Q) {
  extern __shared__ double2 sm[];
  double2* vs = sm;                        // KB_RS_ROWS x mq (row-major, padded)
  double2* qs = sm + KB_RS_ROWS * (mq + 1);  // mq x q column-major
  const int i0 = blockIdx.x * KB_RS_ROWS;
  const int len = min(KB_RS_ROWS, n - i0);
  for (int e = threadIdx.x; e < mq * q; e += blockDim.x) qs[e] = Q[e];
  for (int e = threadIdx.x; e < mq * KB_RS_ROWS; e += blockDim.x) {
    int c = e / KB_RS_ROWS, i = e % KB_RS_ROWS;
    if (i < len) vs[i * (mq + 1) + c] = V[(size_t)c * ldv + i0 + i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < q * KB_RS_ROWS; e += blockDim.x) {
    int c = e / KB_RS_ROWS, i = e % KB_RS_ROWS;
    if (i >= len) continue;
    double2 acc = zmake(0.0, 0.0);
    for (int k = 0; k < mq; ++k) zfma(acc, vs[i * (mq + 1) + k], qs[(size_t)c * mq + k]);
    V[(size_t)c * ldv + i0 + i] = acc;
  }
}

// The same with Q read through the read-only cache (basis sizes whose Q does not fit shared memory
// next to the row block: ncv beyond ~90)
__global__ void __launch_bounds__(256)
kb_restart_update_gq(int n, int mq, int q, double2* __restrict__ V, int64_t ldv, const double2* __restrict__ Q) {
  extern __shared__ double2 sm[];
  double2* vs = sm;  // KB_RS_ROWS x mq (row-major, padded)
  const int i0 = blockIdx.x * KB_RS_ROWS;
  const int len = min(KB_RS_ROWS, n - i0);
  for (int e = threadIdx.x; e < mq * KB_RS_ROWS; e += blockDim.x) {
    int c = e / KB_RS_ROWS, i = e % KB_RS_ROWS;
    if (i < len) vs[i * (mq + 1) + c] = V[(size_t)c * ldv + i0 + i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < q * KB_RS_ROWS; e += blockDim.x) {
    int c = e / KB_RS_ROWS, i = e % KB_RS_ROWS;
    if (i >= len) continue;
    double2 acc = zmake(0.0, 0.0);
    for (int k = 0; k < mq; ++k) zfma(acc, vs[i * (mq + 1) + k], __ldg(&Q[(size_t)c * mq + k]));
    V[(size_t)c * ldv + i0 + i] = acc;
  }
}

// partial sums for residuals: out[3*b+0] = |ax - lam bx|^2, +1 = |bx|^2, +2 = |x|^2
__global__ void kb_resid_partial(int n, const double2* __restrict__ ax, const double2* __restrict__ bx,
                                 const double2* __restrict__ x, double2 lam, double* __restrict__ out) {
  __shared__ double sh[3][8];
  double a0 = 0, a1 = 0, a2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double2 r = ax[i];
    zfms(r, lam, bx[i]);
    a0 += zabs2(r);
    a1 += zabs2(bx[i]);
    a2 += zabs2(x[i]);
  }
  for (int s = 16; s > 0; s >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, s);
    a1 += __shfl_xor_sync(0xffffffffu, a1, s);
    a2 += __shfl_xor_sync(0xffffffffu, a2, s);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = a0;
    sh[1][threadIdx.x >> 5] = a1;
    sh[2][threadIdx.x >> 5] = a2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      t0 += sh[0][k];
      t1 += sh[1][k];
      t2 += sh[2][k];
    }
    out[3 * blockIdx.x + 0] = t0;
    out[3 * blockIdx.x + 1] = t1;
    out[3 * blockIdx.x + 2] = t2;
  }
}


namespace {

typedef kbd::Z Z;

double which_key(Z lam, int which, Z tau) {
  switch (which) {
    case KB_WHICH_LM: return -std::abs(lam);
    case KB_WHICH_SM: return std::abs(lam);
    case KB_WHICH_LR: return -lam.real();
    case KB_WHICH_SR: return lam.real();
    case KB_WHICH_LI: return -lam.imag();
    case KB_WHICH_SI: return lam.imag();
    case KB_WHICH_TM: return std::abs(lam - tau);
    case KB_WHICH_TR: return std::fabs((lam - tau).real());
    case KB_WHICH_TI: return std::fabs((lam - tau).imag());
  }
  return 0.0;
}

struct Krylov {
  kb_context* h;
  int n, ncv;
  int64_t ldv;
  int nchunks, nblocks;
  double2* V;
  double2 *w, *hdev, *hsum, *hpart, *Qdev;
  double* normpart;
  double* beta_dev;
  // pinned host mirrors
  double2* h_host = nullptr;  // (ncv+1) entries of hsum per column -> S column
  double* beta_host = nullptr;
  // l-sharded pencil: this rank keeps rows [r0, r0 + nl) of every basis vector (its segment of the
  // chain); nchunks / nblocks count the LOCAL rows; coefficients and norms are summed over the ranks
  bool sharded = false;
  int r0 = 0, nl = 0;
  double2* hscratch = nullptr;

  double2* col(int j) { return V + (size_t)j * ldv; }

  // beta_dev <- ||v|| over all ranks from the per-block sums of squares in normpart
  int finish_norm(int nparts) {
    if (sharded) return kbi_shard_norm(h, nparts, normpart, beta_dev);
    kb_norm_finish<<<1, 32, 0, h->stream>>>(nparts, normpart, beta_dev);
    return KB_OK;
  }

  // v <- v / ||v||  (local rows)
  int normalize(double2* v) {
    cudaStream_t s = h->stream;
    const int nb = std::min(nblocks, 512);
    kb_norm2_partial<<<nb, 256, 0, s>>>(nl, v + r0, normpart);
    KB_TRY(finish_norm(nb));
    kb_scale_by_recip<<<nblk(nl, 256), 256, 0, s>>>(nl, v + r0, beta_dev, v + r0);
    h->launches += 3;
    KB_LAUNCH_CHECK(h);
    return KB_OK;
  }

  // w <- OP v_j ; orthogonalise against v_0..v_j (CGS2); S[:j+2, j]; v_{j+1} = w/beta
  int arnoldi_step(int j, int refine) {
    cudaStream_t s = h->stream;
    KB_TRY(kbi_apply_op_chain(h, col(j), w, refine));
    for (int pass = 0; pass < 2; ++pass) {
      kb_multidot_partial<<<nchunks, KB_DOT_THREADS, 0, s>>>(nl, j + 1, V + r0, ldv, w + r0, hpart);
      if (!sharded) {
        kb_reduce_cols<<<j + 1, 32, 0, s>>>(j + 1, nchunks, hpart, hdev, hsum, pass);
      } else {
        KB_TRY(kbi_shard_reduce_cols(h, j + 1, nchunks, hpart, hdev, hsum, hscratch, pass));
      }
      kb_multiaxpy<<<nblocks, 256, 0, s>>>(nl, j + 1, V + r0, ldv, hdev, w + r0, normpart);
    }
    KB_TRY(finish_norm(nblocks));
    kb_scale_by_recip<<<nblk(nl, 256), 256, 0, s>>>(nl, w + r0, beta_dev, col(j + 1) + r0);
    h->launches += 8;
    KB_LAUNCH_CHECK(h);
    KB_CUDA(h, cudaMemcpyAsync(h_host + (size_t)j * (ncv + 1), hsum, (j + 1) * sizeof(double2),
                               cudaMemcpyDeviceToHost, s));
    KB_CUDA(h, cudaMemcpyAsync(beta_host + j, beta_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
    return KB_OK;
  }
};

// V[:, 0:q] <- V[:, 0:mq] Q on the device (Q mq x q column-major, already uploaded)
int restart_update(kb_context* h, int n, int mq, int q, double2* V, int64_t ldv, const double2* Qdev) {
  const size_t smem_full = (size_t)(KB_RS_ROWS * (mq + 1) + mq * q) * sizeof(double2);
  const size_t smem_rows = (size_t)(KB_RS_ROWS * (mq + 1)) * sizeof(double2);
  if (smem_full <= 200 * 1024) {
    KB_CUDA(h, cudaFuncSetAttribute(kb_restart_update, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kb_restart_update<<<nblk(n, KB_RS_ROWS), 256, smem_full, h->stream>>>(n, mq, q, V, ldv, Qdev);
  } else {
    KB_CUDA(h, cudaFuncSetAttribute(kb_restart_update_gq, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kb_restart_update_gq<<<nblk(n, KB_RS_ROWS), 256, smem_rows, h->stream>>>(n, mq, q, V, ldv, Qdev);
  }
  h->launches++;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

int normalize_dev(kb_context* h, int n, double2* v, double* normpart, double* beta_dev, int nblocks) {
  // v <- v / ||v||
  cudaStream_t s = h->stream;
  kb_norm2_partial<<<nblocks, 256, 0, s>>>(n, v, normpart);
  kb_norm_finish<<<1, 32, 0, s>>>(nblocks, normpart, beta_dev);
  kb_scale_by_recip<<<nblk(n, 256), 256, 0, s>>>(n, v, beta_dev, v);
  h->launches += 3;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

}  // namespace

// SLEPc's default basis size max(2 nev, nev + 15) stays within KB_MAX_NCV for nev <= 63
// (the coefficient vectors of kb_multiaxpy / kb_lincomb live in 128-entry shared arrays).
#define KB_MAX_NCV 127
// repeated purification of the extracted eigenvectors (see the extraction loop)
#define KB_PURIFY_TARGET 1e-11
#define KB_PURIFY_EXTRA 3
// KB_OPT_REFINE_EIGS = -1 (automatic): one refinement step inside every operator application when the
// first correction of the start vector's solve is larger than this, relative to the solution
#ifndef KB_REFINE_AUTO
#define KB_REFINE_AUTO 1e-10
#endif

static int eigs_impl(kb_handle h, int nev, int ncv, double tol, int maxit, int which,
                     const double* target, int true_residual, const double* v0, int max_pairs,
                     double* evals, double* evecs, int* nconv_out, int* its_out, double* resid);

extern "C" int kb_eigs(kb_handle h, int nev, int ncv, double tol, int maxit, int which,
                       const double* target, int true_residual, const double* v0, int max_pairs,
                       double* evals, double* evecs, int* nconv_out, int* its_out, double* resid) {
  if (!h || !evals || !nconv_out) return KB_EINVAL;
  // a chain sweep whose device-side wait expired sends the handle to safe mode (refactored with
  // the per-node kernels); the eigensolve is then repeated from its start vector, once
  int rc = eigs_impl(h, nev, ncv, tol, maxit, which, target, true_residual, v0, max_pairs, evals, evecs, nconv_out,
                     its_out, resid);
  if (rc == KB_EPROTOCOL_RETRY)
    rc = eigs_impl(h, nev, ncv, tol, maxit, which, target, true_residual, v0, max_pairs, evals, evecs, nconv_out,
                   its_out, resid);
  return rc == KB_EPROTOCOL_RETRY ? kb_fail(h, KB_ECUDA, "chain sweep failed twice") : rc;
}

static int eigs_impl(kb_handle h, int nev, int ncv, double tol, int maxit, int which,
                     const double* target, int true_residual, const double* v0, int max_pairs,
                     double* evals, double* evecs, int* nconv_out, int* its_out, double* resid) {
  if (!h->factored) return kb_fail(h, KB_EINVAL, "kb_factor must succeed before kb_eigs");
  if (!h->B.present) return kb_fail(h, KB_EINVAL, "kb_eigs needs a B matrix");
  if (which < KB_WHICH_LM || which > KB_WHICH_TI) return kb_fail(h, KB_EINVAL, "bad `which`");
  if (nev < 1) return kb_fail(h, KB_EINVAL, "nev must be >= 1");
  KB_CUDA(h, cudaSetDevice(h->device));
  const int n = (int)h->n;
  if (ncv <= 0) ncv = std::max(2 * nev, nev + 15);
  if (ncv > n) ncv = n;
  if (ncv > KB_MAX_NCV)
    return kb_fail(h, KB_EINVAL, "ncv = %d exceeds the supported basis size %d (nev <= 63 with SLEPc's default ncv)",
                   ncv, KB_MAX_NCV);
  if (nev >= ncv) return kb_fail(h, KB_EINVAL, "nev must be < ncv");
  if (maxit < 1) maxit = 1;
  const Z tau = target ? Z(target[0], target[1]) : h->sigma;
  const Z sigma = h->sigma;
  cudaStream_t s = h->stream;
  const int m = ncv;

  KbEventPair ev;
  KB_CUDA(h, ev.create());
  KB_CUDA(h, cudaEventRecord(ev.e0, s));
  h->stats.op_applies = 0;
  h->stats.solve_calls = 0;
  h->stats.eigs_solve_ms = 0.0;
  h->time_sweeps = true;
  struct SweepTimerGuard {
    kb_context* h;
    ~SweepTimerGuard() {
      h->time_sweeps = false;
      h->keep_sharded = false;
      for (cudaEvent_t e : h->sweep_events) cudaEventDestroy(e);
      h->sweep_events.clear();
    }
  } sweep_guard{h};

  KB_TRY(kbi_solve_workspace(h));

  Krylov K;
  K.h = h;
  K.n = n;
  K.ncv = ncv;
  K.ldv = ((int64_t)n + 15) / 16 * 16;
  K.sharded = h->nranks > 1;
  K.r0 = 0;
  K.nl = n;
  if (K.sharded) {
    // row-sharded basis: vectors live on the rank's segment, solves do not publish their result
    int64_t r0, r1;
    kbi_shard_rows(h, &r0, &r1);
    K.r0 = (int)r0;
    K.nl = (int)(r1 - r0);
    h->keep_sharded = true;
  }
  K.nchunks = (K.nl + KB_DOT_CHUNK - 1) / KB_DOT_CHUNK;
  K.nblocks = (int)nblk(K.nl, 256);
  KB_CUDA(h, h->d_V.alloc((size_t)K.ldv * (ncv + 1)));
  KB_CUDA(h, h->d_w.alloc(n));
  KB_CUDA(h, h->d_w2.alloc((size_t)3 * n));
  KB_CUDA(h, h->d_h.alloc(4 * (ncv + 2)));
  KB_CUDA(h, h->d_hpart.alloc((size_t)(ncv + 1) * K.nchunks));
  KB_CUDA(h, h->d_Q.alloc((size_t)ncv * ncv));
  DevBuf<double>& d_normpart = h->d_normpart;
  DevBuf<double>& d_beta = h->d_beta;
  KB_CUDA(h, d_normpart.alloc(std::max(K.nblocks, 3 * 512)));
  KB_CUDA(h, d_beta.alloc(4));
  K.V = h->d_V.p;
  K.w = h->d_w.p;
  K.hdev = h->d_h.p;
  K.hsum = h->d_h.p + (ncv + 2);
  K.hscratch = h->d_h.p + 2 * (ncv + 2);
  K.hpart = h->d_hpart.p;
  K.Qdev = h->d_Q.p;
  K.normpart = d_normpart.p;
  K.beta_dev = d_beta.p;
  // pinned mirrors live on the context, sized for the largest ncv seen
  if (h->pinned_ncv < ncv) {
    if (h->pinned_h) cudaFreeHost(h->pinned_h);
    if (h->pinned_beta) cudaFreeHost(h->pinned_beta);
    h->pinned_h = h->pinned_beta = nullptr;
    h->pinned_ncv = 0;
    KB_CUDA(h, cudaMallocHost((void**)&h->pinned_h, (size_t)(ncv + 2) * (ncv + 2) * sizeof(double2)));
    KB_CUDA(h, cudaMallocHost((void**)&h->pinned_beta, (size_t)(ncv + 2) * sizeof(double)));
    h->pinned_ncv = ncv;
  }
  if (!h->pinned_err) KB_CUDA(h, cudaMallocHost((void**)&h->pinned_err, sizeof(int)));
  K.h_host = (double2*)h->pinned_h;
  K.beta_host = (double*)h->pinned_beta;

  // ---- start vector: v0 (or seeded random) mapped to chain order, pushed through
  //      the operator once so that it lies in range(OP) (no null(B) component), unit norm
  int refine_eigs = h->opt_refine_eigs;  // < 0: decided below
  {
    std::vector<double2> hv(n);
    if (v0) {
      for (int i = 0; i < n; ++i) hv[i] = zmake(v0[2 * i], v0[2 * i + 1]);
    } else {
      std::mt19937_64 gen((uint64_t)h->opt_seed);
      std::normal_distribution<double> nd(0.0, 1.0);
      for (int i = 0; i < n; ++i) {
        double a = nd(gen), b = nd(gen);
        hv[i] = zmake(a, b);
      }
    }
    KB_CUDA(h, cudaMemcpyAsync(h->d_w2.p, hv.data(), (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, s));
    KB_TRY(kbi_to_chain(h, h->d_w2.p, K.w));
    KB_CUDA(h, cudaStreamSynchronize(s));
    // Accuracy of the operator applications.  Every Ritz residual inherits the error of the solves
    // that built the basis, amplified by |theta_max / theta_i|; on a well-conditioned pencil (the
    // benchmark's synthetic one) the plain sweep is accurate to 1e-15 and nothing is needed, on the
    // reference-assembled E = 1e-8 pencil (shift 2.7e-4 from an eigenvalue) it is not, and the pairs
    // furthest from the target came out at 6e-10 -- above the 1e-10 bar -- however often they were
    // purified afterwards (measured: profiles/r2A_assembly_bench.json).  Automatic mode therefore
    // solves for the start vector WITH one refinement step and looks at the size of that correction:
    // the forward error of the unrefined solve.
    if (refine_eigs < 0) {
      refine_eigs = 0;
      if (!K.sharded) {
        KB_TRY(kbi_apply_op_chain(h, K.w, K.col(0), 1));
        const int nbk = std::min((int)nblk(n, 256), 512);
        double nrm[2] = {0.0, 0.0};
        const double2* vec[2] = {h->d_x0.p, h->d_y.p};  // last correction, corrected solution (scaled chain space)
        for (int q = 0; q < 2; ++q) {
          kb_norm2_partial<<<nbk, 256, 0, s>>>(n, vec[q], K.normpart);
          kb_norm_finish<<<1, 32, 0, s>>>(nbk, K.normpart, K.beta_dev);
          h->launches += 2;
          KB_CUDA(h, cudaMemcpyAsync(&nrm[q], K.beta_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
          KB_CUDA(h, cudaStreamSynchronize(s));
        }
        h->stats.refine_resid = nrm[1] > 0.0 ? nrm[0] / nrm[1] : 0.0;
        if (!(h->stats.refine_resid <= KB_REFINE_AUTO)) refine_eigs = 1;
      } else {
        KB_TRY(kbi_apply_op_chain(h, K.w, K.col(0), 0));
      }
    } else {
      KB_TRY(kbi_apply_op_chain(h, K.w, K.col(0), refine_eigs));
    }
    h->stats.op_applies++;
    KB_TRY(K.normalize(K.col(0)));
  }

  kbd::Mat S(m + 1, m);  // projected matrix with residual row
  int nconv = 0, k = 0, its = 0;
  kbd::Mat Qlast;
  int nconv_before_last = 0;
  bool failed_schur = false;

  auto ritz_key = [&](Z theta) { return which_key(sigma + 1.0 / theta, which, tau); };

  while (true) {
    ++its;
    for (int j = k; j < m; ++j) {
      KB_TRY(K.arnoldi_step(j, refine_eigs));
      h->stats.op_applies++;
    }
    // the sweep error flag travels with the projected matrix: a failed exchange is seen after
    // one restart at the latest, not after maxit restarts of garbage
    *h->pinned_err = 0;
    if (h->d_sweep_err.p)
      KB_CUDA(h, cudaMemcpyAsync(h->pinned_err, h->d_sweep_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    KB_TRY(kbi_sync(h));
    if (*h->pinned_err != 0) return kbi_enter_safe_mode(h, "sweep", *h->pinned_err);
    for (int j = k; j < m; ++j) {
      for (int i = 0; i <= j; ++i) {
        double2 v = K.h_host[(size_t)j * (ncv + 1) + i];
        S(i, j) = Z(v.x, v.y);
      }
      S(j + 1, j) = Z(K.beta_host[j], 0.0);
      if (!(K.beta_host[j] == K.beta_host[j]))
        return kb_fail(h, KB_ESINGULAR, "non-finite Arnoldi vector (operator application produced NaN)");
    }
    const double beta = K.beta_host[m - 1];

    // ---- Schur form of the active window, wanted Ritz values first
    const int na = m - nconv;
    kbd::Mat Ha(na, na), Q;
    for (int j = 0; j < na; ++j)
      for (int i = 0; i < na; ++i) Ha(i, j) = S(nconv + i, nconv + j);
    if (!kbd::schur(Ha, Q)) {
      failed_schur = true;
      break;
    }
    kbd::sort_schur(Ha, Q, ritz_key);
    // coupling with the locked block and write back
    if (nconv > 0) {
      kbd::Mat C(nconv, na);
      for (int j = 0; j < na; ++j)
        for (int i = 0; i < nconv; ++i) {
          Z acc(0, 0);
          for (int l = 0; l < na; ++l) acc += S(i, nconv + l) * Q(l, j);
          C(i, j) = acc;
        }
      for (int j = 0; j < na; ++j)
        for (int i = 0; i < nconv; ++i) S(i, nconv + j) = C(i, j);
    }
    for (int j = 0; j < na; ++j)
      for (int i = 0; i < na; ++i) S(nconv + i, nconv + j) = Ha(i, j);
    std::vector<Z> bres(m, Z(0, 0));
    for (int j = 0; j < na; ++j) bres[nconv + j] = beta * Q(na - 1, j);

    // ---- convergence of the leading Ritz pairs
    kbd::Mat Tm(m, m);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i <= j; ++i) Tm(i, j) = S(i, j);
    int kc = nconv;
    for (int i = nconv; i < m; ++i) {
      std::vector<Z> z = kbd::tri_eigvec(Tm, i);
      Z dot(0, 0);
      for (int l = 0; l <= i; ++l) dot += bres[l] * z[l];
      double err = std::abs(dot) / std::abs(Tm(i, i));
      if (err < tol && true_residual) {
        // -eps_true_residual: ||A x - lambda B x|| <= tol |lambda| ||x|| on the Ritz vector itself.
        // Coefficients of x in the un-rotated basis V[:, 0:m]: locked part as is, active part through Q.
        std::vector<double2> c(m, zmake(0.0, 0.0));
        for (int l = 0; l < nconv && l <= i; ++l) c[l] = zmake(z[l].real(), z[l].imag());
        for (int r = 0; r < na; ++r) {
          Z acc(0, 0);
          for (int l = nconv; l <= i; ++l) acc += Q(r, l - nconv) * z[l];
          c[nconv + r] = zmake(acc.real(), acc.imag());
        }
        double2* xv = h->d_w2.p;
        double2* axv = h->d_w2.p + n;
        double2* bxv = h->d_w2.p + 2 * (size_t)n;
        KB_CUDA(h, cudaMemcpyAsync(K.hdev, c.data(), m * sizeof(double2), cudaMemcpyHostToDevice, s));
        kb_lincomb<<<nblk(K.nl, 256), 256, 0, s>>>(K.nl, m, K.V + K.r0, K.ldv, K.hdev, xv + K.r0);
        if (K.sharded) KB_TRY(kbi_shard_gather_segments(h, xv));
        KB_TRY(kbi_spmv_A_chain(h, xv, axv));
        KB_TRY(kbi_spmv_B_chain(h, xv, bxv, false));
        const Z lam = sigma + 1.0 / Tm(i, i);
        const int rbk = 256;
        kb_resid_partial<<<rbk, 256, 0, s>>>(n, axv, bxv, xv, zmake(lam.real(), lam.imag()), K.normpart);
        h->launches += 2;
        std::vector<double> rp(3 * rbk);
        KB_CUDA(h, cudaMemcpyAsync(rp.data(), K.normpart, 3 * rbk * sizeof(double), cudaMemcpyDeviceToHost, s));
        KB_CUDA(h, cudaStreamSynchronize(s));
        double r0 = 0, r2 = 0;
        for (int bb = 0; bb < rbk; ++bb) {
          r0 += rp[3 * bb];
          r2 += rp[3 * bb + 2];
        }
        err = std::sqrt(r0) / (std::abs(lam) * std::sqrt(r2));
      }
      if (err < tol)
        kc = i + 1;
      else
        break;
    }
    Qlast = Q;
    nconv_before_last = nconv;
    if (kc >= nev || its >= maxit) {
      nconv = kc;
      break;
    }
    // ---- restart: keep the converged ones plus half of the rest
    int l = std::max(1, (int)((m - kc) * 0.5));
    int newk = kc + l;
    if (newk >= m) newk = m - 1;
    {
      const int q = newk - nconv;
      std::vector<double2> qh((size_t)na * q);
      for (int j = 0; j < q; ++j)
        for (int i = 0; i < na; ++i) qh[(size_t)j * na + i] = zmake(Q(i, j).real(), Q(i, j).imag());
      KB_CUDA(h, cudaMemcpyAsync(K.Qdev, qh.data(), qh.size() * sizeof(double2), cudaMemcpyHostToDevice, s));
      KB_TRY(restart_update(h, K.nl, na, q, K.col(nconv) + K.r0, K.ldv, K.Qdev));
      KB_CUDA(h, cudaMemcpyAsync(K.col(newk), K.col(m), (size_t)n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
      KB_CUDA(h, cudaStreamSynchronize(s));
    }
    for (int j = 0; j < m; ++j)
      for (int i = 0; i <= m; ++i)
        if (i >= newk || j >= newk) S(i, j) = Z(0, 0);
    for (int j = kc; j < newk; ++j) S(newk, j) = bres[j];
    nconv = kc;
    k = newk;
  }
  if (failed_schur) return kb_fail(h, KB_EINVAL, "QR iteration on the projected matrix did not converge");

  // ---- rotate the active part of the basis into the Schur basis
  {
    const int na = m - nconv_before_last;
    std::vector<double2> qh((size_t)na * na);
    for (int j = 0; j < na; ++j)
      for (int i = 0; i < na; ++i) qh[(size_t)j * na + i] = zmake(Qlast(i, j).real(), Qlast(i, j).imag());
    KB_CUDA(h, cudaMemcpyAsync(K.Qdev, qh.data(), qh.size() * sizeof(double2), cudaMemcpyHostToDevice, s));
    KB_TRY(restart_update(h, K.nl, na, na, K.col(nconv_before_last) + K.r0, K.ldv, K.Qdev));
    KB_CUDA(h, cudaStreamSynchronize(s));
  }

  // ---- extract eigenpairs
  int nret = std::min(nconv, max_pairs);
  kbd::Mat Tm(m, m);
  for (int j = 0; j < m; ++j)
    for (int i = 0; i <= j; ++i) Tm(i, j) = S(i, j);
  double2* x = h->d_w2.p;
  double2* ax = h->d_w2.p + n;
  double2* bx = h->d_w2.p + 2 * (size_t)n;
  DevBuf<double2>& d_xo = h->d_xo;
  KB_CUDA(h, d_xo.alloc(n));
  const int rb = 256;
  std::vector<double> rpart(3 * rb);
  for (int i = 0; i < nret; ++i) {
    std::vector<Z> z = kbd::tri_eigvec(Tm, i);
    std::vector<double2> zh(i + 1);
    for (int l = 0; l <= i; ++l) zh[l] = zmake(z[l].real(), z[l].imag());
    KB_CUDA(h, cudaMemcpyAsync(K.hdev, zh.data(), zh.size() * sizeof(double2), cudaMemcpyHostToDevice, s));
    kb_lincomb<<<nblk(K.nl, 256), 256, 0, s>>>(K.nl, i + 1, K.V + K.r0, K.ldv, K.hdev, x + K.r0);
    h->launches++;
    KB_CUDA(h, cudaStreamSynchronize(s));
    Z theta = Tm(i, i);
    Z lam = sigma + 1.0 / theta;
    evals[2 * i] = lam.real();
    evals[2 * i + 1] = lam.imag();
    // Purification x <- OP x (SLEPc's EPSSetPurify) is one step of inverse iteration with the
    // factorisation at hand.  On pencils whose sections differ by many orders of magnitude (thermal
    // runs: |Bx| ~ 1e-11 |x|) one step leaves ||Ax - lam Bx|| / (|lam| ||Bx||) at 1e-9, the next one at
    // 1e-11 .. 1e-12 (measured on the reference-assembled dormy case at N = 150, oracle and GPU
    // alike), so the step is repeated -- with one refinement step of the linear solve -- while the
    // residual is above KB_PURIFY_TARGET and still falling, at most KB_PURIFY_EXTRA more times.
    // Well-scaled pencils (the benchmark's: 1e-14 after the first step) pay nothing.
    double rnorm = 0.0, prev = 1e300;
    for (int round = 0;; ++round) {
      if (h->opt_purify) {
        const int refine = (round > 0 && !K.sharded) ? 1 : refine_eigs;
        KB_TRY(kbi_apply_op_chain(h, x, K.w, refine));
        h->stats.op_applies++;
        KB_CUDA(h, cudaMemcpyAsync(x, K.w, (size_t)n * sizeof(double2), cudaMemcpyDeviceToDevice, s));
      }
      // the vector leaves the sharded world here: every rank holds all of it from now on
      if (K.sharded) KB_TRY(kbi_shard_gather_segments(h, x));
      KB_TRY(normalize_dev(h, n, x, K.normpart, K.beta_dev, std::min((int)nblk(n, 256), 512)));
      // residual ||A x - lam B x|| / (|lam| ||B x||)
      KB_TRY(kbi_spmv_A_chain(h, x, ax));
      KB_TRY(kbi_spmv_B_chain(h, x, bx, false));
      kb_resid_partial<<<rb, 256, 0, s>>>(n, ax, bx, x, zmake(lam.real(), lam.imag()), K.normpart);
      h->launches++;
      KB_CUDA(h, cudaMemcpyAsync(rpart.data(), K.normpart, 3 * rb * sizeof(double), cudaMemcpyDeviceToHost, s));
      KB_CUDA(h, cudaStreamSynchronize(s));
      double r0 = 0, r1 = 0;
      for (int b = 0; b < rb; ++b) {
        r0 += rpart[3 * b];
        r1 += rpart[3 * b + 1];
      }
      rnorm = std::sqrt(r0) / (std::abs(lam) * std::sqrt(r1));
      if (!h->opt_purify || !(rnorm > KB_PURIFY_TARGET) || round >= KB_PURIFY_EXTRA || rnorm > 0.5 * prev) break;
      prev = rnorm;
    }
    if (evecs) {
      KB_TRY(kbi_from_chain(h, x, d_xo.p));
      KB_CUDA(h, cudaMemcpyAsync(evecs + (size_t)2 * n * i, d_xo.p, (size_t)n * sizeof(double2),
                                 cudaMemcpyDeviceToHost, s));
      KB_CUDA(h, cudaStreamSynchronize(s));
    }
    if (resid) resid[i] = rnorm;
  }
  *nconv_out = nret;
  if (its_out) *its_out = its;
  KB_CUDA(h, cudaEventRecord(ev.e1, s));
  KB_TRY(kbi_sync(h));
  h->stats.eigs_ms = ev.ms();
  KB_TRY(kbi_check_sweep_error(h));
  for (size_t i = 0; i + 1 < h->sweep_events.size(); i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->sweep_events[i], h->sweep_events[i + 1]) == cudaSuccess)
      h->stats.eigs_solve_ms += t;
  }
  return KB_OK;
}

// Host-only hook used by the CPU test-suite to validate the projected-problem
// kernels (complex Schur + ordering) against LAPACK through numpy.
extern "C" int kb_dbg_schur(int m, const double* Hin, int which, const double* sigma, const double* tau,
                            double* Tout, double* Qout) {
  kbd::Mat H(m, m), Q;
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) H(i, j) = Z(Hin[2 * (i + (size_t)j * m)], Hin[2 * (i + (size_t)j * m) + 1]);
  if (!kbd::schur(H, Q)) return KB_EINVAL;
  if (which >= 0) {
    Z sg(sigma[0], sigma[1]), ta(tau[0], tau[1]);
    kbd::sort_schur(H, Q, [&](Z th) { return which_key(sg + 1.0 / th, which, ta); });
  }
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) {
      Tout[2 * (i + (size_t)j * m)] = H(i, j).real();
      Tout[2 * (i + (size_t)j * m) + 1] = H(i, j).imag();
      Qout[2 * (i + (size_t)j * m)] = Q(i, j).real();
      Qout[2 * (i + (size_t)j * m) + 1] = Q(i, j).imag();
    }
  return KB_OK;
}
