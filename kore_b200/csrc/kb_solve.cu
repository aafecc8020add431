// Application of the factored chain: forward / backward block substitution,
// iterative refinement, the B (and A) SpMV and the public solve entry points.
//
// Replaces: KSP preonly + PC lu solve phase of MUMPS (K.solve(bvec,x),
// /root/reference/bin/solve.py:227, and every ST application inside
// E.solve(), solve.py:123) and PETSc MatMult (w = B v inside EPSSolve).
//
//   forward :  y_p = M_p ( r_p - L_{p,p-1} y_{p-1} )          p = 0 .. P-1
//   backward:  x_p = y_p - M_p ( U_{p,p+1} x_{p+1} )          p = P-2 .. 0
// Both sweeps are HBM-bound: each reads every explicit inverse M_p once
// (16 b_p^2 bytes per node per sweep).
#include <math.h>
#include <stdio.h>

#include <chrono>
#include <thread>

#include "kb_internal.cuh"

#define KB_NODE_WARPS 4

// Sparse coupling of one node (one warp per row, lanes over the row's entries):
//   MODE 0 (forward):  t_i = r_i - sum_k L_{p,p-1}[i,k] y[k]
//   MODE 1 (backward): t_i =       sum_k U_{p,p+1}[i,k] y[k]
template <int MODE>
__global__ void __launch_bounds__(256)
kb_node_tvec(int b, int o, const double2* __restrict__ r, const double2* __restrict__ y,
             double2* __restrict__ t, const int64_t* __restrict__ rowptr,
             const int64_t* __restrict__ dstart, const int64_t* __restrict__ ustart,
             const int* __restrict__ col, const double2* __restrict__ T) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= b) return;
  const int gi = o + i;
  int64_t k0, k1;
  if (MODE == 0) {
    k0 = rowptr[gi];
    k1 = dstart[gi];
  } else {
    k0 = ustart[gi];
    k1 = rowptr[gi + 1];
  }
  double2 acc = zmake(0.0, 0.0);
  for (int64_t k = k0 + lane; k < k1; k += 32) zfma(acc, T[k], y[col[k]]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
  }
  if (lane == 0) t[gi] = (MODE == 0) ? zsub(r[gi], acc) : acc;
}

// Dense part of one node.  A CTA of 8 warps owns KB_NODE_ROWS rows of M_p; two
// warps share a row (half each) and every lane keeps 8 independent 16-byte loads
// in flight, so one CTA per SM has ~32 KB outstanding (HBM latency x bandwidth).
//   MODE 0: y_p = M_p t_p          MODE 1: y_p -= M_p t_p
#define KB_NODE_ROWS 4
template <int MODE>
__global__ void __launch_bounds__(256)
kb_node_gemv(const double2* __restrict__ M, int b, int o, const double2* __restrict__ t,
             double2* __restrict__ y) {
  extern __shared__ double2 tvec[];  // b entries
  __shared__ double2 part[8];
  for (int i = threadIdx.x; i < b; i += blockDim.x) tvec[i] = t[o + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * KB_NODE_ROWS + (wid >> 1);
  const int half = wid & 1;
  const int hb = (b + 1) >> 1;
  const int j0 = half ? hb : 0, j1 = half ? b : hb;
  double2 acc = zmake(0.0, 0.0);
  if (row < b) {
    const double2* Mrow = M + (size_t)row * b;
    for (int j = j0 + lane; j < j1; j += 256) {
      double2 m[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        int jj = j + 32 * u;
        m[u] = jj < j1 ? __ldcs(&Mrow[jj]) : zmake(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        int jj = j + 32 * u;
        if (jj < j1) zfma(acc, m[u], tvec[jj]);
      }
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
  }
  if (lane == 0) part[wid] = acc;
  __syncthreads();
  if (threadIdx.x < KB_NODE_ROWS) {
    int rr = blockIdx.x * KB_NODE_ROWS + threadIdx.x;
    if (rr < b) {
      double2 v = zadd(part[2 * threadIdx.x], part[2 * threadIdx.x + 1]);
      if (MODE == 0)
        y[o + rr] = v;
      else
        y[o + rr] = zsub(y[o + rr], v);
    }
  }
}

// y = alpha_rows .* (CSR x)   (one warp per row).  VT = double or double2 values.
template <typename VT>
__device__ __forceinline__ double2 kb_valmul(VT v, double2 x);
template <>
__device__ __forceinline__ double2 kb_valmul<double>(double v, double2 x) { return zscale(x, v); }
template <>
__device__ __forceinline__ double2 kb_valmul<double2>(double2 v, double2 x) { return zmul(v, x); }

// MODE 0: y = s .* (M x); MODE 1: y = r - M x.  LPR lanes per row (a power of two <= 32): Kore's B
// has ~10 entries per row and A ~35, so a full warp per row leaves most lanes idle.
template <typename VT, int MODE, int LPR = 32>
__global__ void kb_spmv(int n, const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                        const VT* __restrict__ val, const double2* __restrict__ x,
                        const double* __restrict__ rowscale, const double2* __restrict__ r,
                        double2* __restrict__ y) {
  int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR);
  int lane = threadIdx.x & (LPR - 1);
  if (row >= n) return;
  double2 acc = zmake(0.0, 0.0);
  for (int64_t k = rowptr[row] + lane, e = rowptr[row + 1]; k < e; k += LPR)
    acc = zadd(acc, kb_valmul<VT>(val[k], x[col[k]]));
#pragma unroll
  for (int s = LPR / 2; s > 0; s >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
  }
  if (lane == 0) {
    if (MODE == 0) {
      double sc = rowscale ? rowscale[row] : 1.0;
      y[row] = zscale(acc, sc);
    } else {
      y[row] = zsub(r[row], acc);
    }
  }
}

// out[k] = scale[k] * in[perm[k]]   (original -> chain)
__global__ void kb_gather_scale(int n, const int* __restrict__ perm, const double* __restrict__ scale,
                                const double2* __restrict__ in, double2* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double s = scale ? scale[k] : 1.0;
  out[k] = zscale(in[perm[k]], s);
}
// out[perm[k]] = scale[k] * in[k]   (chain -> original)
__global__ void kb_scatter_scale(int n, const int* __restrict__ perm, const double* __restrict__ scale,
                                 const double2* __restrict__ in, double2* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double s = scale ? scale[k] : 1.0;
  out[perm[k]] = zscale(in[k], s);
}
__global__ void kb_scale_vec(int n, const double* __restrict__ scale, const double2* __restrict__ in,
                             double2* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = zscale(in[k], scale[k]);
}
__global__ void kb_axpy1(int n, const double2* __restrict__ dx, double2* __restrict__ x) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  x[k] = zadd(x[k], dx[k]);
}
// partial sums of |v|^2 per block -> out[blockIdx]
__global__ void kb_norm2_partial(int n, const double2* __restrict__ v, double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) acc += zabs2(v[k]);
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
  }
}


// one forward + backward sweep: y <- T'^{-1} r   (scaled chain space), raw launches
static int chain_sweeps_launch(kb_context* h, const double2* r, double2* y) {
  cudaStream_t s = h->stream;
  const int64_t P = h->P;
  double2* t = h->d_t.p;
  for (int64_t p = 0; p < P; ++p) {
    int o = (int)h->nodeptr[p], b = (int)(h->nodeptr[p + 1] - h->nodeptr[p]);
    kb_node_tvec<0><<<(b + 7) / 8, 256, 0, s>>>(b, o, r, y, t, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                                h->d_col.p, h->d_Tval.p);
    kb_node_gemv<0><<<(b + KB_NODE_ROWS - 1) / KB_NODE_ROWS, 256, b * sizeof(double2), s>>>(
        h->d_M.p + h->Moff[p], b, o, t, y);
  }
  for (int64_t p = P - 2; p >= 0; --p) {
    int o = (int)h->nodeptr[p], b = (int)(h->nodeptr[p + 1] - h->nodeptr[p]);
    kb_node_tvec<1><<<(b + 7) / 8, 256, 0, s>>>(b, o, r, y, t, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                                h->d_col.p, h->d_Tval.p);
    kb_node_gemv<1><<<(b + KB_NODE_ROWS - 1) / KB_NODE_ROWS, 256, b * sizeof(double2), s>>>(
        h->d_M.p + h->Moff[p], b, o, t, y);
  }
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

void kbi_drop_graphs(kb_context* h) {
  for (auto& g : h->sweep_graphs)
    if (g.exec) cudaGraphExecDestroy((cudaGraphExec_t)g.exec);
  h->sweep_graphs.clear();
}

// The 2(2P-1) launches of a sweep pair are captured once per (r, y) buffer pair
// and replayed as a CUDA graph: the chain is latency-bound, so per-launch CPU
// and driver overhead is what the graph removes.
static int chain_sweeps(kb_context* h, const double2* r, double2* y) {
  cudaStream_t s = h->stream;
  if (h->nranks > 1) return kbi_sharded_sweeps(h, r, y);
  if (h->inject_fault == 1 && h->d_sweep_err.p) {
    // tests: behave as if a wait of this sweep had expired (every wait of the launch gives up)
    static const int code = KB_WERR_WATCHDOG;
    KB_CUDA(h, cudaMemcpyAsync(h->d_sweep_err.p, &code, sizeof(int), cudaMemcpyHostToDevice, s));
    h->inject_fault = 0;
  }
  if (h->fold_ready) return kbi_sweep_fold(h, r, y);
  if (h->M_transposed) return kbi_sweep_onehop(h, r, y);
  if (h->opt_sweep == 1 && h->sweep_grid > 0) return kbi_sweep_dataflow(h, r, y);
  if (h->opt_sweep == 2 && h->sweep_grid > 0) return kbi_sweep_persistent(h, r, y);
  h->launches += 2 * (2 * h->P - 1);
  for (auto& g : h->sweep_graphs)
    if (g.r == r && g.y == y) {
      KB_CUDA(h, cudaGraphLaunch((cudaGraphExec_t)g.exec, s));
      return KB_OK;
    }
  if (h->sweep_graphs.size() >= 8) return chain_sweeps_launch(h, r, y);
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  KB_CUDA(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = chain_sweeps_launch(h, r, y);
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  if (rc != KB_OK) return rc;
  if (e != cudaSuccess) return kb_fail(h, KB_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return kb_fail(h, KB_ECUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  kb_context::SweepGraph g;
  g.r = r;
  g.y = y;
  g.exec = (void*)exec;
  h->sweep_graphs.push_back(g);
  KB_CUDA(h, cudaGraphLaunch(exec, s));
  return KB_OK;
}

// Switch the handle to the kernels that have no device-side waits (one kernel pair per node for
// the sweeps, per-step kernels for the factorisation), refactor at the current shift and tell the
// caller to repeat its operation.  Reached when a wait of a persistent kernel has expired
// (wait_code says which, kb_internal.cuh) or the host watchdog has fired; never in a healthy run.
int kbi_enter_safe_mode(kb_context* h, const char* what, int wait_code) {
  h->stats.protocol_fallbacks++;
  h->stats.wait_error = wait_code;
  fprintf(stderr,
          "libkoreb200: WARNING: a device-side wait of the persistent %s kernel expired (code %d, CTA %d); "
          "falling back to the per-node kernels for this handle and repeating the call\n",
          what, wait_code & 255, wait_code >> 8);
  if (h->safe_mode)
    return kb_fail(h, KB_ECUDA, "persistent-kernel time-out reported in safe mode (code %d)", wait_code);
  if (h->nranks > 1)
    return kb_fail(h, KB_ECUDA, "persistent-kernel time-out on rank %d of an l-sharded pencil (code %d, CTA %d); "
                   "set KB_SHARD_GENERAL=1 to use the per-node kernels", h->rank, wait_code & 255, wait_code >> 8);
  h->safe_mode = true;
  h->opt_factor = 0;
  h->opt_sweep = 0;
  h->opt_fold = 0;
  if (h->d_sweep_err.p) KB_CUDA(h, cudaMemsetAsync(h->d_sweep_err.p, 0, sizeof(int), h->stream));
  const bool was_factored = h->factored || what[0] == 'f';
  h->factored = false;
  if (was_factored) KB_TRY(kbi_factor(h, h->sigma));
  return KB_EPROTOCOL_RETRY;
}

int kbi_sync(kb_context* h) {
  cudaStream_t s = h->stream;
  cudaError_t q = cudaStreamQuery(s);
  if (q == cudaSuccess) return KB_OK;
  const double limit_ms = h->watchdog_ms > 0.0 ? h->watchdog_ms : 3.0 * (double)h->wait_ns * 1e-6 + 5000.0;
  for (int phase = 0; phase < 2; ++phase) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      q = cudaStreamQuery(s);
      if (q == cudaSuccess) return KB_OK;
      if (q != cudaErrorNotReady)
        return kb_fail(h, KB_ECUDA, "stream failed: %s", cudaGetErrorString(q));
      const double el = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (el > limit_ms) break;
      if (el > 1.0) std::this_thread::sleep_for(std::chrono::microseconds(el > 50.0 ? 200 : 40));
    }
    if (phase == 1) break;
    // deadline: raise the flags every device-side wait looks at, from a side stream
    fprintf(stderr, "libkoreb200: WARNING: kernel still running after %.0f ms; watchdog raises the abort flags\n",
            limit_ms);
    if (!h->side_stream) KB_CUDA(h, cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    static const int code = KB_WERR_WATCHDOG;
    if (h->d_sweep_err.p)
      KB_CUDA(h, cudaMemcpyAsync(h->d_sweep_err.p, &code, sizeof(int), cudaMemcpyHostToDevice, h->side_stream));
    if (h->d_kfsync.p && h->d_kfsync.count > KF_ERR_WORD)
      KB_CUDA(h, cudaMemcpyAsync(h->d_kfsync.p + KF_ERR_WORD, &code, sizeof(int), cudaMemcpyHostToDevice,
                                 h->side_stream));
    KB_CUDA(h, cudaStreamSynchronize(h->side_stream));
  }
  return kb_fail(h, KB_ECUDA, "a kernel does not react to the watchdog flag; the device needs a reset");
}

// Call after kbi_sync.  A raised flag means some chain sweep since the last check produced
// garbage: the handle goes to safe mode and the caller repeats (KB_EPROTOCOL_RETRY).
int kbi_check_sweep_error(kb_context* h) {
  if (!h->d_sweep_err.p) return KB_OK;
  int e = 0;
  KB_CUDA(h, cudaMemcpy(&e, h->d_sweep_err.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) return kbi_enter_safe_mode(h, "sweep", e);
  return KB_OK;
}

int kbi_solve_workspace(kb_context* h) {
  const int64_t n = h->n;
  KB_CUDA(h, h->d_r.alloc(n));
  // solution buffers carry a zero sentinel at [n] (padding target of the ELL couplings); the
  // buffers survive a change of pencil, so the sentinel is rewritten whenever n changes
  if (h->d_y.count < (size_t)n + 1 || h->d_x0.count < (size_t)n + 1 || h->ws_n != n) {
    KB_CUDA(h, h->d_y.alloc(n + 1));
    KB_CUDA(h, h->d_x0.alloc(n + 1));
    KB_CUDA(h, cudaMemsetAsync(h->d_y.p, 0, (size_t)(n + 1) * sizeof(double2), h->stream));
    KB_CUDA(h, cudaMemsetAsync(h->d_x0.p, 0, (size_t)(n + 1) * sizeof(double2), h->stream));
    h->ws_n = n;
  }
  KB_CUDA(h, h->d_res.alloc(n));
  KB_CUDA(h, h->d_t.alloc(n));
  KB_CUDA(h, h->d_in.alloc(n));
  KB_CUDA(h, h->d_out.alloc(n));
  KB_CUDA(h, h->d_partial.alloc(1024));
  return KB_OK;
}

int kbi_chain_solve(kb_context* h, const double2* r_dev, double2* x_dev, int refine) {
  if (!h->factored) return kb_fail(h, KB_EINVAL, "kb_factor must succeed before solving");
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (h->time_sweeps) {
    KB_CUDA(h, cudaEventCreate(&ev0));
    KB_CUDA(h, cudaEventCreate(&ev1));
    h->sweep_events.push_back(ev0);
    h->sweep_events.push_back(ev1);
    KB_CUDA(h, cudaEventRecord(ev0, s));
  }
  KB_TRY(chain_sweeps(h, r_dev, x_dev));
  h->stats.solve_calls++;
  for (int it = 0; it < refine; ++it) {
    kb_spmv<double2, 1, 16><<<nblk((int64_t)n * 16, 256), 256, 0, s>>>(n, h->d_rowptr.p, h->d_col.p, h->d_Tval.p,
                                                                   x_dev, nullptr, r_dev, h->d_res.p);
    KB_TRY(chain_sweeps(h, h->d_res.p, h->d_x0.p));
    kb_axpy1<<<nblk(n, 256), 256, 0, s>>>(n, h->d_x0.p, x_dev);
    h->launches += 2;
    h->stats.solve_calls++;
  }
  if (ev1) KB_CUDA(h, cudaEventRecord(ev1, s));
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

int kbi_spmv_B_chain(kb_context* h, const double2* x, double2* y, bool scale_rows) {
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  if (!h->B.present) return kb_fail(h, KB_EINVAL, "pencil has no B matrix");
  const double* rs = scale_rows ? h->d_rscale.p : nullptr;
  if (h->b_is_complex)
    kb_spmv<double2, 0, 8><<<nblk((int64_t)n * 8, 256), 256, 0, s>>>(n, h->d_browptr.p, h->d_bcol.p, h->d_bval_c.p,
                                                                   x, rs, nullptr, y);
  else
    kb_spmv<double, 0, 8><<<nblk((int64_t)n * 8, 256), 256, 0, s>>>(n, h->d_browptr.p, h->d_bcol.p, h->d_bval_r.p, x,
                                                                  rs, nullptr, y);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

int kbi_spmv_A_chain(kb_context* h, const double2* x, double2* y) {
  const int n = (int)h->n;
  kb_spmv<double2, 0, 16><<<nblk((int64_t)n * 16, 256), 256, 0, h->stream>>>(n, h->d_rowptr.p, h->d_col.p,
                                                                         h->d_Aval.p, x, nullptr, nullptr, y);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

int kbi_to_chain(kb_context* h, const double2* x_orig, double2* x_chain) {
  const int n = (int)h->n;
  kb_gather_scale<<<nblk(n, 256), 256, 0, h->stream>>>(n, h->d_perm.p, nullptr, x_orig, x_chain);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}
int kbi_from_chain(kb_context* h, const double2* x_chain, double2* x_orig) {
  const int n = (int)h->n;
  kb_scatter_scale<<<nblk(n, 256), 256, 0, h->stream>>>(n, h->d_perm.p, nullptr, x_chain, x_orig);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

// out = C T'^{-1} R B in      (chain order, unscaled vectors)
// With h->keep_sharded (l-sharded eigensolve) `in` and `out` carry this rank's segment only: one
// node of halo from each neighbour for the B product, the rows of the segment, the sharded solve
// without its final publication, the column scaling of the segment.
int kbi_apply_op_chain(kb_context* h, const double2* in_chain, double2* out_chain, int refine) {
  const int n = (int)h->n;
  KB_TRY(kbi_solve_workspace(h));
  if (h->nranks > 1 && h->keep_sharded) {
    cudaStream_t s = h->stream;
    int64_t r0, r1;
    kbi_shard_rows(h, &r0, &r1);
    const int nl = (int)(r1 - r0);
    KB_TRY(kbi_shard_halo(h, const_cast<double2*>(in_chain)));
    if (h->b_is_complex)
      kb_spmv<double2, 0, 8><<<nblk((int64_t)nl * 8, 256), 256, 0, s>>>(nl, h->d_browptr.p + r0, h->d_bcol.p, h->d_bval_c.p,
                                                                      in_chain, h->d_rscale.p + r0, nullptr,
                                                                      h->d_r.p + r0);
    else
      kb_spmv<double, 0, 8><<<nblk((int64_t)nl * 8, 256), 256, 0, s>>>(nl, h->d_browptr.p + r0, h->d_bcol.p, h->d_bval_r.p,
                                                                     in_chain, h->d_rscale.p + r0, nullptr,
                                                                     h->d_r.p + r0);
    KB_TRY(kbi_chain_solve(h, h->d_r.p, h->d_y.p, 0));
    kb_scale_vec<<<nblk(nl, 256), 256, 0, s>>>(nl, h->d_cscale.p + r0, h->d_y.p + r0, out_chain + r0);
    h->launches += 2;
    KB_LAUNCH_CHECK(h);
    return KB_OK;
  }
  KB_TRY(kbi_spmv_B_chain(h, in_chain, h->d_r.p, true));
  KB_TRY(kbi_chain_solve(h, h->d_r.p, h->d_y.p, refine));
  kb_scale_vec<<<nblk(n, 256), 256, 0, h->stream>>>(n, h->d_cscale.p, h->d_y.p, out_chain);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

// ---------------------------------------------------------------------------
// public entry points
// ---------------------------------------------------------------------------
static int solve_dev_once(kb_context* h, const double2* rhs_dev, double2* x_dev, int nrhs) {
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  KB_TRY(kbi_solve_workspace(h));
  KbEventPair ev;
  KB_CUDA(h, ev.create());
  KB_CUDA(h, cudaEventRecord(ev.e0, s));
  for (int c = 0; c < nrhs; ++c) {
    kb_gather_scale<<<nblk(n, 256), 256, 0, s>>>(n, h->d_perm.p, h->d_rscale.p, rhs_dev + (size_t)c * n,
                                                 h->d_r.p);
    KB_TRY(kbi_chain_solve(h, h->d_r.p, h->d_y.p, h->opt_refine));
    kb_scatter_scale<<<nblk(n, 256), 256, 0, s>>>(n, h->d_perm.p, h->d_cscale.p, h->d_y.p,
                                                  x_dev + (size_t)c * n);
    h->launches += 2;
  }
  KB_CUDA(h, cudaEventRecord(ev.e1, s));
  KB_TRY(kbi_sync(h));
  h->stats.solve_ms = ev.ms() / (nrhs > 0 ? nrhs : 1);
  double bytes = 0.0;
  for (int64_t p = 0; p < h->P; ++p) {
    double b = (double)(h->nodeptr[p + 1] - h->nodeptr[p]);
    bytes += 2.0 * 16.0 * b * b;
  }
  h->stats.solve_bytes = bytes + 3.0 * 16.0 * n;
  return kbi_check_sweep_error(h);
}

// one repetition after a fall-back to safe mode (kbi_enter_safe_mode has refactored)
static int solve_dev_impl(kb_context* h, const double2* rhs_dev, double2* x_dev, int nrhs) {
  int rc = solve_dev_once(h, rhs_dev, x_dev, nrhs);
  if (rc == KB_EPROTOCOL_RETRY) rc = solve_dev_once(h, rhs_dev, x_dev, nrhs);
  return rc == KB_EPROTOCOL_RETRY ? kb_fail(h, KB_ECUDA, "chain sweep failed twice") : rc;
}

extern "C" int kb_solve_dev(kb_handle h, const double* rhs_dev, double* x_dev, int nrhs) {
  if (!h || !rhs_dev || !x_dev || nrhs < 1) return KB_EINVAL;
  if (!h->factored) return kb_fail(h, KB_EINVAL, "kb_factor must succeed before kb_solve");
  KB_CUDA(h, cudaSetDevice(h->device));
  return solve_dev_impl(h, (const double2*)rhs_dev, (double2*)x_dev, nrhs);
}

// staging buffers of the host-pointer entry points (kept on the handle: a cudaMalloc / cudaFree
// pair per call costs more than a sweep)
static int io_buffers(kb_context* h, size_t cnt) {
  KB_CUDA(h, h->d_io_a.alloc(cnt));
  KB_CUDA(h, h->d_io_b.alloc(cnt));
  return KB_OK;
}

extern "C" int kb_solve(kb_handle h, const double* rhs, double* x, int nrhs) {
  if (!h || !rhs || !x || nrhs < 1) return KB_EINVAL;
  if (!h->factored) return kb_fail(h, KB_EINVAL, "kb_factor must succeed before kb_solve");
  KB_CUDA(h, cudaSetDevice(h->device));
  const size_t cnt = (size_t)h->n * nrhs;
  KB_TRY(io_buffers(h, cnt));
  KB_CUDA(h, cudaMemcpyAsync(h->d_io_a.p, rhs, cnt * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  KB_TRY(solve_dev_impl(h, h->d_io_a.p, h->d_io_b.p, nrhs));
  KB_CUDA(h, cudaMemcpyAsync(x, h->d_io_b.p, cnt * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}

extern "C" int kb_apply_op(kb_handle h, const double* x, double* y) {
  if (!h || !x || !y) return KB_EINVAL;
  if (!h->factored) return kb_fail(h, KB_EINVAL, "kb_factor must succeed before kb_apply_op");
  KB_CUDA(h, cudaSetDevice(h->device));
  const int n = (int)h->n;
  for (int attempt = 0; attempt < 2; ++attempt) {
    KB_TRY(kbi_solve_workspace(h));
    KB_TRY(io_buffers(h, n));
    KB_CUDA(h, cudaMemcpyAsync(h->d_io_a.p, x, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    KB_TRY(kbi_to_chain(h, h->d_io_a.p, h->d_in.p));
    KB_TRY(kbi_apply_op_chain(h, h->d_in.p, h->d_out.p, h->opt_refine));
    KB_TRY(kbi_from_chain(h, h->d_out.p, h->d_io_b.p));
    KB_CUDA(h, cudaMemcpyAsync(y, h->d_io_b.p, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    KB_TRY(kbi_sync(h));
    const int rc = kbi_check_sweep_error(h);
    if (rc != KB_EPROTOCOL_RETRY) return rc;
  }
  return kb_fail(h, KB_ECUDA, "chain sweep failed twice");
}

extern "C" int kb_matvec(kb_handle h, int which, const double* x, double* y) {
  if (!h || !x || !y) return KB_EINVAL;
  if (!h->chain_set) return kb_fail(h, KB_EINVAL, "kb_set_chain must be called before kb_matvec");
  KB_CUDA(h, cudaSetDevice(h->device));
  const int n = (int)h->n;
  KB_TRY(kbi_solve_workspace(h));
  KB_TRY(io_buffers(h, n));
  KB_CUDA(h, cudaMemcpyAsync(h->d_io_a.p, x, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  KB_TRY(kbi_to_chain(h, h->d_io_a.p, h->d_in.p));
  if (which == 0)
    KB_TRY(kbi_spmv_A_chain(h, h->d_in.p, h->d_out.p));
  else
    KB_TRY(kbi_spmv_B_chain(h, h->d_in.p, h->d_out.p, false));
  KB_TRY(kbi_from_chain(h, h->d_out.p, h->d_io_b.p));
  KB_CUDA(h, cudaMemcpyAsync(y, h->d_io_b.p, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}
