// l-sharded factor / solve: ONE pencil across the GPUs of a node, one rank per GPU.
//
// Replaces what PETSc's row ownership (/root/reference/bin/solve.py:50, 76) and
// MUMPS' distributed fronts / solve phase do across MPI ranks.
//
// The chain of P nodes is cut into G contiguous segments (rank g owns nodes
// [lo_g, hi_g)).  The last node of every segment but the final one is a SEPARATOR;
// the others are the rank's INTERIOR.  Each rank eliminates its interior by block
// Thomas exactly as on one GPU, and additionally carries the fill towards the
// separator t above it ("spikes"):
//     V_p = M_p F_p        F_{p+1} = -L_{p+1,p} V_p        (block (p, t))
//     H_p = G_p M_p        G_{p+1} = -H_p U_{p,p+1}        (block (t, p))
//     Acc_t = sum_p G_p V_p                                 (Schur update of D_t)
// The (G-1)-node reduced interface system (diagonal blocks D_sep - ..., dense
// couplings) is exchanged with ONE ncclAllGather and factored redundantly on every
// rank.  A solve is: local forward sweep, ncclAllGather of two b-vectors per rank,
// redundant reduced solve, local backward sweep with the spike correction, and a
// grouped ncclBroadcast that leaves the full solution on every rank (the Krylov
// basis is replicated; only the operator is sharded).
// tests/shard_model.py is the numpy statement of the same algebra.
#include <dlfcn.h>
#include <nccl.h>

#include "kb_internal.cuh"

// ---------------------------------------------------------------------------
// NCCL through dlopen (the library must load on boxes without NCCL)
// ---------------------------------------------------------------------------
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;

bool nccl_load() {
  if (g_nccl.ok) return true;
  if (!g_nccl.handle) g_nccl.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.handle) g_nccl.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.handle) return false;
#define KB_SYM(field, name)                                          \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);             \
  if (!g_nccl.field) return false;
  KB_SYM(GetUniqueId, "ncclGetUniqueId");
  KB_SYM(CommInitRank, "ncclCommInitRank");
  KB_SYM(CommDestroy, "ncclCommDestroy");
  KB_SYM(AllGather, "ncclAllGather");
  KB_SYM(Broadcast, "ncclBroadcast");
  KB_SYM(GroupStart, "ncclGroupStart");
  KB_SYM(GroupEnd, "ncclGroupEnd");
  KB_SYM(GetErrorString, "ncclGetErrorString");
#undef KB_SYM
  g_nccl.ok = true;
  return true;
}
}  // namespace

#define KB_NCCL(h, expr)                                                                          \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess)                                                                        \
      return kb_fail((h), KB_ENCCL, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r),    \
                     __FILE__, __LINE__);                                                         \
  } while (0)

void kbi_nccl_destroy(kb_context* h) {
  if (h->nccl_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
  h->nccl_comm = nullptr;
}

extern "C" int kb_nccl_unique_id(void* id128) {
  if (!id128) return KB_EINVAL;
  if (!nccl_load()) return KB_ENCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return KB_ENCCL;
  memcpy(id128, &id, sizeof(id));
  return KB_OK;
}

extern "C" int kb_set_sharding(kb_handle h, int rank, int nranks, const void* id128) {
  if (!h) return KB_EINVAL;
  if (nranks < 1 || rank < 0 || rank >= nranks) return kb_fail(h, KB_EINVAL, "bad rank/nranks");
  if (!h->chain_set) return kb_fail(h, KB_EINVAL, "kb_set_chain must be called before kb_set_sharding");
  h->factored = false;
  if (nranks == 1) {
    h->rank = 0;
    h->nranks = 1;
    return KB_OK;
  }
  if (h->P < 2 * nranks)
    return kb_fail(h, KB_EINVAL, "chain of %lld nodes is too short for %d ranks (need >= 2 per rank)",
                   (long long)h->P, nranks);
  if (!id128) return kb_fail(h, KB_EINVAL, "nccl unique id missing");
  if (!nccl_load()) return kb_fail(h, KB_ENCCL, "cannot load libnccl.so.2");
  KB_CUDA(h, cudaSetDevice(h->device));
  kbi_nccl_destroy(h);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  KB_NCCL(h, g_nccl.CommInitRank(&comm, nranks, id, rank));
  h->nccl_comm = (void*)comm;
  h->rank = rank;
  h->nranks = nranks;
  // contiguous, near-equal node ranges (same rule as kore_b200/chain.py split_ranges)
  h->seg_lo.assign(nranks, 0);
  h->seg_hi.assign(nranks, 0);
  int64_t base = h->P / nranks, rem = h->P % nranks, lo = 0;
  for (int r = 0; r < nranks; ++r) {
    int64_t hi = lo + base + (r < rem ? 1 : 0);
    h->seg_lo[r] = lo;
    h->seg_hi[r] = hi;
    lo = hi;
  }
  h->top_sep = rank > 0 ? h->seg_lo[rank] - 1 : -1;
  h->bot_sep = rank < nranks - 1 ? h->seg_hi[rank] - 1 : -1;
  h->int_lo = h->seg_lo[rank];
  h->int_hi = rank < nranks - 1 ? h->seg_hi[rank] - 1 : h->seg_hi[rank];
  return KB_OK;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// Out[i, j] = sign * sum_e Dn[i, ridx[e] - orow] * T[pos[e]],  e over column (ocol + j) of a
// coupling block stored by columns.  Dn is (gridDim.x x kdim), Out is (gridDim.x x nc).
__global__ void kb_dense_spcols(const double2* __restrict__ Dn, int kdim, int orow, double2* __restrict__ Out,
                                int nc, int ocol, const int64_t* __restrict__ cptr,
                                const int* __restrict__ ridx, const int64_t* __restrict__ pos,
                                const double2* __restrict__ T, double sign) {
  extern __shared__ double2 drow[];
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < kdim; j += blockDim.x) drow[j] = Dn[(size_t)i * kdim + j];
  __syncthreads();
  for (int j = threadIdx.x; j < nc; j += blockDim.x) {
    int c = ocol + j;
    double2 acc = zmake(0.0, 0.0);
    for (int64_t e = cptr[c]; e < cptr[c + 1]; ++e) zfma(acc, drow[ridx[e] - orow], T[pos[e]]);
    Out[(size_t)i * nc + j] = zscale(acc, sign);
  }
}

// Out[i, :] (+)= sign * sum_{k in part(row o+i)} T[k] * Dn[col[k] - ocol, :]
// part 0: sub-diagonal (L) entries of the row, part 1: super-diagonal (U) entries.
__global__ void kb_sprows_dense(double2* __restrict__ Out, int nc, int o, int part,
                                const double2* __restrict__ Dn, int ocol,
                                const int64_t* __restrict__ rowptr, const int64_t* __restrict__ dstart,
                                const int64_t* __restrict__ ustart, const int* __restrict__ col,
                                const double2* __restrict__ T, double sign, int accumulate) {
  const int i = blockIdx.x;
  const int gi = o + i;
  const int64_t k0 = part == 0 ? rowptr[gi] : ustart[gi];
  const int64_t k1 = part == 0 ? dstart[gi] : rowptr[gi + 1];
  for (int j = threadIdx.x; j < nc; j += blockDim.x) {
    double2 acc = zmake(0.0, 0.0);
    for (int64_t k = k0; k < k1; ++k) zfma(acc, __ldg(&T[k]), Dn[(size_t)(__ldg(&col[k]) - ocol) * nc + j]);
    acc = zscale(acc, sign);
    if (accumulate) acc = zadd(acc, Out[(size_t)i * nc + j]);
    Out[(size_t)i * nc + j] = acc;
  }
}

// dense copy of a coupling block: Out (b x nc) = part(rows o..o+b) with columns offset by ocol
__global__ void kb_scatter_part(double2* __restrict__ Out, int nc, int o, int part, int ocol,
                                const int64_t* __restrict__ rowptr, const int64_t* __restrict__ dstart,
                                const int64_t* __restrict__ ustart, const int* __restrict__ col,
                                const double2* __restrict__ T) {
  const int i = blockIdx.x;
  const int gi = o + i;
  for (int j = threadIdx.x; j < nc; j += blockDim.x) Out[(size_t)i * nc + j] = zmake(0.0, 0.0);
  __syncthreads();
  const int64_t k0 = part == 0 ? rowptr[gi] : ustart[gi];
  const int64_t k1 = part == 0 ? dstart[gi] : rowptr[gi + 1];
  for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) Out[(size_t)i * nc + (col[k] - ocol)] = T[k];
}

// C (m x n) = alpha A (m x k) B (k x n) + beta C, row-major.  64 x 64 tile, 16-deep, 4 x 4 per thread.
__global__ void __launch_bounds__(256)
kb_zgemm(int m, int n, int k, double2 alpha, const double2* __restrict__ A, int lda,
         const double2* __restrict__ B, int ldb, double2 beta, double2* __restrict__ C, int ldc) {
  __shared__ double2 As[16][64 + 1];
  __shared__ double2 Bs[16][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64;
  double2 acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = zmake(0.0, 0.0);
  for (int kk = 0; kk < k; kk += 16) {
    for (int e = tid; e < 64 * 16; e += 256) {
      int i = e / 16, q = e % 16;  // A tile: 64 rows x 16 k
      int gi = row0 + i, gk = kk + q;
      As[q][i] = (gi < m && gk < k) ? A[(size_t)gi * lda + gk] : zmake(0.0, 0.0);
    }
    for (int e = tid; e < 16 * 64; e += 256) {
      int q = e / 64, j = e % 64;  // B tile: 16 k x 64 cols
      int gk = kk + q, gj = col0 + j;
      Bs[q][j] = (gk < k && gj < n) ? B[(size_t)gk * ldb + gj] : zmake(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      double2 av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[q][ty + 16 * a];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = Bs[q][tx + 16 * c];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) zfma(acc[a][c], av[a], bv[c]);
    }
    __syncthreads();
  }
  const bool has_beta = beta.x != 0.0 || beta.y != 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int gi = row0 + ty + 16 * a;
    if (gi >= m) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int gj = col0 + tx + 16 * c;
      if (gj >= n) continue;
      double2 v = zmul(alpha, acc[a][c]);
      if (has_beta) v = zadd(v, zmul(beta, C[(size_t)gi * ldc + gj]));
      C[(size_t)gi * ldc + gj] = v;
    }
  }
}

// y (op)= Mat (m x k, ld) x      mode 0: y = Mx, 1: y -= Mx, 2: y += Mx   (one warp per row)
__global__ void kb_dense_gemv(const double2* __restrict__ Mat, int m, int k, int ld,
                              const double2* __restrict__ x, double2* __restrict__ y, int mode) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= m) return;
  const double2* Mr = Mat + (size_t)row * ld;
  double2 acc = zmake(0.0, 0.0);
  for (int j = lane; j < k; j += 32) zfma(acc, Mr[j], x[j]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
  }
  if (lane == 0) {
    if (mode == 0)
      y[row] = acc;
    else if (mode == 1)
      y[row] = zsub(y[row], acc);
    else
      y[row] = zadd(y[row], acc);
  }
}

__global__ void kb_sub(int64_t n, const double2* __restrict__ a, const double2* __restrict__ b,
                       double2* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = zsub(a[i], b[i]);
}

// first Gauss-Jordan panel of an n x n matrix, column-major, for kb_gj_panel
__global__ void kb_first_panel(const double2* __restrict__ S, int n, int nb0, double2* __restrict__ PT) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int j = 0; j < nb0 && j < n; ++j) PT[(size_t)j * n + i] = S[(size_t)i * n + j];
}

static void zgemm(kb_context* h, int m, int n, int k, double alpha, const double2* A, int lda, const double2* B,
                  int ldb, double beta, double2* C, int ldc) {
  dim3 grid((n + 63) / 64, (m + 63) / 64);
  kb_zgemm<<<grid, 256, 0, h->stream>>>(m, n, k, zmake(alpha, 0.0), A, lda, B, ldb, zmake(beta, 0.0), C, ldc);
  h->launches++;
}

static inline int nsize(const kb_context* h, int64_t p) { return (int)(h->nodeptr[p + 1] - h->nodeptr[p]); }
static inline int noff(const kb_context* h, int64_t p) { return (int)h->nodeptr[p]; }

// ---------------------------------------------------------------------------
// factor
// ---------------------------------------------------------------------------
int kbi_factor_sharded(kb_context* h, zcomplex sigma) {
  if (!h->chain_set) return kb_fail(h, KB_EINVAL, "kb_set_chain must be called before kb_factor");
  if (!h->nccl_comm) return kb_fail(h, KB_EINVAL, "kb_set_sharding must be called before kb_factor");
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int64_t P = h->P, bmax = h->bmax;
  const int G = h->nranks, g = h->rank;
  h->factored = false;
  h->sigma = sigma;
  kbi_drop_graphs(h);

  cudaEvent_t e0, e1;
  KB_CUDA(h, cudaEventCreate(&e0));
  KB_CUDA(h, cudaEventCreate(&e1));
  KB_CUDA(h, cudaEventRecord(e0, s));
  KB_TRY(kbi_build_T(h, sigma));
  KB_TRY(kbi_factor_workspace(h));

  const int64_t lo = h->int_lo, hi = h->int_hi;
  const bool has_top = h->top_sep >= 0, has_bot = h->bot_sep >= 0;
  const int bt = has_top ? nsize(h, h->top_sep) : 0, ot = has_top ? noff(h, h->top_sep) : 0;

  // storage: inverses of the interior Schur blocks, spikes
  h->Moff.assign(P + 1, 0);
  h->Voff.assign(P + 1, 0);
  int64_t mtot = 0, vtot = 0;
  for (int64_t p = lo; p < hi; ++p) {
    int64_t b = nsize(h, p);
    h->Moff[p] = mtot;
    mtot += b * b;
    h->Voff[p] = vtot;
    vtot += b * (int64_t)bt;
  }
  if (h->d_M.alloc((size_t)mtot) != cudaSuccess ||
      (has_top && (h->d_Vsp.alloc((size_t)vtot) != cudaSuccess || h->d_Gsp.alloc((size_t)vtot) != cudaSuccess)))
    return kb_fail(h, KB_ENOMEM, "cannot allocate %.2f GB for this rank's chain factors",
                   (mtot + 2 * vtot) * 16.0 / 1e9);
  KB_CUDA(h, h->d_F.alloc((size_t)bmax * bmax));
  KB_CUDA(h, h->d_H.alloc((size_t)bmax * bmax));
  KB_CUDA(h, h->d_Acc.alloc((size_t)bmax * bmax));
  const size_t slot = (size_t)bmax * bmax;
  KB_CUDA(h, h->d_contrib.alloc(4 * slot));
  KB_CUDA(h, h->d_contrib_all.alloc(4 * slot * G));
  KB_CUDA(h, cudaMemsetAsync(h->d_contrib.p, 0, 4 * slot * sizeof(double2), s));

  double flops = 0.0;
  for (int64_t p = lo; p < hi; ++p) {
    const int o = noff(h, p), b = nsize(h, p);
    const int oprev = p > lo ? noff(h, p - 1) : 0;
    kb_schur_row<<<b, 128, 0, s>>>(h->d_S0.p, h->d_PT.p, kbi_panel_width(h, b), b, o,
                                   p > lo ? h->d_W.p : nullptr, oprev, nullptr, 0, h->d_rowptr.p, h->d_dstart.p,
                                   h->d_ustart.p, h->d_col.p, h->d_Tval.p);
    h->launches++;
    double2* X = nullptr;
    KB_TRY(gj_invert(h, kbi_ws_main(h), b, &X));
    double2* Mp = h->d_M.p + h->Moff[p];
    kb_store_inverse<<<b, 128, 0, s>>>(X, b, h->d_orig.p, Mp);
    h->launches++;
    flops += 8.0 * (double)b * b * b;
    if (has_top) {
      double2* Vp = h->d_Vsp.p + h->Voff[p];
      double2* Gp = h->d_Gsp.p + h->Voff[p];
      if (p == lo) {
        // V = M_p L_{p,t};  G = U_{t,p} (dense copy);  H = G M_p;  Acc = G V
        kb_dense_spcols<<<b, 128, b * sizeof(double2), s>>>(Mp, b, o, Vp, bt, ot, h->d_lcptr.p, h->d_lrow.p,
                                                            h->d_lpos.p, h->d_Tval.p, 1.0);
        kb_scatter_part<<<bt, 128, 0, s>>>(Gp, b, ot, 1, o, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                           h->d_col.p, h->d_Tval.p);
        kb_sprows_dense<<<bt, 128, 0, s>>>(h->d_H.p, b, ot, 1, Mp, o, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                           h->d_col.p, h->d_Tval.p, 1.0, 0);
        kb_sprows_dense<<<bt, 128, 0, s>>>(h->d_Acc.p, bt, ot, 1, Vp, o, h->d_rowptr.p, h->d_dstart.p,
                                           h->d_ustart.p, h->d_col.p, h->d_Tval.p, 1.0, 0);
        h->launches += 4;
      } else {
        const int bp = nsize(h, p - 1);
        const double2* Vprev = h->d_Vsp.p + h->Voff[p - 1];
        // F = -L_{p,p-1} V_{p-1};  V_p = M_p F
        kb_sprows_dense<<<b, 128, 0, s>>>(h->d_F.p, bt, o, 0, Vprev, oprev, h->d_rowptr.p, h->d_dstart.p,
                                          h->d_ustart.p, h->d_col.p, h->d_Tval.p, -1.0, 0);
        zgemm(h, b, bt, b, 1.0, Mp, b, h->d_F.p, bt, 0.0, Vp, bt);
        // G_p = -H_{p-1} U_{p-1,p};  H_p = G_p M_p;  Acc += G_p V_p
        kb_dense_spcols<<<bt, 128, bp * sizeof(double2), s>>>(h->d_H.p, bp, oprev, Gp, b, o, h->d_ucptr.p,
                                                              h->d_urow.p, h->d_upos.p, h->d_Tval.p, -1.0);
        zgemm(h, bt, b, b, 1.0, Gp, b, Mp, b, 0.0, h->d_H.p, b);
        zgemm(h, bt, bt, b, 1.0, Gp, b, Vp, bt, 1.0, h->d_Acc.p, bt);
        h->launches += 2;
        flops += 8.0 * ((double)b * bt * b + (double)bt * b * b + (double)bt * bt * b);
      }
    }
    if (p + 1 < hi) {
      const int onext = noff(h, p + 1), bnext = nsize(h, p + 1);
      kb_w_rows<<<b, 128, b * sizeof(double2), s>>>(Mp, b, o, h->d_W.p, bnext, onext, h->d_ucptr.p, h->d_urow.p,
                                                    h->d_upos.p, h->d_Tval.p);
      h->launches++;
    }
    KB_LAUNCH_CHECK(h);
  }
  // ---- contributions to the reduced (separator) system
  //   slot 0: D_bot - L_{bot,e} M_e U_{e,bot}   slot 1: Acc_top
  //   slot 2: block (bot, top) = -L_{bot,e} V_e  slot 3: block (top, bot) = -H_e U_{e,bot}
  {
    const int64_t e = hi - 1;
    const int oe = noff(h, e), be = nsize(h, e);
    const double2* Me = h->d_M.p + h->Moff[e];
    if (has_bot) {
      const int ob = noff(h, h->bot_sep), bb = nsize(h, h->bot_sep);
      kb_w_rows<<<be, 128, be * sizeof(double2), s>>>(Me, be, oe, h->d_W.p, bb, ob, h->d_ucptr.p, h->d_urow.p,
                                                      h->d_upos.p, h->d_Tval.p);
      kb_schur_row<<<bb, 128, 0, s>>>(h->d_contrib.p, h->d_PT.p, 0, bb, ob, h->d_W.p, oe, nullptr, 0,
                                      h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p, h->d_col.p, h->d_Tval.p);
      h->launches += 2;
      if (has_top) {
        const double2* Ve = h->d_Vsp.p + h->Voff[e];
        kb_sprows_dense<<<bb, 128, 0, s>>>(h->d_contrib.p + 2 * slot, bt, ob, 0, Ve, oe, h->d_rowptr.p,
                                           h->d_dstart.p, h->d_ustart.p, h->d_col.p, h->d_Tval.p, -1.0, 0);
        kb_dense_spcols<<<bt, 128, be * sizeof(double2), s>>>(h->d_H.p, be, oe, h->d_contrib.p + 3 * slot, bb, ob,
                                                              h->d_ucptr.p, h->d_urow.p, h->d_upos.p, h->d_Tval.p,
                                                              -1.0);
        h->launches += 2;
      }
    }
    if (has_top)
      KB_CUDA(h, cudaMemcpyAsync(h->d_contrib.p + slot, h->d_Acc.p, (size_t)bt * bt * sizeof(double2),
                                 cudaMemcpyDeviceToDevice, s));
    KB_LAUNCH_CHECK(h);
  }
  KB_NCCL(h, g_nccl.AllGather(h->d_contrib.p, h->d_contrib_all.p, 4 * slot * 2, ncclDouble,
                              (ncclComm_t)h->nccl_comm, s));

  // ---- reduced system, factored redundantly on every rank: node j = separator of rank j
  KB_CUDA(h, h->d_Mr.alloc(slot * (G - 1)));
  for (int j = 0; j < G - 1; ++j) {
    const int bs = nsize(h, h->seg_hi[j] - 1);
    const double2* Rabove = h->d_contrib_all.p + (size_t)j * 4 * slot;
    const double2* Acc = h->d_contrib_all.p + (size_t)(j + 1) * 4 * slot + slot;
    kb_sub<<<nblk((int64_t)bs * bs, 256), 256, 0, s>>>((int64_t)bs * bs, Rabove, Acc, h->d_S0.p);
    h->launches++;
    if (j > 0) {
      const int bsp = nsize(h, h->seg_hi[j - 1] - 1);
      const double2* Csub = h->d_contrib_all.p + (size_t)j * 4 * slot + 2 * slot;  // (bs x bsp)
      const double2* Csup = h->d_contrib_all.p + (size_t)j * 4 * slot + 3 * slot;  // (bsp x bs)
      zgemm(h, bsp, bs, bsp, 1.0, h->d_Mr.p + (size_t)(j - 1) * slot, bsp, Csup, bs, 0.0, h->d_F.p, bs);
      zgemm(h, bs, bs, bsp, -1.0, Csub, bsp, h->d_F.p, bs, 1.0, h->d_S0.p, bs);
    }
    kb_first_panel<<<nblk(bs, 128), 128, 0, s>>>(h->d_S0.p, bs, kbi_panel_width(h, bs), h->d_PT.p);
    double2* X = nullptr;
    KB_TRY(gj_invert(h, kbi_ws_main(h), bs, &X));
    kb_store_inverse<<<bs, 128, 0, s>>>(X, bs, h->d_orig.p, h->d_Mr.p + (size_t)j * slot);
    h->launches += 2;
    flops += 8.0 * (double)bs * bs * bs;
    KB_LAUNCH_CHECK(h);
  }
  KB_CUDA(h, h->d_sepvec.alloc(2 * bmax));
  KB_CUDA(h, h->d_sepvec_all.alloc((size_t)2 * bmax * G));
  KB_CUDA(h, h->d_redz.alloc((size_t)2 * bmax * G));

  KB_CUDA(h, cudaEventRecord(e1, s));
  int info = 0;
  KB_CUDA(h, cudaMemcpyAsync(&info, h->d_info.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaStreamSynchronize(s));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  h->stats.factor_ms = ms;
  h->stats.factor_flops = flops;
  h->stats.factor_bytes = (mtot + 2 * vtot + (int64_t)slot * (G - 1)) * 16;
  if (info != 0)
    return kb_fail(h, KB_ESINGULAR, "zero or non-finite pivot in a Schur block on rank %d", g);
  h->factored = true;
  return KB_OK;
}

// ---------------------------------------------------------------------------
// solve: y <- T'^{-1} r, full vectors (chain order, scaled space) on every rank
// ---------------------------------------------------------------------------
int kbi_sharded_sweeps(kb_context* h, const double2* r, double2* y) {
  cudaStream_t s = h->stream;
  const int G = h->nranks, g = h->rank;
  const int64_t lo = h->int_lo, hi = h->int_hi, bmax = h->bmax, P = h->P;
  const bool has_top = h->top_sep >= 0, has_bot = h->bot_sep >= 0;
  const int bt = has_top ? nsize(h, h->top_sep) : 0, ot = has_top ? noff(h, h->top_sep) : 0;
  const size_t slot = (size_t)bmax * bmax;
  double2* t = h->d_t.p;
  double2* sv = h->d_sepvec.p;  // [0,bmax): a_top = sum_p G_p y_p ; [bmax,2bmax): r_bot - L_{bot,e} y_e
  KB_CUDA(h, cudaMemsetAsync(sv, 0, 2 * bmax * sizeof(double2), s));

  // ---- forward over the interior
  for (int64_t p = lo; p < hi; ++p) {
    const int o = noff(h, p), b = nsize(h, p);
    if (p == lo && has_top) {
      KB_CUDA(h, cudaMemcpyAsync(t + o, r + o, (size_t)b * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    } else {
      kb_node_tvec<0><<<(b + 7) / 8, 256, 0, s>>>(b, o, r, y, t, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                                  h->d_col.p, h->d_Tval.p);
    }
    kb_node_gemv<0><<<(b + KB_NODE_ROWS - 1) / KB_NODE_ROWS, 256, b * sizeof(double2), s>>>(
        h->d_M.p + h->Moff[p], b, o, t, y);
    if (has_top)
      kb_dense_gemv<<<(bt + 7) / 8, 256, 0, s>>>(h->d_Gsp.p + h->Voff[p], bt, b, b, y + o, sv, 2);
    h->launches += 3;
  }
  if (has_bot) {
    const int ob = noff(h, h->bot_sep), bb = nsize(h, h->bot_sep);
    kb_node_tvec<0><<<(bb + 7) / 8, 256, 0, s>>>(bb, ob, r, y, t, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                                 h->d_col.p, h->d_Tval.p);
    KB_CUDA(h, cudaMemcpyAsync(sv + bmax, t + ob, (size_t)bb * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    h->launches++;
  }
  KB_LAUNCH_CHECK(h);
  KB_NCCL(h, g_nccl.AllGather(sv, h->d_sepvec_all.p, 2 * bmax * 2, ncclDouble, (ncclComm_t)h->nccl_comm, s));

  // ---- reduced solve (redundant): z_j = Mr_j (rho_j - Csub_j z_{j-1});  x_j = z_j - Mr_j Csup_{j+1} x_{j+1}
  double2* z = h->d_redz.p;             // G-1 vectors of bmax
  double2* tmp = h->d_redz.p + (size_t)bmax * G;
  for (int j = 0; j < G - 1; ++j) {
    const int bs = nsize(h, h->seg_hi[j] - 1);
    const double2* tb = h->d_sepvec_all.p + (size_t)j * 2 * bmax + bmax;     // from rank j
    const double2* at = h->d_sepvec_all.p + (size_t)(j + 1) * 2 * bmax;      // from rank j+1
    kb_sub<<<nblk(bs, 256), 256, 0, s>>>(bs, tb, at, tmp);
    if (j > 0) {
      const int bsp = nsize(h, h->seg_hi[j - 1] - 1);
      const double2* Csub = h->d_contrib_all.p + (size_t)j * 4 * slot + 2 * slot;
      kb_dense_gemv<<<(bs + 7) / 8, 256, 0, s>>>(Csub, bs, bsp, bsp, z + (size_t)(j - 1) * bmax, tmp, 1);
    }
    kb_dense_gemv<<<(bs + 7) / 8, 256, 0, s>>>(h->d_Mr.p + (size_t)j * slot, bs, bs, bs, tmp, z + (size_t)j * bmax, 0);
    h->launches += 3;
  }
  for (int j = G - 3; j >= 0; --j) {
    const int bs = nsize(h, h->seg_hi[j] - 1), bsn = nsize(h, h->seg_hi[j + 1] - 1);
    const double2* Csup = h->d_contrib_all.p + (size_t)(j + 1) * 4 * slot + 3 * slot;  // (bs x bsn)
    kb_dense_gemv<<<(bs + 7) / 8, 256, 0, s>>>(Csup, bs, bsn, bsn, z + (size_t)(j + 1) * bmax, tmp, 0);
    kb_dense_gemv<<<(bs + 7) / 8, 256, 0, s>>>(h->d_Mr.p + (size_t)j * slot, bs, bs, bs, tmp, z + (size_t)j * bmax, 1);
    h->launches += 2;
  }
  for (int j = 0; j < G - 1; ++j) {
    const int64_t sp = h->seg_hi[j] - 1;
    KB_CUDA(h, cudaMemcpyAsync(y + noff(h, sp), z + (size_t)j * bmax, (size_t)nsize(h, sp) * sizeof(double2),
                               cudaMemcpyDeviceToDevice, s));
  }

  // ---- backward over the interior, with the spike correction
  for (int64_t p = hi - 1; p >= lo; --p) {
    const int o = noff(h, p), b = nsize(h, p);
    if (p + 1 < P) {
      kb_node_tvec<1><<<(b + 7) / 8, 256, 0, s>>>(b, o, r, y, t, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                                  h->d_col.p, h->d_Tval.p);
      kb_node_gemv<1><<<(b + KB_NODE_ROWS - 1) / KB_NODE_ROWS, 256, b * sizeof(double2), s>>>(
          h->d_M.p + h->Moff[p], b, o, t, y);
      h->launches += 2;
    }
    if (has_top) {
      kb_dense_gemv<<<(b + 7) / 8, 256, 0, s>>>(h->d_Vsp.p + h->Voff[p], b, bt, bt, y + ot, y + o, 1);
      h->launches++;
    }
  }
  KB_LAUNCH_CHECK(h);

  // ---- every rank publishes its interior; separators are already everywhere
  KB_NCCL(h, g_nccl.GroupStart());
  for (int q = 0; q < G; ++q) {
    int64_t qlo = h->seg_lo[q], qhi = q < G - 1 ? h->seg_hi[q] - 1 : h->seg_hi[q];
    int64_t off = h->nodeptr[qlo], cnt = h->nodeptr[qhi] - off;
    KB_NCCL(h, g_nccl.Broadcast(y + off, y + off, (size_t)cnt * 2, ncclDouble, q, (ncclComm_t)h->nccl_comm, s));
  }
  KB_NCCL(h, g_nccl.GroupEnd());
  (void)g;
  return KB_OK;
}

int kbi_chain_solve_sharded(kb_context* h, const double2* r, double2* x, int refine) {
  (void)refine;
  return kbi_sharded_sweeps(h, r, x);
}
