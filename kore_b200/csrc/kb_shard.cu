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
#include <stdio.h>
#include <stdlib.h>
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
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
  KB_SYM(AllReduce, "ncclAllReduce");
  KB_SYM(Send, "ncclSend");
  KB_SYM(Recv, "ncclRecv");
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

static void mailbox_close(kb_context* h);

void kbi_nccl_destroy(kb_context* h) {
  mailbox_close(h);
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

static int shard_ranges(kb_context* h);
static int mailbox_setup(kb_context* h);
static void mailbox_close(kb_context* h);

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
  KB_TRY(mailbox_setup(h));
  return shard_ranges(h);
}

// Map every rank's mailbox buffer into every other rank (CUDA IPC; the handles travel through one
// ncclAllGather).  Any failure leaves mb_ready == false on ALL ranks (they agree through an
// all-reduce) and the small collectives stay on NCCL.
static void mailbox_close(kb_context* h) {
  for (int q = 0; q < KB_MB_MAXRANKS; ++q) {
    if (h->mb_peer[q] && q != h->rank) cudaIpcCloseMemHandle(h->mb_peer[q]);
    h->mb_peer[q] = nullptr;
  }
  h->mb_ready = false;
}

static int mailbox_setup(kb_context* h) {
  mailbox_close(h);
  const int G = h->nranks, g = h->rank;
  cudaStream_t s = h->stream;
  int ok = (G <= KB_MB_MAXRANKS && !getenv("KB_SHARD_NCCL_ONLY")) ? 1 : 0;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    ok = h->d_mbox.alloc(KB_MB_BYTES) == cudaSuccess && cudaMemsetAsync(h->d_mbox.p, 0, KB_MB_BYTES, s) == cudaSuccess &&
         cudaStreamSynchronize(s) == cudaSuccess && cudaIpcGetMemHandle(&mine, h->d_mbox.p) == cudaSuccess;
    if (!ok) cudaGetLastError();
  }
  DevBuf<unsigned char> dh;
  DevBuf<int> dflag;
  KB_CUDA(h, dh.alloc(sizeof(mine) * (size_t)G));
  KB_CUDA(h, dflag.alloc(1));
  KB_CUDA(h, cudaMemcpyAsync(dh.p + sizeof(mine) * (size_t)g, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
  KB_NCCL(h, g_nccl.AllGather(dh.p + sizeof(mine) * (size_t)g, dh.p, sizeof(mine), ncclChar, (ncclComm_t)h->nccl_comm, s));
  std::vector<cudaIpcMemHandle_t> all(G);
  KB_CUDA(h, cudaMemcpyAsync(all.data(), dh.p, sizeof(mine) * (size_t)G, cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaStreamSynchronize(s));
  if (ok) {
    h->mb_peer[g] = h->d_mbox.p;
    for (int q = 0; q < G && ok; ++q) {
      if (q == g) continue;
      void* pp = nullptr;
      if (cudaIpcOpenMemHandle(&pp, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      } else {
        h->mb_peer[q] = pp;
      }
    }
  }
  // agree: the mailboxes are used only if every rank mapped every peer
  KB_CUDA(h, cudaMemcpyAsync(dflag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, s));
  KB_NCCL(h, g_nccl.AllReduce(dflag.p, dflag.p, 1, ncclInt, ncclMin, (ncclComm_t)h->nccl_comm, s));
  int all_ok = 0;
  KB_CUDA(h, cudaMemcpyAsync(&all_ok, dflag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaStreamSynchronize(s));
  if (!all_ok) {
    mailbox_close(h);
    return KB_OK;
  }
  h->mb_ready = true;
  h->mb_epoch = 0;
  return KB_OK;
}

static KbMailbox mailbox_of(const kb_context* h) {
  KbMailbox m;
  m.rank = h->rank;
  m.nranks = h->nranks;
  for (int q = 0; q < KB_MB_MAXRANKS; ++q) m.base[q] = (unsigned char*)h->mb_peer[q];
  return m;
}

// contiguous, near-equal node ranges (same rule as kore_b200/chain.py split_ranges); derived again at
// every factorisation, so the chain may change under an established communicator
static int shard_ranges(kb_context* h) {
  const int nranks = h->nranks, rank = h->rank;
  if (h->P < 2 * nranks)
    return kb_fail(h, KB_EINVAL, "chain of %lld nodes is too short for %d ranks (need >= 2 per rank)",
                   (long long)h->P, nranks);
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

// ---------------------------------------------------------------------------
// Batched dense products of the fast l-sharded factorisation (below): up to four independent
// C = beta C + alpha op(A) B, one problem per blockIdx.z, row-major.  32 x 64 tiles (190 per
// 600 x 600 product: two to four products spread evenly over 148 SMs x 4 resident CTAs), 8 deep, two
// shared-memory stages filled by 16-byte cp.async copies (zero-filled outside the matrices) so that
// a k-tile costs one wait and two barriers.  op(A) = A^T reads A (k x m, row-major) by columns -- the explicit inverses
// are stored transposed.  FP64 FMA issue is the bound: 64 DFMA per thread and k-step against 8
// 16-byte shared-memory loads, two of which are warp broadcasts.
// ---------------------------------------------------------------------------
struct KbGemmProb {
  int m, n, k, transA;
  const double2* A;
  int lda;
  const double2* B;
  int ldb;
  double2* C;
  int ldc;
  double alpha, beta;
};
struct KbGemmBatch {
  KbGemmProb p[4];
};

#define ZG_BM 32
#define ZG_BN 64
#define ZG_BK 8
#define ZG_THREADS 128
#define ZG_PA (ZG_BK + 4)   // pitch of an A-tile row (complex elements): = 4 (mod 8), conflict-free 16-byte fragment loads
#define ZG_PB (ZG_BN + 2)   // pitch of a B-tile row: = 2 (mod 8)

// 16-byte asynchronous global -> shared copy; ok == false writes zeros (src-size 0)
__device__ __forceinline__ void zg_cp16(void* dst, const void* src, bool ok) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(kb_smem_addr(dst)), "l"(src), "r"(ok ? 16 : 0)
               : "memory");
}
// FP64 tensor-core product D(8x8) += A(8x4) B(4x8): lane = 4 g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1]
__device__ __forceinline__ void zg_dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

// The products are contraction-bound (ncu on the SIMT version of this kernel: FP64 pipe 62.5 % active,
// math-pipe throttle the first stall reason; a DFMA with three distinct register operands issues
// every ~3 cycles, not 2), so they run on the FP64 tensor cores: DMMA m8n8k4 measures 37.1 TFLOP/s on
// this part, the same peak as the DFMA pipe, at an eighth of the instructions
// (profiles/r2p_microbench_dmma_vs_dfma_rate.txt).  A complex product is four real ones:
//   Cre += Are Bre + (-Aim) Bim,   Cim += Are Bim + Aim Bre.
// CTA = 4 warps side by side (one per SM sub-partition, so that the last CTAs of a launch still use
// the whole SM), each a 32 x 16 complex tile = 4 x 2 DMMA blocks: 32 DMMAs and 6 fragment loads
// (16 bytes: re and im together) per k-step of 4.
__global__ void __launch_bounds__(ZG_THREADS, 4) kb_zgemm_batch(KbGemmBatch batch) {
  const KbGemmProb& g = batch.p[blockIdx.z];
  const int m = g.m, n = g.n, k = g.k;
  const int row0 = blockIdx.y * ZG_BM, col0 = blockIdx.x * ZG_BN;
  if (row0 >= m || col0 >= n) return;
  __shared__ __align__(16) double2 As[2][ZG_BM][ZG_PA];  // [row][k]
  __shared__ __align__(16) double2 Bs[2][ZG_BK][ZG_PB];  // [k][col]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const double2* __restrict__ A = g.A;
  const double2* __restrict__ B = g.B;
  const int lda = g.lda, ldb = g.ldb, transA = g.transA;
  auto issue = [&](int kk, int st) {
#pragma unroll
    for (int u = 0; u < (ZG_BM * ZG_BK) / ZG_THREADS; ++u) {
      const int e = tid + ZG_THREADS * u;
      int i, q;
      if (transA) {  // A is k x m: consecutive threads along i
        q = e / ZG_BM;
        i = e % ZG_BM;
      } else {       // A is m x k: consecutive threads along k
        i = e / ZG_BK;
        q = e % ZG_BK;
      }
      const int gi = row0 + i, gk = kk + q;
      const bool ok = gi < m && gk < k;
      const double2* src = ok ? (transA ? A + (size_t)gk * lda + gi : A + (size_t)gi * lda + gk) : A;
      zg_cp16(&As[st][i][q], src, ok);
    }
#pragma unroll
    for (int u = 0; u < (ZG_BK * ZG_BN) / ZG_THREADS; ++u) {
      const int e = tid + ZG_THREADS * u;
      const int q = e / ZG_BN, j = e % ZG_BN;
      const int gk = kk + q, gj = col0 + j;
      const bool ok = gk < k && gj < n;
      zg_cp16(&Bs[st][q][j], ok ? B + (size_t)gk * ldb + gj : B, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double cre[4][2][2], cim[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) cre[a][c][0] = cre[a][c][1] = cim[a][c][0] = cim[a][c][1] = 0.0;
  issue(0, 0);
  int st = 0;
  for (int kk = 0; kk < k; kk += ZG_BK) {
    const bool more = kk + ZG_BK < k;
    if (more) issue(kk + ZG_BK, st ^ 1);  // the other stage was released by the barrier that ended the last trip
    if (more)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    else
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int kq = 0; kq < ZG_BK / 4; ++kq) {
      double are[4], aim[4], nim[4], bre[2], bim[2];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double2 v = As[st][8 * a + gq][4 * kq + tq];
        are[a] = v.x;
        aim[a] = v.y;
        nim[a] = -v.y;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const double2 v = Bs[st][4 * kq + tq][16 * wid + 8 * c + gq];
        bre[c] = v.x;
        bim[c] = v.y;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          zg_dmma(cre[a][c], are[a], bre[c]);
          zg_dmma(cim[a][c], are[a], bim[c]);
          zg_dmma(cre[a][c], nim[a], bim[c]);
          zg_dmma(cim[a][c], aim[a], bre[c]);
        }
    }
    __syncthreads();
    st ^= 1;
  }
  double2* __restrict__ C = g.C;
  const int ldc = g.ldc;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = row0 + 8 * a + gq;
    if (gi >= m) continue;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int gj = col0 + 16 * wid + 8 * c + 2 * tq + u;
        if (gj >= n) continue;
        double2 v = zmake(cre[a][c][u] * g.alpha, cim[a][c][u] * g.alpha);
        if (g.beta != 0.0) v = zadd(v, zscale(C[(size_t)gi * ldc + gj], g.beta));
        C[(size_t)gi * ldc + gj] = v;
      }
    }
  }
}

struct GemmQueue {
  KbGemmBatch b;
  int count = 0, mmax = 0, nmax = 0;
  void add(int m, int n, int k, double alpha, const double2* A, int lda, bool transA, const double2* B, int ldb,
           double beta, double2* C, int ldc) {
    if (m <= 0 || n <= 0) return;
    KbGemmProb& g = b.p[count++];
    g.m = m;
    g.n = n;
    g.k = k;
    g.transA = transA ? 1 : 0;
    g.A = A;
    g.lda = lda;
    g.B = B;
    g.ldb = ldb;
    g.C = C;
    g.ldc = ldc;
    g.alpha = alpha;
    g.beta = beta;
    if (m > mmax) mmax = m;
    if (n > nmax) nmax = n;
  }
  void flush(kb_context* h) {
    if (!count) return;
    dim3 grid((nmax + ZG_BN - 1) / ZG_BN, (mmax + ZG_BM - 1) / ZG_BM, count);
    kb_zgemm_batch<<<grid, ZG_THREADS, 0, h->stream>>>(b);
    h->launches++;
    count = mmax = nmax = 0;
  }
};

// ---------------------------------------------------------------------------
// Reduced (separator) solve in ONE cooperative launch.  With E_j = Mr_j Csub_j and
// Fm_j = Mr_j Csup_{j+1} formed at factor time the block-Thomas recurrences on the G-1 separators are
//     z_j = Mr_j rho_j - E_j z_{j-1}   (j = 0 .. G-2),      rho_j = tb_j - at_{j+1} (gathered contributions)
//     x_j = z_j - Fm_j x_{j+1}         (j = G-3 .. 0),      x_{G-2} = z_{G-2}
// 2 (G-1) - 1 dependent dense matrix-vector steps: one warp per row, a grid barrier per step (the
// per-step kernels this replaces cost 0.35 ms per solve at G = 8, a third of the l-sharded solve).
// ---------------------------------------------------------------------------
#define KB_MAXSEP 16
struct KbRedParams {
  int nsep, bmax;
  int bs[KB_MAXSEP];
  const double2* Mr[KB_MAXSEP];
  const double2* E[KB_MAXSEP];   // bs_j x bs_{j-1}; unused for j = 0
  const double2* Fm[KB_MAXSEP];  // bs_j x bs_{j+1}; unused for j = nsep-1
  double2* xout[KB_MAXSEP];      // where x_j goes (the solution vector at separator j)
  const double2* sepvec_all;     // per rank `sepstride` entries apart: [at | tb] (2 bmax entries)
  size_t sepstride;
  // mailbox mode: CTA 0 first sends this rank's contribution `sv` to every peer, every CTA waits for
  // all of them, and sepvec_all is the own mailbox (the all-gather is part of this launch)
  int use_mb;
  KbMailbox mb;
  unsigned mb_epoch;
  const double2* sv;
  double2* z;                    // nsep x bmax
  unsigned* ctr;
  unsigned epoch0;
  int* err;
  unsigned long long wait_ns;
};

__device__ __forceinline__ void kb_red_sync(unsigned* ctr, unsigned epoch, int* err, unsigned long long wait_ns) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned target = epoch * gridDim.x;
    unsigned v;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    KbSpin sp;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if ((int)(v - target) >= 0) break;
      if (kb_spin_expired(sp, err, KB_WERR_SWEEP, wait_ns)) break;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) kb_reduced_solve_kernel(KbRedParams q) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
  unsigned epoch = q.epoch0;
  const size_t bm = (size_t)q.bmax;
  if (q.use_mb) {
    if (blockIdx.x == 0) {
      const int par = (int)(q.mb_epoch & 1u);
      for (int r = 0; r < q.mb.nranks; ++r) {
        double2* dst = kb_mb_data(q.mb, r, par, q.mb.rank);
        for (int i = threadIdx.x; i < 2 * q.bmax; i += blockDim.x) dst[i] = q.sv[i];
      }
      kb_mb_signal_all(q.mb, q.mb_epoch);
    }
    kb_mb_wait_all(q.mb, q.mb_epoch, q.err, q.wait_ns);
  }
  for (int j = 0; j < q.nsep; ++j) {
    const int bs = q.bs[j], bsp = j > 0 ? q.bs[j - 1] : 0;
    const double2* tb = q.sepvec_all + (size_t)j * q.sepstride + bm;
    const double2* at = q.sepvec_all + (size_t)(j + 1) * q.sepstride;
    const double2* zp = q.z + (size_t)(j > 0 ? j - 1 : 0) * bm;
    for (int row = gw; row < bs; row += nw) {
      double2 acc = zmake(0.0, 0.0);
      const double2* Mrow = q.Mr[j] + (size_t)row * bs;
#pragma unroll 4
      for (int c = lane; c < bs; c += 32) zfma(acc, Mrow[c], zsub(__ldcg(&tb[c]), __ldcg(&at[c])));
      if (j > 0) {
        const double2* Erow = q.E[j] + (size_t)row * bsp;
#pragma unroll 4
        for (int c = lane; c < bsp; c += 32) zfms(acc, Erow[c], __ldcg(&zp[c]));
      }
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
      }
      if (lane == 0) {
        q.z[(size_t)j * bm + row] = acc;
        if (j == q.nsep - 1) q.xout[j][row] = acc;
      }
    }
    if (q.nsep > 1) kb_red_sync(q.ctr, ++epoch, q.err, q.wait_ns);
  }
  for (int j = q.nsep - 2; j >= 0; --j) {
    const int bs = q.bs[j], bsn = q.bs[j + 1];
    const double2* xn = q.xout[j + 1];
    for (int row = gw; row < bs; row += nw) {
      double2 acc = zmake(0.0, 0.0);
      const double2* Frow = q.Fm[j] + (size_t)row * bsn;
#pragma unroll 4
      for (int c = lane; c < bsn; c += 32) zfma(acc, Frow[c], __ldcg(&xn[c]));
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
      }
      if (lane == 0) q.xout[j][row] = zsub(__ldcg(&q.z[(size_t)j * bm + row]), acc);
    }
    if (j > 0) kb_red_sync(q.ctr, ++epoch, q.err, q.wait_ns);
  }
}

// one product C = alpha A B + beta C
static void zgemm(kb_context* h, int m, int n, int k, double alpha, const double2* A, int lda, const double2* B,
                  int ldb, double beta, double2* C, int ldc) {
  GemmQueue q;
  q.add(m, n, k, alpha, A, lda, false, B, ldb, beta, C, ldc);
  q.flush(h);
}

// debug / microbenchmark (tools/dev_zgemm.py): C = alpha op(A) B + beta C on host arrays, `batch`
// identical problems per launch, `reps` launches timed; returns the mean ms per launch
extern "C" int kb_dbg_zgemm(kb_handle h, int m, int n, int k, int transA, const double* A, const double* B,
                            double* C, double alpha, double beta, int batch, int reps, double* ms_out) {
  if (!h || batch < 1 || batch > 4) return KB_EINVAL;
  KB_CUDA(h, cudaSetDevice(h->device));
  const size_t na = (size_t)m * k, nb = (size_t)k * n, nc = (size_t)m * n;
  DevBuf<double2> dA, dB, dC;
  KB_CUDA(h, dA.alloc(na));
  KB_CUDA(h, dB.alloc(nb));
  KB_CUDA(h, dC.alloc(nc * batch));
  KB_CUDA(h, cudaMemcpy(dA.p, A, na * sizeof(double2), cudaMemcpyHostToDevice));
  KB_CUDA(h, cudaMemcpy(dB.p, B, nb * sizeof(double2), cudaMemcpyHostToDevice));
  for (int i = 0; i < batch; ++i) KB_CUDA(h, cudaMemcpy(dC.p + i * nc, C, nc * sizeof(double2), cudaMemcpyHostToDevice));
  KbEventPair ev;
  KB_CUDA(h, ev.create());
  for (int r = -1; r < reps; ++r) {
    if (r == 0) KB_CUDA(h, cudaEventRecord(ev.e0, h->stream));
    GemmQueue q;
    for (int i = 0; i < batch; ++i)
      q.add(m, n, k, alpha, dA.p, transA ? m : k, transA != 0, dB.p, n, r <= 0 ? beta : 0.0, dC.p + i * nc, n);
    q.flush(h);
  }
  KB_LAUNCH_CHECK(h);
  KB_CUDA(h, cudaEventRecord(ev.e1, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (ms_out) *ms_out = reps > 0 ? ev.ms() / reps : 0.0;
  KB_CUDA(h, cudaMemcpy(C, dC.p + (batch - 1) * nc, nc * sizeof(double2), cudaMemcpyDeviceToHost));
  return KB_OK;
}

__global__ void kb_set_identity(double2* __restrict__ A, int b) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (int64_t)b * b) A[e] = zmake((e / b) == (e % b) ? 1.0 : 0.0, 0.0);
}

// Right-hand side of the second pass of the l-sharded solve, rows of this rank's interior:
//   out = r - L_{f,top} y_top   on the first interior node (rows [of, of+bf)), if there is a separator above,
//   out = r - U_{l,bot} y_bot   on the last interior node  (rows [ol, ol+bl)), if there is one below,
//   out = r                      elsewhere.  One warp per row.
__global__ void __launch_bounds__(256)
kb_shard_rhs(int row_lo, int row_hi, int of, int bf, int use_l, int ol, int bl, int use_u,
             const double2* __restrict__ r, const double2* __restrict__ y, double2* __restrict__ out,
             const int64_t* __restrict__ rowptr, const int64_t* __restrict__ dstart,
             const int64_t* __restrict__ ustart, const int* __restrict__ col, const double2* __restrict__ T) {
  const int lane = threadIdx.x & 31;
  const int gi = row_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gi >= row_hi) return;
  double2 acc = zmake(0.0, 0.0);
  if (use_l && gi >= of && gi < of + bf)
    for (int64_t k = rowptr[gi] + lane; k < dstart[gi]; k += 32) zfma(acc, T[k], y[col[k]]);
  if (use_u && gi >= ol && gi < ol + bl)
    for (int64_t k = ustart[gi] + lane; k < rowptr[gi + 1]; k += 32) zfma(acc, T[k], y[col[k]]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
  }
  if (lane == 0) out[gi] = zsub(r[gi], acc);
}

static inline int nsize(const kb_context* h, int64_t p) { return (int)(h->nodeptr[p + 1] - h->nodeptr[p]); }
static inline int noff(const kb_context* h, int64_t p) { return (int)h->nodeptr[p]; }

// ---------------------------------------------------------------------------
// factor
// ---------------------------------------------------------------------------
// per-phase device times of the fast path on stderr when KB_SHARD_TIMING is set
struct PhaseTimer {
  bool on;
  cudaStream_t s;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
  PhaseTimer(cudaStream_t st) : on(getenv("KB_SHARD_TIMING") != nullptr), s(st) { mark("start"); }
  ~PhaseTimer() {
    for (auto e : ev) cudaEventDestroy(e);
  }
  void mark(const char* name) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    ev.push_back(e);
    names.push_back(name);
  }
  void report(int rank) {
    if (!on) return;
    cudaStreamSynchronize(s);
    std::string line = "libkoreb200: shard timing rank " + std::to_string(rank) + ":";
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      char buf[64];
      snprintf(buf, sizeof(buf), " %s %.2f ms", names[i], ms);
      line += buf;
    }
    fprintf(stderr, "%s\n", line.c_str());
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

static int reduced_factor(kb_context* h, double* flops);
static int factor_sharded_fast(kb_context* h, double* flops, int64_t* bytes);
static int factor_sharded_general(kb_context* h, double* flops, int64_t* bytes);

// The fast path needs the strip kernel and the folded sweep on this rank's interior
static bool shard_fast_supported(kb_context* h) {
  if (getenv("KB_SHARD_GENERAL")) return false;
  if (h->safe_mode || h->opt_sweep != 1 || h->opt_fold == 0 || h->opt_factor == 0) return false;
  if (!kbi_chainfac_supported(h)) return false;
  const bool two = h->int_hi - h->int_lo >= 4;
  const bool was = h->M_transposed;
  h->M_transposed = true;
  const bool ok = kbi_fold_supported(h, kbi_fold_grid(h, two), two, nullptr, nullptr);
  h->M_transposed = was;
  return ok;
}

// Everything of an l-sharded factorisation that precedes the elimination and needs no communication:
// T = A - sigma B, workspaces, storage of the interior inverses, node tables / ELL couplings / sweep
// grid on the device, this rank's (zeroed) contribution blocks.
static int shard_prepare(kb_context* h, zcomplex sigma) {
  KB_TRY(shard_ranges(h));
  cudaStream_t s = h->stream;
  const int64_t P = h->P, bmax = h->bmax;
  h->factored = false;
  h->fold_ready = false;
  h->M_transposed = false;
  h->shard_fast = false;
  h->sigma = sigma;
  kbi_drop_graphs(h);
  KB_TRY(kbi_build_T(h, sigma));
  KB_TRY(kbi_factor_workspace(h));
  h->Moff.assign(P + 1, 0);
  int64_t mt = 0;
  for (int64_t p = h->int_lo; p < h->int_hi; ++p) {
    const int64_t b = h->nodeptr[p + 1] - h->nodeptr[p];
    h->Moff[p] = mt;
    mt += b * b;
  }
  if (h->d_M.alloc((size_t)mt) != cudaSuccess)
    return kb_fail(h, KB_ENOMEM, "cannot allocate %.2f GB for this rank's chain factors", mt * 16.0 / 1e9);
  KB_TRY(kbi_sweep_prepare(h));
  KB_CUDA(h, h->d_contrib.alloc(4 * (size_t)bmax * bmax));
  KB_CUDA(h, h->d_contrib_all.alloc(4 * (size_t)bmax * bmax * h->nranks));
  KB_CUDA(h, cudaMemsetAsync(h->d_contrib.p, 0, 4 * (size_t)bmax * bmax * sizeof(double2), s));
  return KB_OK;
}

// General path: per-node kernels, one-sided elimination of the interior with spike recurrences
// towards the separator above (nodes of any size).
static int factor_sharded_general(kb_context* h, double* flops_io, int64_t* bytes_io) {
  cudaStream_t s = h->stream;
  const int64_t P = h->P, bmax = h->bmax;
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
  *flops_io += flops;
  *bytes_io = (mtot + 2 * vtot) * 16;
  return KB_OK;
}

int kbi_factor_sharded(kb_context* h, zcomplex sigma) {
  if (!h->chain_set) return kb_fail(h, KB_EINVAL, "kb_set_chain must be called before kb_factor");
  if (!h->nccl_comm) return kb_fail(h, KB_EINVAL, "kb_set_sharding must be called before kb_factor");
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int64_t bmax = h->bmax;
  const int G = h->nranks, g = h->rank;
  KbEventPair ev;
  KB_CUDA(h, ev.create());
  KB_CUDA(h, cudaEventRecord(ev.e0, s));
  KB_TRY(shard_prepare(h, sigma));
  const bool fast = shard_fast_supported(h);
  double flops = 0.0;
  int64_t bytes = 0;
  if (fast)
    KB_TRY(factor_sharded_fast(h, &flops, &bytes));
  else
    KB_TRY(factor_sharded_general(h, &flops, &bytes));
  PhaseTimer pt(s);
  KB_NCCL(h, g_nccl.AllGather(h->d_contrib.p, h->d_contrib_all.p, 4 * (size_t)bmax * bmax * 2, ncclDouble,
                              (ncclComm_t)h->nccl_comm, s));
  pt.mark("allgather");
  KB_TRY(reduced_factor(h, &flops));
  pt.mark("reduced-factor");
  pt.report(g);
  KB_CUDA(h, cudaEventRecord(ev.e1, s));
  KB_TRY(kbi_sync(h));
  int info = 0, kerr = 0;
  KB_CUDA(h, cudaMemcpy(&info, h->d_info.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (fast) KB_CUDA(h, cudaMemcpy(&kerr, h->d_kfsync.p + KF_ERR_WORD, sizeof(int), cudaMemcpyDeviceToHost));
  h->stats.factor_ms = ev.ms();
  h->stats.factor_flops = flops;
  h->stats.factor_bytes = bytes + (int64_t)bmax * bmax * (G - 1) * 16;
  h->stats.shard_path = fast ? 2 : 1;
  if (kerr != 0)
    return kb_fail(h, KB_ECUDA, "a device-side wait of the strip factorisation expired on rank %d (code %d, CTA %d)",
                   g, kerr & 255, kerr >> 8);
  if (info != 0) return kb_fail(h, KB_ESINGULAR, "zero or non-finite pivot in a Schur block on rank %d", g);
  h->shard_fast = fast;
  h->factored = true;
  return KB_OK;
}

// Test hook (tests/test_gpu_parity.py, one GPU, no communicator): the elimination rank `rank` of
// `nranks` would do on this pencil and the four blocks it would contribute to the reduced system
// (4 x bmax x bmax complex128, slots as in factor_sharded_fast), by the fast (path = 2) or the general
// (path = 1) kernels.  The handle is left unfactored and unsharded.
extern "C" int kb_dbg_shard_segment(kb_handle h, int rank, int nranks, const double* sigma, int path, double* blocks) {
  if (!h || !sigma || !blocks || nranks < 2 || rank < 0 || rank >= nranks) return KB_EINVAL;
  if (!h->chain_set) return kb_fail(h, KB_EINVAL, "kb_set_chain must be called first");
  if (h->nccl_comm) return kb_fail(h, KB_EINVAL, "kb_dbg_shard_segment is for unsharded handles");
  KB_CUDA(h, cudaSetDevice(h->device));
  const int rank0 = h->rank, nranks0 = h->nranks;
  h->rank = rank;
  h->nranks = nranks;
  double flops = 0.0;
  int64_t bytes = 0;
  int rc = shard_prepare(h, zcomplex(sigma[0], sigma[1]));
  if (rc == KB_OK) {
    if (path == 2 && !shard_fast_supported(h))
      rc = kb_fail(h, KB_EINVAL, "the fast l-sharded path does not support this pencil");
    else
      rc = path == 2 ? factor_sharded_fast(h, &flops, &bytes) : factor_sharded_general(h, &flops, &bytes);
  }
  if (rc == KB_OK) {
    cudaError_t e = cudaMemcpyAsync(blocks, h->d_contrib.p, 4 * (size_t)h->bmax * h->bmax * sizeof(double2),
                                    cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = kb_fail(h, KB_ECUDA, "copy of the contribution blocks failed: %s", cudaGetErrorString(e));
  }
  h->rank = rank0;
  h->nranks = nranks0;
  h->factored = false;
  h->fold_ready = false;
  h->shard_fast = false;
  h->M_transposed = false;
  return rc;
}

// ---- reduced (separator) system, factored redundantly on every rank: node j = separator of rank j
static int reduced_factor(kb_context* h, double* flops_io) {
  cudaStream_t s = h->stream;
  const int64_t bmax = h->bmax;
  const int G = h->nranks;
  const size_t slot = (size_t)bmax * bmax;
  double flops = 0.0;
  KB_CUDA(h, h->d_F.alloc(slot));
  KB_CUDA(h, h->d_Mr.alloc(slot * (G - 1)));
  // separator blocks are inverted by the strip kernel as one-node chains when it supports the size
  const bool strip = kbi_chainfac_supported(h) && !h->safe_mode && !getenv("KB_SHARD_GENERAL");
  if (strip) {
    std::vector<int64_t> tab(2 * (G - 1) + 1, 0);
    for (int j = 0; j < G - 1; ++j) tab[2 * j + 1] = nsize(h, h->seg_hi[j] - 1);
    KB_CUDA(h, h->d_redtab.alloc(tab.size()));
    KB_CUDA(h, cudaMemcpyAsync(h->d_redtab.p, tab.data(), tab.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    KB_CUDA(h, cudaStreamSynchronize(s));  // (pageable source on the stack of this call)
  }
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
    if (strip) {
      KbDenseNode dn;
      dn.S = h->d_S0.p;
      dn.out = h->d_Mr.p + (size_t)j * slot;
      dn.nodeptr = h->d_redtab.p + 2 * j;
      dn.moff = h->d_redtab.p + 2 * (G - 1);
      KB_TRY(kbi_chainfac_run(h, false, false, 0, 1, &dn));
    } else {
      kb_first_panel<<<nblk(bs, 128), 128, 0, s>>>(h->d_S0.p, bs, kbi_panel_width(h, bs), h->d_PT.p);
      double2* X = nullptr;
      KB_TRY(gj_invert(h, kbi_ws_main(h), bs, &X));
      kb_store_inverse<<<bs, 128, 0, s>>>(X, bs, h->d_orig.p, h->d_Mr.p + (size_t)j * slot);
      h->launches += 2;
    }
    flops += 8.0 * (double)bs * bs * bs;
    KB_LAUNCH_CHECK(h);
  }
  // E_j = Mr_j Csub_j, Fm_j = Mr_j Csup_{j+1}: the fused reduced solve (kb_reduced_solve_kernel)
  KB_CUDA(h, h->d_Csub.alloc(slot * (G - 1)));
  KB_CUDA(h, h->d_Csup.alloc(slot * (G - 1)));
  {
    GemmQueue q;
    for (int j = 0; j < G - 1; ++j) {
      const int bs = nsize(h, h->seg_hi[j] - 1);
      const double2* Mr = h->d_Mr.p + (size_t)j * slot;
      if (j > 0) {
        const int bsp = nsize(h, h->seg_hi[j - 1] - 1);
        q.add(bs, bsp, bs, 1.0, Mr, bs, false, h->d_contrib_all.p + (size_t)j * 4 * slot + 2 * slot, bsp, 0.0,
              h->d_Csub.p + (size_t)j * slot, bsp);
        flops += 8.0 * (double)bs * bsp * bs;
        if (q.count == 4) q.flush(h);
      }
      if (j < G - 2) {
        const int bsn = nsize(h, h->seg_hi[j + 1] - 1);
        q.add(bs, bsn, bs, 1.0, Mr, bs, false, h->d_contrib_all.p + (size_t)(j + 1) * 4 * slot + 3 * slot, bsn, 0.0,
              h->d_Csup.p + (size_t)j * slot, bsn);
        flops += 8.0 * (double)bs * bsn * bs;
        if (q.count == 4) q.flush(h);
      }
    }
    q.flush(h);
    KB_LAUNCH_CHECK(h);
  }
  KB_CUDA(h, h->d_sepvec.alloc(2 * bmax));
  KB_CUDA(h, h->d_sepvec_all.alloc((size_t)2 * bmax * G));
  KB_CUDA(h, h->d_redz.alloc((size_t)2 * bmax * G));

  *flops_io += flops;
  return KB_OK;
}

// ---------------------------------------------------------------------------
// Fast l-sharded factorisation: the rank's interior [lo, hi) is a block-tridiagonal system of
// its own and goes through the SAME kernels as a whole pencil on one GPU -- two-sided strip
// factorisation (kb_chainfac.cu), folded couplings (kb_sweep2.cu).  What the separators need from
// it are the four corner blocks of G = T_I^{-1} (f / l = first / last interior node):
//     G_ff, G_lf = column block f of G at rows f and l;   G_fl, G_ll = column block l.
// A column block of G is the interior solve with an identity block as right-hand side, i.e. the
// folded sweep with b right-hand sides: chains of dense products with the folded couplings,
//     forward   Tf_f = I, Tf_{p+1} = -FL_p Tf_p (p < m);    Tl_l = I, Tl_{p-1} = -FU_p Tl_p (p > m)
//     backward  U_{p-1} = T_{p-1} - FU_p U_p  (up),  U_{p+1} = T_{p+1} - FL_p U_p  (down),  U_m = T_m
//     corner    G_{f,.} = M_f U_f,   G_{l,.} = M_l U_l
// (T = 0 on the half of the chain the identity block is not in), three b x b x b products per
// interior node in all, issued as batches of independent products (kb_zgemm_batch).  The reduced
// system then receives
//     slot 0: D_bot - L_{bot,l} G_ll U_{l,bot}      slot 1: U_{top,f} G_ff L_{f,top}
//     slot 2: -L_{bot,l} G_lf L_{f,top}             slot 3: -U_{top,f} G_fl U_{l,bot}
// exactly as from the general path above.
// ---------------------------------------------------------------------------
static int factor_sharded_fast(kb_context* h, double* flops_io, int64_t* bytes_io) {
  cudaStream_t s = h->stream;
  PhaseTimer pt(s);
  const int64_t P = h->P, bmax = h->bmax;
  const int64_t lo = h->int_lo, hi = h->int_hi, nI = hi - lo;
  const bool has_top = h->top_sep >= 0, has_bot = h->bot_sep >= 0;
  const size_t slot = (size_t)bmax * bmax;
  const bool two = nI >= 4;
  h->ch_lo = lo;
  h->ch_hi = hi;
  h->mid = two ? lo + nI / 2 : hi - 1;
  h->M_transposed = true;
  KB_TRY(kbi_chainfac_run(h, two, true, lo, hi));
  pt.mark("strip-factor");
  KB_TRY(kbi_fold_prepare(h));
  pt.mark("fold");
  if (!h->fold_ready) return kb_fail(h, KB_ENOMEM, "cannot set up the folded sweep on rank %d", h->rank);
  double flops = 0.0;
  int64_t mtot = 0;
  for (int64_t p = lo; p < hi; ++p) {
    const double b = (double)nsize(h, p);
    flops += 8.0 * b * b * b;
    mtot += (int64_t)b * (int64_t)b;
  }
  *bytes_io = 3 * mtot * 16;
  if (!has_top && !has_bot) {
    *flops_io += flops;
    return KB_OK;
  }

  const int64_t f = lo, l = hi - 1, m = h->mid;
  const int bf = nsize(h, f), bl = nsize(h, l);
  // Tf_p (b_p x bf), p = f .. m;  Tl_p (b_p x bl), p = m .. l;  two ping-pong pairs;  four corners
  std::vector<int64_t> tfo(P, -1), tlo(P, -1);
  int64_t tot = 0;
  for (int64_t p = f; p <= m; ++p) {
    tfo[p] = tot;
    tot += (int64_t)nsize(h, p) * bf;
  }
  for (int64_t p = m; p <= l; ++p) {
    tlo[p] = tot;
    tot += (int64_t)nsize(h, p) * bl;
  }
  if (h->d_spk.alloc((size_t)tot + 8 * slot) != cudaSuccess)
    return kb_fail(h, KB_ENOMEM, "cannot allocate %.2f GB of spike workspace", (tot + 8 * slot) * 16.0 / 1e9);
  double2* base = h->d_spk.p;
  double2* pp = base + tot;  // ping-pong buffers [0..3], corners [4..7]
  auto Tf = [&](int64_t p) { return base + tfo[p]; };
  auto Tl = [&](int64_t p) { return base + tlo[p]; };
  const double2* F = h->d_fold.p;
  auto gflops = [&](int mm, int nn, int kk) { flops += 8.0 * (double)mm * nn * kk; };

  kb_set_identity<<<nblk((int64_t)bf * bf, 256), 256, 0, s>>>(Tf(f), bf);
  kb_set_identity<<<nblk((int64_t)bl * bl, 256), 256, 0, s>>>(Tl(l), bl);
  h->launches += 2;
  // ---- forward
  {
    int64_t pf = f, pl = l;
    while (pf < m || pl > m) {
      GemmQueue q;
      if (pf < m) {
        q.add(nsize(h, pf + 1), bf, nsize(h, pf), -1.0, F + h->FLoff[pf], nsize(h, pf), false, Tf(pf), bf, 0.0,
              Tf(pf + 1), bf);
        gflops(nsize(h, pf + 1), bf, nsize(h, pf));
        ++pf;
      }
      if (pl > m) {
        q.add(nsize(h, pl - 1), bl, nsize(h, pl), -1.0, F + h->FUoff[pl], nsize(h, pl), false, Tl(pl), bl, 0.0,
              Tl(pl - 1), bl);
        gflops(nsize(h, pl - 1), bl, nsize(h, pl));
        --pl;
      }
      q.flush(h);
    }
  }
  pt.mark("spike-forward");
  // ---- backward: four independent chains in lock step
  const double2* uf_l = Tf(m);  // U of column block f at node l (down chain; = Tf_m when m == l)
  const double2* ul_f = Tl(m);  // U of column block l at node f (up chain;   = Tl_m when m == f)
  {
    int64_t up = m, dn = m;     // next step: up uses FU_up (-> node up-1), down uses FL_dn (-> node dn+1)
    int tog = 0;
    while (up > f || dn < l) {
      GemmQueue q;
      if (up > f) {
        const int bo = nsize(h, up - 1), bi = nsize(h, up);
        // column block f: Tf_{up-1} <- Tf_{up-1} - FU_up U_up   (U_up already sits in Tf_up)
        q.add(bo, bf, bi, -1.0, F + h->FUoff[up], bi, false, Tf(up), bf, 1.0, Tf(up - 1), bf);
        // column block l: U_{up-1} = -FU_up U_up
        double2* out = pp + (size_t)(tog ? 1 : 0) * slot;
        q.add(bo, bl, bi, -1.0, F + h->FUoff[up], bi, false, ul_f, bl, 0.0, out, bl);
        ul_f = out;
        gflops(bo, bf, bi);
        gflops(bo, bl, bi);
        --up;
      }
      if (dn < l) {
        const int bo = nsize(h, dn + 1), bi = nsize(h, dn);
        // column block l: Tl_{dn+1} <- Tl_{dn+1} - FL_dn U_dn
        q.add(bo, bl, bi, -1.0, F + h->FLoff[dn], bi, false, Tl(dn), bl, 1.0, Tl(dn + 1), bl);
        // column block f: U_{dn+1} = -FL_dn U_dn
        double2* out = pp + (size_t)(tog ? 3 : 2) * slot;
        q.add(bo, bf, bi, -1.0, F + h->FLoff[dn], bi, false, uf_l, bf, 0.0, out, bf);
        uf_l = out;
        gflops(bo, bl, bi);
        gflops(bo, bf, bi);
        ++dn;
      }
      q.flush(h);
      tog ^= 1;
    }
  }
  // ---- corners (the inverses are stored transposed)
  double2 *Gff = pp + 4 * slot, *Glf = pp + 5 * slot, *Gfl = pp + 6 * slot, *Gll = pp + 7 * slot;
  {
    const double2* Mf = h->d_M.p + h->Moff[f];
    const double2* Ml = h->d_M.p + h->Moff[l];
    GemmQueue q;
    q.add(bf, bf, bf, 1.0, Mf, bf, true, Tf(f), bf, 0.0, Gff, bf);
    q.add(bl, bf, bl, 1.0, Ml, bl, true, uf_l, bf, 0.0, Glf, bf);
    q.add(bf, bl, bf, 1.0, Mf, bf, true, ul_f, bl, 0.0, Gfl, bl);
    q.add(bl, bl, bl, 1.0, Ml, bl, true, Tl(l), bl, 0.0, Gll, bl);
    q.flush(h);
    gflops(bf, bf, bf);
    gflops(bl, bf, bl);
    gflops(bf, bl, bf);
    gflops(bl, bl, bl);
  }
  KB_LAUNCH_CHECK(h);
  pt.mark("spike-backward+corners");
  // ---- contributions to the reduced system
  const int of = noff(h, f), ol = noff(h, l);
  double2* X = h->d_W.p;  // bmax x bmax scratch
  if (has_bot) {
    const int ob = noff(h, h->bot_sep), bb = nsize(h, h->bot_sep);
    // X = G_ll U_{l,bot};  slot 0 = D_bot - L_{bot,l} X
    kb_dense_spcols<<<bl, 128, bl * sizeof(double2), s>>>(Gll, bl, ol, X, bb, ob, h->d_ucptr.p, h->d_urow.p,
                                                          h->d_upos.p, h->d_Tval.p, 1.0);
    kb_schur_row<<<bb, 128, 0, s>>>(h->d_contrib.p, h->d_PT.p, 0, bb, ob, X, ol, nullptr, 0, h->d_rowptr.p,
                                    h->d_dstart.p, h->d_ustart.p, h->d_col.p, h->d_Tval.p);
    h->launches += 2;
  }
  if (has_top) {
    const int ot = noff(h, h->top_sep), bt = nsize(h, h->top_sep);
    // X = G_ff L_{f,top};  slot 1 = U_{top,f} X
    kb_dense_spcols<<<bf, 128, bf * sizeof(double2), s>>>(Gff, bf, of, X, bt, ot, h->d_lcptr.p, h->d_lrow.p,
                                                          h->d_lpos.p, h->d_Tval.p, 1.0);
    kb_sprows_dense<<<bt, 128, 0, s>>>(h->d_contrib.p + slot, bt, ot, 1, X, of, h->d_rowptr.p, h->d_dstart.p,
                                       h->d_ustart.p, h->d_col.p, h->d_Tval.p, 1.0, 0);
    h->launches += 2;
  }
  if (has_top && has_bot) {
    const int ot = noff(h, h->top_sep), bt = nsize(h, h->top_sep);
    const int ob = noff(h, h->bot_sep), bb = nsize(h, h->bot_sep);
    // X = G_lf L_{f,top} (bl x bt);  slot 2 = -L_{bot,l} X  (bb x bt)
    kb_dense_spcols<<<bl, 128, bf * sizeof(double2), s>>>(Glf, bf, of, X, bt, ot, h->d_lcptr.p, h->d_lrow.p,
                                                          h->d_lpos.p, h->d_Tval.p, 1.0);
    kb_sprows_dense<<<bb, 128, 0, s>>>(h->d_contrib.p + 2 * slot, bt, ob, 0, X, ol, h->d_rowptr.p, h->d_dstart.p,
                                       h->d_ustart.p, h->d_col.p, h->d_Tval.p, -1.0, 0);
    // X = G_fl U_{l,bot} (bf x bb);  slot 3 = -U_{top,f} X  (bt x bb)
    kb_dense_spcols<<<bf, 128, bl * sizeof(double2), s>>>(Gfl, bl, ol, h->d_S1.p, bb, ob, h->d_ucptr.p, h->d_urow.p,
                                                          h->d_upos.p, h->d_Tval.p, 1.0);
    kb_sprows_dense<<<bt, 128, 0, s>>>(h->d_contrib.p + 3 * slot, bb, ot, 1, h->d_S1.p, of, h->d_rowptr.p,
                                       h->d_dstart.p, h->d_ustart.p, h->d_col.p, h->d_Tval.p, -1.0, 0);
    h->launches += 4;
  }
  KB_LAUNCH_CHECK(h);
  pt.mark("contributions");
  pt.report(h->rank);
  *flops_io += flops;
  return KB_OK;
}

// ---------------------------------------------------------------------------
// solve: y <- T'^{-1} r, full vectors (chain order, scaled space) on every rank
// ---------------------------------------------------------------------------
// reduced solve (redundant on every rank) from the gathered separator contributions:
//   z_j = Mr_j (rho_j - Csub_j z_{j-1});  x_j = z_j - Mr_j Csup_{j+1} x_{j+1};  x_j -> y at separator j
static bool sepvec_by_mailbox(const kb_context* h) {
  return h->mb_ready && (size_t)2 * h->bmax * sizeof(double2) <= KB_MB_SLOT;
}

static int reduced_solve(kb_context* h, double2* y) {
  cudaStream_t s = h->stream;
  const int G = h->nranks;
  const int64_t bmax = h->bmax;
  const size_t slot = (size_t)bmax * bmax;
  if (G - 1 > KB_MAXSEP) return kb_fail(h, KB_EINVAL, "more than %d separators", KB_MAXSEP);
  KbRedParams q;
  memset(&q, 0, sizeof(q));
  q.nsep = G - 1;
  q.bmax = (int)bmax;
  for (int j = 0; j < G - 1; ++j) {
    const int64_t sp = h->seg_hi[j] - 1;
    q.bs[j] = nsize(h, sp);
    q.Mr[j] = h->d_Mr.p + (size_t)j * slot;
    q.E[j] = h->d_Csub.p + (size_t)j * slot;
    q.Fm[j] = h->d_Csup.p + (size_t)j * slot;
    q.xout[j] = y + noff(h, sp);
  }
  q.sepvec_all = h->d_sepvec_all.p;
  q.sepstride = (size_t)2 * bmax;
  q.use_mb = 0;
  if (h->mb_ready && (size_t)2 * bmax * sizeof(double2) <= KB_MB_SLOT) {
    q.use_mb = 1;
    q.mb = mailbox_of(h);
    q.mb_epoch = ++h->mb_epoch;
    q.sv = h->d_sepvec.p;
    q.sepvec_all = (const double2*)(h->d_mbox.p + (size_t)(q.mb_epoch & 1u) * KB_MB_MAXRANKS * KB_MB_SLOT);
    q.sepstride = KB_MB_SLOT / sizeof(double2);
  }
  q.z = h->d_redz.p;
  int sms = 0;
  KB_CUDA(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  int grid = (int)((bmax + 7) / 8);  // one warp per row
  if (grid > sms) grid = sms;
  if (h->red_grid != grid || !h->d_redctr.p) {
    KB_CUDA(h, h->d_redctr.alloc(32));
    KB_CUDA(h, cudaMemsetAsync(h->d_redctr.p, 0, 32 * sizeof(unsigned), s));
    h->red_grid = grid;
    h->red_epoch = 0;
  }
  q.ctr = h->d_redctr.p;
  q.epoch0 = h->red_epoch;
  // exactly the barriers of one launch (nsep forward + nsep - 2 backward for nsep >= 2): every barrier
  // adds gridDim.x to the counter, so the epochs of consecutive launches must be consecutive
  h->red_epoch += (unsigned)(G - 1 > 1 ? 2 * (G - 1) - 2 : 0);
  q.err = h->d_sweep_err.p;
  q.wait_ns = h->wait_ns;
  void* args[] = {(void*)&q};
  KB_CUDA(h, cudaLaunchCooperativeKernel((const void*)kb_reduced_solve_kernel, dim3(grid), dim3(256), args, 0, s));
  h->launches++;
  return KB_OK;
}

// every rank publishes its interior; the separators are already everywhere
static int publish_interiors(kb_context* h, double2* y) {
  const int G = h->nranks;
  // Segments of equal length (interior + the separator below it, whose value every rank already
  // holds): ONE in-place all-gather.  Otherwise a grouped broadcast per rank.
  const int64_t cnt0 = h->nodeptr[h->seg_hi[0]] - h->nodeptr[h->seg_lo[0]];
  bool equal = true;
  for (int q = 1; q < G; ++q) equal = equal && (h->nodeptr[h->seg_hi[q]] - h->nodeptr[h->seg_lo[q]] == cnt0);
  if (equal) {
    KB_NCCL(h, g_nccl.AllGather(y + (size_t)h->rank * cnt0, y, (size_t)cnt0 * 2, ncclDouble, (ncclComm_t)h->nccl_comm,
                                h->stream));
    return KB_OK;
  }
  KB_NCCL(h, g_nccl.GroupStart());
  for (int q = 0; q < G; ++q) {
    int64_t qlo = h->seg_lo[q], qhi = q < G - 1 ? h->seg_hi[q] - 1 : h->seg_hi[q];
    int64_t off = h->nodeptr[qlo], cnt = h->nodeptr[qhi] - off;
    KB_NCCL(h, g_nccl.Broadcast(y + off, y + off, (size_t)cnt * 2, ncclDouble, q, (ncclComm_t)h->nccl_comm,
                                h->stream));
  }
  KB_NCCL(h, g_nccl.GroupEnd());
  return KB_OK;
}

// Fast l-sharded solve (factors from factor_sharded_fast):
//   pass 1  y_I = T_I^{-1} r_I by the folded sweep on the interior; only the end nodes' values are formed
//   gather  rank g contributes r_bot - L_{bot,l} y_l to its lower separator and U_{top,f} y_f to its upper one
//   reduced solve (redundant) -> the separator values
//   pass 2  x_I = T_I^{-1} (r_I - T_{I,sep} x_sep): the folded sweep again on the corrected right-hand side
//   publish interiors
static int sharded_sweeps_fast(kb_context* h, const double2* r, double2* y) {
  cudaStream_t s = h->stream;
  const int64_t lo = h->int_lo, hi = h->int_hi, bmax = h->bmax;
  const bool has_top = h->top_sep >= 0, has_bot = h->bot_sep >= 0;
  const int64_t f = lo, l = hi - 1;
  const int of = noff(h, f), bf = nsize(h, f), ol = noff(h, l), bl = nsize(h, l);
  double2* t = h->d_t.p;
  double2* sv = h->d_sepvec.p;
  static int timed_solves = 0;
  PhaseTimer pt(s);
  if (timed_solves >= 3) pt.on = false;
  KB_CUDA(h, cudaMemsetAsync(sv, 0, 2 * bmax * sizeof(double2), s));
  KB_TRY(kbi_sweep_fold(h, r, y, 1));
  pt.mark("pass1");
  if (has_top) {
    const int ot = noff(h, h->top_sep), bt = nsize(h, h->top_sep);
    kb_node_tvec<1><<<(bt + 7) / 8, 256, 0, s>>>(bt, ot, r, y, t, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                                 h->d_col.p, h->d_Tval.p);
    KB_CUDA(h, cudaMemcpyAsync(sv, t + ot, (size_t)bt * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    h->launches++;
  }
  if (has_bot) {
    const int ob = noff(h, h->bot_sep), bb = nsize(h, h->bot_sep);
    kb_node_tvec<0><<<(bb + 7) / 8, 256, 0, s>>>(bb, ob, r, y, t, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                                 h->d_col.p, h->d_Tval.p);
    KB_CUDA(h, cudaMemcpyAsync(sv + bmax, t + ob, (size_t)bb * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    h->launches++;
  }
  KB_LAUNCH_CHECK(h);
  pt.mark("sep-rhs");
  if (!sepvec_by_mailbox(h))
    KB_NCCL(h, g_nccl.AllGather(sv, h->d_sepvec_all.p, 2 * bmax * 2, ncclDouble, (ncclComm_t)h->nccl_comm, s));
  pt.mark("allgather");
  KB_TRY(reduced_solve(h, y));
  pt.mark("reduced-solve");
  if (has_top || has_bot) {
    const int row_lo = of, row_hi = ol + bl;
    kb_shard_rhs<<<nblk((int64_t)(row_hi - row_lo) * 32, 256), 256, 0, s>>>(
        row_lo, row_hi, of, bf, has_top ? 1 : 0, ol, bl, has_bot ? 1 : 0, r, y, t, h->d_rowptr.p, h->d_dstart.p,
        h->d_ustart.p, h->d_col.p, h->d_Tval.p);
    h->launches++;
    KB_TRY(kbi_sweep_fold(h, t, y, 0));
  }
  KB_LAUNCH_CHECK(h);
  pt.mark("pass2");
  if (!h->keep_sharded) KB_TRY(publish_interiors(h, y));
  pt.mark("publish");
  if (pt.on) {
    ++timed_solves;
    pt.report(h->rank);
  }
  return KB_OK;
}

int kbi_sharded_sweeps(kb_context* h, const double2* r, double2* y) {
  if (h->shard_fast) return sharded_sweeps_fast(h, r, y);
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
  if (!sepvec_by_mailbox(h))
    KB_NCCL(h, g_nccl.AllGather(sv, h->d_sepvec_all.p, 2 * bmax * 2, ncclDouble, (ncclComm_t)h->nccl_comm, s));

  KB_TRY(reduced_solve(h, y));

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

  (void)g;
  (void)G;
  (void)slot;
  return h->keep_sharded ? KB_OK : publish_interiors(h, y);
}

int kbi_chain_solve_sharded(kb_context* h, const double2* r, double2* x, int refine) {
  (void)refine;
  return kbi_sharded_sweeps(h, r, x);
}

// ---------------------------------------------------------------------------
// Row-sharded Krylov basis (kb_eigs.cu with nranks > 1): every rank keeps the rows of its own
// segment [seg_lo, seg_hi) of every basis vector; what crosses the ranks are the Gram-Schmidt
// coefficients and norms (all-reduce of <= ncv + 1 numbers), one chain node of halo on each side for
// the B product, and the separator contributions of the solve -- not vectors of length n.
// ---------------------------------------------------------------------------
__global__ void kb_reduce_cols_local(int ncols, int nchunks, const double2* __restrict__ partial, double2* __restrict__ h) {
  const int c = blockIdx.x, lane = threadIdx.x;
  double2 acc = zmake(0.0, 0.0);
  for (int b = lane; b < nchunks; b += 32) acc = zadd(acc, partial[(size_t)c * nchunks + b]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
  }
  if (lane == 0) h[c] = acc;
}
__global__ void kb_accum_cols(int ncols, const double2* __restrict__ h, double2* __restrict__ hsum, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncols) hsum[c] = accumulate ? zadd(hsum[c], h[c]) : h[c];
}
__global__ void kb_sumsq_local(int nparts, const double* __restrict__ normpart, double* __restrict__ out) {
  const int lane = threadIdx.x;
  double acc = 0.0;
  for (int b = lane; b < nparts; b += 32) acc += normpart[b];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) out[0] = acc;
}
__global__ void kb_sqrt1(double* __restrict__ v) { v[0] = sqrt(v[0]); }

void kbi_shard_rows(const kb_context* h, int64_t* row_lo, int64_t* row_hi) {
  *row_lo = h->nodeptr[h->seg_lo[h->rank]];
  *row_hi = h->nodeptr[h->seg_hi[h->rank]];
}

// h[c] = sum over the ranks of (sum_b partial[c, b]);  hsum (+)= h.  One CTA: the local reduction,
// the exchange through the mailboxes and the sum in rank order (bit-identical on every rank).
__global__ void __launch_bounds__(256) kb_reduce_cols_mb(KbMailbox m, unsigned epoch, int ncols, int nchunks,
                                                         const double2* __restrict__ partial, double2* __restrict__ h,
                                                         double2* __restrict__ hsum, int accumulate, int* err,
                                                         unsigned long long wait_ns) {
  __shared__ double2 loc[128];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int c = wid; c < ncols; c += 8) {
    double2 acc = zmake(0.0, 0.0);
    for (int b = lane; b < nchunks; b += 32) acc = zadd(acc, partial[(size_t)c * nchunks + b]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
    }
    if (lane == 0) loc[c] = acc;
  }
  __syncthreads();
  const int par = (int)(epoch & 1u);
  for (int r = 0; r < m.nranks; ++r) {
    double2* dst = kb_mb_data(m, r, par, m.rank);
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) dst[c] = loc[c];
  }
  kb_mb_signal_all(m, epoch);
  kb_mb_wait_all(m, epoch, err, wait_ns);
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    double2 sum = zmake(0.0, 0.0);
    for (int r = 0; r < m.nranks; ++r) sum = zadd(sum, __ldcg(kb_mb_data(m, m.rank, par, r) + c));
    h[c] = sum;
    hsum[c] = accumulate ? zadd(hsum[c], sum) : sum;
  }
}

// beta = sqrt(sum over the ranks of sum_b normpart[b])
__global__ void __launch_bounds__(32) kb_norm_mb(KbMailbox m, unsigned epoch, int nparts,
                                                 const double* __restrict__ normpart, double* __restrict__ beta, int* err,
                                                 unsigned long long wait_ns) {
  const int lane = threadIdx.x;
  double acc = 0.0;
  for (int b = lane; b < nparts; b += 32) acc += normpart[b];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  const int par = (int)(epoch & 1u);
  if (lane < m.nranks) kb_mb_data(m, lane, par, m.rank)[0] = zmake(acc, 0.0);
  kb_mb_signal_all(m, epoch);
  kb_mb_wait_all(m, epoch, err, wait_ns);
  if (lane == 0) {
    double t = 0.0;
    for (int r = 0; r < m.nranks; ++r) t += __ldcg(kb_mb_data(m, m.rank, par, r)).x;
    beta[0] = sqrt(t);
  }
}

// halo of a row-sharded vector: first node to the rank above, last node to the rank below; every
// rank signals every rank (the lock-step invariant of the two-parity mailboxes)
__global__ void __launch_bounds__(256) kb_halo_mb(KbMailbox m, unsigned epoch, double2* __restrict__ x, int o_first,
                                                  int b_first, int o_last, int b_last, int o_above, int b_above,
                                                  int o_below, int b_below, int* err, unsigned long long wait_ns) {
  const int par = (int)(epoch & 1u);
  if (m.rank > 0) {
    double2* dst = kb_mb_data(m, m.rank - 1, par, m.rank);
    for (int i = threadIdx.x; i < b_first; i += blockDim.x) dst[i] = x[o_first + i];
  }
  if (m.rank < m.nranks - 1) {
    double2* dst = kb_mb_data(m, m.rank + 1, par, m.rank);
    for (int i = threadIdx.x; i < b_last; i += blockDim.x) dst[i] = x[o_last + i];
  }
  kb_mb_signal_all(m, epoch);
  kb_mb_wait_all(m, epoch, err, wait_ns);
  if (m.rank > 0) {
    const double2* src = kb_mb_data(m, m.rank, par, m.rank - 1);
    for (int i = threadIdx.x; i < b_above; i += blockDim.x) x[o_above + i] = __ldcg(src + i);
  }
  if (m.rank < m.nranks - 1) {
    const double2* src = kb_mb_data(m, m.rank, par, m.rank + 1);
    for (int i = threadIdx.x; i < b_below; i += blockDim.x) x[o_below + i] = __ldcg(src + i);
  }
}

int kbi_shard_reduce_cols(kb_context* h, int ncols, int nchunks, const double2* hpart, double2* hdev, double2* hsum,
                          double2* scratch, int accumulate) {
  cudaStream_t s = h->stream;
  if (h->mb_ready && ncols <= 128) {
    kb_reduce_cols_mb<<<1, 256, 0, s>>>(mailbox_of(h), ++h->mb_epoch, ncols, nchunks, hpart, hdev, hsum, accumulate,
                                        h->d_sweep_err.p, h->wait_ns);
    h->launches++;
    KB_LAUNCH_CHECK(h);
    return KB_OK;
  }
  kb_reduce_cols_local<<<ncols, 32, 0, s>>>(ncols, nchunks, hpart, hdev);
  KB_NCCL(h, g_nccl.AllReduce(hdev, hdev, (size_t)2 * ncols, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, s));
  kb_accum_cols<<<1, 128, 0, s>>>(ncols, hdev, hsum, accumulate);
  (void)scratch;
  h->launches += 2;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

int kbi_shard_norm(kb_context* h, int nparts, const double* normpart, double* beta_dev) {
  cudaStream_t s = h->stream;
  if (h->mb_ready) {
    kb_norm_mb<<<1, 32, 0, s>>>(mailbox_of(h), ++h->mb_epoch, nparts, normpart, beta_dev, h->d_sweep_err.p, h->wait_ns);
    h->launches++;
    KB_LAUNCH_CHECK(h);
    return KB_OK;
  }
  kb_sumsq_local<<<1, 32, 0, s>>>(nparts, normpart, beta_dev);
  KB_NCCL(h, g_nccl.AllReduce(beta_dev, beta_dev, 1, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, s));
  kb_sqrt1<<<1, 1, 0, s>>>(beta_dev);
  h->launches += 2;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

// in-place sum over the ranks of `count` doubles
int kbi_shard_allreduce(kb_context* h, double* buf, size_t count) {
  KB_NCCL(h, g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
  return KB_OK;
}

// x holds this rank's segment; fetch the chain node above it and the node below it from the
// neighbouring ranks (block-tridiagonal B and A reach no further)
int kbi_shard_halo(kb_context* h, double2* x) {
  const int g = h->rank, G = h->nranks;
  const int64_t lo = h->seg_lo[g], hi = h->seg_hi[g];
  if (h->mb_ready && (size_t)h->bmax * sizeof(double2) <= KB_MB_SLOT) {
    kb_halo_mb<<<1, 256, 0, h->stream>>>(mailbox_of(h), ++h->mb_epoch, x, noff(h, lo), nsize(h, lo), noff(h, hi - 1),
                                         nsize(h, hi - 1), g > 0 ? noff(h, lo - 1) : 0, g > 0 ? nsize(h, lo - 1) : 0,
                                         g < G - 1 ? noff(h, hi) : 0, g < G - 1 ? nsize(h, hi) : 0, h->d_sweep_err.p,
                                         h->wait_ns);
    h->launches++;
    KB_LAUNCH_CHECK(h);
    return KB_OK;
  }
  KB_NCCL(h, g_nccl.GroupStart());
  if (g > 0) {
    KB_NCCL(h, g_nccl.Send(x + noff(h, lo), (size_t)nsize(h, lo) * 2, ncclDouble, g - 1, (ncclComm_t)h->nccl_comm,
                           h->stream));
    KB_NCCL(h, g_nccl.Recv(x + noff(h, lo - 1), (size_t)nsize(h, lo - 1) * 2, ncclDouble, g - 1,
                           (ncclComm_t)h->nccl_comm, h->stream));
  }
  if (g < G - 1) {
    KB_NCCL(h, g_nccl.Send(x + noff(h, hi - 1), (size_t)nsize(h, hi - 1) * 2, ncclDouble, g + 1,
                           (ncclComm_t)h->nccl_comm, h->stream));
    KB_NCCL(h, g_nccl.Recv(x + noff(h, hi), (size_t)nsize(h, hi) * 2, ncclDouble, g + 1, (ncclComm_t)h->nccl_comm,
                           h->stream));
  }
  KB_NCCL(h, g_nccl.GroupEnd());
  return KB_OK;
}

// every rank's segment of x to every rank (extraction of the eigenvectors)
int kbi_shard_gather_segments(kb_context* h, double2* x) {
  const int G = h->nranks;
  const int64_t cnt0 = h->nodeptr[h->seg_hi[0]] - h->nodeptr[h->seg_lo[0]];
  bool equal = true;
  for (int q = 1; q < G; ++q) equal = equal && (h->nodeptr[h->seg_hi[q]] - h->nodeptr[h->seg_lo[q]] == cnt0);
  if (equal) {
    KB_NCCL(h, g_nccl.AllGather(x + (size_t)h->rank * cnt0, x, (size_t)cnt0 * 2, ncclDouble, (ncclComm_t)h->nccl_comm,
                                h->stream));
    return KB_OK;
  }
  KB_NCCL(h, g_nccl.GroupStart());
  for (int q = 0; q < G; ++q) {
    const int64_t off = h->nodeptr[h->seg_lo[q]], cnt = h->nodeptr[h->seg_hi[q]] - off;
    KB_NCCL(h, g_nccl.Broadcast(x + off, x + off, (size_t)cnt * 2, ncclDouble, q, (ncclComm_t)h->nccl_comm, h->stream));
  }
  KB_NCCL(h, g_nccl.GroupEnd());
  return KB_OK;
}
