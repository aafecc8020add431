// Device-side build of the chain (l-major, block-tridiagonal) layout from the caller's CSR.
//
// Replaces: PETSc Mat create / setValuesCSR / assembly for A and B
// (/root/reference/bin/solve.py:43-59, 69-85) plus the analysis phase of the sparse direct
// solver, for the data-layout part: the raw CSR triplets go to the device once, and the
// union pattern of (A, B) in chain order -- rows sorted by chain column, B's own CSR and its
// position map into the union pattern, per-row split points (sub-diagonal | diagonal |
// super-diagonal node), the couplings by column -- is built there:
//   1. key = chain_row * n + chain_col for every entry of A, then of B (payload = entry index,
//      top bit = "from B");
//   2. one stable radix sort (cub) of the n_A + n_B keys: rows in chain order, columns
//      ascending inside a row, A before B on equal keys;
//   3. heads of equal-key runs = the union pattern (exclusive scan -> positions); a second
//      scan over the "from B" flags gives B's positions in its own chain-ordered CSR;
//   4. row pointers from per-row counts, split points by binary search inside each row,
//      coupling width and structure check (a nonzero more than one node away) by reductions;
//   5. the couplings by column: keys column * n + row of the off-diagonal-node entries,
//      sorted the same way.
// The host implementation in kb_setup.cu (two threaded passes + std::sort per row) is kept as
// the cross-check (KB_HOST_LAYOUT=1; tests compare the two bit for bit through the solver).
#include <cub/cub.cuh>

#include "kb_internal.cuh"

namespace {

__global__ void kl_iperm(int n, const int* __restrict__ perm, int* __restrict__ iperm, int* __restrict__ bad) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int o = perm[k];
  if (o < 0 || o >= n) {
    atomicExch(bad, 1);
    return;
  }
  if (atomicCAS(&iperm[o], -1, k) != -1) atomicExch(bad, 1);
}

__global__ void kl_node_of(int P, const int64_t* __restrict__ nodeptr, int* __restrict__ node_of) {
  // one CTA per node
  int p = blockIdx.x;
  for (int64_t i = nodeptr[p] + threadIdx.x; i < nodeptr[p + 1]; i += blockDim.x) node_of[i] = p;
}

// one warp per ORIGINAL row: keys of its entries
template <typename IT>
__global__ void kl_keys(int n, const int64_t* __restrict__ indptr, const IT* __restrict__ indices,
                        const int* __restrict__ iperm, unsigned long long* __restrict__ keys,
                        unsigned* __restrict__ payload, int64_t base, unsigned tag, int* __restrict__ bad) {
  int o = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (o >= n) return;
  const unsigned long long row = (unsigned long long)iperm[o];
  for (int64_t e = indptr[o] + lane; e < indptr[o + 1]; e += 32) {
    const long long c = (long long)indices[e];
    if (c < 0 || c >= n) {
      atomicExch(bad, 2);
      keys[base + e] = row * (unsigned long long)n;
    } else {
      keys[base + e] = row * (unsigned long long)n + (unsigned long long)iperm[c];
    }
    payload[base + e] = (unsigned)e | tag;
  }
}

__global__ void kl_heads(int64_t N, const unsigned long long* __restrict__ keys, const unsigned* __restrict__ payload,
                         int* __restrict__ head, int* __restrict__ isb) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
  isb[i] = (payload[i] >> 31) ? 1 : 0;
}

// per sorted entry: union slot k = (inclusive scan of heads) - 1
template <typename BT>
__global__ void kl_scatter(int64_t N, int n, const unsigned long long* __restrict__ keys,
                           const unsigned* __restrict__ payload, const int* __restrict__ head,
                           const int* __restrict__ hscan, const int* __restrict__ bscan,
                           const double2* __restrict__ aval_raw, const BT* __restrict__ bval_raw,
                           const int* __restrict__ node_of, int* __restrict__ col, double2* __restrict__ aval,
                           int* __restrict__ rowcnt, int* __restrict__ bcol, int* __restrict__ bmap,
                           BT* __restrict__ bval, int* __restrict__ browcnt, int* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned long long key = keys[i];
  const int row = (int)(key / (unsigned long long)n), c = (int)(key - (unsigned long long)row * n);
  const int k = hscan[i] + head[i] - 1;  // hscan is exclusive
  if (head[i]) {
    col[k] = c;
    atomicAdd(&rowcnt[row], 1);
    const int d = node_of[c] - node_of[row];
    if (d > 1 || d < -1) atomicMax(bad, 3);
    // the A entries of the run (normally one), summed in order: deterministic
    double2 acc = make_double2(0.0, 0.0);
    for (int64_t j = i; j < N && keys[j] == key; ++j) {
      const unsigned pl = payload[j];
      if (!(pl >> 31)) {
        const double2 v = aval_raw[pl];
        acc.x += v.x;
        acc.y += v.y;
      }
    }
    aval[k] = acc;
  }
  const unsigned pl = payload[i];
  if (pl >> 31) {
    // kb_sub_sigma_B gives every B entry its own read-modify-write of its union slot: two B entries
    // of one (row, column) would race there
    if (i > 0 && keys[i - 1] == key && (payload[i - 1] >> 31)) atomicMax(bad, 4);
    const int kb = bscan[i];
    bcol[kb] = c;
    bmap[kb] = k;
    bval[kb] = bval_raw[pl & 0x7fffffffu];
    atomicAdd(&browcnt[row], 1);
  }
}

__global__ void kl_widen(int n, const int* __restrict__ cnt, int64_t* __restrict__ ptr) {
  // ptr[i] = exclusive scan result (int) widened; ptr[n] set by the caller
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) ptr[i] = (int64_t)cnt[i];
}

// split points of every row, coupling widths
__global__ void kl_splits(int n, const int64_t* __restrict__ rowptr, const int* __restrict__ col,
                          const int* __restrict__ node_of, int64_t* __restrict__ dstart,
                          int64_t* __restrict__ ustart, int* __restrict__ wl, int* __restrict__ wu) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pi = node_of[i];
  const int64_t a = rowptr[i], b = rowptr[i + 1];
  int64_t lo = a, hi = b;
  while (lo < hi) {  // first entry with node >= pi
    int64_t m = (lo + hi) >> 1;
    if (node_of[col[m]] < pi) lo = m + 1; else hi = m;
  }
  const int64_t ds = lo;
  hi = b;
  while (lo < hi) {  // first entry with node > pi
    int64_t m = (lo + hi) >> 1;
    if (node_of[col[m]] <= pi) lo = m + 1; else hi = m;
  }
  const int64_t us = lo;
  dstart[i] = ds;
  ustart[i] = us;
  atomicMax(wl, (int)(ds - a));
  atomicMax(wu, (int)(b - us));
}

// keys of the coupling entries by column: which = 0: U (column node = row node + 1), 1: L
__global__ void kl_coupling_keys(int n, int which, const int64_t* __restrict__ rowptr,
                                 const int64_t* __restrict__ dstart, const int64_t* __restrict__ ustart,
                                 const int* __restrict__ col, unsigned long long* __restrict__ keys,
                                 unsigned* __restrict__ pos, int* __restrict__ colcnt,
                                 unsigned long long* __restrict__ counter) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t a = which == 0 ? ustart[i] : rowptr[i];
  const int64_t b = which == 0 ? rowptr[i + 1] : dstart[i];
  if (b <= a) return;
  const unsigned long long at = atomicAdd(counter, (unsigned long long)(b - a));
  for (int64_t k = a; k < b; ++k) {
    const int c = col[k];
    keys[at + (k - a)] = (unsigned long long)c * (unsigned long long)n + (unsigned long long)i;
    pos[at + (k - a)] = (unsigned)k;
    atomicAdd(&colcnt[c], 1);
  }
}

__global__ void kl_coupling_out(int64_t N, int n, const unsigned long long* __restrict__ keys,
                                const unsigned* __restrict__ pos, int* __restrict__ rowout,
                                int64_t* __restrict__ posout) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  rowout[i] = (int)(keys[i] % (unsigned long long)n);
  posout[i] = (int64_t)pos[i];
}

// Temporary device buffer from the stream-ordered pool (the pool keeps the memory between
// calls: cudaMalloc / cudaFree of the ~0.7 GB of sort buffers per layout build cost more
// than the build itself)
template <typename T>
struct PoolBuf {
  T* p = nullptr;
  cudaStream_t s = nullptr;
  PoolBuf() {}
  PoolBuf(const PoolBuf&) = delete;
  PoolBuf& operator=(const PoolBuf&) = delete;
  ~PoolBuf() {
    if (p) cudaFreeAsync(p, s);
  }
  cudaError_t alloc(size_t n, cudaStream_t st) {
    s = st;
    if (n == 0) n = 1;
    return cudaMallocAsync((void**)&p, n * sizeof(T), st);
  }
};

struct Tmp {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t s = nullptr;
  ~Tmp() {
    if (p) cudaFreeAsync(p, s);
  }
  cudaError_t need(size_t b) {
    if (b <= bytes) return cudaSuccess;
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMallocAsync(&p, b, s);
    if (e == cudaSuccess) bytes = b;
    return e;
  }
};

int bits_for(unsigned long long maxkey) {
  int b = 1;
  while (b < 64 && (maxkey >> b)) ++b;
  return b;
}

// exclusive scan of n+1 ints in place (cnt[n] must be 0 on entry; cnt[n] = total on exit)
int scan_counts(kb_context* h, PoolBuf<int>& cnt, int n, Tmp& tmp, cudaStream_t s) {
  size_t tb = 0;
  KB_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, cnt.p, n + 1, s));
  KB_CUDA(h, tmp.need(tb));
  KB_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, cnt.p, n + 1, s));
  return KB_OK;
}

}  // namespace

// Raw CSR of the caller on the device (kb_set_pencil), consumed by kbi_layout_device.
int kbi_upload_raw(kb_context* h, KbRawCSR& M, int64_t n, int index_bytes, const void* indptr, const void* indices,
                   const void* values, bool is_complex, const char* name) {
  cudaStream_t s = h->stream;
  std::vector<int64_t> ptr(n + 1);
  if (index_bytes == 4) {
    const int32_t* p = (const int32_t*)indptr;
    for (int64_t i = 0; i <= n; ++i) ptr[i] = p[i];
  } else {
    const int64_t* p = (const int64_t*)indptr;
    for (int64_t i = 0; i <= n; ++i) ptr[i] = p[i];
  }
  if (ptr[0] != 0) return kb_fail(h, KB_EINVAL, "%s: indptr[0] != 0", name);
  for (int64_t i = 0; i < n; ++i)
    if (ptr[i + 1] < ptr[i]) return kb_fail(h, KB_EINVAL, "%s: indptr not monotone", name);
  const int64_t nnz = ptr[n];
  if (nnz >= 0x7fffffff) return kb_fail(h, KB_EINVAL, "%s: more than 2^31 nonzeros are not supported", name);
  M.n = n;
  M.nnz = nnz;
  M.index_bytes = index_bytes;
  M.is_complex = is_complex;
  KB_CUDA(h, M.indptr.alloc(n + 1));
  KB_CUDA(h, cudaMemcpyAsync(M.indptr.p, ptr.data(), (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  KB_CUDA(h, M.indices.alloc((size_t)(nnz > 0 ? nnz : 1) * index_bytes));
  KB_CUDA(h, M.values.alloc((size_t)(nnz > 0 ? nnz : 1) * (is_complex ? 16 : 8)));
  if (nnz > 0) {
    KB_CUDA(h, cudaMemcpyAsync(M.indices.p, indices, (size_t)nnz * index_bytes, cudaMemcpyHostToDevice, s));
    KB_CUDA(h, cudaMemcpyAsync(M.values.p, values, (size_t)nnz * (is_complex ? 16 : 8), cudaMemcpyHostToDevice, s));
  }
  // the caller may free its arrays as soon as kb_set_pencil returns
  KB_CUDA(h, cudaStreamSynchronize(s));
  M.present = true;
  return KB_OK;
}

// Builds every device array kb_set_chain promises from the raw CSR on the device.
// h->perm / h->nodeptr / h->P / h->bmax are set by the caller.
int kbi_layout_device(kb_context* h) {
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  const int64_t P = h->P;
  KbRawCSR& A = h->rawA;
  KbRawCSR& B = h->rawB;
  const int64_t NA = A.nnz, NB = B.present ? B.nnz : 0, N = NA + NB;
  const int thr = 256;
  Tmp tmp;
  tmp.s = s;

  // ---- permutation, node of every chain position
  std::vector<int> perm32(n);
  for (int k = 0; k < n; ++k) {
    if (h->perm[k] < 0 || h->perm[k] >= n) return kb_fail(h, KB_EINVAL, "perm is not a permutation");
    perm32[k] = (int)h->perm[k];
  }
  KB_CUDA(h, h->d_perm.alloc(n));
  KB_CUDA(h, cudaMemcpyAsync(h->d_perm.p, perm32.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
  KB_CUDA(h, h->d_nodeptr.alloc(P + 1));
  KB_CUDA(h, cudaMemcpyAsync(h->d_nodeptr.p, h->nodeptr.data(), (P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  PoolBuf<int> iperm, node_of, bad;
  KB_CUDA(h, iperm.alloc(n, s));
  KB_CUDA(h, node_of.alloc(n, s));
  KB_CUDA(h, bad.alloc(1, s));
  KB_CUDA(h, cudaMemsetAsync(iperm.p, 0xff, (size_t)n * sizeof(int), s));
  KB_CUDA(h, cudaMemsetAsync(bad.p, 0, sizeof(int), s));
  kl_iperm<<<nblk(n, thr), thr, 0, s>>>(n, h->d_perm.p, iperm.p, bad.p);
  kl_node_of<<<(unsigned)P, 128, 0, s>>>((int)P, h->d_nodeptr.p, node_of.p);

  // ---- keys + one stable sort
  PoolBuf<unsigned long long> keys, keys2;
  PoolBuf<unsigned> pay, pay2;
  KB_CUDA(h, keys.alloc(N, s));
  KB_CUDA(h, keys2.alloc(N, s));
  KB_CUDA(h, pay.alloc(N, s));
  KB_CUDA(h, pay2.alloc(N, s));
  const unsigned wblk = nblk((int64_t)n * 32, thr);
  if (A.index_bytes == 4)
    kl_keys<int32_t><<<wblk, thr, 0, s>>>(n, A.indptr.p, (const int32_t*)A.indices.p, iperm.p, keys.p, pay.p, 0, 0u, bad.p);
  else
    kl_keys<int64_t><<<wblk, thr, 0, s>>>(n, A.indptr.p, (const int64_t*)A.indices.p, iperm.p, keys.p, pay.p, 0, 0u, bad.p);
  if (NB > 0) {
    if (B.index_bytes == 4)
      kl_keys<int32_t><<<wblk, thr, 0, s>>>(n, B.indptr.p, (const int32_t*)B.indices.p, iperm.p, keys.p, pay.p, NA,
                                            0x80000000u, bad.p);
    else
      kl_keys<int64_t><<<wblk, thr, 0, s>>>(n, B.indptr.p, (const int64_t*)B.indices.p, iperm.p, keys.p, pay.p, NA,
                                            0x80000000u, bad.p);
  }
  const int kbits = bits_for((unsigned long long)n * (unsigned long long)n);
  {
    size_t tb = 0;
    KB_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, keys2.p, pay.p, pay2.p, (int)N, 0, kbits, s));
    KB_CUDA(h, tmp.need(tb));
    KB_CUDA(h, cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys2.p, pay.p, pay2.p, (int)N, 0, kbits, s));
  }
  // sorted: keys2 / pay2
  PoolBuf<int> head, isb, hscan, bscan;
  KB_CUDA(h, head.alloc(N + 1, s));
  KB_CUDA(h, isb.alloc(N + 1, s));
  KB_CUDA(h, hscan.alloc(N + 1, s));
  KB_CUDA(h, bscan.alloc(N + 1, s));
  KB_CUDA(h, cudaMemsetAsync(head.p + N, 0, sizeof(int), s));
  KB_CUDA(h, cudaMemsetAsync(isb.p + N, 0, sizeof(int), s));
  kl_heads<<<nblk(N, thr), thr, 0, s>>>(N, keys2.p, pay2.p, head.p, isb.p);
  {
    size_t tb = 0;
    KB_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tb, head.p, hscan.p, (int)N + 1, s));
    KB_CUDA(h, tmp.need(tb));
    KB_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp.p, tb, head.p, hscan.p, (int)N + 1, s));
    KB_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp.p, tb, isb.p, bscan.p, (int)N + 1, s));
  }
  int tot[2] = {0, 0}, hbad = 0;
  KB_CUDA(h, cudaMemcpyAsync(&tot[0], hscan.p + N, sizeof(int), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaMemcpyAsync(&tot[1], bscan.p + N, sizeof(int), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaStreamSynchronize(s));
  if (hbad == 1) return kb_fail(h, KB_EINVAL, "perm is not a permutation");
  if (hbad == 2) return kb_fail(h, KB_EINVAL, "column index out of range");
  const int64_t nnz = tot[0], nnzB = tot[1];
  h->nnz = nnz;
  h->nnzB = nnzB;

  // ---- union pattern, B's CSR, row counts
  KB_CUDA(h, h->d_col.alloc(nnz > 0 ? nnz : 1));
  KB_CUDA(h, h->d_Aval.alloc(nnz > 0 ? nnz : 1));
  KB_CUDA(h, h->d_Tval.alloc(nnz > 0 ? nnz : 1));
  KB_CUDA(h, h->d_bcol.alloc(nnzB > 0 ? nnzB : 1));
  KB_CUDA(h, h->d_bmap.alloc(nnzB > 0 ? nnzB : 1));
  if (h->b_is_complex)
    KB_CUDA(h, h->d_bval_c.alloc(nnzB > 0 ? nnzB : 1));
  else
    KB_CUDA(h, h->d_bval_r.alloc(nnzB > 0 ? nnzB : 1));
  PoolBuf<int> rowcnt, browcnt;
  KB_CUDA(h, rowcnt.alloc(n + 1, s));
  KB_CUDA(h, browcnt.alloc(n + 1, s));
  KB_CUDA(h, cudaMemsetAsync(rowcnt.p, 0, (size_t)(n + 1) * sizeof(int), s));
  KB_CUDA(h, cudaMemsetAsync(browcnt.p, 0, (size_t)(n + 1) * sizeof(int), s));
  if (h->b_is_complex)
    kl_scatter<double2><<<nblk(N, thr), thr, 0, s>>>(N, n, keys2.p, pay2.p, head.p, hscan.p, bscan.p,
                                                     (const double2*)A.values.p, (const double2*)B.values.p, node_of.p,
                                                     h->d_col.p, h->d_Aval.p, rowcnt.p, h->d_bcol.p, h->d_bmap.p,
                                                     h->d_bval_c.p, browcnt.p, bad.p);
  else
    kl_scatter<double><<<nblk(N, thr), thr, 0, s>>>(N, n, keys2.p, pay2.p, head.p, hscan.p, bscan.p,
                                                    (const double2*)A.values.p, (const double*)B.values.p, node_of.p,
                                                    h->d_col.p, h->d_Aval.p, rowcnt.p, h->d_bcol.p, h->d_bmap.p,
                                                    h->d_bval_r.p, browcnt.p, bad.p);
  KB_TRY(scan_counts(h, rowcnt, n, tmp, s));
  KB_TRY(scan_counts(h, browcnt, n, tmp, s));
  KB_CUDA(h, h->d_rowptr.alloc(n + 1));
  KB_CUDA(h, h->d_browptr.alloc(n + 1));
  kl_widen<<<nblk(n + 1, thr), thr, 0, s>>>(n, rowcnt.p, h->d_rowptr.p);
  kl_widen<<<nblk(n + 1, thr), thr, 0, s>>>(n, browcnt.p, h->d_browptr.p);

  // ---- split points, coupling widths
  KB_CUDA(h, h->d_dstart.alloc(n));
  KB_CUDA(h, h->d_ustart.alloc(n));
  PoolBuf<int> wbuf;
  KB_CUDA(h, wbuf.alloc(2, s));
  KB_CUDA(h, cudaMemsetAsync(wbuf.p, 0, 2 * sizeof(int), s));
  kl_splits<<<nblk(n, thr), thr, 0, s>>>(n, h->d_rowptr.p, h->d_col.p, node_of.p, h->d_dstart.p, h->d_ustart.p,
                                         wbuf.p, wbuf.p + 1);

  // ---- couplings by column (U: row node p, column node p+1; L: column node p-1)
  PoolBuf<unsigned long long> counter;
  KB_CUDA(h, counter.alloc(1, s));
  for (int which = 0; which < 2; ++which) {
    PoolBuf<int> colcnt;
    KB_CUDA(h, colcnt.alloc(n + 1, s));
    KB_CUDA(h, cudaMemsetAsync(colcnt.p, 0, (size_t)(n + 1) * sizeof(int), s));
    KB_CUDA(h, cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), s));
    // the sort buffers of step 2 are large enough (couplings are a subset of the union pattern)
    kl_coupling_keys<<<nblk(n, thr), thr, 0, s>>>(n, which, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p, h->d_col.p,
                                                  keys.p, pay.p, colcnt.p, counter.p);
    unsigned long long nc = 0;
    KB_CUDA(h, cudaMemcpyAsync(&nc, counter.p, sizeof(nc), cudaMemcpyDeviceToHost, s));
    KB_CUDA(h, cudaStreamSynchronize(s));
    if (nc > 0) {
      size_t tb = 0;
      KB_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, keys2.p, pay.p, pay2.p, (int)nc, 0, kbits, s));
      KB_CUDA(h, tmp.need(tb));
      KB_CUDA(h, cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys2.p, pay.p, pay2.p, (int)nc, 0, kbits, s));
    }
    KB_TRY(scan_counts(h, colcnt, n, tmp, s));
    DevBuf<int64_t>& cptr = which == 0 ? h->d_ucptr : h->d_lcptr;
    DevBuf<int>& crow = which == 0 ? h->d_urow : h->d_lrow;
    DevBuf<int64_t>& cpos = which == 0 ? h->d_upos : h->d_lpos;
    KB_CUDA(h, cptr.alloc(n + 1));
    KB_CUDA(h, crow.alloc(nc > 0 ? nc : 1));
    KB_CUDA(h, cpos.alloc(nc > 0 ? nc : 1));
    kl_widen<<<nblk(n + 1, thr), thr, 0, s>>>(n, colcnt.p, cptr.p);
    if (nc > 0) kl_coupling_out<<<nblk((int64_t)nc, thr), thr, 0, s>>>((int64_t)nc, n, keys2.p, pay2.p, crow.p, cpos.p);
    if (which == 0) h->nnzU = (int64_t)nc;
  }

  int wh[2] = {0, 0};
  KB_CUDA(h, cudaMemcpyAsync(wh, wbuf.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, h->d_rscale.alloc(n));
  KB_CUDA(h, h->d_cscale.alloc(n));
  KB_CUDA(h, h->d_maxbits.alloc(n));
  KB_CUDA(h, cudaStreamSynchronize(s));
  KB_LAUNCH_CHECK(h);
  if (hbad == 4)
    return kb_fail(h, KB_EINVAL,
                   "B holds more than one entry for some (row, column): sum the duplicates first "
                   "(scipy: B.sum_duplicates())");
  if (hbad == 3)
    return kb_fail(h, KB_ESTRUCTURE,
                   "pencil is not block tridiagonal under the given chain (a nonzero couples nodes "
                   "more than one apart)");
  h->WL = wh[0];
  h->WU = wh[1];
  return KB_OK;
}
