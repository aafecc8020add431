// Device-side assembly of Kore's pencil from an assembly program.
//
// Replaces /root/reference/bin/assemble.py:432-1171 for every set-up kore_b200/assembly.py writes a
// program for (hydrodynamic, thermal, anelastic, degree-1 magnetic): the host writes down what every
// N1 x N1 block of A and B is
// -- groups of (coefficient, radial operator) terms with their scalar factors, plus the dense
// boundary-condition rows -- and the kernels below evaluate that program straight into the raw
// CSR (original Kore ordering, canonical: rows and columns ascending, exact zeros dropped as
// utils.py:164 does) that kb_set_chain's layout build consumes.  The matrices never exist on the
// host: what crosses PCIe is the radial operators as bands (nop x N1 x (2H+1) doubles) and a
// few hundred KB of tables instead of the 250 MB of CSR triplets of the E = 1e-8 pencil.
//
// Bit compatibility with the reference: an entry of the reference's matrices is the result of a
// fixed sequence of IEEE double multiplications and additions (scipy.sparse scalar products and
// sums, evaluated left to right by the Python expressions of operators.py).  ka_eval performs
// that sequence with __dmul_rn / __dadd_rn, which the compiler never contracts into fused
// multiply-adds, so the CSR equals the reference's to the bit (tests/test_assembly.py compares
// with the reference-assembled fixtures; tests/assembly_model.py is the NumPy model of this file).
//
// Two passes over the rows, one warp per row: count the nonzeros, exclusive scan (cub), fill.
// Candidates of an operator row are the (block, band offset) pairs of its block row in column
// order; a boundary row is one dense row of the diagonal block.  HBM-bound and tiny next to the
// factorisation (the E = 1e-8 pencil is 1.2e7 entries); nothing here is worth a tensor core.
#include <cub/cub.cuh>

#include "kb_internal.cuh"

namespace {

struct KaProg {
  int N1, nblockrows, H, W, is_complex, use_final;
  double final_scale;
  const double* ops;
  const double* bc;
  const int* br_chop;
  const int* br_bc;
  const int* blk_ptr;
  const int* blk_col;
  const int* blk_grp;
  const int* grp_part;
  const int* grp_sign;
  const int* grp_nsc;
  const double* grp_sc;
  const int* grp_term;
  const double* term_coef;
  const int* term_op;
};

// value of entry (i, i + d - H) of block `blk`
__device__ __forceinline__ void ka_eval(const KaProg& p, int blk, int i, int d, double& re, double& im) {
  re = 0.0;
  im = 0.0;
  bool hre = false, him = false;
  for (int g = p.blk_grp[blk]; g < p.blk_grp[blk + 1]; ++g) {
    double lin = 0.0;
    const int t0 = p.grp_term[g], t1 = p.grp_term[g + 1];
    for (int t = t0; t < t1; ++t) {
      const double x = p.ops[((size_t)p.term_op[t] * p.N1 + i) * p.W + d];
      const double pr = __dmul_rn(p.term_coef[t], x);
      lin = t == t0 ? pr : __dadd_rn(lin, pr);
    }
    const int ns = p.grp_nsc[g];
    for (int k = 0; k < ns; ++k) lin = __dmul_rn(p.grp_sc[g * 4 + k], lin);
    if (p.grp_sign[g] < 0) lin = -lin;
    if (p.grp_part[g] == 0) {
      re = hre ? __dadd_rn(re, lin) : lin;
      hre = true;
    } else {
      im = him ? __dadd_rn(im, lin) : lin;
      him = true;
    }
  }
  if (p.use_final) {
    re = __dmul_rn(re, p.final_scale);
    im = __dmul_rn(im, p.final_scale);
  }
}

// one warp per row.  FILL = false: cnt[row] = nonzeros of the row; FILL = true: indices and
// values of the row at rowptr[row], columns ascending.
template <bool FILL>
__global__ void __launch_bounds__(256) ka_rows(KaProg p, int64_t* __restrict__ cnt, const int64_t* __restrict__ rowptr,
                                               int* __restrict__ indices, double* __restrict__ values) {
  const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int n = p.N1 * p.nblockrows;
  if (row >= n) return;
  const int br = row / p.N1, i = row - br * p.N1;
  const int chop = p.br_chop[br];
  int64_t at = FILL ? rowptr[row] : 0;
  int total = 0;
  const unsigned below = (1u << lane) - 1u;
  if (i < chop) {
    // boundary row: dense, diagonal block
    const double* src = p.bc + (size_t)(p.br_bc[br] + i) * p.N1;
    for (int j0 = 0; j0 < p.N1; j0 += 32) {
      const int j = j0 + lane;
      double v = j < p.N1 ? src[j] : 0.0;
      if (p.use_final) v = __dmul_rn(v, p.final_scale);
      const bool nz = v != 0.0;
      const unsigned mask = __ballot_sync(0xffffffffu, nz);
      if (FILL && nz) {
        const int64_t pos = at + total + __popc(mask & below);
        indices[pos] = br * p.N1 + j;
        if (p.is_complex) {
          values[2 * pos] = v;
          values[2 * pos + 1] = 0.0;
        } else {
          values[pos] = v;
        }
      }
      total += __popc(mask);
    }
  } else {
    const int b0 = p.blk_ptr[br], nb = p.blk_ptr[br + 1] - b0;
    const int ncand = nb * p.W;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
      const int c = c0 + lane;
      bool nz = false;
      double re = 0.0, im = 0.0;
      int col = 0;
      if (c < ncand) {
        const int kb = c / p.W, d = c - kb * p.W;
        const int j = i + d - p.H;
        if (j >= 0 && j < p.N1) {
          ka_eval(p, b0 + kb, i, d, re, im);
          nz = re != 0.0 || im != 0.0;
          col = p.blk_col[b0 + kb] * p.N1 + j;
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, nz);
      if (FILL && nz) {
        const int64_t pos = at + total + __popc(mask & below);
        indices[pos] = col;
        if (p.is_complex) {
          values[2 * pos] = re;
          values[2 * pos + 1] = im;
        } else {
          values[pos] = re;
        }
      }
      total += __popc(mask);
    }
  }
  if (!FILL && lane == 0) cnt[row] = total;
}

// every table of a program in one device allocation
struct KaDeviceProg {
  DevBuf<unsigned char> blob;
  KaProg p;
};

template <typename T>
void ka_put(std::vector<unsigned char>& host, std::vector<size_t>& offs, const T* src, size_t count) {
  size_t at = (host.size() + 15) & ~(size_t)15;
  host.resize(at + count * sizeof(T));
  if (count) memcpy(host.data() + at, src, count * sizeof(T));
  offs.push_back(at);
}

int ka_check(kb_context* h, const kb_asm_program* q, const char* name) {
  if (q->N1 < 1 || q->nblockrows < 1 || q->H < 0 || 2 * q->H + 1 > 1023)
    return kb_fail(h, KB_EINVAL, "%s program: bad sizes (N1 %d, block rows %d, half band %d)", name, q->N1,
                   q->nblockrows, q->H);
  if ((int64_t)q->N1 * q->nblockrows > 0x7fffffff) return kb_fail(h, KB_EINVAL, "%s program: n out of range", name);
  if (q->nop < 1 || q->nblk < 0 || q->ngrp < 0 || q->nterm < 0 || q->nbc < 0)
    return kb_fail(h, KB_EINVAL, "%s program: negative table size", name);
  if (!q->ops || !q->br_chop || !q->br_bc || !q->blk_ptr || !q->blk_col || !q->blk_grp || !q->grp_part ||
      !q->grp_sign || !q->grp_nsc || !q->grp_sc || !q->grp_term || !q->term_coef || !q->term_op || (q->nbc > 0 && !q->bc))
    return kb_fail(h, KB_EINVAL, "%s program: null table", name);
  // the tables index each other: check every index on the host once (they are tiny)
  if (q->blk_ptr[0] != 0 || q->blk_ptr[q->nblockrows] != q->nblk)
    return kb_fail(h, KB_EINVAL, "%s program: blk_ptr does not span the blocks", name);
  for (int r = 0; r < q->nblockrows; ++r) {
    if (q->blk_ptr[r + 1] < q->blk_ptr[r]) return kb_fail(h, KB_EINVAL, "%s program: blk_ptr not monotone", name);
    for (int b = q->blk_ptr[r]; b < q->blk_ptr[r + 1]; ++b) {
      if (q->blk_col[b] < 0 || q->blk_col[b] >= q->nblockrows)
        return kb_fail(h, KB_EINVAL, "%s program: block column out of range", name);
      if (b > q->blk_ptr[r] && q->blk_col[b] <= q->blk_col[b - 1])
        return kb_fail(h, KB_EINVAL, "%s program: block columns of a block row must ascend", name);
    }
    const int chop = q->br_chop[r];
    if (chop < 0 || chop > q->N1) return kb_fail(h, KB_EINVAL, "%s program: bad boundary-row count", name);
    if (chop > 0 && (q->br_bc[r] < 0 || q->br_bc[r] + chop > q->nbc))
      return kb_fail(h, KB_EINVAL, "%s program: boundary rows out of range", name);
  }
  if (q->blk_grp[0] != 0 || q->blk_grp[q->nblk] != q->ngrp || q->grp_term[0] != 0 || q->grp_term[q->ngrp] != q->nterm)
    return kb_fail(h, KB_EINVAL, "%s program: group / term tables do not span", name);
  for (int b = 0; b < q->nblk; ++b)
    if (q->blk_grp[b + 1] < q->blk_grp[b]) return kb_fail(h, KB_EINVAL, "%s program: blk_grp not monotone", name);
  for (int g = 0; g < q->ngrp; ++g) {
    if (q->grp_term[g + 1] < q->grp_term[g]) return kb_fail(h, KB_EINVAL, "%s program: grp_term not monotone", name);
    if (q->grp_nsc[g] < 0 || q->grp_nsc[g] > 4) return kb_fail(h, KB_EINVAL, "%s program: more than 4 scalar factors", name);
    if (q->grp_part[g] != 0 && (q->grp_part[g] != 1 || !q->is_complex))
      return kb_fail(h, KB_EINVAL, "%s program: bad component in a group", name);
  }
  for (int t = 0; t < q->nterm; ++t)
    if (q->term_op[t] < 0 || q->term_op[t] >= q->nop) return kb_fail(h, KB_EINVAL, "%s program: operator index out of range", name);
  return KB_OK;
}

int ka_upload(kb_context* h, const kb_asm_program* q, KaDeviceProg& d) {
  std::vector<unsigned char> host;
  std::vector<size_t> o;
  const size_t W = 2 * (size_t)q->H + 1;
  ka_put(host, o, q->ops, (size_t)q->nop * q->N1 * W);
  ka_put(host, o, q->bc, (size_t)q->nbc * q->N1);
  ka_put(host, o, q->br_chop, (size_t)q->nblockrows);
  ka_put(host, o, q->br_bc, (size_t)q->nblockrows);
  ka_put(host, o, q->blk_ptr, (size_t)q->nblockrows + 1);
  ka_put(host, o, q->blk_col, (size_t)q->nblk);
  ka_put(host, o, q->blk_grp, (size_t)q->nblk + 1);
  ka_put(host, o, q->grp_part, (size_t)q->ngrp);
  ka_put(host, o, q->grp_sign, (size_t)q->ngrp);
  ka_put(host, o, q->grp_nsc, (size_t)q->ngrp);
  ka_put(host, o, q->grp_sc, (size_t)q->ngrp * 4);
  ka_put(host, o, q->grp_term, (size_t)q->ngrp + 1);
  ka_put(host, o, q->term_coef, (size_t)q->nterm);
  ka_put(host, o, q->term_op, (size_t)q->nterm);
  KB_CUDA(h, d.blob.alloc(host.size() + 16));
  KB_CUDA(h, cudaMemcpyAsync(d.blob.p, host.data(), host.size(), cudaMemcpyHostToDevice, h->stream));
  // pageable source: the tables must be on their way before `host` goes out of scope
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  const unsigned char* base = d.blob.p;
  KaProg& p = d.p;
  p.N1 = q->N1;
  p.nblockrows = q->nblockrows;
  p.H = q->H;
  p.W = (int)W;
  p.is_complex = q->is_complex != 0;
  p.use_final = q->use_final != 0;
  p.final_scale = q->final_scale;
  p.ops = (const double*)(base + o[0]);
  p.bc = (const double*)(base + o[1]);
  p.br_chop = (const int*)(base + o[2]);
  p.br_bc = (const int*)(base + o[3]);
  p.blk_ptr = (const int*)(base + o[4]);
  p.blk_col = (const int*)(base + o[5]);
  p.blk_grp = (const int*)(base + o[6]);
  p.grp_part = (const int*)(base + o[7]);
  p.grp_sign = (const int*)(base + o[8]);
  p.grp_nsc = (const int*)(base + o[9]);
  p.grp_sc = (const double*)(base + o[10]);
  p.grp_term = (const int*)(base + o[11]);
  p.term_coef = (const double*)(base + o[12]);
  p.term_op = (const int*)(base + o[13]);
  return KB_OK;
}

int ka_run(kb_context* h, const kb_asm_program* q, KbRawCSR& M, const char* name) {
  KB_TRY(ka_check(h, q, name));
  cudaStream_t s = h->stream;
  KaDeviceProg d;
  KB_TRY(ka_upload(h, q, d));
  const int64_t n = (int64_t)q->N1 * q->nblockrows;
  M.present = false;
  KB_CUDA(h, M.indptr.alloc(n + 1));
  DevBuf<int64_t> cnt;
  KB_CUDA(h, cnt.alloc(n + 1));
  KB_CUDA(h, cudaMemsetAsync(cnt.p + n, 0, sizeof(int64_t), s));
  const int thr = 256;
  const unsigned grid = nblk(n * 32, thr);
  ka_rows<false><<<grid, thr, 0, s>>>(d.p, cnt.p, nullptr, nullptr, nullptr);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  {
    size_t tb = 0;
    KB_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, M.indptr.p, (int)(n + 1), s));
    DevBuf<unsigned char> tmp;
    KB_CUDA(h, tmp.alloc(tb));
    KB_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, M.indptr.p, (int)(n + 1), s));
    int64_t nnz = 0;
    KB_CUDA(h, cudaMemcpyAsync(&nnz, M.indptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    KB_CUDA(h, cudaStreamSynchronize(s));  // also keeps `tmp` alive until the scan is done
    if (nnz >= 0x7fffffff) return kb_fail(h, KB_EINVAL, "%s: more than 2^31 nonzeros are not supported", name);
    M.nnz = nnz;
  }
  const bool cplx = q->is_complex != 0;
  KB_CUDA(h, M.indices.alloc((size_t)(M.nnz > 0 ? M.nnz : 1) * 4));
  KB_CUDA(h, M.values.alloc((size_t)(M.nnz > 0 ? M.nnz : 1) * (cplx ? 16 : 8)));
  ka_rows<true><<<grid, thr, 0, s>>>(d.p, nullptr, M.indptr.p, (int*)M.indices.p, (double*)M.values.p);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  KB_CUDA(h, cudaStreamSynchronize(s));  // the program tables (`d`) are freed on return
  M.n = n;
  M.index_bytes = 4;
  M.is_complex = cplx;
  M.present = true;
  return KB_OK;
}

}  // namespace

extern "C" int kb_assemble(kb_handle h, const kb_asm_program* A, const kb_asm_program* B) {
  if (!h) return KB_EINVAL;
  if (!A && !B) return kb_fail(h, KB_EINVAL, "kb_assemble: no program given");
  if (A && !A->is_complex) return kb_fail(h, KB_EINVAL, "kb_assemble: the program of A must be complex");
  const int64_t n = A ? (int64_t)A->N1 * A->nblockrows : (int64_t)B->N1 * B->nblockrows;
  if (A && B && (int64_t)B->N1 * B->nblockrows != n)
    return kb_fail(h, KB_EINVAL, "kb_assemble: the programs of A and B differ in size");
  KB_CUDA(h, cudaSetDevice(h->device));
  // a new pencil (exactly the matrices given here): whatever was ingested, laid out or factored
  // belongs to the old one
  h->chain_set = false;
  h->factored = false;
  h->A = HostCSR();
  h->B = HostCSR();
  h->rawA.present = false;
  h->rawB.present = false;
  h->n = n;
  if (A) KB_TRY(ka_run(h, A, h->rawA, "A"));
  if (B) {
    KB_TRY(ka_run(h, B, h->rawB, "B"));
    h->b_is_complex = B->is_complex != 0;
  }
  h->A.present = h->rawA.present;
  h->B.present = h->rawB.present;
  return KB_OK;
}

extern "C" int kb_get_assembled(kb_handle h, int which, int64_t* nnz, int64_t* indptr, int32_t* indices,
                                double* values) {
  if (!h || (which != 0 && which != 1)) return KB_EINVAL;
  const KbRawCSR& M = which == 0 ? h->rawA : h->rawB;
  if (!M.present) return kb_fail(h, KB_EINVAL, "kb_get_assembled: no %s on the device", which == 0 ? "A" : "B");
  if (M.index_bytes != 4) return kb_fail(h, KB_EINVAL, "kb_get_assembled: 64-bit column indices");
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  if (nnz) *nnz = M.nnz;
  if (indptr) KB_CUDA(h, cudaMemcpyAsync(indptr, M.indptr.p, (size_t)(M.n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  if (indices && M.nnz > 0)
    KB_CUDA(h, cudaMemcpyAsync(indices, M.indices.p, (size_t)M.nnz * 4, cudaMemcpyDeviceToHost, s));
  if (values && M.nnz > 0)
    KB_CUDA(h, cudaMemcpyAsync(values, M.values.p, (size_t)M.nnz * (M.is_complex ? 16 : 8), cudaMemcpyDeviceToHost, s));
  KB_CUDA(h, cudaStreamSynchronize(s));
  return KB_OK;
}
