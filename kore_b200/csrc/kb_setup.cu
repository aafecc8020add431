// Context life-cycle, pencil ingestion and the chain (l-major) layout build.
//
// Replaces: PETSc Mat create / setValuesCSR / assembly for A and B
// (/root/reference/bin/solve.py:43-59, 69-85) and the analysis (ordering)
// phase of the sparse direct solver (PCSetUp(LU) inside E.solve(),
// solve.py:123).  The ordering is not computed: the caller hands in Kore's own
// l-major chain (kore_b200/chain.py), which makes A - sigma B block tridiagonal.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <thread>

#include "kb_internal.cuh"

int kb_fail(kb_context* h, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

static thread_local std::string g_create_err;

extern "C" int kb_create(kb_handle* out, int device) {
  if (!out) return KB_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = "no CUDA device visible (libkoreb200 has no CPU fallback)";
    return KB_ENODEVICE;
  }
  if (device < 0 || device >= ndev) {
    g_create_err = "device index out of range";
    return KB_ENODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
    g_create_err = "device is not sm_100 class; this library is built for sm_100a only";
    return KB_ENODEVICE;
  }
  kb_context* h = new kb_context();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_err = "cannot create CUDA stream";
    delete h;
    return KB_ECUDA;
  }
  {
    // temporaries of the layout build come from the stream-ordered pool; keep what it has
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  *out = h;
  return KB_OK;
}

extern "C" int kb_destroy(kb_handle h) {
  if (!h) return KB_OK;
  cudaSetDevice(h->device);
  if (h->stream) {
    // bounded: a kernel that does not even react to the watchdog flag must not block the host in
    // here for ever (cudaFree would); the context is leaked and the error returned instead
    if (kbi_sync(h) != KB_OK) {
      fprintf(stderr, "libkoreb200: kb_destroy: the handle's stream does not drain; leaking the context\n");
      return KB_ECUDA;
    }
    cudaStreamDestroy(h->stream);
  }
  if (h->stream2) {
    cudaStreamSynchronize(h->stream2);
    cudaStreamDestroy(h->stream2);
  }
  kbi_nccl_destroy(h);
  kbi_drop_graphs(h);
  if (h->pinned_h) cudaFreeHost(h->pinned_h);
  if (h->pinned_beta) cudaFreeHost(h->pinned_beta);
  if (h->pinned_err) cudaFreeHost(h->pinned_err);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  delete h;
  return KB_OK;
}

extern "C" const char* kb_last_error(kb_handle h) {
  if (!h) return g_create_err.c_str();
  return h->err.c_str();
}

extern "C" int kb_set_option(kb_handle h, int option, int64_t value) {
  if (!h) return KB_EINVAL;
  switch (option) {
    case KB_OPT_EQUILIBRATE: h->opt_equil = value != 0; break;
    case KB_OPT_REFINE: h->opt_refine = (int)std::max<int64_t>(0, value); break;
    case KB_OPT_PURIFY: h->opt_purify = value != 0; break;
    case KB_OPT_SEED: h->opt_seed = value; break;
    case KB_OPT_PANEL: h->opt_panel = (int)value; break;
    case KB_OPT_REFINE_EIGS: h->opt_refine_eigs = (int)std::max<int64_t>(-1, value); break;
    // The factors are laid out for the sweep that will read them (two-sided or one-sided,
    // transposed or row-major, folded couplings or not): changing any of these invalidates them.
    case KB_OPT_SWEEP:
      if (value < 0 || value > 2) return kb_fail(h, KB_EINVAL, "KB_OPT_SWEEP must be 0, 1 or 2");
      if (h->opt_sweep != (int)value) h->factored = false;
      h->opt_sweep = (int)value;
      break;
    case KB_OPT_FACTOR:
      if (h->opt_factor != (int)(value != 0)) h->factored = false;
      h->opt_factor = value != 0;
      break;
    case KB_OPT_FOLD:
      if (h->opt_fold != (int)(value != 0)) h->factored = false;
      h->opt_fold = value != 0;
      break;
    case KB_OPT_WAIT_MS:
      if (value < 1) return kb_fail(h, KB_EINVAL, "KB_OPT_WAIT_MS must be >= 1");
      h->wait_ns = (unsigned long long)value * 1000000ull;
      break;
    case KB_OPT_INJECT_FAULT: h->inject_fault = (int)value; break;
    default: return kb_fail(h, KB_EINVAL, "unknown option %d", option);
  }
  return KB_OK;
}

extern "C" int kb_get_stats(kb_handle h, kb_stats* out) {
  if (!h || !out) return KB_EINVAL;
  h->stats.kernel_launches = h->launches;
  *out = h->stats;
  return KB_OK;
}

extern "C" int kb_stream(kb_handle h, void** s) {
  if (!h || !s) return KB_EINVAL;
  *s = (void*)h->stream;
  return KB_OK;
}

template <typename F>
static void parallel_for(int64_t n, F f) {
  const int nthr = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < (1 << 16)) {
    f(0, n);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < nthr; ++t) pool.emplace_back([=]() { f(n * t / nthr, n * (t + 1) / nthr); });
  for (auto& th : pool) th.join();
}

static int ingest(kb_context* h, HostCSR& M, int64_t n, int index_bytes, const void* indptr,
                  const void* indices, const void* values, bool is_complex, const char* name) {
  M.n = n;
  M.indptr.resize(n + 1);
  if (index_bytes == 4) {
    const int32_t* p = (const int32_t*)indptr;
    for (int64_t i = 0; i <= n; ++i) M.indptr[i] = p[i];
  } else {
    const int64_t* p = (const int64_t*)indptr;
    for (int64_t i = 0; i <= n; ++i) M.indptr[i] = p[i];
  }
  if (M.indptr[0] != 0) return kb_fail(h, KB_EINVAL, "%s: indptr[0] != 0", name);
  for (int64_t i = 0; i < n; ++i)
    if (M.indptr[i + 1] < M.indptr[i]) return kb_fail(h, KB_EINVAL, "%s: indptr not monotone", name);
  const int64_t nnz = M.indptr[n];
  M.indices.resize(nnz);
  M.values.resize(nnz);
  std::vector<int> badv(64, 0);
  int64_t* idst = M.indices.data();
  zcomplex* vdst = M.values.data();
  int* badp = badv.data();
  parallel_for(nnz, [=](int64_t lo, int64_t hi) {
    int bad = 0;
    if (index_bytes == 4) {
      const int32_t* p = (const int32_t*)indices;
      for (int64_t k = lo; k < hi; ++k) {
        idst[k] = p[k];
        bad |= (p[k] < 0 || p[k] >= n);
      }
    } else {
      const int64_t* p = (const int64_t*)indices;
      for (int64_t k = lo; k < hi; ++k) {
        idst[k] = p[k];
        bad |= (p[k] < 0 || p[k] >= n);
      }
    }
    const double* v = (const double*)values;
    if (is_complex)
      for (int64_t k = lo; k < hi; ++k) vdst[k] = zcomplex(v[2 * k], v[2 * k + 1]);
    else
      for (int64_t k = lo; k < hi; ++k) vdst[k] = zcomplex(v[k], 0.0);
    if (bad) badp[(lo >> 10) & 63] = 1;
  });
  for (int b : badv)
    if (b) return kb_fail(h, KB_EINVAL, "%s: column index out of range", name);
  M.present = true;
  return KB_OK;
}

extern "C" int kb_set_pencil(kb_handle h, int64_t n, int index_bytes, const void* a_indptr,
                             const void* a_indices, const double* a_values, const void* b_indptr,
                             const void* b_indices, const void* b_values, int b_is_complex) {
  if (!h) return KB_EINVAL;
  if (n <= 0 || n > 0x7fffffff) return kb_fail(h, KB_EINVAL, "n out of range");
  if (index_bytes != 4 && index_bytes != 8) return kb_fail(h, KB_EINVAL, "index_bytes must be 4 or 8");
  if (!a_indptr || !a_indices || !a_values) return kb_fail(h, KB_EINVAL, "A is required");
  h->n = n;
  h->chain_set = false;
  h->factored = false;
  h->B = HostCSR();
  h->A = HostCSR();
  h->rawA.present = false;
  h->rawB.present = false;
  h->b_is_complex = b_is_complex != 0;
  if (b_indptr && (!b_indices || !b_values)) return kb_fail(h, KB_EINVAL, "B indices/values missing");
  if (!getenv("KB_HOST_LAYOUT")) {
    // the CSR triplets go to the device as they are; kb_set_chain builds the layout there
    KB_CUDA(h, cudaSetDevice(h->device));
    KB_TRY(kbi_upload_raw(h, h->rawA, n, index_bytes, a_indptr, a_indices, a_values, true, "A"));
    h->A.present = true;
    if (b_indptr) {
      KB_TRY(kbi_upload_raw(h, h->rawB, n, index_bytes, b_indptr, b_indices, b_values, b_is_complex != 0, "B"));
      h->B.present = true;
    }
    return KB_OK;
  }
  KB_TRY(ingest(h, h->A, n, index_bytes, a_indptr, a_indices, a_values, true, "A"));
  if (b_indptr) KB_TRY(ingest(h, h->B, n, index_bytes, b_indptr, b_indices, b_values, b_is_complex != 0, "B"));
  return KB_OK;
}

namespace {
struct Ent {
  int col;
  int src;  // 0: A entry, 1: B entry
  zcomplex v;
};

template <typename T>
cudaError_t upload(DevBuf<T>& d, const std::vector<T>& v, cudaStream_t s) {
  cudaError_t e = d.alloc(v.size());
  if (e != cudaSuccess) return e;
  if (v.empty()) return cudaSuccess;
  return cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

template <typename F>
void parallel_rows(int64_t n, int nthr, F f) {
  std::vector<std::thread> pool;
  for (int t = 0; t < nthr; ++t)
    pool.emplace_back([=]() {
      int64_t lo = n * t / nthr, hi = n * (t + 1) / nthr;
      f(t, lo, hi);
    });
  for (auto& th : pool) th.join();
}
}  // namespace

extern "C" int kb_set_chain(kb_handle h, const int64_t* perm, const int64_t* nodeptr, int64_t nnodes) {
  if (!h) return KB_EINVAL;
  if (!h->A.present) return kb_fail(h, KB_EINVAL, "kb_set_pencil must be called first");
  const int64_t n = h->n;
  if (!perm || !nodeptr || nnodes <= 0) return kb_fail(h, KB_EINVAL, "bad chain arguments");
  if (nodeptr[0] != 0 || nodeptr[nnodes] != n) return kb_fail(h, KB_EINVAL, "nodeptr must span [0,n]");
  KB_CUDA(h, cudaSetDevice(h->device));
  h->P = nnodes;
  h->nodeptr.assign(nodeptr, nodeptr + nnodes + 1);
  h->perm.assign(perm, perm + n);
  h->bmax = 0;
  h->bmin = n;
  for (int64_t p = 0; p < nnodes; ++p) {
    int64_t b = nodeptr[p + 1] - nodeptr[p];
    if (b <= 0) return kb_fail(h, KB_EINVAL, "empty chain node %lld", (long long)p);
    h->bmax = std::max(h->bmax, b);
    h->bmin = std::min(h->bmin, b);
  }
  if (h->rawA.present) {
    KB_TRY(kbi_layout_device(h));
    h->chain_set = true;
    h->factored = false;
    return KB_OK;
  }
  std::vector<int> iperm(n, -1);
  for (int64_t k = 0; k < n; ++k) {
    if (perm[k] < 0 || perm[k] >= n || iperm[perm[k]] != -1)
      return kb_fail(h, KB_EINVAL, "perm is not a permutation");
    iperm[perm[k]] = (int)k;
  }
  std::vector<int> node_of(n);
  for (int64_t p = 0; p < nnodes; ++p)
    for (int64_t i = nodeptr[p]; i < nodeptr[p + 1]; ++i) node_of[i] = (int)p;

  const HostCSR& A = h->A;
  const HostCSR& B = h->B;
  const int nthr = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));

  // merged (union of A and B), column-sorted row i of the chain-ordered pencil
  auto merged_row = [&](int64_t i, std::vector<Ent>& buf) {
    const int64_t o = perm[i];
    buf.clear();
    for (int64_t k = A.indptr[o]; k < A.indptr[o + 1]; ++k) buf.push_back(Ent{iperm[A.indices[k]], 0, A.values[k]});
    if (B.present)
      for (int64_t k = B.indptr[o]; k < B.indptr[o + 1]; ++k) buf.push_back(Ent{iperm[B.indices[k]], 1, B.values[k]});
    std::sort(buf.begin(), buf.end(), [](const Ent& x, const Ent& y) {
      return x.col < y.col || (x.col == y.col && x.src < y.src);
    });
  };

  // ---- pass 1: row lengths of the union pattern, structure check
  std::vector<int64_t> rowptr(n + 1, 0), browptr(n + 1, 0);
  std::vector<int> bad(nthr, 0);
  parallel_rows(n, nthr, [&](int t, int64_t lo, int64_t hi) {
    // union length without sorting: a per-thread marker remembers which chain columns the
    // row's A entries (and earlier B entries) have already claimed
    std::vector<int64_t> mark(n, -1);
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t o = perm[i];
      const int pi = node_of[i];
      int64_t cnt = 0, cntb = 0;
      for (int64_t k = A.indptr[o]; k < A.indptr[o + 1]; ++k) {
        const int c = iperm[A.indices[k]];
        if (mark[c] != i) {
          mark[c] = i;
          ++cnt;
          int d = node_of[c] - pi;
          if (d > 1 || d < -1) bad[t] = 1;
        }
      }
      if (B.present)
        for (int64_t k = B.indptr[o]; k < B.indptr[o + 1]; ++k) {
          const int c = iperm[B.indices[k]];
          ++cntb;
          if (mark[c] != i) {
            mark[c] = i;
            ++cnt;
            int d = node_of[c] - pi;
            if (d > 1 || d < -1) bad[t] = 1;
          }
        }
      rowptr[i + 1] = cnt;
      browptr[i + 1] = cntb;
    }
  });
  for (int t = 0; t < nthr; ++t)
    if (bad[t])
      return kb_fail(h, KB_ESTRUCTURE,
                     "pencil is not block tridiagonal under the given chain (a nonzero couples nodes "
                     "more than one apart)");
  for (int64_t i = 0; i < n; ++i) {
    rowptr[i + 1] += rowptr[i];
    browptr[i + 1] += browptr[i];
  }
  const int64_t nnz = rowptr[n], nnzB = browptr[n];
  h->nnz = nnz;
  h->nnzB = nnzB;

  // ---- pass 2: fill
  std::vector<int> col((size_t)nnz), bcol((size_t)nnzB), bmap((size_t)nnzB);
  std::vector<double2> aval((size_t)nnz);
  std::vector<double2> bvc;
  std::vector<double> bvr;
  if (h->b_is_complex)
    bvc.resize((size_t)nnzB);
  else
    bvr.resize((size_t)nnzB);
  std::vector<int64_t> dstart(n), ustart(n);
  std::vector<int64_t> wlt(nthr, 0), wut(nthr, 0);
  parallel_rows(n, nthr, [&](int t, int64_t lo, int64_t hi) {
    std::vector<Ent> buf;
    buf.reserve(256);
    for (int64_t i = lo; i < hi; ++i) {
      merged_row(i, buf);
      int64_t k = rowptr[i] - 1, kb = browptr[i];
      int last = -1;
      const int pi = node_of[i];
      int64_t ds = -1, us = -1;
      for (const Ent& e : buf) {
        if (e.col != last) {
          ++k;
          last = e.col;
          col[k] = e.col;
          aval[k] = make_double2(0.0, 0.0);
          const int q = node_of[e.col];
          if (q >= pi && ds < 0) ds = k;
          if (q > pi && us < 0) us = k;
        }
        if (e.src == 0) {
          aval[k].x += e.v.real();
          aval[k].y += e.v.imag();
        } else {
          bcol[kb] = e.col;
          bmap[kb] = (int)k;
          if (h->b_is_complex)
            bvc[kb] = make_double2(e.v.real(), e.v.imag());
          else
            bvr[kb] = e.v.real();
          ++kb;
        }
      }
      if (us < 0) us = rowptr[i + 1];
      if (ds < 0) ds = us;
      dstart[i] = ds;
      ustart[i] = us;
      wlt[t] = std::max<int64_t>(wlt[t], ds - rowptr[i]);
      wut[t] = std::max<int64_t>(wut[t], rowptr[i + 1] - us);
    }
  });
  int64_t wl = 0, wu = 0;
  for (int t = 0; t < nthr; ++t) {
    wl = std::max(wl, wlt[t]);
    wu = std::max(wu, wut[t]);
  }
  if (nnz > 0x7fffffff) return kb_fail(h, KB_EINVAL, "more than 2^31 nonzeros are not supported");

  // ---- couplings by column (U: row node p, column node p+1; L: column node p-1)
  std::vector<int64_t> ucount(n + 1, 0), lcount(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    for (int64_t k = rowptr[i]; k < dstart[i]; ++k) lcount[col[k] + 1]++;
    for (int64_t k = ustart[i]; k < rowptr[i + 1]; ++k) ucount[col[k] + 1]++;
  }
  for (int64_t c = 0; c < n; ++c) {
    ucount[c + 1] += ucount[c];
    lcount[c + 1] += lcount[c];
  }
  const int64_t nnzU = ucount[n], nnzL = lcount[n];
  h->nnzU = nnzU;
  std::vector<int> urow((size_t)nnzU), lrow((size_t)nnzL);
  std::vector<int64_t> upos((size_t)nnzU), lpos((size_t)nnzL);
  {
    std::vector<int64_t> fu(ucount.begin(), ucount.end() - 1), fl(lcount.begin(), lcount.end() - 1);
    for (int64_t i = 0; i < n; ++i) {
      for (int64_t k = rowptr[i]; k < dstart[i]; ++k) {
        int64_t dst = fl[col[k]]++;
        lrow[dst] = (int)i;
        lpos[dst] = k;
      }
      for (int64_t k = ustart[i]; k < rowptr[i + 1]; ++k) {
        int64_t dst = fu[col[k]]++;
        urow[dst] = (int)i;
        upos[dst] = k;
      }
    }
  }

  std::vector<int> perm32(n);
  for (int64_t k = 0; k < n; ++k) perm32[k] = (int)perm[k];

  cudaStream_t s = h->stream;
  KB_CUDA(h, upload(h->d_perm, perm32, s));
  KB_CUDA(h, upload(h->d_rowptr, rowptr, s));
  KB_CUDA(h, upload(h->d_col, col, s));
  KB_CUDA(h, upload(h->d_Aval, aval, s));
  KB_CUDA(h, h->d_Tval.alloc(nnz));
  KB_CUDA(h, upload(h->d_dstart, dstart, s));
  KB_CUDA(h, upload(h->d_ustart, ustart, s));
  KB_CUDA(h, upload(h->d_ucptr, ucount, s));
  KB_CUDA(h, upload(h->d_urow, urow, s));
  KB_CUDA(h, upload(h->d_upos, upos, s));
  KB_CUDA(h, upload(h->d_lcptr, lcount, s));
  KB_CUDA(h, upload(h->d_lrow, lrow, s));
  KB_CUDA(h, upload(h->d_lpos, lpos, s));
  KB_CUDA(h, upload(h->d_browptr, browptr, s));
  KB_CUDA(h, upload(h->d_bcol, bcol, s));
  KB_CUDA(h, upload(h->d_bmap, bmap, s));
  if (h->b_is_complex)
    KB_CUDA(h, upload(h->d_bval_c, bvc, s));
  else
    KB_CUDA(h, upload(h->d_bval_r, bvr, s));
  KB_CUDA(h, h->d_rscale.alloc(n));
  KB_CUDA(h, h->d_cscale.alloc(n));
  KB_CUDA(h, h->d_maxbits.alloc(n));
  KB_CUDA(h, cudaStreamSynchronize(s));

  h->WL = (int)wl;
  h->WU = (int)wu;
  h->chain_set = true;
  h->factored = false;
  return KB_OK;
}
