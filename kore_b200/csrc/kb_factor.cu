// Numeric factorisation of T = A - sigma B in chain (block-tridiagonal) layout.
//
// Replaces: STSetUp (T = A - sigma B, MatAXPY) + KSPSetUp + PCSetUp(LU) numeric
// phase of MUMPS / SuperLU_DIST, i.e. the dominant cost inside E.solve()
// (/root/reference/bin/solve.py:123) and K.solve (solve.py:227).
//
// Algorithm (block Thomas with explicit inverses):
//     S_0 = D_0,   M_p = S_p^{-1},   S_{p+1} = D_{p+1} - L_{p+1,p} (M_p U_{p,p+1})
// D_p / L / U are the diagonal / sub / super blocks of the chain CSR; the
// couplings L, U are sparse (banded), so the only O(b^3) work per node is the
// in-place blocked Gauss-Jordan inversion of the dense Schur block S_p with
// partial (row) pivoting inside the block:  8 b^3 real flops per node.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "kb_internal.cuh"

// ---------------------------------------------------------------------------
// T = A - sigma B, equilibration
// ---------------------------------------------------------------------------
// T -= sigma B on the union pattern (T starts as a copy of A; every B entry has its own slot)
template <typename VT>
__global__ void kb_sub_sigma_B(int64_t nnzB, const VT* __restrict__ bval, const int* __restrict__ bmap,
                               double2 sigma, double2* __restrict__ T) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnzB) return;
  double2 t = T[bmap[k]];
  zfms(t, sigma, kb_as_complex(bval[k]));
  T[bmap[k]] = t;
}

__device__ __forceinline__ double kb_pow2_recip(double m) {
  // largest power of two r with r*m < 1 (m > 0), r = 1 for m == 0 / non-finite
  if (!(m > 0.0) || isinf(m)) return 1.0;
  int e;
  frexp(m, &e);
  return ldexp(1.0, -e);
}

// one warp per row: r = 2^-e(max|T_row|), row scaled in place
__global__ void kb_row_scale(int n, const int64_t* __restrict__ rowptr, double2* __restrict__ T,
                             double* __restrict__ rscale, int enable) {
  int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (row >= n) return;
  int64_t a = rowptr[row], b = rowptr[row + 1];
  double m = 0.0;
  for (int64_t k = a + lane; k < b; k += 32) {
    double2 v = T[k];
    m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  double r = enable ? kb_pow2_recip(m) : 1.0;
  if (lane == 0) rscale[row] = r;
  if (r != 1.0)
    for (int64_t k = a + lane; k < b; k += 32) T[k] = zscale(T[k], r);
}

__global__ void kb_col_max(int64_t nnz, const int* __restrict__ col, const double2* __restrict__ T,
                           unsigned long long* __restrict__ maxbits) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  double2 v = T[k];
  double m = fmax(fabs(v.x), fabs(v.y));
  atomicMax(&maxbits[col[k]], (unsigned long long)__double_as_longlong(m));
}

__global__ void kb_col_scale_vec(int n, const unsigned long long* __restrict__ maxbits,
                                 double* __restrict__ cscale, int enable) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m = __longlong_as_double((long long)maxbits[i]);
  cscale[i] = enable ? kb_pow2_recip(m) : 1.0;
}

__global__ void kb_col_apply(int64_t nnz, const int* __restrict__ col, const double* __restrict__ cscale,
                             double2* __restrict__ T) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  double c = cscale[col[k]];
  if (c != 1.0) T[k] = zscale(T[k], c);
}

// ---------------------------------------------------------------------------
// Schur block (one CTA per row):
//   S[i,:] = D_p[i,:] - sum_k L_{p,p-1}[i,k] W1[k,:] - sum_k U_{p,p+1}[i,k] W2[k,:]
// W1 = M_{p-1} U_{p-1,p} (downward chain), W2 = M'_{p+1} L_{p+1,p} (upward chain);
// either may be null.  The first Gauss-Jordan panel is also written column-major.
// ---------------------------------------------------------------------------
__global__ void kb_schur_row(double2* __restrict__ S, double2* __restrict__ PT, int nb0, int b, int o,
                             const double2* __restrict__ W1, int o1, const double2* __restrict__ W2, int o2,
                             const int64_t* __restrict__ rowptr, const int64_t* __restrict__ dstart,
                             const int64_t* __restrict__ ustart, const int* __restrict__ col,
                             const double2* __restrict__ T) {
  const int i = blockIdx.x;
  const int gi = o + i;
  const int64_t l0 = rowptr[gi], d0 = dstart[gi], u0 = ustart[gi], e0 = rowptr[gi + 1];
  double2* Srow = S + (size_t)i * b;
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    double2 acc = zmake(0.0, 0.0);
    if (W1) {
      for (int64_t k = l0; k < d0; ++k) {
        double2 l = __ldg(&T[k]);
        int c = __ldg(&col[k]) - o1;
        zfms(acc, l, W1[(size_t)c * b + j]);
      }
    }
    if (W2) {
      for (int64_t k = u0; k < e0; ++k) {
        double2 u = __ldg(&T[k]);
        int c = __ldg(&col[k]) - o2;
        zfms(acc, u, W2[(size_t)c * b + j]);
      }
    }
    Srow[j] = acc;
  }
  __syncthreads();
  for (int64_t k = d0 + threadIdx.x; k < u0; k += blockDim.x) {
    int c = col[k] - o;
    Srow[c] = zadd(Srow[c], T[k]);
  }
  // first Gauss-Jordan panel, column-major, for kb_gj_panel
  __syncthreads();
  for (int j = threadIdx.x; j < nb0 && j < b; j += blockDim.x) PT[(size_t)j * b + i] = Srow[j];
}

// W = M_p U_{p,p+1}: W[i,j] = sum_{k in col j of U} M[i,k] U[k,j]   (one CTA per row i)
__global__ void kb_w_rows(const double2* __restrict__ M, int b, int o, double2* __restrict__ W,
                          int bnext, int onext, const int64_t* __restrict__ ucptr,
                          const int* __restrict__ urow, const int64_t* __restrict__ upos,
                          const double2* __restrict__ T) {
  extern __shared__ double2 mrow[];
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < b; j += blockDim.x) mrow[j] = M[(size_t)i * b + j];
  __syncthreads();
  for (int j = threadIdx.x; j < bnext; j += blockDim.x) {
    int c = onext + j;
    double2 acc = zmake(0.0, 0.0);
    for (int64_t e = ucptr[c]; e < ucptr[c + 1]; ++e) zfma(acc, mrow[urow[e] - o], T[upos[e]]);
    W[(size_t)i * bnext + j] = acc;
  }
}

// ---------------------------------------------------------------------------
// Blocked in-place Gauss-Jordan inversion with partial pivoting
// ---------------------------------------------------------------------------
// Panel step: columns [k0, k0+nbv) of all n rows live in registers, one (or RPT)
// row(s) per thread; they arrive in the column-major side buffer PT (NB x n) that
// the previous update step wrote, so loads and stores are coalesced.  On exit GpT
// (NB x n) holds the nontrivial columns
// of the composite Gauss-Jordan transform, srcrow[i] the pre-panel row that now
// sits at position i, and orig[] the running row permutation of the whole
// inversion.
//
// One column step = pivot search (integer keys: the high word of |a|^2, reduced
// with redux.sync -- any entry within 2^-20 of the column maximum is an equally
// good partial pivot), publish the pivot row through shared memory, eliminate.
// The column loop is unrolled by two only and the register row is rotated by two
// slots per trip, so the loop body stays small enough for the instruction cache
// (a fully unrolled NB = 16 body is 140 KB of SASS and ran 2x slower).
// Shared state of the panel kernel.  prow/swp/best are double-buffered by column parity
// so that a column needs only TWO block barriers (after the pivot vote, after the
// pivot-row publication): the buffers written for column k+1 are not the ones column k
// is still reading.
template <int NB>
struct KbPanelShared {
  double2 prow[2][NB];
  double2 swp[2][NB];
  unsigned long long best[2];  // (key << 32) | ~row : atomicMax picks the largest key, lowest row
};

template <int NB, int RPT, int C>
__device__ __forceinline__ void kb_gj_column(double2 (&a)[RPT][NB], const int gk, const int n, const int t,
                                             const int T, const int lane, KbPanelShared<NB>& sh, int* s_src,
                                             int* s_orig, int* info) {
  const int par = gk & 1;
  // ---- pivot vote over positions >= gk: integer key = high word of |a|^2 (+1), packed with
  //      the row so that one shared-memory atomicMax per warp elects the pivot
  unsigned key = 0u;
  int bi = 0x7fffffff;
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int i = t + r * T;
    if (i >= gk && i < n) {
      unsigned kk = (unsigned)__double2hiint(zabs2(a[r][C])) + 1u;
      if (kk > key) {
        key = kk;
        bi = i;
      }
    }
  }
#ifdef KB_ABL_NOSEARCH
  key = (t == 0) ? 5u : 0u;
  bi = gk;
#endif
  {
    unsigned long long packed = ((unsigned long long)key << 32) | (unsigned)(~(unsigned)bi);
    // warp-level max of the packed value (two 32-bit redux steps keep it exact)
    unsigned wmax = __reduce_max_sync(0xffffffffu, key);
    unsigned lo = (key == wmax) ? (unsigned)(~(unsigned)bi) : 0u;
    unsigned wlo = __reduce_max_sync(0xffffffffu, lo);
    packed = ((unsigned long long)wmax << 32) | wlo;
    if (lane == 0 && wmax != 0u) atomicMax(&sh.best[par], packed);
  }
  __syncthreads();
  const unsigned long long best = sh.best[par];
  const unsigned gmax = (unsigned)(best >> 32);
  const int rp = (int)(~(unsigned)(best & 0xffffffffu));
  // key 0: no candidate; 1: |pivot|^2 is zero/denormal; >= 0x7ff00001: inf or NaN
  const bool broken = gmax <= 1u || gmax >= 0x7ff00001u || rp >= n || rp < gk;
  const int rps = broken ? gk : rp;  // keep going on breakdown; host reports KB_ESINGULAR
  // ---- publish the (unscaled) pivot row, and the row being displaced from position gk
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int i = t + r * T;
    if (i == rps) {
#pragma unroll
      for (int j = 0; j < NB; ++j) sh.prow[par][j] = a[r][j];
    }
    if (i == gk && rps != gk) {
#pragma unroll
      for (int j = 0; j < NB; ++j) sh.swp[par][j] = a[r][j];
    }
  }
  // bookkeeping by the last thread (its warp owns the fewest live rows); it also re-arms the
  // vote slot that column gk+1 will use
  if (t == T - 1) {
    if (broken) atomicExch(info, gk + 1);
    int tmp = s_src[gk];
    s_src[gk] = s_src[rps];
    s_src[rps] = tmp;
    int to = s_orig[gk];
    s_orig[gk] = s_orig[rps];
    s_orig[rps] = to;
    sh.best[par ^ 1] = 0ull;
  }
  __syncthreads();
  // ---- eliminate.  Every thread forms 1/pivot itself (no serial section) and the pivot
  //      row goes through the same arithmetic as the others with (base, g) = (0, -1/pivot),
  //      so there is no divergent branch:  a[j] = base[j] - g * prow[j]
  const double2 pinv = zinv_fast(sh.prow[par][C]);
  double2 g[RPT];
  bool live[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int i = t + r * T;
    live[r] = i < n;
    const bool is_piv = (i == gk);
    if (i == rps && rps != gk) {
#pragma unroll
      for (int j = 0; j < NB; ++j) a[r][j] = sh.swp[par][j];
    }
    g[r] = is_piv ? zneg(pinv) : zmul(a[r][C], pinv);
    if (is_piv) {
#pragma unroll
      for (int j = 0; j < NB; ++j) a[r][j] = zmake(0.0, 0.0);
    }
  }
#ifndef KB_ABL_NOELIM
  // the pivot row is read in chunks of four entries (a compiler barrier between chunks keeps
  // ptxas from hoisting all sixteen loads and spilling the register rows)
#pragma unroll
  for (int j0 = 0; j0 < NB; j0 += 4) {
    double2 pr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) pr[u] = sh.prow[par][j0 + u];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (live[r]) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u != C) zfms(a[r][j0 + u], g[r], pr[u]);
      }
    }
    asm volatile("" ::: "memory");
  }
#endif
#pragma unroll
  for (int r = 0; r < RPT; ++r)
    if (live[r]) a[r][C] = zneg(g[r]);
}

// Body of the panel step (see kb_gj_column).  COHERENT: the inputs were written by other
// CTAs of the SAME kernel (fused update + look-ahead panel), so they are read through L2.
template <int NB, int RPT, bool COHERENT>
__device__ __forceinline__ void kb_panel_body(const double2* PT, int n, int k0, int nbv, double2* __restrict__ GpT,
                                              int* orig, int* __restrict__ srcrow, int* __restrict__ info,
                                              long long* __restrict__ dbg, int* s_src) {
  const long long c0 = clock64();
  int* s_orig = s_src + n;
  __shared__ KbPanelShared<NB> sh;
  const int T = blockDim.x;
  const int t = threadIdx.x;
  const int lane = t & 31;

  double2 a[RPT][NB];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int i = t + r * T;
#pragma unroll
    for (int j = 0; j < NB; ++j)
      a[r][j] = (i < n && j < nbv) ? (COHERENT ? __ldcg(&PT[(size_t)j * n + i]) : PT[(size_t)j * n + i]) : zmake(0.0, 0.0);
  }
  for (int i = t; i < n; i += T) {
    s_src[i] = i;
    s_orig[i] = (k0 == 0) ? i : (COHERENT ? __ldcg(&orig[i]) : orig[i]);
  }
  if (t == 0) {
    sh.best[0] = 0ull;
    sh.best[1] = 0ull;
  }
  __syncthreads();

  const long long c1 = clock64();
  int rot = 0;
#pragma unroll 1
  for (int k = 0; k < nbv; k += 2) {
    kb_gj_column<NB, RPT, 0>(a, k0 + k, n, t, T, lane, sh, s_src, s_orig, info);
    if (k + 1 < nbv)
      kb_gj_column<NB, RPT, 1>(a, k0 + k + 1, n, t, T, lane, sh, s_src, s_orig, info);
    // rotate the register row left by two: the next active columns are slots 0 and 1 again
#ifndef KB_ABL_NOROT
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      double2 f0 = a[r][0], f1 = a[r][1];
#pragma unroll
      for (int j = 0; j + 2 < NB; ++j) a[r][j] = a[r][j + 2];
      a[r][NB - 2] = f0;
      a[r][NB - 1] = f1;
    }
#endif
    rot += 2;
  }
  __syncthreads();
  const long long c2 = clock64();
  // register slot j now holds panel column (j + rot) mod NB
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int i = t + r * T;
    if (i < n) {
#pragma unroll
      for (int j = 0; j < NB; ++j) GpT[(size_t)((j + rot) % NB) * n + i] = a[r][j];
    }
  }
  for (int i = t; i < n; i += T) {
    srcrow[i] = s_src[i];
    orig[i] = s_orig[i];
  }
  if (dbg && t == 0) {
    const long long c3 = clock64();
    atomicAdd((unsigned long long*)&dbg[0], (unsigned long long)(c1 - c0));
    atomicAdd((unsigned long long*)&dbg[1], (unsigned long long)(c2 - c1));
    atomicAdd((unsigned long long*)&dbg[2], (unsigned long long)(c3 - c2));
    atomicAdd((unsigned long long*)&dbg[3], 1ULL);
  }
}

template <int NB, int RPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
kb_gj_panel(const double2* __restrict__ PT, int n, int k0, int nbv, double2* __restrict__ GpT,
            int* __restrict__ orig, int* __restrict__ srcrow, int* __restrict__ info,
            long long* __restrict__ dbg) {
  extern __shared__ int s_dyn[];  // n ints: pre-panel row at each position; then n ints: orig
  kb_panel_body<NB, RPT, false>(PT, n, k0, nbv, GpT, orig, srcrow, info, dbg, s_dyn);
}

#ifndef KB_UPDATE_TM
#define KB_UPDATE_TM 64
#endif
// Update step (out of place): for every non-panel column j
//   Aout[i,j] = (i in panel rows ? 0 : Ain[src(i),j]) + sum_k Gp[i,k] Ain[src(k0+k),j]
// and Aout[i, panel] = Gp[i,:].   Tile TM x TN per CTA, 256 threads, (TM/16) x 4 per thread.
template <int NB, int TM>
struct KbTileShared {
  double2 Gs[TM][NB];
  double2 Rs[NB][64];
  int src_s[TM];
};

// One TM x 64 tile of the rank-NB update; called by all threads of the CTA (it contains a
// block barrier), worked on by the first 256.
template <int NB, int TM>
__device__ __forceinline__ void kb_update_tile(const double2* __restrict__ Ain, double2* __restrict__ Aout, int n,
                                               int k0, int nbv, const double2* __restrict__ GpT,
                                               const int* __restrict__ srcrow, double2* __restrict__ PTnext,
                                               int row0, int col0, KbTileShared<NB, TM>& ts) {
  constexpr int TN = 64;
  constexpr int RM = TM / 16;  // rows per thread
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = (tid >> 4) & 15;
  if (tid < 256) {
    for (int e = tid; e < TM * NB; e += 256) {
      int k = e / TM, i = e % TM;
      int gi = row0 + i;
      ts.Gs[i][k] = (gi < n && k < nbv) ? GpT[(size_t)k * n + gi] : zmake(0.0, 0.0);
    }
    for (int e = tid; e < NB * TN; e += 256) {
      int k = e / TN, j = e % TN;
      int gj = col0 + j;
      ts.Rs[k][j] = (k < nbv && gj < n) ? Ain[(size_t)srcrow[k0 + k] * n + gj] : zmake(0.0, 0.0);
    }
    for (int i = tid; i < TM; i += 256) ts.src_s[i] = (row0 + i < n) ? srcrow[row0 + i] : 0;
  }
  __syncthreads();
  if (tid >= 256) return;

  double2 acc[RM][4];
#pragma unroll
  for (int a = 0; a < RM; ++a) {
    int i = ty + 16 * a;
    int gi = row0 + i;
    bool prow_ = (gi >= k0 && gi < k0 + nbv);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int gj = col0 + tx + 16 * c;
      acc[a][c] = (gi < n && gj < n && !prow_) ? Ain[(size_t)ts.src_s[i] * n + gj] : zmake(0.0, 0.0);
    }
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    double2 g[RM], rr[4];
#pragma unroll
    for (int a = 0; a < RM; ++a) g[a] = ts.Gs[ty + 16 * a][k];
#pragma unroll
    for (int c = 0; c < 4; ++c) rr[c] = ts.Rs[k][tx + 16 * c];
#pragma unroll
    for (int a = 0; a < RM; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) zfma(acc[a][c], g[a], rr[c]);
  }
#pragma unroll
  for (int a = 0; a < RM; ++a) {
    int i = ty + 16 * a;
    int gi = row0 + i;
    if (gi >= n) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int gj = col0 + tx + 16 * c;
      if (gj >= n) continue;
      double2 v = acc[a][c];
      if (gj >= k0 && gj < k0 + nbv) v = ts.Gs[i][gj - k0];
      Aout[(size_t)gi * n + gj] = v;
      // the next panel's columns also go, column-major, to the panel kernel's input buffer
      if (gj >= k0 + NB && gj < k0 + 2 * NB) PTnext[(size_t)(gj - k0 - NB) * n + gi] = v;
    }
  }
}

template <int NB, int TM>
__global__ void __launch_bounds__(256)
kb_gj_update(const double2* __restrict__ Ain, double2* __restrict__ Aout, int n, int k0, int nbv,
             const double2* __restrict__ GpT, const int* __restrict__ srcrow, double2* __restrict__ PTnext) {
  __shared__ KbTileShared<NB, TM> ts;
  kb_update_tile<NB, TM>(Ain, Aout, n, k0, nbv, GpT, srcrow, PTnext, blockIdx.y * TM, blockIdx.x * 64, ts);
}

// Fused step with look-ahead: the tiles of update k (Gp_k) and, as one extra CTA, the panel of
// step k+1.  The tiles of the column block that holds the next panel are scheduled first and
// count themselves in `ready`; the panel CTA starts as soon as they are done, so the rest of
// the update overlaps the (latency-bound, single-SM) panel instead of preceding it.
template <int NB, int RPT, int TM, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
kb_gj_fused(const double2* __restrict__ Ain, double2* __restrict__ Aout, int n, int k0, int nbv,
            const double2* __restrict__ GpT, const int* __restrict__ srcrow, double2* PT, int has_next,
            int nbv_next, double2* __restrict__ GpT_next, int* __restrict__ srcrow_next, int* orig,
            int* __restrict__ info, unsigned* ready, long long* __restrict__ dbg) {
  extern __shared__ int s_dyn[];
  __shared__ KbTileShared<NB, TM> ts;
  const int nrt = (n + TM - 1) / TM, nct = (n + 63) / 64, ntiles = nrt * nct;
  const int bid = blockIdx.x;
  if (bid < ntiles) {
    const int ct_next = (k0 + NB) / 64;
    int tc, tr;
    const bool prio = has_next && ct_next < nct;
    if (prio) {
      if (bid < nrt) {
        tc = ct_next;
        tr = bid;
      } else {
        int idx = bid - nrt;
        tc = idx / nrt;
        if (tc >= ct_next) ++tc;
        tr = idx % nrt;
      }
    } else {
      tc = bid / nrt;
      tr = bid % nrt;
    }
    kb_update_tile<NB, TM>(Ain, Aout, n, k0, nbv, GpT, srcrow, PT, tr * TM, tc * 64, ts);
    if (prio && tc == ct_next) {
      // all 256 workers of this tile have stored their part of the next panel
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ready, 1u);
      }
    }
  } else {
    if (threadIdx.x == 0) {
      int spins = 0;
      while (*(volatile unsigned*)ready < (unsigned)nrt && ++spins < (1 << 24)) {
      }
      if (*(volatile unsigned*)ready < (unsigned)nrt) atomicExch(info, n + 1);
      __threadfence();
    }
    __syncthreads();
    kb_panel_body<NB, RPT, true>(PT, n, k0 + NB, nbv_next, GpT_next, orig, srcrow_next, info, dbg, s_dyn);
    if (threadIdx.x == 0) *ready = 0u;
  }
}

// M[r, orig[i]] = X[r, i]  (undo the row pivoting as a column permutation)
__global__ void kb_store_inverse(const double2* __restrict__ X, int n, const int* __restrict__ orig,
                                 double2* __restrict__ M) {
  int r = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) M[(size_t)r * n + orig[i]] = X[(size_t)r * n + i];
}

// ---------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------
// Workspace + stream of one elimination chain (two chains run concurrently in the
// two-sided factorisation).
struct GjWs {
  cudaStream_t st;
  double2 *S0, *S1, *W, *Gp, *PT;
  int *orig, *srcrow;
  double2* Gp2;     // second panel buffer (look-ahead)
  int* srcrow2;
  unsigned* ready;  // next-panel tile counter of the fused step
};

GjWs kbi_ws_main(kb_context* h) {
  GjWs w;
  w.st = h->stream;
  w.S0 = h->d_S0.p;
  w.S1 = h->d_S1.p;
  w.W = h->d_W.p;
  w.Gp = h->d_Gp.p;
  w.PT = h->d_PT.p;
  w.orig = h->d_orig.p;
  w.srcrow = h->d_srcrow.p;
  w.Gp2 = h->d_Gp.p + (size_t)h->bmax * 16;
  w.srcrow2 = h->d_srcrow.p + h->bmax;
  w.ready = h->d_ready.p;
  return w;
}

template <int NB, int RPT, int MAXT>
static void launch_panel(kb_context* h, const GjWs& w, int n, int k0, int nbv) {
  int T = (n + RPT - 1) / RPT;
  T = ((T + 31) / 32) * 32;
  if (T > MAXT) T = MAXT;
  kb_gj_panel<NB, RPT, MAXT><<<1, T, 2 * n * sizeof(int), w.st>>>(w.PT, n, k0, nbv, w.Gp, w.orig, w.srcrow,
                                                                 h->d_info.p, h->d_sweep_timing.p);
}

template <int NB>
static void launch_update(kb_context* h, const GjWs& w, const double2* in, double2* out, int n, int k0, int nbv) {
  constexpr int TM = KB_UPDATE_TM;
  dim3 grid((n + 63) / 64, (n + TM - 1) / TM);
  kb_gj_update<NB, TM><<<grid, 256, 0, w.st>>>(in, out, n, k0, nbv, w.Gp, w.srcrow, w.PT);
  (void)h;
}

int kbi_panel_width(const kb_context* h, int n) {
  if (h->opt_panel == 4 || h->opt_panel == 8 || h->opt_panel == 16) {
    int nb = h->opt_panel;
    if (nb == 16 && n <= 640) return 16;
    if (nb >= 8 && n <= 1024) return 8;
    return 4;
  }
  if (n <= 640) return 16;
  if (n <= 1024) return 8;
  return 4;
}

// In-place inverse of the n x n row-major matrix in w.S0 (w.S1 = scratch of the same
// size; w.PT must hold the first panel column-major).  Returns the buffer that holds
// (Pi S)^{-1}; w.orig holds Pi.
static int gj_invert(kb_context* h, const GjWs& w, int n, double2** result) {
  if (n > 4096) return kb_fail(h, KB_EINVAL, "chain node of %d rows exceeds the supported 4096", n);
  const int NB = kbi_panel_width(h, n);
  double2* in = w.S0;
  double2* out = w.S1;
  if (NB == 16 && !getenv("KB_NO_LOOKAHEAD")) {
    // first panel alone, then one fused kernel per step: update k + look-ahead panel k+1
    constexpr int TM = KB_UPDATE_TM;
    const int nrt = (n + TM - 1) / TM, nct = (n + 63) / 64;
    int T = (n + 1) / 2;
    T = ((T + 31) / 32) * 32;
    if (T < 256) T = 256;
    if (T > 320) T = 320;
    launch_panel<16, 2, 320>(h, w, n, 0, n < NB ? n : NB);
    h->launches++;
    double2* gp[2] = {w.Gp, w.Gp2};
    int* sr[2] = {w.srcrow, w.srcrow2};
    int sidx = 0;
    for (int k0 = 0; k0 < n; k0 += NB, ++sidx) {
      const int nbv = n - k0 < NB ? n - k0 : NB;
      const int knext = k0 + NB;
      const int has_next = knext < n ? 1 : 0;
      const int nbv_next = has_next ? (n - knext < NB ? n - knext : NB) : 0;
      kb_gj_fused<16, 2, TM, 320><<<nrt * nct + has_next, T, 2 * n * sizeof(int), w.st>>>(
          in, out, n, k0, nbv, gp[sidx & 1], sr[sidx & 1], w.PT, has_next, nbv_next, gp[(sidx + 1) & 1],
          sr[(sidx + 1) & 1], w.orig, h->d_info.p, w.ready, h->d_sweep_timing.p);
      h->launches++;
      double2* t = in;
      in = out;
      out = t;
    }
    KB_LAUNCH_CHECK(h);
    *result = in;
    return KB_OK;
  }
  for (int k0 = 0; k0 < n; k0 += NB) {
    int nbv = n - k0 < NB ? n - k0 : NB;
    if (NB == 16) {
      launch_panel<16, 2, 320>(h, w, n, k0, nbv);
      launch_update<16>(h, w, in, out, n, k0, nbv);
    } else if (NB == 8) {
      launch_panel<8, 2, 512>(h, w, n, k0, nbv);
      launch_update<8>(h, w, in, out, n, k0, nbv);
    } else {
      if (n <= 1024)
        launch_panel<4, 2, 512>(h, w, n, k0, nbv);
      else if (n <= 2048)
        launch_panel<4, 4, 512>(h, w, n, k0, nbv);
      else
        launch_panel<4, 4, 1024>(h, w, n, k0, nbv);
      launch_update<4>(h, w, in, out, n, k0, nbv);
    }
    h->launches += 2;
    double2* t = in;
    in = out;
    out = t;
  }
  KB_LAUNCH_CHECK(h);
  *result = in;
  return KB_OK;
}

// T = A - sigma B on the union pattern, then power-of-two row and column equilibration
int kbi_build_T(kb_context* h, zcomplex sigma) {
  cudaStream_t s = h->stream;
  const int64_t n = h->n, nnz = h->nnz;
  {
    int thr = 256;
    int64_t blk = (nnz + thr - 1) / thr;
    KB_CUDA(h, cudaMemcpyAsync(h->d_Tval.p, h->d_Aval.p, (size_t)nnz * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    if (h->B.present && h->nnzB > 0) {
      const double2 sg = zmake(sigma.real(), sigma.imag());
      unsigned bblk = (unsigned)((h->nnzB + thr - 1) / thr);
      if (h->b_is_complex)
        kb_sub_sigma_B<double2><<<bblk, thr, 0, s>>>(h->nnzB, h->d_bval_c.p, h->d_bmap.p, sg, h->d_Tval.p);
      else
        kb_sub_sigma_B<double><<<bblk, thr, 0, s>>>(h->nnzB, h->d_bval_r.p, h->d_bmap.p, sg, h->d_Tval.p);
    }
    int64_t wblk = (n * 32 + thr - 1) / thr;
    kb_row_scale<<<(unsigned)wblk, thr, 0, s>>>((int)n, h->d_rowptr.p, h->d_Tval.p, h->d_rscale.p,
                                                h->opt_equil);
    KB_CUDA(h, cudaMemsetAsync(h->d_maxbits.p, 0, n * sizeof(unsigned long long), s));
    kb_col_max<<<(unsigned)blk, thr, 0, s>>>(nnz, h->d_col.p, h->d_Tval.p, h->d_maxbits.p);
    kb_col_scale_vec<<<(unsigned)((n + thr - 1) / thr), thr, 0, s>>>((int)n, h->d_maxbits.p,
                                                                     h->d_cscale.p, h->opt_equil);
    kb_col_apply<<<(unsigned)blk, thr, 0, s>>>(nnz, h->d_col.p, h->d_cscale.p, h->d_Tval.p);
    h->launches += 5;
    KB_LAUNCH_CHECK(h);
  }
  return KB_OK;
}

int kbi_factor_workspace(kb_context* h) {
  cudaStream_t s = h->stream;
  const int64_t bmax = h->bmax;
  KB_CUDA(h, h->d_S0.alloc((size_t)bmax * bmax));
  KB_CUDA(h, h->d_S1.alloc((size_t)bmax * bmax));
  KB_CUDA(h, h->d_W.alloc((size_t)bmax * bmax));
  KB_CUDA(h, h->d_Gp.alloc((size_t)bmax * 32));
  KB_CUDA(h, h->d_PT.alloc((size_t)bmax * 16));
  KB_CUDA(h, h->d_orig.alloc(bmax));
  KB_CUDA(h, h->d_srcrow.alloc(2 * bmax));
  KB_CUDA(h, h->d_ready.alloc(4));
  KB_CUDA(h, cudaMemsetAsync(h->d_ready.p, 0, 4 * sizeof(unsigned), s));
  KB_CUDA(h, h->d_info.alloc(1));
  KB_CUDA(h, cudaMemsetAsync(h->d_info.p, 0, sizeof(int), s));
  if (getenv("KB_SWEEP_TIMING")) {
    KB_CUDA(h, h->d_sweep_timing.alloc(256 * 8));
    KB_CUDA(h, cudaMemsetAsync(h->d_sweep_timing.p, 0, 256 * 8 * sizeof(long long), s));
  }
  return KB_OK;
}

// One node of an elimination chain: Schur block, inversion, store, coupling product for
// the next node of the chain.  dir = +1: downward (uses L_{p,p-1} W, produces M_p U_{p,p+1});
// dir = -1: upward (uses U_{p,p+1} W, produces M'_p L_{p,p-1}).
static int factor_node(kb_context* h, const GjWs& w, int64_t p, int dir, bool first, bool produce_next) {
  const int o = (int)h->nodeptr[p];
  const int b = (int)(h->nodeptr[p + 1] - h->nodeptr[p]);
  const int64_t q = p - dir;  // node eliminated just before p on this chain
  const int oq = (!first) ? (int)h->nodeptr[q] : 0;
  kb_schur_row<<<b, 128, 0, w.st>>>(w.S0, w.PT, kbi_panel_width(h, b), b, o,
                                    (!first && dir > 0) ? w.W : nullptr, oq,
                                    (!first && dir < 0) ? w.W : nullptr, oq, h->d_rowptr.p, h->d_dstart.p,
                                    h->d_ustart.p, h->d_col.p, h->d_Tval.p);
  h->launches++;
  double2* X = nullptr;
  KB_TRY(gj_invert(h, w, b, &X));
  double2* Mp = h->d_M.p + h->Moff[p];
  kb_store_inverse<<<b, 128, 0, w.st>>>(X, b, w.orig, Mp);
  h->launches++;
  if (produce_next) {
    const int64_t nx = p + dir;
    const int onext = (int)h->nodeptr[nx];
    const int bnext = (int)(h->nodeptr[nx + 1] - h->nodeptr[nx]);
    if (dir > 0)
      kb_w_rows<<<b, 128, b * sizeof(double2), w.st>>>(Mp, b, o, w.W, bnext, onext, h->d_ucptr.p, h->d_urow.p,
                                                       h->d_upos.p, h->d_Tval.p);
    else
      kb_w_rows<<<b, 128, b * sizeof(double2), w.st>>>(Mp, b, o, w.W, bnext, onext, h->d_lcptr.p, h->d_lrow.p,
                                                       h->d_lpos.p, h->d_Tval.p);
    h->launches++;
  }
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

int kbi_factor(kb_context* h, zcomplex sigma) {
  if (!h->chain_set) return kb_fail(h, KB_EINVAL, "kb_set_chain must be called before kb_factor");
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int64_t P = h->P;
  const int64_t bmax = h->bmax;
  h->factored = false;
  h->M_transposed = false;
  h->fold_ready = false;
  h->fold_required = false;
  h->rng_valid = false;
  h->sigma = sigma;
  h->ch_lo = 0;
  h->ch_hi = P;
  h->stats.shard_path = 0;
  kbi_drop_graphs(h);

  KbEventPair ev;
  KB_CUDA(h, ev.create());
  KB_CUDA(h, cudaEventRecord(ev.e0, s));

  KB_TRY(kbi_build_T(h, sigma));

  // ---- storage for the explicit inverses
  h->Moff.assign(P + 1, 0);
  for (int64_t p = 0; p < P; ++p) {
    int64_t b = h->nodeptr[p + 1] - h->nodeptr[p];
    h->Moff[p + 1] = h->Moff[p] + b * b;
  }
  if (h->d_M.alloc((size_t)h->Moff[P]) != cudaSuccess)
    return kb_fail(h, KB_ENOMEM, "cannot allocate %.2f GB for the chain factors",
                   h->Moff[P] * 16.0 / 1e9);
  const bool chainfac = kbi_chainfac_supported(h);
  if (!chainfac) KB_TRY(kbi_factor_workspace(h));
  GjWs top = kbi_ws_main(h);

  // Two-sided ("burn at both ends") elimination: nodes 0..mid-1 are eliminated downward
  // and nodes P-1..mid+1 upward on two streams; the chains are independent until the
  // middle node, so every latency-bound step of one overlaps the other's.  The solve
  // must use the matching two-sided sweep (kb_sweep.cu); the per-node and barrier
  // sweeps and the l-sharded path need the one-sided factors (mid = P-1).
  const bool two_sided = h->nranks == 1 && h->opt_sweep == 1 && P >= 4;
  const int64_t mid = two_sided ? P / 2 : P - 1;
  h->mid = mid;
  double flops = 0.0;
  for (int64_t p = 0; p < P; ++p) {
    double b = (double)(h->nodeptr[p + 1] - h->nodeptr[p]);
    flops += 8.0 * b * b * b;
  }
  if (chainfac) {
    // one persistent launch for the whole elimination (kb_chainfac.cu)
    KB_CUDA(h, h->d_nodeptr.alloc(P + 1));
    KB_CUDA(h, h->d_Moff.alloc(P + 1));
    KB_CUDA(h, cudaMemcpyAsync(h->d_nodeptr.p, h->nodeptr.data(), (P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    KB_CUDA(h, cudaMemcpyAsync(h->d_Moff.p, h->Moff.data(), (P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    // transposed factors feed the one-hop sweep (kb_sweep1.cu); every other sweep reads M_p row-major
    int sms = 0;
    KB_CUDA(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    int sg = sms < 256 ? sms : 256;
    if (getenv("KB_SWEEP_GRID")) {
      int gsz = atoi(getenv("KB_SWEEP_GRID"));
      if (gsz >= 1 && gsz < sg) sg = gsz;
    }
    h->M_transposed = h->opt_sweep == 1 && kbi_onehop_supported(h, sg, two_sided, nullptr, nullptr);
    if (!h->M_transposed && h->opt_sweep == 1 && h->opt_fold) {
      // nodes too wide for the one-hop sweep's two stages (641 .. 704 rows) still fit the folded
      // sweep with a single slice stage; it reads the transposed factors as well
      h->sweep_grid_hint = sg;
      h->M_transposed = true;
      const int64_t cap = two_sided ? 2 * h->bmin : h->bmin;
      const bool ok = kbi_fold_supported(h, (int)(sg < cap ? sg : cap), two_sided, nullptr, nullptr);
      h->M_transposed = ok;
      h->fold_required = ok;
    }
    KB_TRY(kbi_chainfac_run(h, two_sided, h->M_transposed));
  } else if (two_sided) {
    if (!h->stream2) KB_CUDA(h, cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    KB_CUDA(h, h->d_S0b.alloc((size_t)bmax * bmax));
    KB_CUDA(h, h->d_S1b.alloc((size_t)bmax * bmax));
    KB_CUDA(h, h->d_Wb.alloc((size_t)bmax * bmax));
    KB_CUDA(h, h->d_Gpb.alloc((size_t)bmax * 32));
    KB_CUDA(h, h->d_PTb.alloc((size_t)bmax * 16));
    KB_CUDA(h, h->d_origb.alloc(bmax));
    KB_CUDA(h, h->d_srcrowb.alloc(2 * bmax));
    GjWs bot;
    bot.st = h->stream2;
    bot.S0 = h->d_S0b.p;
    bot.S1 = h->d_S1b.p;
    bot.W = h->d_Wb.p;
    bot.Gp = h->d_Gpb.p;
    bot.PT = h->d_PTb.p;
    bot.orig = h->d_origb.p;
    bot.srcrow = h->d_srcrowb.p;
    bot.Gp2 = h->d_Gpb.p + (size_t)bmax * 16;
    bot.srcrow2 = h->d_srcrowb.p + bmax;
    bot.ready = h->d_ready.p + 1;
    cudaEvent_t fork, join;
    KB_CUDA(h, cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    KB_CUDA(h, cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    KB_CUDA(h, cudaEventRecord(fork, s));
    KB_CUDA(h, cudaStreamWaitEvent(h->stream2, fork, 0));
    // interleave the host-side launches of the two chains so both streams stay fed
    int64_t pt = 0, pb = P - 1;
    while (pt < mid || pb > mid) {
      if (pt < mid) {
        KB_TRY(factor_node(h, top, pt, +1, pt == 0, true));
        ++pt;
      }
      if (pb > mid) {
        KB_TRY(factor_node(h, bot, pb, -1, pb == P - 1, true));
        --pb;
      }
    }
    KB_CUDA(h, cudaEventRecord(join, h->stream2));
    KB_CUDA(h, cudaStreamWaitEvent(s, join, 0));
    // middle node: couplings from both sides
    {
      const int o = (int)h->nodeptr[mid], b = (int)(h->nodeptr[mid + 1] - h->nodeptr[mid]);
      kb_schur_row<<<b, 128, 0, s>>>(top.S0, top.PT, kbi_panel_width(h, b), b, o, top.W,
                                     (int)h->nodeptr[mid - 1], bot.W, (int)h->nodeptr[mid + 1], h->d_rowptr.p,
                                     h->d_dstart.p, h->d_ustart.p, h->d_col.p, h->d_Tval.p);
      double2* X = nullptr;
      KB_TRY(gj_invert(h, top, b, &X));
      kb_store_inverse<<<b, 128, 0, s>>>(X, b, top.orig, h->d_M.p + h->Moff[mid]);
      h->launches += 2;
      KB_LAUNCH_CHECK(h);
    }
    cudaEventDestroy(fork);
    cudaEventDestroy(join);
  } else {
    for (int64_t p = 0; p < P; ++p) KB_TRY(factor_node(h, top, p, +1, p == 0, p + 1 < P));
  }
  KB_TRY(kbi_sweep_prepare(h));
  if (h->opt_fold) KB_TRY(kbi_fold_prepare(h));
  if (h->fold_required && !h->fold_ready)
    return kb_fail(h, KB_ENOMEM, "cannot allocate the folded couplings (%.2f GB) that nodes of %lld rows need",
                   2.0 * h->Moff[P] * 16.0 / 1e9, (long long)bmax);
  KB_CUDA(h, cudaEventRecord(ev.e1, s));
  KB_TRY(kbi_sync(h));
  int info = 0;
  KB_CUDA(h, cudaMemcpy(&info, h->d_info.p, sizeof(int), cudaMemcpyDeviceToHost));
  h->stats.factor_ms = ev.ms();
  h->stats.factor_flops = flops;
  h->stats.factor_bytes = h->Moff[P] * 16;
  if (chainfac) {
    int kerr = 0;
    KB_CUDA(h, cudaMemcpy(&kerr, h->d_kfsync.p + KF_ERR_WORD, sizeof(int), cudaMemcpyDeviceToHost));
    if (h->inject_fault == 2) {
      kerr = KB_WERR_WATCHDOG;  // tests: as if a wait of the factorisation kernel had expired
      h->inject_fault = 0;
    }
    if (kerr != 0) {
      // a wait of the persistent kernel expired: the factors are garbage.  Per-step kernels from
      // here on (kbi_enter_safe_mode refactors through this function with opt_factor = 0)
      const int rc = kbi_enter_safe_mode(h, "factorisation", kerr);
      return rc == KB_EPROTOCOL_RETRY ? KB_OK : rc;
    }
  }
  if (info != 0)
    return kb_fail(h, KB_ESINGULAR,
                   "zero or non-finite pivot in a Schur block (A - sigma B is singular to working "
                   "precision at this shift)");
  h->factored = true;
  return KB_OK;
}

extern "C" int kb_factor(kb_handle h, const double* sigma) {
  if (!h || !sigma) return KB_EINVAL;
  if (h->nranks > 1) {
    return kbi_factor_sharded(h, zcomplex(sigma[0], sigma[1]));
  }
  return kbi_factor(h, zcomplex(sigma[0], sigma[1]));
}
