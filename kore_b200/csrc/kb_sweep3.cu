// Column-split chain sweep with a one-hop tagged exchange: the algebra of kb_sweep1.cu (the
// factors M_p^T are the only dense data read: twice per solve, forward and backward) with the
// exchange machinery of kb_sweep2.cu (32-byte tagged elements, gather warps that keep all of a
// thread's polls in flight, the dense operand in registers before the step's input exists).
//
// Replaces the MUMPS solve phase inside every ST application of E.solve()
// (/root/reference/bin/solve.py:123) and K.solve (solve.py:227):
//     forward   y_p = M_p (r_p - C_{p,q} y_q)          q = node eliminated before p
//     middle    x_m = M_m (r_m - L y_{m-1} - U y_{m+1})
//     backward  x_p = y_p - M_p (C_{p,q'} x_{q'})      q' = node eliminated after p
// CTA c of a chain group owns the columns C_c of every M_p (rows C_c of M_p^T, one contiguous
// slice, TMA-staged two steps ahead).  A step:
//   gather warps (4):  poll the b-vectors of partial products the producers published in the
//       previous step on the rows this CTA's coupling rows touch (~23), all polls of a thread
//       in flight together, sum them over the producers in a fixed order (deterministic), apply
//       the sparse coupling rows C_c  ->  t[C_c] in shared memory;
//   compute warps (8): hold their part of the slice in registers (3 outputs x <= 10 columns per
//       thread, loaded from the stage before t exists), partial_c = M_p[:, C_c] t[C_c] for all
//       b rows, published as tagged elements.
// Against kb_sweep2.cu: 2 x 16 sum b^2 bytes of HBM traffic per solve instead of 3 x, no folded
// buffers, no solution pass; the price is an exchange of b values per CTA and step through L2
// (19 KB published, ~54 KB polled) instead of <= 10.
#include <stdlib.h>

#include "kb_internal.cuh"

#define K3_CW 8                        // compute warps
#define K3_GW 4                        // gather warps
#define K3_THREADS ((K3_CW + K3_GW) * 32)
#define K3_CT (K3_CW * 32)
#define K3_GT (K3_GW * 32)
#define K3_RING 8
#define K3_OPT 3                       // outputs per compute thread: nodes up to K3_OPT * K3_CT = 768 rows
#define K3_XR 10                       // most columns of a node a CTA may own
#define K3_NPOLL 10                    // polls in flight per gather thread and batch
#define K3_AHEAD 4

struct K3Params {
  const double2* MT;
  const int64_t* Moff;
  const int64_t* nodeptr;
  int P, mid;
  const double2* r;
  K2Elem* yft;     // forward results y as tagged elements (n): read by OTHER CTAs in the backward pass
  double ytag;     // their tag: unique per solve
  double2* x;
  K2Elem* ring[2];
  K2Elem* xchg[2];
  const double2* Lval;
  const int* Lcol;
  int WL;
  const double2* Uval;
  const int* Ucol;
  int WU;
  double tag0[2];  // publication of step s of group g carries tag0[g] + s + 1
  int* err;
  long long* timing;
  int bmax;
  int G0;
  const int4* rng;
  int smax;
};

__device__ __forceinline__ void k3_bar_compute() { asm volatile("bar.sync 2, %0;" ::"n"(K3_CT) : "memory"); }
__device__ __forceinline__ void k3_bar_gather() { asm volatile("bar.sync 4, %0;" ::"n"(K3_GT) : "memory"); }
__device__ __forceinline__ void k3_bar_t_arrive() { asm volatile("bar.arrive 3, %0;" ::"n"(K3_THREADS) : "memory"); }
__device__ __forceinline__ void k3_bar_t_wait() { asm volatile("bar.sync 3, %0;" ::"n"(K3_THREADS) : "memory"); }

// dst[j - lo] = base ? base[qo + j] - sum_c' part[c'][j] : sum_c' part[c'][j]   for j in [lo, hi),
// and store[qo + j] (plain) / storet[qo + j] (tagged) = that value for j in [slo, shi).  `base` and
// `storet` are the tagged forward results: they are written by one CTA and read by its
// neighbours many steps later, and a tag check is the only ordering this kernel has between
// CTAs (no fences anywhere).  All K3_GT gather threads call it.  Thread
// (pg = warp, jl = lane) sums the producers pg, pg + K3_GW, ... of row j0 + jl in ascending order;
// the K3_GW partial sums of a row meet in shared memory in a fixed order: deterministic.
__device__ __noinline__ void k3_collect(const K2Elem* part, int ld, int nprod, double tag, int qo, int lo, int hi,
                                        const K2Elem* base, double2* dst, double2* store, K2Elem* storet, double ytag,
                                        int slo, int shi, double2* red, int* err) {
  const int g = threadIdx.x - K3_CT;
  const int jl = g & 31, pg = g >> 5;
  for (int j0 = lo; j0 < hi; j0 += 32) {
    const int j = j0 + jl;
    const bool rowok = j < hi;
    double2 acc = zmake(0.0, 0.0);
    for (int cb = pg; cb < nprod; cb += K3_GW * K3_NPOLL) {
      double re[K3_NPOLL], t0[K3_NPOLL], im[K3_NPOLL], t1[K3_NPOLL];
      unsigned pend = 0u;
#pragma unroll
      for (int u = 0; u < K3_NPOLL; ++u) {
        re[u] = im[u] = 0.0;
        if (rowok && cb + K3_GW * u < nprod) pend |= 1u << u;
      }
      int spins = 0;
      while (pend) {
#pragma unroll
        for (int u = 0; u < K3_NPOLL; ++u)
          if (pend & (1u << u)) {
            asm volatile("ld.relaxed.gpu.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                         : "=d"(re[u]), "=d"(t0[u]), "=d"(im[u]), "=d"(t1[u])
                         : "l"(part + (size_t)(cb + K3_GW * u) * ld + j)
                         : "memory");
          }
#pragma unroll
        for (int u = 0; u < K3_NPOLL; ++u)
          if ((pend & (1u << u)) && t0[u] == tag && t1[u] == tag) pend &= ~(1u << u);
        if (pend && (++spins & 255) == 0 && (*(volatile int*)err != 0 || spins > KB_SPIN_LIMIT)) {
          atomicExch(err, 1);
          break;
        }
      }
#pragma unroll
      for (int u = 0; u < K3_NPOLL; ++u) acc = zadd(acc, zmake(re[u], im[u]));
    }
    red[pg * 32 + jl] = acc;
    k3_bar_gather();
    if (pg == 0 && rowok) {
      double2 sum = red[jl];
#pragma unroll
      for (int w = 1; w < K3_GW; ++w) sum = zadd(sum, red[w * 32 + jl]);
      const double2 val = base ? zsub(k2_poll(base + qo + j, ytag, err), sum) : sum;
      dst[j - lo] = val;
      if (j >= slo && j < shi) {
        if (store) store[qo + j] = val;
        if (storet) k2_publish(storet + qo + j, val, ytag);
      }
    }
    k3_bar_gather();
  }
}

// tl[i - c0] (-)= sum_k val[i][k] v[col[i][k] - qo - lo] for rows i in [c0, c1) of node p; the
// gather warps take the rows round-robin, one row per warp pass (half a warp when W <= 16).
__device__ __noinline__ void k3_couple(const double2* val, const int* col, int W, int o, int c0, int c1, int qo,
                                       int lo, int hi, const double2* v, double2* tl, bool subtract) {
  const int g = threadIdx.x - K3_CT;
  const int lane = g & 31, wid = g >> 5;
  const bool halfw = W <= 16;
  const int hl = halfw ? (lane & 15) : lane;
  const int hsel = halfw ? (lane >> 4) : 0;
  const int rpp = halfw ? 2 * K3_GW : K3_GW;
  for (int ib = c0; ib < c1; ib += rpp) {
    const int i = ib + (halfw ? 2 * wid + hsel : wid);
    double2 acc = zmake(0.0, 0.0);
    if (i < c1) {
      for (int k = hl; k < W; k += (halfw ? 16 : 32)) {
        const size_t e = (size_t)(o + i) * W + k;
        const int cc = __ldg(&col[e]) - qo;
        if (cc >= lo && cc < hi) zfma(acc, __ldg(&val[e]), v[cc - lo]);
      }
    }
    if (!halfw) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    }
#pragma unroll
    for (int sft = 8; sft > 0; sft >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
    }
    if (i < c1 && hl == 0) tl[i - c0] = subtract ? zsub(tl[i - c0], acc) : zadd(tl[i - c0], acc);
  }
}

__global__ void __launch_bounds__(K3_THREADS, 1) kb_sweep_tagged(K3Params q, int slice_elems) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* stage0 = (double2*)smem_raw;
  const int bpad = (q.bmax + 7) & ~7;
  double2* va = stage0 + 2 * (size_t)slice_elems;   // collected previous result
  double2* vb = va + bpad;                          // second input of the middle node
  double2* tl0 = vb + bpad;                         // t on this CTA's columns, 2 x 16 by step parity
  double2* red = tl0 + 32;                          // K3_GW x 32
  int64_t* s_moff = (int64_t*)(red + K3_GW * 32);
  int* s_nptr = (int*)(s_moff + (q.P + 1));
  __shared__ __align__(8) uint64_t mbar[2];
  const int tid = threadIdx.x;
  const int P = q.P, mid = q.mid;
  const int group = ((int)blockIdx.x < q.G0) ? 0 : 1;
  const int gsz[2] = {q.G0, (int)gridDim.x - q.G0};
  const int gsize = gsz[group];
  const int grank = group == 0 ? (int)blockIdx.x : (int)blockIdx.x - q.G0;
  const int nbot = P - 1 - mid;
  const int S = group == 0 ? 2 * mid + 1 : 2 * nbot;
  const int ld = (q.bmax + 7) & ~7;

  for (int i = tid; i <= P; i += K3_THREADS) {
    s_moff[i] = q.Moff[i];
    s_nptr[i] = (int)q.nodeptr[i];
  }
  if (tid == 0) {
    kb_mbar_init(&mbar[0], 1);
    kb_mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // where the partial products of step sp of group gp are published
  auto part_of = [&](int gp, int sp) -> K2Elem* {
    if (gp == 0 && sp == mid) return q.xchg[0];
    if (gp == 1 && sp == nbot - 1) return q.xchg[1];
    return q.ring[gp] + (size_t)(sp % K3_RING) * gsz[gp] * ld;
  };
  auto tag_of = [&](int gp, int sp) -> double { return q.tag0[gp] + (double)(sp + 1); };

  if (tid >= K3_CT) {
    // ===== gather warps: inputs of every step -> t on this CTA's columns =====
    for (int s = 0; s < S; ++s) {
      int p, mode;
      kb_flow_step(group, s, P, mid, p, mode);
      const int o = s_nptr[p], b = s_nptr[p + 1] - o;
      const int4* tab = q.rng + ((size_t)blockIdx.x * q.smax + s) * 2;
      const int4 rng = __ldg(tab), own = __ldg(tab + 1);
      const int c0 = own.x, c1 = own.y, nc = c1 - c0;
      const bool fwd = mode <= KB_MID;
      double2* tl = tl0 + (s & 1) * 16;
      const int g = tid - K3_CT;
      if (g < nc) tl[g] = fwd ? q.r[o + c0 + g] : zmake(0.0, 0.0);
      int qn = -1;
      bool prev_base = false;
      if (s > 0) {
        int pm;
        kb_flow_step(group, s - 1, P, mid, qn, pm);
        prev_base = (pm == KB_BWD_U || pm == KB_BWD_L);
      }
      const bool cross = (group == 1 && s == nbot);
      const bool useL = (mode == KB_FWD_L || mode == KB_MID || mode == KB_BWD_L);
      int lo = rng.x, hi = rng.y, qo = 0;
      bool have = false;
      if (fwd ? (qn >= 0) : true) {
        const int qq = fwd ? qn : (mode == KB_BWD_U ? p + 1 : p - 1);
        qo = s_nptr[qq];
        const int bq = s_nptr[qq + 1] - qo;
        const int q0 = own.z, q1 = own.w;
        if (q1 > q0) {
          lo = min(lo, q0);
          hi = max(hi, q1);
        }
        have = lo < hi;
        if (have) {
          if (cross)
            k3_collect(q.xchg[0], ld, min(gsz[0], bq), tag_of(0, mid), qo, lo, hi, nullptr, va, nullptr, nullptr,
                       q.ytag, 0, 0, red, q.err);
          else if (fwd)
            k3_collect(part_of(group, s - 1), ld, min(gsize, bq), tag_of(group, s - 1), qo, lo, hi, nullptr, va,
                       nullptr, q.yft, q.ytag, q0, q1, red, q.err);
          else
            k3_collect(part_of(group, s - 1), ld, min(gsize, bq), tag_of(group, s - 1), qo, lo, hi,
                       prev_base ? q.yft : nullptr, va, q.x, nullptr, q.ytag, q0, q1, red, q.err);
        }
        if (cross && nc > 0) {
          // y_{mid+1} on this CTA's own rows (the forward result of this very node): yf = sum of the
          // last forward partials of this group
          k3_collect(q.xchg[1], ld, min(gsize, b), tag_of(1, nbot - 1), o, c0, c1, nullptr, vb, nullptr, q.yft,
                     q.ytag, c0, c1, red, q.err);
        }
      }
      k3_bar_gather();  // tl initialised, va complete
      if (have && nc > 0)
        k3_couple(useL ? q.Lval : q.Uval, useL ? q.Lcol : q.Ucol, useL ? q.WL : q.WU, o, c0, c1, qo, lo, hi, va, tl,
                  fwd);
      if (mode == KB_MID && mid < P - 1) {
        const int q2 = mid + 1;
        const int qo2 = s_nptr[q2], bq2 = s_nptr[q2 + 1] - qo2;
        const int lo2 = rng.z, hi2 = rng.w;
        if (lo2 < hi2)
          k3_collect(q.xchg[1], ld, min(gsz[1], bq2), tag_of(1, nbot - 1), qo2, lo2, hi2, nullptr, vb, nullptr,
                     nullptr, q.ytag, 0, 0, red, q.err);
        k3_bar_gather();
        if (nc > 0 && lo2 < hi2) k3_couple(q.Uval, q.Ucol, q.WU, o, c0, c1, qo2, lo2, hi2, vb, tl, true);
      }
      __threadfence_block();
      k3_bar_t_arrive();
    }
    // ---- the result of the last step, on this CTA's own rows
    if (S > 0) {
      int p, mode;
      kb_flow_step(group, S - 1, P, mid, p, mode);
      const int o = s_nptr[p], b = s_nptr[p + 1] - o;
      int c0, c1;
      kb_group_rows(b, gsize, grank, c0, c1);
      const bool base = (mode == KB_BWD_U || mode == KB_BWD_L);
      if (c1 > c0)
        k3_collect(part_of(group, S - 1), ld, min(gsize, b), tag_of(group, S - 1), o, c0, c1, base ? q.yft : nullptr,
                   va, q.x, nullptr, q.ytag, c0, c1, red, q.err);
    }
    return;
  }

  // ===== compute warps =====
  auto slice_of = [&](int sn, const double2*& src, unsigned& bytes) {
    bytes = 0;
    src = nullptr;
    if (sn >= S) return;
    int pn, mn;
    kb_flow_step(group, sn, P, mid, pn, mn);
    const int bn = s_nptr[pn + 1] - s_nptr[pn];
    int a0, a1;
    kb_group_rows(bn, gsize, grank, a0, a1);
    bytes = (unsigned)((size_t)(a1 - a0) * bn * sizeof(double2));
    src = q.MT + s_moff[pn] + (size_t)a0 * bn;
  };
  auto issue_copy = [&](int sn) {
    const double2* src;
    unsigned bytes;
    slice_of(sn, src, bytes);
    if (bytes) {
      uint64_t* mb = &mbar[sn & 1];
      kb_mbar_expect_tx(mb, bytes);
      kb_bulk_g2s(stage0 + (size_t)(sn & 1) * slice_elems, src, bytes, mb);
    }
  };
  auto prefetch = [&](int sn) {
    const double2* src;
    unsigned bytes;
    slice_of(sn, src, bytes);
    while (bytes > 0) {
      const unsigned c = bytes > 65536u ? 65536u : bytes;
      kb_prefetch_l2(src, c);
      src = (const double2*)((const char*)src + c);
      bytes -= c;
    }
  };
  if (tid == 0) {
    issue_copy(0);
    issue_copy(1);
  }
  if (tid == 1)
    for (int sn = 2; sn < K3_AHEAD; ++sn) prefetch(sn);

  unsigned uses0 = 0u, uses1 = 0u;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tc0 = clock64();
#define K3_TICK(k)              \
  do {                          \
    if (q.timing) {             \
      long long _t = clock64(); \
      tacc[k] += _t - tc0;      \
      tc0 = _t;                 \
    }                           \
  } while (0)
  int c_b = -1, c_c0 = 0, c_nc = 0;
  for (int s = 0; s < S; ++s) {
    int p, mode;
    kb_flow_step(group, s, P, mid, p, mode);
    const int o = s_nptr[p], b = s_nptr[p + 1] - o;
    (void)o;
    if (b != c_b) {
      int c1;
      kb_group_rows(b, gsize, grank, c_c0, c1);
      c_nc = c1 - c_c0;
      c_b = b;
    }
    const int nc = c_nc;
    const double2* tl = tl0 + (s & 1) * 16;
    // ---- the slice of M_p^T into registers: outputs i = tid + u K3_CT, all nc columns
    double2 m[K3_OPT][K3_XR];
    if (nc > 0) {
      if (s & 1) {
        kb_mbar_wait(&mbar[1], uses1 & 1u);
        uses1++;
      } else {
        kb_mbar_wait(&mbar[0], uses0 & 1u);
        uses0++;
      }
      const double2* Ms = stage0 + (size_t)(s & 1) * slice_elems;
#pragma unroll
      for (int u = 0; u < K3_OPT; ++u) {
        const int i = tid + u * K3_CT;
#pragma unroll
        for (int j = 0; j < K3_XR; ++j) m[u][j] = (i < b && j < nc) ? Ms[(size_t)j * b + i] : zmake(0.0, 0.0);
      }
    }
    k3_bar_compute();  // the stage of step s may be refilled
    if (tid == 0) issue_copy(s + 2);
    if (tid == 1) prefetch(s + K3_AHEAD);
    K3_TICK(0);
    k3_bar_t_wait();
    K3_TICK(1);
    if (nc > 0) {
      double2 acc[K3_OPT];
#pragma unroll
      for (int u = 0; u < K3_OPT; ++u) acc[u] = zmake(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < K3_XR; ++j) {
        const double2 tj = j < nc ? tl[j] : zmake(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < K3_OPT; ++u) zfma(acc[u], m[u][j], tj);
      }
      K2Elem* out = part_of(group, s) + (size_t)grank * ld;
      const double tag = tag_of(group, s);
#pragma unroll
      for (int u = 0; u < K3_OPT; ++u) {
        const int i = tid + u * K3_CT;
        if (i < b) k2_publish(out + i, acc[u], tag);
      }
    }
    K3_TICK(2);
  }
  if (q.timing && tid == 0)
    for (int k = 0; k < 8; ++k) q.timing[blockIdx.x * 8 + k] = tacc[k];
#undef K3_TICK
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool kbi_tagged_supported(const kb_context* h, int G, bool two_sided, int* slice_elems_out, size_t* smem_out) {
  if (!getenv("KB_SWEEP_TAGGED")) return false;  // experimental: opt-in
  if (!h->M_transposed || G < 2) return false;
  const int gmin = two_sided ? G / 2 : G;
  if (gmin < 1) return false;
  const int64_t rpc = (h->bmax + gmin - 1) / gmin;
  if (rpc > K3_XR || h->bmax > K3_OPT * K3_CT) return false;
  const int64_t slice_elems = (rpc * h->bmax + 7) & ~(int64_t)7;
  const size_t bpad = (size_t)((h->bmax + 7) & ~(int64_t)7);
  const size_t smem = 2 * (size_t)slice_elems * sizeof(double2) + 2 * bpad * sizeof(double2) +
                      (32 + K3_GW * 32) * sizeof(double2) + (size_t)(h->P + 1) * (sizeof(int64_t) + sizeof(int)) + 16;
  if (smem > 220 * 1024) return false;
  if (slice_elems_out) *slice_elems_out = (int)slice_elems;
  if (smem_out) *smem_out = smem;
  return true;
}

int kbi_onehop_build_ranges(kb_context* h, int G, int G0);  // kb_sweep1.cu

// y <- T'^{-1} r on TRANSPOSED two-sided factors.  y has n+1 entries, y[n] == 0.
int kbi_sweep_tagged(kb_context* h, const double2* r, double2* y) {
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  const int G = h->sweep_grid;
  const bool two = h->mid < h->P - 1;
  int slice_elems = 0;
  size_t smem = 0;
  if (!kbi_tagged_supported(h, G, two, &slice_elems, &smem))
    return kb_fail(h, KB_EINVAL, "chain does not fit the tagged sweep kernel");
  const int G0 = two ? (G + 1) / 2 : G;
  const size_t ldp = (size_t)((h->bmax + 7) & ~(int64_t)7);
  const size_t ring_elems = (size_t)K3_RING * G * ldp;
  const size_t xchg_elems = 2 * (size_t)G0 * ldp;
  const size_t yft_elems = (size_t)n + 8;
  const size_t need = (ring_elems + xchg_elems + yft_elems) * sizeof(K2Elem);
  if (h->d_tagring.count < need) {
    KB_CUDA(h, h->d_tagring.alloc(need));
    KB_CUDA(h, cudaMemsetAsync(h->d_tagring.p, 0, need, s));
    h->tag_epoch[0] = h->tag_epoch[1] = 0;
  }
  KB_TRY(kbi_onehop_build_ranges(h, G, G0));
  K3Params q;
  q.MT = h->d_M.p;
  q.Moff = h->d_Moff.p;
  q.nodeptr = h->d_nodeptr.p;
  q.P = (int)h->P;
  q.mid = (int)h->mid;
  q.r = r;
  q.x = y;
  K2Elem* base = (K2Elem*)h->d_tagring.p;
  q.yft = base + ring_elems + xchg_elems;
  q.ytag = (double)h->tag_epoch[0] + 0.5;  // unique per solve (the epochs grow by >= 2 per solve)
  q.ring[0] = base;
  q.ring[1] = base + (size_t)K3_RING * G0 * ldp;
  q.xchg[0] = base + ring_elems;
  q.xchg[1] = base + ring_elems + (size_t)G0 * ldp;
  q.Lval = h->d_Lval.p;
  q.Lcol = h->d_Lcol.p;
  q.WL = h->WL > 0 ? h->WL : 1;
  q.Uval = h->d_Uval.p;
  q.Ucol = h->d_Ucol.p;
  q.WU = h->WU > 0 ? h->WU : 1;
  const unsigned long long s0 = (unsigned long long)(2 * h->mid + 1), s1 = (unsigned long long)(2 * (h->P - 1 - h->mid));
  q.tag0[0] = (double)h->tag_epoch[0];
  q.tag0[1] = (double)h->tag_epoch[1];
  h->tag_epoch[0] += s0 + 1ull;
  h->tag_epoch[1] += s1 + 1ull;
  q.err = h->d_sweep_err.p;
  q.timing = h->d_sweep_timing.p;
  q.bmax = (int)h->bmax;
  q.G0 = G0;
  q.rng = h->d_rng.p;
  q.smax = (int)(2 * h->mid + 1);
  const void* fn = (const void*)kb_sweep_tagged;
  if (smem > 48 * 1024) KB_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&q, (void*)&slice_elems};
  KB_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(K3_THREADS), args, smem, s));
  h->launches += 1;
  return KB_OK;
}
