// One-hop chain sweep: the two-sided forward + backward block substitution of one solve in
// ONE cooperative launch, with ONE cross-CTA exchange per chain step.
//
// Replaces the MUMPS solve phase inside every ST application of E.solve()
// (/root/reference/bin/solve.py:123) and K.solve (solve.py:227); same algebra as
// kb_sweep.cu:
//     forward   y_p = M_p (r_p - C_{p,q} y_q)          q = node eliminated before p
//     middle    x_m = M_m (r_m - L y_{m-1} - U y_{m+1})
//     backward  x_p = y_p - M_p (C_{p,q'} x_{q'})      q' = node eliminated after p
// but the dense product is split by COLUMNS of M_p instead of rows.  CTA c of a chain
// group owns the columns C_c of every node of its chain, i.e. rows C_c of the stored
// transpose M_p^T (written that way by kb_chainfac.cu), one contiguous slice that arrives
// in shared memory by a 1-D bulk (TMA) copy issued two steps ahead.  A step is then
//   1. sum, over the producers c', the partial products of the previous step on the few
//      rows this CTA needs (its own rows and the band neighbourhood its coupling rows
//      touch): the ONLY cross-CTA dependency of the step;
//   2. t[C_c] = r[C_c] -/+ sparse coupling rows C_c against those values (CTA-local);
//   3. partial_c = M_p[:, C_c] t[C_c]  (all b rows, CTA-local, conflict-free shared-memory
//      reads, no cross-thread reduction) -> written with coalesced 16-byte stores.
// kb_sweep.cu needs two exchanges per step (all-gather of t, then of y); here the
// reduction of the partials IS the exchange.
//
// Publication protocol: one step counter per group.  The service warp of a CTA waits on a
// named barrier for the compute warps' stores of step s, then adds one to the counter with
// red.release.gpu (bar.sync + release is cumulative over the CTA's earlier writes).  A
// consumer polls the counter with one thread (ld.acquire) until every CTA of the group has
// added its one for the step, passes a block barrier and reads the partials ONCE with plain
// L2 loads.  The partial buffers form a ring of K1_RING step slots; a slot is rewritten only
// after its owner has seen every CTA publish a later step, i.e. after every reader is done
// with it.  Two other protocols were measured first: polling the data itself
// (sentinel values, as kb_sweep.cu does) and one flag per CTA; both cost more than the
// counter because every poll round touches tens of L2 lines per CTA.
#include <stdlib.h>

#include "kb_internal.cuh"

#define K1_THREADS 256
#define K1_RING 8
#define K1_CW 7   // compute warps; warp K1_CW is the service warp (publication, bulk copies)
#define K1_FSTRIDE 64  // words between two CTAs' flags (256 bytes: one L2 slice-hash granule)
#define K1_V 20   // polling loads in flight per lane (80 sources x 8 rows per warp pass)

struct K1Params {
  const double2* MT;
  const int64_t* Moff;
  const int64_t* nodeptr;
  int P, mid;
  const double2* r;
  double2* yf;        // forward results (n)
  double2* x;         // solution (n + zero slot)
  double2* ring[2];   // per group: K1_RING x (CTAs of the group) x bmax partial products
  double2* xchg[2];   // [0] partials of the middle node (group 0); [1] of group 1's last forward node
  const double2* Lval;
  const int* Lcol;
  int WL;
  const double2* Uval;
  const int* Ucol;
  int WU;
  unsigned* flags;    // step counters of the two groups (words 0 and K1_FSTRIDE)
  unsigned epoch0, epoch1;  // their values before this solve
  int* err;
  long long* timing;
  int bmax;
  int G0;
  int dbgflags;       // timing experiments only: 1 = no L2 prefetch
  const int4* rng;    // per CTA and step: [lo, hi) of the coupling columns (x, y), second coupling of the middle node (z, w)
  int smax;
};

// named barriers: 1 = compute warps only; 2 = compute warps arrive / service warp waits (the
// partial products of the step are stored, its staged slice is free)
__device__ __forceinline__ void k1_bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(K1_CW * 32) : "memory"); }
__device__ __forceinline__ void k1_bar_arrive(int id) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(K1_THREADS) : "memory");
}
__device__ __forceinline__ void k1_bar_wait(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(K1_THREADS) : "memory");
}

// Wait until the step counter of a group has reached `target` (every CTA of the group adds one
// per step, after its partial products are stored).  One thread polls, then everybody passes
// the compute barrier.  tools/microbench/hop_latency.cu: a counter round among 74 CTAs costs
// ~1.9 K cycles on B200, polling the data itself (all-to-all, 16 bytes per pair) ~3.1 K.
__device__ __forceinline__ void k1_wait(const unsigned* ctr, unsigned target, int* err) {
  if (threadIdx.x == 0) {
    KbSpin sp;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if ((int)(v - target) >= 0) break;
      if (kb_spin_expired(sp, err, KB_WERR_SWEEP, KB_WAIT_NS_DEFAULT)) break;
    }
  }
  k1_bar_compute();
}

// dst[j - lo] = (base ? base[qo + j] : 0) -/+ sum_{c' < nprod} part[c' * ld + j]   for j in [lo, hi)
// and store[qo + j] = that value for j in [slo, shi).  The producers have been waited for.
// Every load instruction of a warp covers 8 consecutive rows (one 128-byte line) of 4
// producers.  Warp w takes the 8-row chunks w, w + NW, ...; lane (sub = lane >> 3, row =
// lane & 7) sums the sources sub, sub + 4, ... in ascending order (the base vector counts as
// the last source), then a fixed two-stage shuffle tree: deterministic.
__device__ __noinline__ void k1_collect(const double2* part, int ld, int nprod, int qo, int lo, int hi,
                                        const double2* base, bool neg, double2* dst, double2* store, int slo,
                                        int shi) {
  constexpr int NW = K1_CW;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane >> 3;
  const int lo8 = lo & ~7;
  const int nchunks = (hi - lo8 + 7) >> 3;
  const int ntot = nprod + (base ? 1 : 0);
  for (int ch = wid; ch < nchunks; ch += NW) {
    const int j = lo8 + ch * 8 + (lane & 7);
    const bool rowok = j >= lo && j < hi;
    double2 acc = zmake(0.0, 0.0);
    for (int cb = 0; cb < ntot; cb += 4 * K1_V) {
      double2 v[K1_V];
#pragma unroll
      for (int u = 0; u < K1_V; ++u) {
        const int src = cb + 4 * u + sub;
        v[u] = (rowok && src < ntot) ? __ldcg(src < nprod ? part + (size_t)src * ld + j : base + qo + j)
                                     : zmake(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < K1_V; ++u) acc = (cb + 4 * u + sub < nprod) ? zadd(acc, v[u]) : zsub(acc, v[u]);
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 8);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 8);
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    if (rowok && sub == 0) {
      const double2 val = neg ? zneg(acc) : acc;
      dst[j - lo] = val;
      if (store && j >= slo && j < shi) store[qo + j] = val;
    }
  }
}

// Column range [lo, hi) of node q (local indices) referenced by the coupling rows [c0, c1) of
// node p (global rows o + c0 ..), from the ELL copy; empty rows give lo >= hi.
__device__ __noinline__ void k1_range(const int* col, int W, int o, int c0, int c1, int qo, int bq, int* s_lo,
                                         int* s_hi) {
  // called by all threads; result in *s_lo / *s_hi after the barrier
  if (threadIdx.x == 0) {
    *s_lo = 0x7fffffff;
    *s_hi = 0;
  }
  __syncthreads();
  int lo = 0x7fffffff, hi = 0;
  for (int e = threadIdx.x; e < (c1 - c0) * W; e += K1_THREADS) {
    const int cc = __ldg(&col[(size_t)(o + c0) * W + e]) - qo;
    if (cc >= 0 && cc < bq) {
      lo = min(lo, cc);
      hi = max(hi, cc + 1);
    }
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(s_lo, lo);
    atomicMax(s_hi, hi);
  }
  __syncthreads();
}

// t[i - c0] (+)= sgn * sum_k val[i][k] v[col[i][k] - qo - lo]  for rows i in [c0, c1) of node p
// (half-warp per row when W <= 16).  Accumulates into tl (shared); caller initialises tl.
__device__ __noinline__ void k1_couple(const double2* val, const int* col, int W, int o, int c0, int c1, int qo,
                                          int lo, int hi, const double2* v, double2* tl, bool neg) {
  constexpr int NW = K1_CW;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool halfw = W <= 16;
  const int hl = halfw ? (lane & 15) : lane;
  const int hsel = halfw ? (lane >> 4) : 0;
  const int rpp = halfw ? 2 * NW : NW;
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
    if (i < c1 && hl == 0) tl[i - c0] = neg ? zsub(tl[i - c0], acc) : zadd(tl[i - c0], acc);  // tl starts at 0 backward
  }
}

__global__ void __launch_bounds__(K1_THREADS, 1) kb_sweep_onehop(K1Params q, int slice_elems) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* stage0 = (double2*)smem_raw;
  double2* va = stage0 + 2 * (size_t)slice_elems;       // collected previous result (range A)
  double2* vb = va + ((q.bmax + 7) & ~7);               // second input of the middle node (range B)
  double2* tl0 = vb + ((q.bmax + 7) & ~7);              // t on this CTA's rows, double-buffered by step parity
  const int tlen = (slice_elems / max(q.bmax, 1) + 8) & ~7;
  int64_t* s_moff = (int64_t*)(tl0 + 2 * tlen);
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
  const int ld = (q.bmax + 7) & ~7;  // partial rows start on 128-byte lines
  const unsigned* gctr[2] = {q.flags, q.flags + K1_FSTRIDE};
  const unsigned ebase[2] = {q.epoch0, q.epoch1};  // counter values before this solve
  unsigned uses[2] = {0u, 0u};
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tc0 = clock64();
#define K1_TICK(k)                 \
  do {                             \
    if (q.timing) {                \
      long long _t = clock64();    \
      tacc[k] += _t - tc0;         \
      tc0 = _t;                    \
    }                              \
  } while (0)

  for (int i = tid; i <= P; i += K1_THREADS) {
    s_moff[i] = q.Moff[i];
    s_nptr[i] = (int)q.nodeptr[i];
  }
  if (tid == 0) {
    kb_mbar_init(&mbar[0], 1);
    kb_mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // slice of M_p^T owned by this CTA at step sn -> stage sn & 1 (thread 0 only)
  auto issue_copy = [&](int sn) {
    if (sn >= S) return;
    int pn, mn;
    kb_flow_step(group, sn, P, mid, pn, mn);
    const int bn = s_nptr[pn + 1] - s_nptr[pn];
    int a0, a1;
    kb_group_rows(bn, gsize, grank, a0, a1);
    const unsigned bytes = (unsigned)((size_t)(a1 - a0) * bn * sizeof(double2));
    if (bytes) {
      uint64_t* mb = &mbar[sn & 1];
      kb_mbar_expect_tx(mb, bytes);
      kb_bulk_g2s(stage0 + (size_t)(sn & 1) * slice_elems, q.MT + s_moff[pn] + (size_t)a0 * bn, bytes, mb);
    }
  };
  auto prefetch_l2 = [&](int sn) {
    if (sn >= S) return;
    int pn, mn;
    kb_flow_step(group, sn, P, mid, pn, mn);
    const int bn = s_nptr[pn + 1] - s_nptr[pn];
    int a0, a1;
    kb_group_rows(bn, gsize, grank, a0, a1);
    const double2* m = q.MT + s_moff[pn] + (size_t)a0 * bn;
    size_t bytes = (size_t)(a1 - a0) * bn * sizeof(double2);
    while (bytes > 0) {
      size_t c = bytes > 65536 ? 65536 : bytes;
      kb_prefetch_l2(m, c);
      m = (const double2*)((const char*)m + c);
      bytes -= c;
    }
  };
  if (tid == K1_CW * 32) {
    issue_copy(0);
    issue_copy(1);
  }
  if (tid == K1_CW * 32 + 1) {
    prefetch_l2(2);
    prefetch_l2(3);
  }

  // where the partial products of step sp of group gp were published, and by how many CTAs
  auto part_of = [&](int gp, int sp) -> double2* {
    if (gp == 0 && sp == mid) return q.xchg[0];
    if (gp == 1 && sp == nbot - 1) return q.xchg[1];
    return q.ring[gp] + (size_t)(sp % K1_RING) * gsz[gp] * ld;
  };

  // ---- software pipeline: everything a step needs besides the other CTAs' partial
  //      products is fetched one step ahead, in the shadow of the exchange latency:
  //      the coupling ranges (table built at factor time), this lane's coupling entry of
  //      the first row pass, the right-hand side of the CTA's rows
  constexpr int NW = K1_CW;
  const int lane = tid & 31, wid = tid >> 5;
  constexpr int CT = K1_CW * 32;  // compute threads

  if (wid == K1_CW) {
    // ===== service warp: one pass per step, behind the compute warps =====
    //   wait until the compute warps have stored the partial products of step s (and are done
    //   with its staged slice), publish the step, refill the stage with the slice of step s+2
    unsigned* myctr = q.flags + (size_t)group * K1_FSTRIDE;
    for (int s = 0; s < S; ++s) {
      k1_bar_wait(2);
      if (lane == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(myctr) : "memory");
        issue_copy(s + 2);
      }
      if (lane == 1 && !(q.dbgflags & 1)) prefetch_l2(s + 4);
    }
    return;
  }

  // ===== compute warps =====
  // software pipeline: everything a step needs besides the other CTAs' partial products is
  // fetched one step ahead: table entries at the top of the previous step, this lane's
  // coupling entry of the first row pass and the right-hand side of the CTA's rows before
  // the previous step's dense product.
  //   rng = {lo, hi, lo2, hi2} coupling column ranges, own = {c0, c1, q0, q1} this CTA's rows
  //   of the step's node and of its input node
  int4 rng = make_int4(0, 0, 0, 0), own = make_int4(0, 0, 0, 0), nrng = rng, nown = own;
  double2 pv = zmake(0.0, 0.0), rv = zmake(0.0, 0.0), npv = pv, nrv = rv;
  int pc = -1, npc = -1;
  auto preload_table = [&](int sn) {
    if (sn >= S) return;
    const int4* tab = q.rng + ((size_t)blockIdx.x * q.smax + sn) * 2;
    nrng = __ldg(tab);
    nown = __ldg(tab + 1);
  };
  auto preload_operands = [&](int sn) {
    npv = zmake(0.0, 0.0);
    nrv = zmake(0.0, 0.0);
    npc = -1;
    if (sn >= S) return;
    int pn, mn;
    kb_flow_step(group, sn, P, mid, pn, mn);
    const int on = s_nptr[pn];
    const int a0 = nown.x, a1 = nown.y;
    if (mn <= KB_MID && tid < a1 - a0) nrv = q.r[on + a0 + tid];
    const bool useL = (mn == KB_FWD_L || mn == KB_MID || mn == KB_BWD_L);
    const int W = useL ? q.WL : q.WU;
    const bool halfw = W <= 16;
    const int hl = halfw ? (lane & 15) : lane;
    const int i = a0 + (halfw ? 2 * wid + (lane >> 4) : wid);
    if (i < a1 && hl < W) {
      const size_t e = (size_t)(on + i) * W + hl;
      npv = (useL ? q.Lval : q.Uval)[e];
      npc = (useL ? q.Lcol : q.Ucol)[e];
    }
  };
  preload_table(0);
  preload_operands(0);

  for (int s = 0; s < S; ++s) {
    int p, mode;
    kb_flow_step(group, s, P, mid, p, mode);
    const int o = s_nptr[p], b = s_nptr[p + 1] - o;
    rng = nrng;
    own = nown;
    pv = npv;
    pc = npc;
    rv = nrv;
    preload_table(s + 1);
    const int c0 = own.x, c1 = own.y;
    const int nc = c1 - c0;
    const bool fwd = mode <= KB_MID;
    // (slow warps may still be reading the other buffer in the previous step's dense product)
    double2* tl = tl0 + (s & 1) * tlen;
    K1_TICK(0);

    // ---- 1. inputs: the previous result of this chain (and, at the middle node and at the
    //         first backward step of group 1, the other chain's) on the rows this CTA needs
    int qn = -1;             // node whose result the previous step of THIS group produced
    bool prev_base = false;  // that result = yf - sum (backward) instead of sum
    if (s > 0) {
      int pm;
      kb_flow_step(group, s - 1, P, mid, qn, pm);
      prev_base = (pm == KB_BWD_U || pm == KB_BWD_L);
    }
    if (tid < nc) tl[tid] = rv;
    for (int i = tid + CT; i < nc; i += CT) tl[i] = fwd ? q.r[o + c0 + i] : zmake(0.0, 0.0);
    const bool cross = (group == 1 && s == nbot);  // first backward step of group 1: x_mid comes from group 0
    const bool useL = (mode == KB_FWD_L || mode == KB_MID || mode == KB_BWD_L);
    int lo = rng.x, hi = rng.y, qo = 0;
    bool have = false;
    if (fwd ? (qn >= 0) : true) {
      const int qq = fwd ? qn : (mode == KB_BWD_U ? p + 1 : p - 1);
      qo = s_nptr[qq];
      const int bq = s_nptr[qq + 1] - qo;
      const int q0 = own.z, q1 = own.w;  // own rows of that node: stored (yf / x); empty at the cross step
      if (q1 > q0) {
        lo = min(lo, q0);
        hi = max(hi, q1);
      }
      have = lo < hi;
      // everything this group published up to its previous step; at the cross step also the
      // middle node of group 0
      k1_wait(gctr[group], ebase[group] + (unsigned)s * (unsigned)gsize, q.err);
      if (cross) k1_wait(gctr[0], ebase[0] + (unsigned)(mid + 1) * (unsigned)gsz[0], q.err);
      if (have) {
        if (cross)
          k1_collect(q.xchg[0], ld, min(gsz[0], bq), qo, lo, hi, nullptr, false, va, nullptr, 0, 0);
        else if (fwd)
          k1_collect(part_of(group, s - 1), ld, min(gsize, bq), qo, lo, hi, nullptr, false, va, q.yf, q0, q1);
        else
          k1_collect(part_of(group, s - 1), ld, min(gsize, bq), qo, lo, hi, prev_base ? q.yf : nullptr, prev_base, va,
                     q.x, q0, q1);
      }
      if (cross && nc > 0) {
        // y_{mid+1} on this CTA's own rows (the forward result of this very node), the base
        // of the next step:  yf = sum of the last forward partials
        k1_collect(q.xchg[1], ld, min(gsize, b), o, c0, c1, nullptr, false, vb, q.yf, c0, c1);
      }
    }
    k1_bar_compute();
    K1_TICK(1);
    // ---- 2. t = r - C y  (forward)  |  t = C x  (backward): first row pass from registers
    if (have && nc > 0) {
      const int W = useL ? q.WL : q.WU;
      const bool halfw = W <= 16;
      const int hl = halfw ? (lane & 15) : lane;
      const int i = c0 + (halfw ? 2 * wid + (lane >> 4) : wid);
      double2 acc = zmake(0.0, 0.0);
      const int cc = pc - qo;
      if (pc >= 0 && cc >= lo && cc < hi) zfma(acc, pv, va[cc - lo]);
      if (!halfw && i < c1) {
        const double2* val = useL ? q.Lval : q.Uval;
        const int* col = useL ? q.Lcol : q.Ucol;
        for (int k = lane + 32; k < W; k += 32) {
          const size_t e = (size_t)(o + i) * W + k;
          const int c2 = __ldg(&col[e]) - qo;
          if (c2 >= lo && c2 < hi) zfma(acc, __ldg(&val[e]), va[c2 - lo]);
        }
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
      }
#pragma unroll
      for (int sft = 8; sft > 0; sft >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
      }
      if (i < c1 && hl == 0) tl[i - c0] = fwd ? zsub(tl[i - c0], acc) : acc;
      // rows beyond the first pass (CTAs owning more than 16 / 8 rows)
      const int rpp = halfw ? 2 * NW : NW;
      if (nc > rpp)
        k1_couple(useL ? q.Lval : q.Uval, useL ? q.Lcol : q.Ucol, W, o, c0 + rpp, c1, qo, lo, hi, va, tl + rpp, fwd);
    }
    if (mode == KB_MID && mid < P - 1) {
      // the other chain's last forward node: y_{mid+1}, coupled through U
      const int q2 = mid + 1;
      const int qo2 = s_nptr[q2], bq2 = s_nptr[q2 + 1] - qo2;
      const int lo2 = rng.z, hi2 = rng.w;
      k1_wait(gctr[1], ebase[1] + (unsigned)nbot * (unsigned)gsz[1], q.err);
      if (lo2 < hi2) k1_collect(q.xchg[1], ld, min(gsz[1], bq2), qo2, lo2, hi2, nullptr, false, vb, nullptr, 0, 0);
      k1_bar_compute();
      if (nc > 0 && lo2 < hi2) k1_couple(q.Uval, q.Ucol, q.WU, o, c0, c1, qo2, lo2, hi2, vb, tl, true);
    }
    k1_bar_compute();
    K1_TICK(2);
    preload_operands(s + 1);  // consumed next step: the latency hides behind the dense product

    // ---- 3. partial_c = M_p[:, C_c] t[C_c] from the staged slice of M_p^T, published
    if (nc > 0) {
      kb_mbar_wait(&mbar[s & 1], uses[s & 1] & 1u, q.err, KB_WAIT_NS_DEFAULT);
      uses[s & 1]++;
      K1_TICK(3);
      const double2* Ms = stage0 + (size_t)(s & 1) * slice_elems;
      double2* out = part_of(group, s) + (size_t)grank * ld;
      // three rows per thread in flight (independent accumulators)
      for (int i0 = tid; i0 < b; i0 += 3 * CT) {
        double2 acc[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) acc[u] = zmake(0.0, 0.0);
        for (int j = 0; j < nc; ++j) {
          const double2 tj = tl[j];
          const double2* Mj = Ms + (size_t)j * b;
#pragma unroll
          for (int u = 0; u < 3; ++u)
            if (i0 + u * CT < b) zfma(acc[u], Mj[i0 + u * CT], tj);
        }
        // strong (gpu-scope) stores: performed at L2 right away instead of lingering in the
        // SM's write path until a fence pushes them out
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (i0 + u * CT < b)
            asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(out + i0 + u * CT), "d"(acc[u].x),
                         "d"(acc[u].y)
                         : "memory");
      }
    }
    K1_TICK(4);
    // ---- 4. hand the step to the service warp (publication, stage refill)
    __threadfence_block();
    k1_bar_arrive(2);
    K1_TICK(5);
  }

  // ---- the result of the last step, on this CTA's own rows
  if (S > 0) {
    int p, mode;
    kb_flow_step(group, S - 1, P, mid, p, mode);
    const int o = s_nptr[p], b = s_nptr[p + 1] - o;
    int c0, c1;
    kb_group_rows(b, gsize, grank, c0, c1);
    const bool base = (mode == KB_BWD_U || mode == KB_BWD_L);
    k1_wait(gctr[group], ebase[group] + (unsigned)S * (unsigned)gsize, q.err);
    if (c1 > c0)
      k1_collect(part_of(group, S - 1), ld, min(gsize, b), o, c0, c1, base ? q.yf : nullptr, base, va, q.x, c0, c1);
  }
  if (q.timing && tid == 0)
    for (int k = 0; k < 8; ++k) q.timing[blockIdx.x * 8 + k] = tacc[k];
#undef K1_TICK
}

// Coupling ranges of every (CTA, step) of the sweep, computed once per factorisation with the
// grid geometry of the sweep kernel.
__global__ void __launch_bounds__(K1_THREADS) kb_onehop_ranges(K1Params q, int4* out) {
  __shared__ int s_rng[4];
  const int P = q.P, mid = q.mid;
  const int group = ((int)blockIdx.x < q.G0) ? 0 : 1;
  const int gsize = group == 0 ? q.G0 : (int)gridDim.x - q.G0;
  const int grank = group == 0 ? (int)blockIdx.x : (int)blockIdx.x - q.G0;
  const int nbot = P - 1 - mid;
  const int S = group == 0 ? 2 * mid + 1 : 2 * nbot;
  for (int s = 0; s < S; ++s) {
    int p, mode;
    kb_flow_step(group, s, P, mid, p, mode);
    const int o = (int)q.nodeptr[p], b = (int)q.nodeptr[p + 1] - o;
    int c0, c1;
    kb_group_rows(b, gsize, grank, c0, c1);
    int4 r = make_int4(0, 0, 0, 0);
    int qq = -1;
    bool useL = true;
    if (mode == KB_FWD_L || mode == KB_MID) {
      qq = p - 1;
    } else if (mode == KB_FWD_U) {
      qq = p + 1;
      useL = false;
    } else if (mode == KB_BWD_U) {
      qq = p + 1;
      useL = false;
    } else {
      qq = p - 1;
    }
    if ((mode == KB_FWD_L || mode == KB_FWD_U) && s == 0) qq = -1;
    if (mode == KB_MID && mid == 0) qq = -1;
    if (qq >= 0 && qq < P) {
      const int qo = (int)q.nodeptr[qq], bq = (int)q.nodeptr[qq + 1] - qo;
      k1_range(useL ? q.Lcol : q.Ucol, useL ? q.WL : q.WU, o, c0, c1, qo, bq, &s_rng[0], &s_rng[1]);
      r.x = s_rng[0];
      r.y = s_rng[1];
    }
    if (mode == KB_MID && mid < P - 1) {
      const int qo = (int)q.nodeptr[mid + 1], bq = (int)q.nodeptr[mid + 2] - qo;
      k1_range(q.Ucol, q.WU, o, c0, c1, qo, bq, &s_rng[2], &s_rng[3]);
      r.z = s_rng[2];
      r.w = s_rng[3];
    }
    if (threadIdx.x == 0) {
      int4 w = make_int4(c0, c1, 0, 0);
      const bool cross = (group == 1 && s == nbot);
      if (qq >= 0 && qq < P && !cross) {
        int q0, q1;
        kb_group_rows((int)q.nodeptr[qq + 1] - (int)q.nodeptr[qq], gsize, grank, q0, q1);
        w.z = q0;
        w.w = q1;
      }
      out[((size_t)blockIdx.x * q.smax + s) * 2] = r;
      out[((size_t)blockIdx.x * q.smax + s) * 2 + 1] = w;
    }
    __syncthreads();
  }
}

// Can the one-hop kernel run this chain?  (two slices of M^T must fit in shared memory)
bool kbi_onehop_supported(const kb_context* h, int G, bool two_sided, int* slice_elems_out, size_t* smem_out) {
  if (getenv("KB_NO_ONEHOP")) return false;
  if (G < 2) return false;
  const int gmin = two_sided ? G / 2 : G;
  if (gmin < 1) return false;
  const int64_t rpc = (h->bmax + gmin - 1) / gmin;
  const int64_t slice_elems = (rpc * h->bmax + 7) & ~(int64_t)7;
  const size_t vec = (size_t)((h->bmax + 7) & ~(int64_t)7) * sizeof(double2);
  const size_t tls = 2 * (size_t)((slice_elems / (h->bmax > 0 ? h->bmax : 1) + 8) & ~(int64_t)7) * sizeof(double2);
  const size_t tabs = (size_t)(h->P + 1) * (sizeof(int64_t) + sizeof(int)) + 16;
  const size_t smem = 2 * (size_t)slice_elems * sizeof(double2) + 2 * vec + tls + tabs;
  if (smem > 220 * 1024) return false;
  if (slice_elems_out) *slice_elems_out = (int)slice_elems;
  if (smem_out) *smem_out = smem;
  return true;
}

// Coupling ranges of every (CTA, step) for a grid of G CTAs, G0 of them in group 0 (built once per
// factorisation; shared with kb_sweep3.cu, which splits the chain the same way).
int kbi_onehop_build_ranges(kb_context* h, int G, int G0) {
  if (h->rng_valid) return KB_OK;
  K1Params q;
  memset(&q, 0, sizeof(q));
  q.nodeptr = h->d_nodeptr.p;
  q.P = (int)h->P;
  q.mid = (int)h->mid;
  q.Lcol = h->d_Lcol.p;
  q.WL = h->WL > 0 ? h->WL : 1;
  q.Ucol = h->d_Ucol.p;
  q.WU = h->WU > 0 ? h->WU : 1;
  q.G0 = G0;
  q.smax = (int)(2 * h->mid + 1);
  KB_CUDA(h, h->d_rng.alloc((size_t)G * q.smax * 2));
  kb_onehop_ranges<<<G, K1_THREADS, 0, h->stream>>>(q, h->d_rng.p);
  h->rng_valid = true;
  h->launches++;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}

// y <- T'^{-1} r on TRANSPOSED two-sided factors.  y has n+1 entries, y[n] == 0.
int kbi_sweep_onehop(kb_context* h, const double2* r, double2* y) {
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  const int G = h->sweep_grid;
  const bool two = h->mid < h->P - 1;
  int slice_elems = 0;
  size_t smem = 0;
  if (!kbi_onehop_supported(h, G, two, &slice_elems, &smem))
    return kb_fail(h, KB_EINVAL, "chain does not fit the one-hop sweep kernel");
  const int G0 = two ? (G + 1) / 2 : G;
  const size_t ldp = (size_t)((h->bmax + 7) & ~(int64_t)7);
  const size_t ring_elems = (size_t)K1_RING * G * ldp;  // both groups
  const size_t xchg_elems = 2 * (size_t)G0 * ldp;
  if (h->d_yf.count < (size_t)n + 1) {
    KB_CUDA(h, h->d_yf.alloc(n + 1));
  }
  KB_CUDA(h, h->d_ring.alloc(ring_elems + xchg_elems));
  if (h->d_k1flags.count < (size_t)G * K1_FSTRIDE) {
    KB_CUDA(h, h->d_k1flags.alloc((size_t)G * K1_FSTRIDE));
    KB_CUDA(h, cudaMemsetAsync(h->d_k1flags.p, 0, (size_t)G * K1_FSTRIDE * sizeof(unsigned), s));
    h->k1_epoch = 0;
    h->k1_epoch1 = 0;
  }
  K1Params q;
  q.MT = h->d_M.p;
  q.Moff = h->d_Moff.p;
  q.nodeptr = h->d_nodeptr.p;
  q.P = (int)h->P;
  q.mid = (int)h->mid;
  q.r = r;
  q.yf = h->d_yf.p;
  q.x = y;
  q.ring[0] = h->d_ring.p;
  q.ring[1] = h->d_ring.p + (size_t)K1_RING * G0 * ldp;
  q.xchg[0] = h->d_ring.p + ring_elems;
  q.xchg[1] = h->d_ring.p + ring_elems + (size_t)G0 * ldp;
  q.Lval = h->d_Lval.p;
  q.Lcol = h->d_Lcol.p;
  q.WL = h->WL > 0 ? h->WL : 1;
  q.Uval = h->d_Uval.p;
  q.Ucol = h->d_Ucol.p;
  q.WU = h->WU > 0 ? h->WU : 1;
  q.flags = h->d_k1flags.p;
  {
    // every CTA of a group adds one per step: 2 mid + 1 steps in group 0, 2 (P - 1 - mid) in group 1
    const unsigned s0 = (unsigned)(2 * h->mid + 1), s1 = (unsigned)(2 * (h->P - 1 - h->mid));
    q.epoch0 = h->k1_epoch;
    q.epoch1 = h->k1_epoch1;
    h->k1_epoch += s0 * (unsigned)G0;
    h->k1_epoch1 += s1 * (unsigned)(G - G0);
  }
  q.err = h->d_sweep_err.p;
  q.timing = h->d_sweep_timing.p;
  q.bmax = (int)h->bmax;
  q.G0 = G0;
  q.dbgflags = getenv("KB_ONEHOP_DBG") ? atoi(getenv("KB_ONEHOP_DBG")) : 0;
  q.smax = (int)(2 * h->mid + 1);
  KB_TRY(kbi_onehop_build_ranges(h, G, G0));
  q.rng = h->d_rng.p;
  const void* fn = (const void*)kb_sweep_onehop;
  if (smem > 48 * 1024) KB_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&q, (void*)&slice_elems};
  KB_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(K1_THREADS), args, smem, s));
  h->launches += 1;
  return KB_OK;
}
