// Persistent chain factorisation: the whole two-sided block-Thomas elimination of
// T = A - sigma B in ONE cooperative launch.
//
// Replaces the numeric factorisation of MUMPS / SuperLU_DIST inside E.solve()
// (/root/reference/bin/solve.py:123) and K.solve (solve.py:227); same algebra as
// kb_factor.cu (explicit inverses M_p = S_p^-1 of the Schur blocks, Gauss-Jordan with
// partial pivoting inside the block), different execution model:
//
//   * the CTAs of the grid form one group per elimination chain (top-down and
//     bottom-up, "burn at both ends"); a group walks its chain node by node;
//   * inside a node, CTA c owns the column strip [c w, (c+1) w) of the b x b Schur
//     block for the whole inversion and keeps it IN REGISTERS, one matrix row per
//     thread (w = 8 or 9 columns = 32/36 registers);
//   * step k of the blocked Gauss-Jordan elimination: the owner of strip k runs the
//     panel (pivot search, row swaps, elimination over its w columns) on its registers
//     and publishes the composite transform G_k (b x w) and the pivot list through
//     L2; every other CTA applies  A <- P_k A + G_k R_k  to its strip, R_k being the
//     pivot rows of its own strip (exchanged through shared memory);
//   * there are no kernel boundaries and no grid-wide barriers inside a node: a
//     release-store of a step counter publishes G_k, consumers poll it; the owner of
//     strip k+1 runs ahead of the others (look-ahead comes for free);
//   * Schur-block formation S = D - C_pq (M_q C_qp) is done strip-wise by the same
//     CTAs from the sparse couplings, so the dense block never exists in memory
//     before it is inverted; only M_p is written (16 b^2 bytes per node).
//
// tests/strip_gj_model.py is the NumPy statement of the strip algebra (net row
// permutation per step, slot exchange, running permutation).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "kb_internal.cuh"

// Two instantiations: <9, 640> for nodes up to 640 rows (strips of <= 9 columns over 74 CTAs per
// chain: every size the benchmark names) and <10, 704> for nodes up to 704 rows (Kore's own
// resolution rule at E = 1e-8 gives N = 676), which has 88 registers per thread and spills a
// little.  NB = columns per strip (compile-time bound of the register row), MAXT = threads per
// CTA = largest supported node.  The row pitch (double2) of the W strip in shared memory is
// NB | 1 (odd: conflict-free).
#define KF_NB 9
#define KF_MAXT 640
#define KF_NB_WIDE 10
#define KF_MAXT_WIDE 704
#define KF_WMIN 8     // narrowest strip used (fewer, fatter steps for small nodes)
#define KF_RING 4     // strips of the column stream in flight (ring slots; a power of two)
#define KF_SYNC_WORDS 512  // sync block: 2 x 96 words of counters, time-out flag at KF_ERR_WORD, ring counters from 192

struct KfParams {
  int plo, P, mid;  // the chain is nodes [plo, P): group 0 walks plo .. mid, group 1 walks P-1 .. mid+1
  const int64_t* nodeptr;
  const int64_t* Moff;
  double2* M;
  const int64_t* rowptr;
  const int64_t* dstart;
  const int64_t* ustart;
  const int* col;
  const double2* T;
  const int64_t* ucptr;
  const int* urow;
  const int64_t* upos;
  const int64_t* lcptr;
  const int* lrow;
  const int64_t* lpos;
  double2* Gbuf[2];   // per group: bmax x bmax, column-major (column = global pivot column)
  int* pivbuf[2];     // per group: 16 ints per step
  unsigned* sync;     // per group 96 words: [0] published steps, [32] barrier arrivals, [64] chain done;
                      // from word 192: one 128-byte line per (group, ring slot): consumers done with the slot
  int* info;
  int* err;
  unsigned long long wait_ns;  // time bound of every wait of the launch
  long long* dbg;     // optional: 16 cycle counters per CTA
  int bmax;
  int G0;             // CTAs of group 0 (the rest form group 1)
  int stagger_ns;     // consumers other than the next owner hold their loads of G_k back by this much
  int transposed;     // store M_p^T (row j of the buffer = column j of M_p) for kb_sweep1.cu
  // per-column stream of the elementary transforms to every other strip: 32-byte tagged
  // elements (re, tag, im, tag), [KF_RING strips][NB columns][bmax rows] per group (strip k uses
  // ring slot k mod KF_RING; its owner does not write before every consumer of strip k - KF_RING
  // has said it is done with the slot); null: consumers wait for the composite G_k
  double* Sbuf[2];
  // non-null: the (single) node's Schur block is this dense row-major b x b matrix instead of being
  // formed from the sparse pencil (separator blocks of the l-sharded reduced system)
  const double2* dense;
};

__device__ __forceinline__ unsigned kf_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void kf_st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Bounded waits without a call and without per-thread state (the strip rows leave no registers
// for a clock, and a call site constrains the register allocation of the whole kernel): every
// wait is measured against a DEADLINE kept in shared memory -- `deadline[0]` for the waits of
// thread 0 on counters (set by thread 0 itself when it starts to spin), `deadline[1]` for the
// polls of the column stream (set by thread 0 at the top of every strip step) -- and looked at
// only every 32nd missed poll, together with the launch's flag; `failed` is the CTA's own copy of
// "the launch has failed" (every 4th missed poll looks at it).
struct KfWait {
  unsigned long long deadline[2];
  int failed;
};

// every 32nd missed poll: true = stop waiting
__device__ __forceinline__ bool kf_expired(int* err, int code, volatile KfWait* w, int which) {
  if (*(volatile int*)err != 0) {
    w->failed = 1;
    return true;
  }
  if (kb_globaltimer() > w->deadline[which]) {
    atomicCAS(err, 0, code | ((int)blockIdx.x << 8));
    w->failed = 1;
    return true;
  }
  return false;
}

// one thread: wait until *ctr >= target
__device__ __forceinline__ void kf_spin_until(const unsigned* ctr, unsigned target, int* err, int code,
                                              unsigned long long wait_ns, volatile KfWait* w) {
  if (kf_ld_acquire(ctr) >= target) return;
  if (w->failed) return;
  w->deadline[0] = kb_globaltimer() + wait_ns;
  int misses = 0;
  while (kf_ld_acquire(ctr) < target)
    if ((++misses & 31) == 0 && kf_expired(err, code, w, 0)) break;
}

// thread 0 waits until *ctr >= target; everybody leaves through the block barrier
__device__ __forceinline__ void kf_wait(const unsigned* ctr, unsigned target, int* err, unsigned long long wait_ns,
                                        volatile KfWait* w) {
  if (threadIdx.x == 0) kf_spin_until(ctr, target, err, KB_WERR_COUNTER, wait_ns, w);
  __syncthreads();
}

__device__ __forceinline__ void kf_group_barrier(unsigned* ctr, unsigned target, int* err, unsigned long long wait_ns,
                                                 volatile KfWait* w) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    kf_spin_until(ctr, target, err, KB_WERR_GROUP_BARRIER, wait_ns, w);
  }
  __syncthreads();
}

// 1/d to ~1 ulp: hardware seed (2^-23) and two Newton steps; d is a sum of squares of
// equilibrated entries, far from the subnormal range the .ftz seed flushes
__device__ __forceinline__ double kf_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Tagged element of the column stream: written with one 256-bit store, polled with 256-bit
// loads.  tag = (running column number within the chain) * 1024 + (pivot row + 1): unique per
// column of a factorisation (the buffer is zeroed before the launch), identical in both 16-byte
// halves so that a torn read is never accepted, and it tells the reader the pivot row.
__device__ __forceinline__ void kf_stream_put(double* p, double2 v, double tag) {
  asm volatile("st.relaxed.gpu.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v.x), "d"(tag), "d"(v.y), "d"(tag)
               : "memory");
}
template <int NB>
struct KfShared {
  double2 slots[2][NB][NB];                  // pivot rows of this CTA's strip for the step being applied
  double2 prow[2][NB + 1];                   // the pivot row of the column being eliminated (+ 1/pivot)
  unsigned long long candkey[2][32];         // (key << 32) | ~row of every warp's candidate
  int piv[16];                               // pivots of the panel being factored
  unsigned consbase[KF_RING];                // consumers that had released each ring slot before this node
  KfWait wait;                               // deadlines and failure flag of the bounded waits
};

// S strip -= C_pq (M_q C_qp)[:, strip].  kind 0: q = p-1 (C_qp = U by column, C_pq = the
// sub-diagonal part of the rows of p); kind 1: q = p+1 (C_qp = L by column, C_pq = the
// super-diagonal part of the rows).
template <int NB>
__device__ __forceinline__ void kf_schur_term(const KfParams& q, double2 (&a)[NB], int kind, int p, int qn, int j0,
                                              int ws, double2* Ws) {
  const int t = threadIdx.x;
  const int o = (int)q.nodeptr[p], b = (int)q.nodeptr[p + 1] - o;
  const int oq = (int)q.nodeptr[qn], bq = (int)q.nodeptr[qn + 1] - oq;
  const double2* Mq = q.M + q.Moff[qn];
  const int64_t* cptr = kind == 0 ? q.ucptr : q.lcptr;
  const int* crow = kind == 0 ? q.urow : q.lrow;
  const int64_t* cpos = kind == 0 ? q.upos : q.lpos;
  // W[k, jj] = sum_e M_q[k, row_e] C_qp[row_e, j0 + jj]   (thread k = row of node q)
  if (t < bq) {
    // element (t, rr) of M_q: row-major, or column-major when the factors are stored transposed
    const double2* Mrow = q.transposed ? Mq + t : Mq + (size_t)t * bq;
    const size_t mstride = q.transposed ? (size_t)bq : 1;
    for (int jj = 0; jj < ws; ++jj) {
      const int gc = o + j0 + jj;
      const int64_t e0 = __ldg(&cptr[gc]), e1 = __ldg(&cptr[gc + 1]);
      double2 acc = zmake(0.0, 0.0);
      for (int64_t e = e0; e < e1; ++e) {
        const int rr = __ldg(&crow[e]) - oq;
        const double2 cv = __ldg(&q.T[__ldg(&cpos[e])]);
        zfma(acc, __ldcg(&Mrow[(size_t)rr * mstride]), cv);
      }
      Ws[(size_t)t * (NB | 1) + jj] = acc;
    }
  }
  __syncthreads();
  if (t < b) {
    const int gi = o + t;
    const int64_t k0 = kind == 0 ? __ldg(&q.rowptr[gi]) : __ldg(&q.ustart[gi]);
    const int64_t k1 = kind == 0 ? __ldg(&q.dstart[gi]) : __ldg(&q.rowptr[gi + 1]);
    for (int64_t k = k0; k < k1; ++k) {
      const int cq = __ldg(&q.col[k]) - oq;
      const double2 v = __ldg(&q.T[k]);
      const double2* wr = Ws + (size_t)cq * (NB | 1);
#pragma unroll
      for (int jj = 0; jj < NB; ++jj)
        if (jj < ws) zfms(a[jj], v, wr[jj]);
    }
  }
  __syncthreads();
}

template <int NB, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) kb_chain_factor(KfParams q) {
  extern __shared__ __align__(16) unsigned char kf_smem[];
  KfShared<NB>& sh = *(KfShared<NB>*)kf_smem;
  int* s_orig = (int*)(kf_smem + sizeof(KfShared<NB>));
  double2* Ws = (double2*)(kf_smem + sizeof(KfShared<NB>) + (((size_t)q.bmax * sizeof(int) + 15) & ~(size_t)15));

  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int NW = (int)blockDim.x >> 5;
  const int group = (int)blockIdx.x < q.G0 ? 0 : 1;
  const int Gc = group == 0 ? q.G0 : (int)gridDim.x - q.G0;
  const int c = group == 0 ? (int)blockIdx.x : (int)blockIdx.x - q.G0;
  const int P = q.P, mid = q.mid;
  const int S = group == 0 ? mid - q.plo + 1 : P - 1 - mid;
  unsigned* pub = q.sync + group * 96;
  unsigned* bar = q.sync + group * 96 + 32;
  unsigned* done1 = q.sync + 96 + 64;
  double2* Gbuf = q.Gbuf[group];
  int* pivbuf = q.pivbuf[group];
  const int ldg = q.bmax;
  unsigned pubbase = 0, barcount = 0;
  long long colbase = 0;  // columns eliminated by this group before the current node
  // ring counters of this group: slot j's line counts the CTAs that are done reading it
  unsigned* cons = q.sync + 192 + group * (KF_RING * 32);
  if (t == 0) {
    for (int j = 0; j < KF_RING; ++j) sh.consbase[j] = 0u;
    sh.wait.failed = 0;
    sh.wait.deadline[0] = sh.wait.deadline[1] = ~0ull;
  }
  __syncthreads();
  volatile KfWait* kw = &sh.wait;
  long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tc = clock64();
  // Cycle counters are compiled in only with -DKB_FACTOR_TIMING (make EXTRA=-DKB_FACTOR_TIMING):
  // even switched off at run time each tick costs ~15 predicated instructions, four per panel
  // column, in a loop that is bound by instruction issue (ncu source page, profiles/).
#ifdef KB_FACTOR_TIMING
#define KF_TICK(k)                \
  do {                            \
    if (q.dbg) {                  \
      long long _n = clock64();   \
      tacc[k] += _n - tc;         \
      tc = _n;                    \
    }                             \
  } while (0)
#else
#define KF_TICK(k) \
  do {             \
  } while (0)
#endif

  for (int s = 0; s < S; ++s) {
    const int p = group == 0 ? q.plo + s : P - 1 - s;
    const int o = (int)q.nodeptr[p], b = (int)q.nodeptr[p + 1] - o;
    int w = (b + Gc - 1) / Gc;
    if (w < KF_WMIN) w = KF_WMIN;
    const int K = (b + w - 1) / w;
    const bool active = c < K;
    const int j0 = c * w;
    const int ws = active ? min(w, b - j0) : 0;

    double2 a[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) a[j] = zmake(0.0, 0.0);

    // the middle node needs the last node of the other chain
    if (group == 0 && p == mid && mid < P - 1) kf_wait(done1, 1u, q.err, q.wait_ns, kw);

    // ---- Schur block, strip-wise:  S = D_p - C_pq M_q C_qp  (one or two eliminated neighbours)
    if (active) {
      if (group == 0 && s > 0) kf_schur_term<NB>(q, a, 0, p, p - 1, j0, ws, Ws);
      if (group == 1 && s > 0) kf_schur_term<NB>(q, a, 1, p, p + 1, j0, ws, Ws);
      if (group == 0 && p == mid && mid < P - 1) kf_schur_term<NB>(q, a, 1, p, p + 1, j0, ws, Ws);
      if (q.dense) {
        if (t < b) {
          const double2* dr = q.dense + (size_t)t * b + j0;
#pragma unroll
          for (int u = 0; u < NB; ++u)
            if (u < ws) a[u] = dr[u];
        }
      } else if (t < b) {
        // D_p[t, strip]: the row's own-node entries are sorted by column; lower bound, then <= ws entries
        const int gi = o + t;
        int64_t lo = __ldg(&q.dstart[gi]), hi = __ldg(&q.ustart[gi]);
        const int64_t end = hi;
        const int c0 = o + j0, c1 = o + j0 + ws;
        while (lo < hi) {
          int64_t m = (lo + hi) >> 1;
          if (__ldg(&q.col[m]) < c0)
            lo = m + 1;
          else
            hi = m;
        }
        for (int64_t k = lo; k < end; ++k) {
          const int cc = __ldg(&q.col[k]);
          if (cc >= c1) break;
          const double2 v = __ldg(&q.T[k]);
          const int jj = cc - c0;
#pragma unroll
          for (int u = 0; u < NB; ++u)
            if (u == jj) a[u] = zadd(a[u], v);
        }
      }
    }
    KF_TICK(0);

    // ---- blocked Gauss-Jordan with IMPLICIT row pivoting, one step per strip.  A row never
    //      moves: thread t keeps original row t; `mycol` < 0 says it has not been a pivot yet,
    //      `mycol` which column it pivoted.  (tests/strip_gj_model.py)
    if (active) {
      int mycol = t < b ? -1 : 0;  // column this row has pivoted; -1: the row is still free
      for (int k = 0; k < K; ++k) {
        const int k0 = k * w;
        const int wk = min(w, b - k0);
        // deadline of this step's polls of the column stream
        if (t == 0 && q.Sbuf[group] && k != c) sh.wait.deadline[1] = kb_globaltimer() + q.wait_ns;
#ifndef KF_NO_BP
        // Back-pressure: strip c reuses the ring slot of strip c - KF_RING; its owner stores
        // nothing before every consumer of that strip has released the slot.  The CTA makes sure
        // of it TWO steps before its panel (and at the first step of a node for strips 0 and 1):
        // the acquire load costs an L2 round trip, which must happen neither in the panel nor
        // while this CTA is the next owner (the hand-over is the critical path); two steps ahead
        // strips c - 4 and c - 3 are complete and their consumers have long let go, so the wait
        // itself practically never spins.  Nobody reads the slot again before strip c is written,
        // so the early check is as good as a late one.  EVERY thread does the (same) load and
        // the (same) wait: a wait by thread 0 alone is a divergent loop inside the strip loop,
        // after which ptxas no longer treats the warp as converged in the panel below (no uniform
        // datapath, convergence barriers around every branch: +14 instructions per panel column,
        // measured +10 % on the whole factorisation).
        if (q.Sbuf[group] && (k + 2 == c || (k == 0 && c < 2))) {
          const unsigned need = sh.consbase[c & (KF_RING - 1)] + (unsigned)(K - 1) * (unsigned)(c / KF_RING);
          kf_spin_until(cons + (c & (KF_RING - 1)) * 32, need, q.err, KB_WERR_BACKPRESSURE, q.wait_ns, kw);
        }
#endif
        if (k == c) {
          // ================= panel: this CTA's own strip =================
          // Two block barriers per column: (1) every warp has published the key of its best
          // free row, (2) the winning row has been published.  Each thread carries 1/|a|^2 of
          // its own entry in the next pivot column (computed under the shadow of the
          // elimination FMAs), so the winner publishes 1/pivot without a reciprocal on the
          // critical path.
          double rinv;
          // column stream of this strip: slot of this row, tag of column 0 with pivot code 0
          double* sput = q.Sbuf[group] ? q.Sbuf[group] + 4 * ((size_t)((k & (KF_RING - 1)) * NB) * ldg + t) : nullptr;
          const double stag = (double)((colbase + k0 + 1) * 1024);
          {
            const double m2 = zabs2(a[0]);
            const unsigned key = mycol < 0 ? (unsigned)__double2hiint(m2) + 1u : 0u;
            const unsigned wmax = __reduce_max_sync(0xffffffffu, key);
            const unsigned wlo = __reduce_max_sync(0xffffffffu, (key == wmax) ? ~(unsigned)t : 0u);
            if (lane == 0) sh.candkey[0][wid] = wmax ? (((unsigned long long)wmax << 32) | wlo) : 0ull;
            rinv = kf_rcp(m2);
          }
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) {
            if (cc < wk) {
              const int par = cc & 1;
              const int gk = k0 + cc;
              KF_TICK(8);
              __syncthreads();
              KF_TICK(9);
              const unsigned long long ck = (lane < NW) ? sh.candkey[par][lane] : 0ull;
              const unsigned khi = (unsigned)(ck >> 32), klo = (unsigned)ck;
              const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
              const unsigned mlo = __reduce_max_sync(0xffffffffu, (khi == mhi) ? klo : 0u);
              int rp = (int)(~mlo);
              // key 0: no candidate; 1: |pivot|^2 zero/denormal; >= 0x7ff00001: inf or NaN
              const bool broken = mhi <= 1u || mhi >= 0x7ff00001u || rp >= b || rp < 0;
              if (broken) rp = -1;  // skip the column; the host reports KB_ESINGULAR
              const bool isp = (t == rp);
              // warp-uniform: only the warp that holds the pivot row executes the publication and
              // the selects of the update below; a predicated-off instruction still takes an issue
              // slot, and the panel is bound by instruction issue
              const bool anyp = (rp >= 0) && ((rp >> 5) == wid);
              if (anyp) {
                if (isp) {
#pragma unroll
                  for (int j = 0; j < NB; ++j) sh.prow[par][j] = a[j];
                  sh.prow[par][NB] = zmake(a[cc].x * rinv, -a[cc].y * rinv);
                }
              }
              if (t == (int)blockDim.x - 1) {
                if (broken) atomicExch(q.info, o + gk + 1);
                sh.piv[cc] = rp;
                s_orig[gk] = broken ? 0 : rp;
              }
              __syncthreads();
              KF_TICK(10);
              const double2* prow = sh.prow[par];
              const double2 pinv = broken ? zmake(0.0, 0.0) : prow[NB];
              // One branch-free update for every row:  a[j] <- base + mult * prow[j]  with
              //   ordinary row:  base = a[j], mult = -a[cc]/pivot          (elimination)
              //   pivot row:     base = 0,    mult = 1/pivot, prow == a    (row / pivot)
              // so the warp of the pivot row does not run a scaling pass of its own while the other
              // nineteen wait at the next barrier (measured: 0.2 K cycles per column)
              const double2 mult = broken ? zmake(0.0, 0.0) : (isp ? pinv : zneg(zmul(a[cc], pinv)));
              // the elementary transform of this column goes to the other strips right away
              // (multiplier of this row, 1/pivot on the pivot row): the owner of the next strip
              // trails the panel by one L2 hop instead of waiting for the whole strip
              if (sput && t < b)
                kf_stream_put(sput + 4 * (size_t)cc * ldg, isp ? mult : zneg(mult), stag + (double)(1024 * cc + rp + 1));
              if (anyp) {
                // the pivot row restarts from zero: a[j] <- 0 + (1/pivot) * a[j]
#pragma unroll
                for (int j = 0; j < NB; ++j)
                  if (isp) a[j] = zmake(0.0, 0.0);
              }
#define KF_UPD(j) zfma(a[j], mult, prow[j])
              // next column first, so that its pivot vote overlaps the rest of the elimination
              if (cc + 1 < NB) {
                KF_UPD(cc + 1);
                if (cc + 1 < wk) {
                  const double m2 = zabs2(a[cc + 1]);
                  const unsigned key = (mycol < 0 && !isp) ? (unsigned)__double2hiint(m2) + 1u : 0u;
                  const unsigned wmax = __reduce_max_sync(0xffffffffu, key);
                  const unsigned wlo = __reduce_max_sync(0xffffffffu, (key == wmax) ? ~(unsigned)t : 0u);
                  if (lane == 0) sh.candkey[par ^ 1][wid] = wmax ? (((unsigned long long)wmax << 32) | wlo) : 0ull;
                  rinv = kf_rcp(m2);
                }
              }
#pragma unroll
              for (int j = 0; j < NB; ++j)
                if (j != cc && j != cc + 1) KF_UPD(j);
#undef KF_UPD
              if (!broken) a[cc] = mult;  // -g, or 1/pivot on the pivot row
              if (isp) mycol = gk;
              KF_TICK(11);
            }
          }
          KF_TICK(8);
          // publish G_k (the strip itself) and the pivots -- only when the consumers apply the
          // composite transform (no column stream)
          if (!q.Sbuf[group]) {
          if (t < b) {
#pragma unroll
            for (int j = 0; j < NB; ++j)
              if (j < wk) Gbuf[(size_t)(k0 + j) * ldg + t] = a[j];
          }
          __syncthreads();
          if (t < 16) pivbuf[k * 16 + t] = (t < wk) ? sh.piv[t] : -1;
          __syncthreads();
          if (t == 0) kf_st_release(pub, pubbase + (unsigned)k + 1u);
          }
          KF_TICK(1);
        } else if (q.Sbuf[group]) {
          // ================= consumer: apply step k column by column from the stream =========
          //   A[i,:] <- A[i,:] - g_i A[piv,:]  (i != piv),   A[piv,:] <- A[piv,:] / pivot
          // The owner of the next strip trails the panel by one L2 hop per column; everybody
          // else finds the columns already there.  Polls go out three columns at a time, so a
          // CTA that is catching up pays one round trip per three columns.
          KF_TICK(2);
          const double* sbase = q.Sbuf[group] + 4 * ((size_t)((k & (KF_RING - 1)) * NB) * ldg + t);
          const double tag0 = (double)((colbase + k0 + 1) * 1024);
#pragma unroll
          for (int c3 = 0; c3 < NB; c3 += 3) {
            double re[3], t0[3], im[3], t1[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
              re[u] = im[u] = 0.0;
              t0[u] = t1[u] = -1.0;
              if (c3 + u < wk && t < b)
                asm volatile("ld.relaxed.gpu.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                             : "=d"(re[u]), "=d"(t0[u]), "=d"(im[u]), "=d"(t1[u])
                             : "l"(sbase + 4 * (size_t)(c3 + u) * ldg)
                             : "memory");
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
              const int cc = c3 + u;
              if (cc < NB && cc < wk) {
                const double tb = tag0 + 1024.0 * cc;
                int rp = -1;
                double2 sv = zmake(re[u], im[u]);
                if (t < b) {
                  if (!(t0[u] == t1[u] && t0[u] >= tb && t0[u] < tb + 1024.0)) {
                    // this CTA has caught up with the panel (it is the next owner, or close):
                    // poll.  Every 4th missed poll looks at the CTA's failure flag in shared memory,
                    // every 32nd at the launch's flag and the clock (out of line).
                    const double* ep = sbase + 4 * (size_t)cc * ldg;
                    int misses = 0;
                    for (;;) {
                      asm volatile("ld.relaxed.gpu.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                                   : "=d"(re[u]), "=d"(t0[u]), "=d"(im[u]), "=d"(t1[u])
                                   : "l"(ep)
                                   : "memory");
                      if (t0[u] == t1[u] && t0[u] >= tb && t0[u] < tb + 1024.0) break;
                      // (a normal wait ends within two or three polls: nothing but the poll until then)
                      if ((++misses & 3) == 0) {
                        if (kw->failed) break;
                        if ((misses & 31) == 0 && kf_expired(q.err, KB_WERR_STREAM, kw, 1)) break;
                      }
                    }
                    sv = zmake(re[u], im[u]);
                  }
                  // (no match after the wait: the launch has failed; the column is skipped)
                  if (t0[u] == t1[u] && t0[u] >= tb && t0[u] < tb + 1024.0) rp = (int)(t0[u] - tb) - 1;
                }
                const bool isp = (t < b) && (t == rp);
                if (isp) {
#pragma unroll
                  for (int j = 0; j < NB; ++j) sh.prow[cc & 1][j] = a[j];
                  s_orig[k0 + cc] = rp;
                  mycol = k0 + cc;
                }
                if (t == 0 && rp < 0) s_orig[k0 + cc] = 0;  // skipped column (singular block)
                KF_TICK(5);
                __syncthreads();
                KF_TICK(6);
                if (t < b && rp >= 0) {
                  const double2* prow = sh.prow[cc & 1];
                  if (isp) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) a[j] = zmul(a[j], sv);
                  } else {
#pragma unroll
                    for (int j = 0; j < NB; ++j) zfms(a[j], sv, prow[j]);
                  }
                }
              }
            }
          }
          // every thread of the CTA had its elements of strip k in registers before it passed the
          // last barrier above: the ring slot may be rewritten as far as this CTA is concerned
          // (relaxed: nothing is published, and a release fence here would sit on the hand-over path)
#ifndef KF_NO_BP
          if (t == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(cons + (k & (KF_RING - 1)) * 32) : "memory");
#endif
          KF_TICK(3);
        } else {
          // ================= consumer: apply step k to this strip =================
          //   A[i,:] <- (i is a pivot row of the step ? 0 : A[i,:]) + sum_c G[i,c] A[piv_c,:]
          kf_wait(pub, pubbase + (unsigned)k + 1u, q.err, q.wait_ns, kw);
          // the owner of the next strip is on the critical path: it reads G_k from an idle L2
          if (q.stagger_ns > 0 && c != k + 1) __nanosleep((unsigned)q.stagger_ns);
          KF_TICK(2);
          int pv[NB];
          {
            const int4* pp = (const int4*)(pivbuf + k * 16);
            const int4 v0 = __ldcg(pp), v1 = __ldcg(pp + 1), v2 = __ldcg(pp + 2);
            const int tmp[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
            for (int j = 0; j < NB; ++j) pv[j] = tmp[j];
          }
          double2 g[NB];
#pragma unroll
          for (int j = 0; j < NB; ++j)
            g[j] = (j < wk && t < b) ? __ldcg(&Gbuf[(size_t)(k0 + j) * ldg + t]) : zmake(0.0, 0.0);
          const int par = k & 1;
          KF_TICK(12);
          bool isp = false;
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) {
            if (cc < wk && pv[cc] == t) {
              isp = true;
              mycol = k0 + cc;
#pragma unroll
              for (int j = 0; j < NB; ++j) sh.slots[par][cc][j] = a[j];
            }
          }
          if (t == (int)blockDim.x - 1) {
#pragma unroll
            for (int cc = 0; cc < NB; ++cc)
              if (cc < wk) s_orig[k0 + cc] = pv[cc] < 0 ? 0 : pv[cc];
          }
          KF_TICK(13);
          __syncthreads();
          KF_TICK(14);
          if (isp) {
#pragma unroll
            for (int j = 0; j < NB; ++j) a[j] = zmake(0.0, 0.0);
          }
#pragma unroll
          for (int cc = 0; cc < NB; ++cc) {
            if (cc < wk && pv[cc] >= 0) {
              const double2* rr = sh.slots[par][cc];
#pragma unroll
              for (int j = 0; j < NB; ++j) zfma(a[j], g[cc], rr[j]);
            }
          }
          KF_TICK(3);
        }
      }
      // ---- the strips now hold Y with Y[piv_c, :] = X[c, :], X = (Pi S)^-1, so
      //      M_p[mycol(t), piv_j] = Y[t, j]   (s_orig[j] = row that pivoted column j)
      __syncthreads();
      if (t < b) {
        double2* Mp = q.M + q.Moff[p];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          if (j < ws) {
            const int cj = s_orig[j0 + j];
            Mp[q.transposed ? (size_t)cj * b + mycol : (size_t)mycol * b + cj] = a[j];
          }
        }
      }
    }
    pubbase += (unsigned)K;
    colbase += b;
    barcount += (unsigned)Gc;
    // releases of this node, per ring slot: K - 1 consumers for every strip k = slot (mod KF_RING)
    if (t == 0 && q.Sbuf[group])
      for (int j = 0; j < KF_RING; ++j)
        if (j < K) sh.consbase[j] += (unsigned)(K - 1) * (unsigned)((K - j + KF_RING - 1) / KF_RING);
    kf_group_barrier(bar, barcount, q.err, q.wait_ns, kw);
    KF_TICK(4);
  }
  if (group == 1 && c == 0 && t == 0 && S > 0) kf_st_release(done1, 1u);
  if (q.dbg && t == 0)
    for (int k = 0; k < 16; ++k) q.dbg[(size_t)blockIdx.x * 16 + k] = tacc[k];
#undef KF_TICK
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool kbi_chainfac_supported(const kb_context* h) {
  if (getenv("KB_NO_CHAINFAC")) return false;
  if (h->opt_factor == 0) return false;
  if (h->bmax > KF_MAXT_WIDE) return false;
  // every strip must fit the register row: ceil(bmax / CTAs per chain) <= KF_NB (KF_NB_WIDE)
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device) != cudaSuccess) return false;
  if (getenv("KB_CHAINFAC_GRID")) {
    int g = atoi(getenv("KB_CHAINFAC_GRID"));
    if (g >= 2 && g < sms) sms = g;
  }
  const int gc = sms / 2;
  if (gc < 1 || (h->bmax + gc - 1) / gc > KF_NB_WIDE) return false;
  return true;
}

// Factor the chain nodes [plo, phi) as a block-tridiagonal system of their own (T must have been
// built; phi < 0: the whole chain).  two_sided: eliminate from both ends towards node h->mid;
// otherwise top-down only (h->mid = phi - 1).
int kbi_chainfac_run(kb_context* h, bool two_sided, bool transposed, int64_t plo, int64_t phi,
                     const KbDenseNode* dense) {
  cudaStream_t s = h->stream;
  const int64_t bmax = h->bmax;
  if (phi < 0) phi = h->P;
  int dev = h->device, sms = 0, coop = 0;
  KB_CUDA(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  KB_CUDA(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return kb_fail(h, KB_ECUDA, "device does not support cooperative launches");
  int G = sms;
  if (getenv("KB_CHAINFAC_GRID")) {
    int g = atoi(getenv("KB_CHAINFAC_GRID"));
    if (g >= 2 && g < G) G = g;
  }
  if (h->nranks == 1 && !dense) {
    // no node has more than ceil(bmax / KF_WMIN) strips: CTAs beyond that would only take part in
    // the barriers (small pencils: forced_small has 5 strips per node; several such launches can
    // then share the GPU, kore_b200/sweep.py).  One-GPU factorisations only: the l-sharded path
    // keeps the grid its 2- to 8-GPU validation ran with.
    const int kmax = (int)((bmax + KF_WMIN - 1) / KF_WMIN);
    const int need = (two_sided && !dense ? 2 : 1) * kmax;
    if (need < G) G = need < 2 ? 2 : need;
  }
  KfParams q;
  q.plo = (int)plo;
  q.P = (int)phi;
  q.mid = (int)h->mid;
  q.nodeptr = h->d_nodeptr.p;
  q.Moff = h->d_Moff.p;
  q.M = h->d_M.p;
  q.dense = nullptr;
  if (dense) {
    // a one-node chain of its own: tables {0, b} / {0} on the device, inverse to dense->out
    q.plo = 0;
    q.P = 1;
    q.mid = 0;
    q.nodeptr = dense->nodeptr;
    q.Moff = dense->moff;
    q.M = dense->out;
    q.dense = dense->S;
    two_sided = false;
  }
  q.rowptr = h->d_rowptr.p;
  q.dstart = h->d_dstart.p;
  q.ustart = h->d_ustart.p;
  q.col = h->d_col.p;
  q.T = h->d_Tval.p;
  q.ucptr = h->d_ucptr.p;
  q.urow = h->d_urow.p;
  q.upos = h->d_upos.p;
  q.lcptr = h->d_lcptr.p;
  q.lrow = h->d_lrow.p;
  q.lpos = h->d_lpos.p;
  const int ngroups = two_sided ? 2 : 1;
  KB_CUDA(h, h->d_kfG.alloc((size_t)ngroups * bmax * bmax));
  const int64_t maxsteps = (bmax + KF_WMIN - 1) / KF_WMIN + 1;
  KB_CUDA(h, h->d_kfpiv.alloc((size_t)ngroups * maxsteps * 16));
  KB_CUDA(h, h->d_kfsync.alloc(KF_SYNC_WORDS));
  if (!dense) {
    KB_CUDA(h, cudaMemsetAsync(h->d_kfsync.p, 0, KF_SYNC_WORDS * sizeof(unsigned), s));
  } else {
    // keep the time-out flag of the factorisation this launch belongs to
    KB_CUDA(h, cudaMemsetAsync(h->d_kfsync.p, 0, KF_ERR_WORD * sizeof(unsigned), s));
    KB_CUDA(h, cudaMemsetAsync(h->d_kfsync.p + KF_ERR_WORD + 1, 0, (KF_SYNC_WORDS - KF_ERR_WORD - 1) * sizeof(unsigned), s));
  }
  KB_CUDA(h, h->d_info.alloc(1));
  if (!dense) KB_CUDA(h, cudaMemsetAsync(h->d_info.p, 0, sizeof(int), s));  // dense nodes add to the flags
  q.Gbuf[0] = h->d_kfG.p;
  q.Gbuf[1] = two_sided ? h->d_kfG.p + (size_t)bmax * bmax : h->d_kfG.p;
  q.pivbuf[0] = h->d_kfpiv.p;
  q.pivbuf[1] = two_sided ? h->d_kfpiv.p + maxsteps * 16 : h->d_kfpiv.p;
  q.sync = h->d_kfsync.p;
  q.info = h->d_info.p;
  q.err = (int*)(h->d_kfsync.p + KF_ERR_WORD);  // time-out flag
  if (dense) q.err = h->d_sweep_err.p;         // (the sync block is cleared per launch; this one is not)
  q.wait_ns = h->wait_ns;
  q.dbg = nullptr;
  if (getenv("KB_SWEEP_TIMING")) {
    KB_CUDA(h, h->d_sweep_timing.alloc(256 * 16));
    KB_CUDA(h, cudaMemsetAsync(h->d_sweep_timing.p, 0, 256 * 16 * sizeof(long long), s));
    q.dbg = h->d_sweep_timing.p;
  }
  q.bmax = (int)bmax;
  q.G0 = two_sided ? (G + 1) / 2 : G;
  q.transposed = transposed ? 1 : 0;
  q.Sbuf[0] = q.Sbuf[1] = nullptr;
  if (!getenv("KB_CHAINFAC_NOSTREAM")) {
    const size_t per = (size_t)KF_RING * KF_NB_WIDE * bmax * 4;  // doubles per group
    KB_CUDA(h, h->d_kfstream.alloc((size_t)ngroups * per));
    KB_CUDA(h, cudaMemsetAsync(h->d_kfstream.p, 0, (size_t)ngroups * per * sizeof(double), s));
    q.Sbuf[0] = h->d_kfstream.p;
    q.Sbuf[1] = two_sided ? h->d_kfstream.p + per : h->d_kfstream.p;
  }
  q.stagger_ns = getenv("KB_CHAINFAC_STAGGER") ? atoi(getenv("KB_CHAINFAC_STAGGER")) : 1000;
  int T = (int)((bmax + 31) / 32) * 32;
  if (T < 64) T = 64;
  // <9, 640> whenever the node and its strips fit, <10, 704> for the wider ones
  const int gcw = (two_sided ? q.G0 : G);
  const bool wide = bmax > KF_MAXT || (bmax + gcw - 1) / gcw > KF_NB;
  const int nbsel = wide ? KF_NB_WIDE : KF_NB;
  const size_t smem = (wide ? sizeof(KfShared<KF_NB_WIDE>) : sizeof(KfShared<KF_NB>)) +
                      (((size_t)bmax * sizeof(int) + 15) & ~(size_t)15) +
                      (size_t)bmax * (nbsel | 1) * sizeof(double2);
  const void* fn = wide ? (const void*)kb_chain_factor<KF_NB_WIDE, KF_MAXT_WIDE>
                        : (const void*)kb_chain_factor<KF_NB, KF_MAXT>;
  KB_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&q};
  KB_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(T), args, smem, s));
  h->launches++;
  return KB_OK;
}

// debug: cycle counters of the last persistent factorisation (needs KB_SWEEP_TIMING=1)
extern "C" int kb_dbg_factor_timing(kb_handle h, long long* out, int max_ctas) {
  if (!h || !h->d_sweep_timing.p || h->d_sweep_timing.count < 256 * 16) return KB_EINVAL;
  cudaStreamSynchronize(h->stream);
  int g = max_ctas < 256 ? max_ctas : 256;
  cudaMemcpy(out, h->d_sweep_timing.p, (size_t)g * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  return g;
}
