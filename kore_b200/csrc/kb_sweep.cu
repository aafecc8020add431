// Persistent chain-sweep kernel: the whole forward + backward block substitution
// of one solve in ONE cooperative launch, one CTA per SM.
//
// Replaces the per-node launches of kb_solve.cu (2(2P-1) kernels per solve) on the
// path that dominates the Krylov phase: the MUMPS solve phase inside every ST
// application of E.solve() (/root/reference/bin/solve.py:123).
//
// The chain is a sequence of 2P-1 dependent node steps; each step is
//   A) t = r_p - L_{p,p-1} y_{p-1}   (fwd)   |   t = U_{p,p+1} x_{p+1}   (bwd)
//   B) y_p = M_p t                   (fwd)   |   x_p = y_p - M_p t       (bwd)
// Every CTA owns a fixed slice of rows of every node.  A needs the whole previous
// result and B the whole t, so each step has two grid-wide synchronisations; they
// are counter barriers (one release-add per CTA, one polling thread per CTA).
// The HBM stream of the explicit inverses M_p is decoupled from that latency
// chain by bulk L2 prefetches (cp.async.bulk.prefetch.L2) issued two steps ahead.
// The couplings are read from an ELL copy (fixed width per pencil, built at
// factor time) so that a CTA's slice is one contiguous range.
#include <stdlib.h>

#include "kb_internal.cuh"

#define KB_SWEEP_THREADS 256
#define KB_FLAG_STRIDE 8  // 32-byte slots

struct KbSweepParams {
  const double2* M;
  const int64_t* Moff;
  const int64_t* nodeptr;
  int P;
  const double2* r;
  double2* y;
  double2* t;
  const double2* Lval;
  const int* Lcol;
  int WL;
  const double2* Uval;
  const int* Ucol;
  int WU;
  unsigned* flags;
  int* err;
  long long* timing;  // optional: 5 accumulated cycle counters per CTA (debug)
  int bmax;
};

__device__ __forceinline__ void kb_st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned kb_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void kb_prefetch_l2(const void* p, size_t bytes) {
  // p 16-byte aligned, bytes a multiple of 16
  if (bytes == 0) return;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((unsigned)bytes) : "memory");
}

// Grid-wide barrier: one release-add per CTA on a single counter, one polling thread
// per CTA (a flag array polled by every CTA was measured 10x slower: ~22 000 polling
// loads per round contend on a handful of L2 slices).  bar.sync + red.release.gpu is
// cumulative over the CTA's earlier writes, ld.acquire.gpu + bar.sync publishes the
// other CTAs' writes to the whole CTA; no extra fences.  `epoch` counts from 1.
__device__ __forceinline__ void kb_grid_sync(unsigned* ctr, unsigned epoch, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned target = epoch * gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    KbSpin sp;
    // fail fast: once any CTA has timed out, nobody waits any more
    while (kb_ld_acquire(ctr) < target)
      if (kb_spin_expired(sp, err, KB_WERR_SWEEP, KB_WAIT_NS_DEFAULT)) break;
  }
  __syncthreads();
}

// rows of a b-row node owned by this CTA
__device__ __forceinline__ void kb_my_rows(int b, int& row0, int& row1) {
  int rpc = (b + gridDim.x - 1) / gridDim.x;
  row0 = min(b, (int)blockIdx.x * rpc);
  row1 = min(b, row0 + rpc);
}

__device__ __forceinline__ void kb_step_node(const KbSweepParams& q, int s, bool& fwd, int& p) {
  fwd = s < q.P;
  p = fwd ? s : 2 * q.P - 2 - s;
}

// STAGED: this CTA's slice of M_p is brought into shared memory by a bulk (TMA)
// copy issued one step ahead (double buffered); otherwise it is read straight from
// global/L2 (slices too large for two stages).
template <bool STAGED>
__global__ void __launch_bounds__(KB_SWEEP_THREADS, 1) kb_sweep_persistent(KbSweepParams q, int slice_elems) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* tvec = (double2*)smem_raw;                       // q.bmax entries
  double2* stage0 = tvec + ((q.bmax + 7) & ~7);             // STAGED: 2 x slice_elems
  // node tables (offset of every node, offset of every M_p) cached in shared memory: they sit
  // on the critical path of every step otherwise
  int64_t* s_moff = (int64_t*)(stage0 + (STAGED ? 2 * (size_t)slice_elems : 0));
  int* s_nptr = (int*)(s_moff + (q.P + 1));
  __shared__ double2 part[8][8];
  __shared__ __align__(8) uint64_t mbar[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int P = q.P;
  const int S = 2 * P - 1;
  unsigned epoch = 0;
  unsigned uses[2] = {0u, 0u};  // completed uses of each stage (mbarrier phase parity), uniform per CTA
  long long tacc[5] = {0, 0, 0, 0, 0};
  long long tc0 = clock64();
#define KB_TICK(slot)         \
  {                           \
    long long _c = clock64(); \
    tacc[slot] += _c - tc0;   \
    tc0 = _c;                 \
  }

  for (int i = tid; i <= q.P; i += KB_SWEEP_THREADS) {
    s_moff[i] = q.Moff[i];
    s_nptr[i] = (int)q.nodeptr[i];
  }
  __syncthreads();
  if (STAGED) {
    if (tid == 0) {
      kb_mbar_init(&mbar[0], 1);
      kb_mbar_init(&mbar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {  // stage the first step's slice
      bool f0;
      int p0;
      kb_step_node(q, 0, f0, p0);
      const int b0 = (s_nptr[p0 + 1] - s_nptr[p0]);
      int a0, a1;
      kb_my_rows(b0, a0, a1);
      unsigned bytes = (unsigned)((size_t)(a1 - a0) * b0 * sizeof(double2));
      if (bytes) {
        kb_mbar_expect_tx(&mbar[0], bytes);
        kb_bulk_g2s(stage0, q.M + s_moff[p0] + (size_t)a0 * b0, bytes, &mbar[0]);
      }
    }
  }

  for (int s = 0; s < S; ++s) {
    bool fwd;
    int p;
    kb_step_node(q, s, fwd, p);
    const int o = s_nptr[p];
    const int b = (s_nptr[p + 1] - s_nptr[p]);
    int row0, row1;
    kb_my_rows(b, row0, row1);

    // ---- next step's slice: bulk copy into the other stage (its last readers finished
    //      before the previous grid barrier), or an L2 prefetch two steps ahead
    if (STAGED) {
      if (tid == 0 && s + 1 < S) {
        bool f1;
        int p1;
        kb_step_node(q, s + 1, f1, p1);
        const int b1 = (s_nptr[p1 + 1] - s_nptr[p1]);
        int a0, a1;
        kb_my_rows(b1, a0, a1);
        unsigned bytes = (unsigned)((size_t)(a1 - a0) * b1 * sizeof(double2));
        if (bytes) {
          uint64_t* mb = &mbar[(s + 1) & 1];
          kb_mbar_expect_tx(mb, bytes);
          kb_bulk_g2s(stage0 + (size_t)((s + 1) & 1) * slice_elems, q.M + s_moff[p1] + (size_t)a0 * b1, bytes, mb);
        }
      }
    }
    if (tid == 32 && s + 2 < S) {
      bool f2;
      int p2;
      kb_step_node(q, s + 2, f2, p2);
      const int o2 = s_nptr[p2];
      const int b2 = (s_nptr[p2 + 1] - s_nptr[p2]);
      int a0, a1;
      kb_my_rows(b2, a0, a1);
      if (a1 > a0) {
        const double2* m = q.M + s_moff[p2] + (size_t)a0 * b2;
        size_t bytes = (size_t)(a1 - a0) * b2 * sizeof(double2);
        while (bytes > 0) {
          size_t c = bytes > 65536 ? 65536 : bytes;
          kb_prefetch_l2(m, c);
          m = (const double2*)((const char*)m + c);
          bytes -= c;
        }
        const int W = f2 ? q.WL : q.WU;
        const double2* v = (f2 ? q.Lval : q.Uval) + (size_t)(o2 + a0) * W;
        kb_prefetch_l2(v, (size_t)(a1 - a0) * W * sizeof(double2));
        const int* c = (f2 ? q.Lcol : q.Ucol) + (size_t)(o2 + a0) * W;
        uintptr_t c0 = (uintptr_t)c & ~(uintptr_t)15;
        uintptr_t c1 = ((uintptr_t)(c + (size_t)(a1 - a0) * W) + 15) & ~(uintptr_t)15;
        kb_prefetch_l2((const void*)c0, (size_t)(c1 - c0));
      }
    }
    KB_TICK(0);

    // ---- phase A: sparse coupling for the rows this CTA owns (one warp per row)
    {
      const int W = fwd ? q.WL : q.WU;
      const double2* val = fwd ? q.Lval : q.Uval;
      const int* col = fwd ? q.Lcol : q.Ucol;
      for (int i = row0 + wid; i < row1; i += KB_SWEEP_THREADS / 32) {
        const int gi = o + i;
        double2 acc = zmake(0.0, 0.0);
        for (int k = lane; k < W; k += 32) {
          double2 v = val[(size_t)gi * W + k];
          int c = col[(size_t)gi * W + k];
          zfma(acc, v, __ldcg(&q.y[c]));
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
        }
        if (lane == 0) q.t[gi] = fwd ? zsub(q.r[gi], acc) : acc;
      }
    }
    KB_TICK(1);
    kb_grid_sync(q.flags, ++epoch, q.err);
    KB_TICK(2);

    // ---- phase B: dense rows of M_p against the full t
    for (int j = tid; j < b; j += KB_SWEEP_THREADS) tvec[j] = __ldcg(&q.t[o + j]);
    if (STAGED && row1 > row0) {
      kb_mbar_wait(&mbar[s & 1], uses[s & 1] & 1u, q.err, KB_WAIT_NS_DEFAULT);
      uses[s & 1]++;
    }
    __syncthreads();
    const double2* Mp = STAGED ? (stage0 + (size_t)(s & 1) * slice_elems) : (q.M + s_moff[p] + (size_t)row0 * b);
    for (int rb = 0; rb < row1 - row0; rb += 8) {
      const int nr = min(8, row1 - row0 - rb);
      double2 acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = zmake(0.0, 0.0);
      for (int j = tid; j < b; j += KB_SWEEP_THREADS) {
        const double2 tj = tvec[j];
        double2 m[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (STAGED)
            m[u] = u < nr ? Mp[(size_t)(rb + u) * b + j] : zmake(0.0, 0.0);
          else
            m[u] = u < nr ? __ldcs(&Mp[(size_t)(rb + u) * b + j]) : zmake(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) zfma(acc[u], m[u], tj);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
          acc[u].x += __shfl_xor_sync(0xffffffffu, acc[u].x, sft);
          acc[u].y += __shfl_xor_sync(0xffffffffu, acc[u].y, sft);
        }
        if (lane == 0) part[u][wid] = acc[u];
      }
      __syncthreads();
      if (tid < nr) {
        double2 v = zmake(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < KB_SWEEP_THREADS / 32; ++w) v = zadd(v, part[tid][w]);
        const int gi = o + row0 + rb + tid;
        q.y[gi] = fwd ? v : zsub(__ldcg(&q.y[gi]), v);
      }
      __syncthreads();
    }
    KB_TICK(3);
    kb_grid_sync(q.flags, ++epoch, q.err);
    KB_TICK(4);
  }
  if (q.timing && tid == 0)
    for (int k = 0; k < 5; ++k) q.timing[blockIdx.x * 5 + k] = tacc[k];
}

// ---------------------------------------------------------------------------
// Dataflow variant: no grid barriers at all.  Every vector entry that crosses CTAs
// (y of the forward sweep, x of the backward sweep, the two t vectors) is written
// exactly once per solve into a buffer pre-filled with a NaN sentinel; consumers
// poll the entry itself (ld.volatile, 16 bytes) until it is no longer the sentinel.
// The datum is its own flag: one L2 round trip per dependency instead of
// release-add + poll + reload, and CTAs that own no rows of a node never wait.
// ---------------------------------------------------------------------------
#define KB_SENTINEL 0x7ff8dead0badbeefLL

__device__ __forceinline__ double2 kb_poll(const double2* p, int* err) {
  double2 v;
  KbSpin sp;
  for (;;) {
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    if (__double_as_longlong(v.x) != KB_SENTINEL && __double_as_longlong(v.y) != KB_SENTINEL) break;
    if (kb_spin_expired(sp, err, KB_WERR_SWEEP, KB_WAIT_NS_DEFAULT)) break;
  }
  return v;
}

__global__ void kb_fill_sentinel(int n, double2* a, double2* b, double2* c, double2* d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sv = __longlong_as_double(KB_SENTINEL);
  const double2 v = make_double2(sv, sv);
  a[i] = v;
  b[i] = v;
  c[i] = v;
  d[i] = v;
}

struct KbFlowParams {
  const double2* M;
  const int64_t* Moff;
  const int64_t* nodeptr;
  int P;
  int mid;       // middle node of the two-sided elimination (P-1: one-sided)
  const double2* r;
  double2* yf;   // forward results (n + zero slot), sentinel-filled
  double2* x;    // final result    (n + zero slot), sentinel-filled
  double2* t1;   // forward t, sentinel-filled
  double2* t2;   // backward t, sentinel-filled
  const double2* Lval;
  const int* Lcol;
  int WL;
  const double2* Uval;
  const int* Ucol;
  int WU;
  int* err;
  long long* timing;
  int bmax;
  int pollw;
};

// Step modes of the two-sided sweep.  Group 0 (top chain) runs FWD_L for nodes
// 0..mid-1, MID for the middle node, then BWD_U for mid-1..0; group 1 (bottom chain)
// runs FWD_U for P-1..mid+1, then BWD_L for mid+1..P-1.  With mid = P-1 group 1 is
// empty and the kernel is the plain one-sided sweep.
enum { KB_FWD_L = 0, KB_FWD_U = 1, KB_MID = 2, KB_BWD_U = 3, KB_BWD_L = 4 };

__device__ __forceinline__ void kb_flow_step(int group, int s, int P, int mid, int& p, int& mode) {
  if (group == 0) {
    if (s < mid) {
      p = s;
      mode = KB_FWD_L;
    } else if (s == mid) {
      p = mid;
      mode = KB_MID;
    } else {
      p = 2 * mid - s;
      mode = KB_BWD_U;
    }
  } else {
    const int nb = P - 1 - mid;
    if (s < nb) {
      p = P - 1 - s;
      mode = KB_FWD_U;
    } else {
      p = mid + 1 + (s - nb);
      mode = KB_BWD_L;
    }
  }
}

// rows of a b-row node owned by CTA `rank` of a group of `size` CTAs (balanced split)
__device__ __forceinline__ void kb_group_rows(int b, int size, int rank, int& row0, int& row1) {
  const int base = b / size, rem = b - base * size;
  row0 = rank * base + min(rank, rem);
  row1 = row0 + base + (rank < rem ? 1 : 0);
}

template <bool STAGED>
__global__ void __launch_bounds__(KB_SWEEP_THREADS, 1) kb_sweep_dataflow(KbFlowParams q, int slice_elems, int G0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* tvec = (double2*)smem_raw;
  double2* stage0 = tvec + ((q.bmax + 7) & ~7);
  int64_t* s_moff = (int64_t*)(stage0 + (STAGED ? 2 * (size_t)slice_elems : 0));
  int* s_nptr = (int*)(s_moff + (q.P + 1));
  __shared__ double2 part[16][8];
  __shared__ __align__(8) uint64_t mbar[2];
  constexpr int NW = KB_SWEEP_THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int P = q.P, mid = q.mid;
  const int group = ((int)blockIdx.x < G0) ? 0 : 1;
  const int gsize = group == 0 ? G0 : (int)gridDim.x - G0;
  const int grank = group == 0 ? (int)blockIdx.x : (int)blockIdx.x - G0;
  const int S = group == 0 ? 2 * mid + 1 : 2 * (P - 1 - mid);
  unsigned uses[2] = {0u, 0u};
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tc0 = clock64();

  for (int i = tid; i <= q.P; i += KB_SWEEP_THREADS) {
    s_moff[i] = q.Moff[i];
    s_nptr[i] = (int)q.nodeptr[i];
  }
  __syncthreads();
  if (STAGED) {
    if (tid == 0) {
      kb_mbar_init(&mbar[0], 1);
      kb_mbar_init(&mbar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && S > 0) {
      int p0, m0;
      kb_flow_step(group, 0, P, mid, p0, m0);
      const int b0 = s_nptr[p0 + 1] - s_nptr[p0];
      int a0, a1;
      kb_group_rows(b0, gsize, grank, a0, a1);
      unsigned bytes = (unsigned)((size_t)(a1 - a0) * b0 * sizeof(double2));
      if (bytes) {
        kb_mbar_expect_tx(&mbar[0], bytes);
        kb_bulk_g2s(stage0, q.M + s_moff[p0] + (size_t)a0 * b0, bytes, &mbar[0]);
      }
    }
  }

  // coupling entry of this lane for the first row this warp handles in the NEXT step
  // (software-pipelined: it does not depend on the chain, only its operand does)
  // Half-warp mode (coupling width <= 16, e.g. 15 for hydro): lanes 0-15 serve row
  // row0 + 2*wid, lanes 16-31 row row0 + 2*wid + 1, so one pass covers 16 rows per CTA.
  const bool halfw = q.WL <= 16 && q.WU <= 16;
  const int hl = halfw ? (lane & 15) : lane;          // lane within the (half-)warp
  const int hsel = halfw ? (lane >> 4) : 0;           // which of the warp's two rows
  const int rows_per_pass = halfw ? 2 * NW : NW;
  double2 pv = zmake(0.0, 0.0);
  int pc = 0;
  auto preload = [&](int sn) {
    pv = zmake(0.0, 0.0);
    pc = s_nptr[P];  // zero slot
    if (sn >= S) return;
    int pn, mn;
    kb_flow_step(group, sn, P, mid, pn, mn);
    if (mn == KB_MID) return;
    const bool useL = (mn == KB_FWD_L || mn == KB_BWD_L);
    const int W = useL ? q.WL : q.WU;
    int r0, r1;
    kb_group_rows(s_nptr[pn + 1] - s_nptr[pn], gsize, grank, r0, r1);
    const int i = r0 + (halfw ? 2 * wid + hsel : wid);
    if (i < r1 && hl < W) {
      const size_t e = (size_t)(s_nptr[pn] + i) * W + hl;
      pv = (useL ? q.Lval : q.Uval)[e];
      pc = (useL ? q.Lcol : q.Ucol)[e];
    }
  };
  preload(0);

  for (int s = 0; s < S; ++s) {
    int p, mode;
    kb_flow_step(group, s, P, mid, p, mode);
    const bool fwd = mode <= KB_MID;
    const int o = s_nptr[p];
    const int b = s_nptr[p + 1] - s_nptr[p];
    int row0, row1;
    kb_group_rows(b, gsize, grank, row0, row1);
    KB_TICK(0);

    // own rows of the forward result, needed by the backward update at the end of the step:
    // fetched now, off the critical path (this CTA wrote them during its forward sweep)
    double2 yown = zmake(0.0, 0.0);
    if (!fwd && tid < row1 - row0 && tid < 2 * NW) yown = kb_poll(&q.yf[o + row0 + tid], q.err);

    // ---- phase A: t rows owned by this CTA; the inputs are polled entry by entry
    {
      const double2* src = fwd ? q.yf : q.x;
      double2* tdst = fwd ? q.t1 : q.t2;
      const bool useL = (mode == KB_FWD_L || mode == KB_BWD_L);
      const int W = useL ? q.WL : q.WU;
      const double2* val = useL ? q.Lval : q.Uval;
      const int* col = useL ? q.Lcol : q.Ucol;
      for (int ib = row0; ib < row1; ib += rows_per_pass) {
        const int i = ib + (halfw ? 2 * wid + hsel : wid);
        const bool active = i < row1;
        const int gi = o + (active ? i : row0);
        double2 acc = zmake(0.0, 0.0);
        if (active) {
          if (mode == KB_MID) {
            for (int k = hl; k < q.WL; k += (halfw ? 16 : 32))
              zfma(acc, q.Lval[(size_t)gi * q.WL + k], kb_poll(&src[q.Lcol[(size_t)gi * q.WL + k]], q.err));
            for (int k = hl; k < q.WU; k += (halfw ? 16 : 32))
              zfma(acc, q.Uval[(size_t)gi * q.WU + k], kb_poll(&src[q.Ucol[(size_t)gi * q.WU + k]], q.err));
          } else if (ib == row0) {
            if (hl < W) zfma(acc, pv, kb_poll(&src[pc], q.err));
            if (!halfw)
              for (int k = lane + 32; k < W; k += 32)
                zfma(acc, val[(size_t)gi * W + k], kb_poll(&src[col[(size_t)gi * W + k]], q.err));
          } else {
            for (int k = hl; k < W; k += (halfw ? 16 : 32))
              zfma(acc, val[(size_t)gi * W + k], kb_poll(&src[col[(size_t)gi * W + k]], q.err));
          }
        }
        // reduce within the half-warp (xor 8..1) or the warp (xor 16..1)
        if (!halfw) {
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
        }
#pragma unroll
        for (int sft = 8; sft > 0; sft >>= 1) {
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
        }
        if (active && hl == 0) tdst[gi] = fwd ? zsub(q.r[gi], acc) : acc;
      }
    }
    KB_TICK(1);

    // ---- off the critical path (the t poll below waits on the other CTAs anyway):
    //      next step's slice of M into the other stage, slices two steps ahead into L2,
    //      next step's coupling entries into registers
    if (STAGED && tid == 0 && s + 1 < S) {
      int p1, m1;
      kb_flow_step(group, s + 1, P, mid, p1, m1);
      const int b1 = s_nptr[p1 + 1] - s_nptr[p1];
      int a0, a1;
      kb_group_rows(b1, gsize, grank, a0, a1);
      unsigned bytes = (unsigned)((size_t)(a1 - a0) * b1 * sizeof(double2));
      if (bytes) {
        uint64_t* mb = &mbar[(s + 1) & 1];
        kb_mbar_expect_tx(mb, bytes);
        kb_bulk_g2s(stage0 + (size_t)((s + 1) & 1) * slice_elems, q.M + s_moff[p1] + (size_t)a0 * b1, bytes, mb);
      }
    }
    if (tid == 32 && s + 2 < S) {
      int p2, m2;
      kb_flow_step(group, s + 2, P, mid, p2, m2);
      const int b2 = s_nptr[p2 + 1] - s_nptr[p2];
      int a0, a1;
      kb_group_rows(b2, gsize, grank, a0, a1);
      if (a1 > a0) {
        const double2* m = q.M + s_moff[p2] + (size_t)a0 * b2;
        size_t bytes = (size_t)(a1 - a0) * b2 * sizeof(double2);
        while (bytes > 0) {
          size_t c = bytes > 65536 ? 65536 : bytes;
          kb_prefetch_l2(m, c);
          m = (const double2*)((const char*)m + c);
          bytes -= c;
        }
      }
    }
    preload(s + 1);

    // ---- phase B: dense rows of M_p against the full t (polled)
    if (row1 > row0) {
      const double2* tsrc = fwd ? q.t1 : q.t2;
      // all-gather of t by `pollw` warps; each lane keeps up to 8 entries in flight and
      // re-polls only the missing ones
      {
        const int pollw = q.pollw;
        if (wid < pollw) {
          for (int j0 = (wid * 32 + lane); j0 < b; j0 += pollw * 32 * 8) {
            double2 v[8];
            unsigned pending = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (j0 + u * pollw * 32 < b) pending |= 1u << u;
            KbSpin sp;
            while (pending) {
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (pending & (1u << u)) {
                  const double2* ptr = &tsrc[o + j0 + u * pollw * 32];
                  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];"
                               : "=d"(v[u].x), "=d"(v[u].y)
                               : "l"(ptr)
                               : "memory");
                }
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if ((pending & (1u << u)) && __double_as_longlong(v[u].x) != KB_SENTINEL &&
                    __double_as_longlong(v[u].y) != KB_SENTINEL) {
                  tvec[j0 + u * pollw * 32] = v[u];
                  pending &= ~(1u << u);
                }
              if (pending && kb_spin_expired(sp, q.err, KB_WERR_SWEEP, KB_WAIT_NS_DEFAULT)) break;
            }
          }
        }
      }
      if (STAGED) {
        kb_mbar_wait(&mbar[s & 1], uses[s & 1] & 1u, q.err, KB_WAIT_NS_DEFAULT);
        uses[s & 1]++;
      }
      __syncthreads();
      KB_TICK(2);
      const double2* Mp = STAGED ? (stage0 + (size_t)(s & 1) * slice_elems) : (q.M + s_moff[p] + (size_t)row0 * b);
      // a warp takes whole rows (row = wid, wid + 8, ...; up to 16 rows per pass), or a
      // 1/nsplit part of a row when the CTA owns fewer rows than warps: one shuffle
      // reduction per row per warp, ONE block barrier per pass
      for (int rb = 0; rb < row1 - row0; rb += 2 * NW) {
        const int nr = min(2 * NW, row1 - row0 - rb);
        const int nsplit = nr >= NW ? 1 : NW / nr;  // >= 1
        for (int rr = wid / nsplit; rr < nr; rr += NW / nsplit) {
          const int mypart = wid % nsplit;
          const int seg = (b + nsplit - 1) / nsplit;
          const int j0 = mypart * seg, j1 = min(b, j0 + seg);
          const double2* Mrow = Mp + (size_t)(rb + rr) * b;
          double2 acc = zmake(0.0, 0.0), acc1 = zmake(0.0, 0.0);
          int j = j0 + lane;
          for (; j + 32 < j1; j += 64) {
            double2 m0 = STAGED ? Mrow[j] : __ldcs(&Mrow[j]);
            double2 m1 = STAGED ? Mrow[j + 32] : __ldcs(&Mrow[j + 32]);
            zfma(acc, m0, tvec[j]);
            zfma(acc1, m1, tvec[j + 32]);
          }
          if (j < j1) zfma(acc, STAGED ? Mrow[j] : __ldcs(&Mrow[j]), tvec[j]);
          acc = zadd(acc, acc1);
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
          }
          if (lane == 0) part[rr][mypart] = acc;
        }
        __syncthreads();
        KB_TICK(3);
        if (tid < nr) {
          double2 v = part[tid][0];
          for (int w = 1; w < nsplit; ++w) v = zadd(v, part[tid][w]);
          const int gi = o + row0 + rb + tid;
          if (mode == KB_FWD_L || mode == KB_FWD_U)
            q.yf[gi] = v;
          else if (mode == KB_MID)
            q.x[gi] = v;
          else
            q.x[gi] = zsub(rb == 0 ? yown : kb_poll(&q.yf[gi], q.err), v);
        }
        __syncthreads();
      }
    }
    KB_TICK(4);
  }
  if (q.timing && tid == 0)
    for (int k = 0; k < 8; ++k) q.timing[blockIdx.x * 8 + k] = tacc[k];
}

// ELL copies of the couplings (built at factor time from the equilibrated T):
// row gi, slot k:  L part = entries [rowptr, dstart), U part = [ustart, rowptr+1).
// Padding: value 0, column n (the vector buffers carry a zero sentinel at [n]).
__global__ void kb_pack_ell(int n, int WL, int WU, const int64_t* __restrict__ rowptr,
                            const int64_t* __restrict__ dstart, const int64_t* __restrict__ ustart,
                            const int* __restrict__ col, const double2* __restrict__ T,
                            double2* __restrict__ Lval, int* __restrict__ Lcol, double2* __restrict__ Uval,
                            int* __restrict__ Ucol) {
  int gi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (gi >= n) return;
  int64_t l0 = rowptr[gi], l1 = dstart[gi], u0 = ustart[gi], u1 = rowptr[gi + 1];
  for (int k = lane; k < WL; k += 32) {
    bool ok = l0 + k < l1;
    Lval[(size_t)gi * WL + k] = ok ? T[l0 + k] : zmake(0.0, 0.0);
    Lcol[(size_t)gi * WL + k] = ok ? col[l0 + k] : n;
  }
  for (int k = lane; k < WU; k += 32) {
    bool ok = u0 + k < u1;
    Uval[(size_t)gi * WU + k] = ok ? T[u0 + k] : zmake(0.0, 0.0);
    Ucol[(size_t)gi * WU + k] = ok ? col[u0 + k] : n;
  }
}

int kbi_sweep_prepare(kb_context* h) {
  // called at the end of kb_factor: ELL copies + device copies of the node tables
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  const int WL = h->WL > 0 ? h->WL : 1, WU = h->WU > 0 ? h->WU : 1;
  KB_CUDA(h, h->d_Lval.alloc((size_t)n * WL));
  KB_CUDA(h, h->d_Lcol.alloc((size_t)n * WL + 4));
  KB_CUDA(h, h->d_Uval.alloc((size_t)n * WU));
  KB_CUDA(h, h->d_Ucol.alloc((size_t)n * WU + 4));
  kb_pack_ell<<<nblk((int64_t)n * 32, 256), 256, 0, s>>>(n, WL, WU, h->d_rowptr.p, h->d_dstart.p, h->d_ustart.p,
                                                         h->d_col.p, h->d_Tval.p, h->d_Lval.p, h->d_Lcol.p,
                                                         h->d_Uval.p, h->d_Ucol.p);
  h->launches++;
  KB_LAUNCH_CHECK(h);
  KB_CUDA(h, h->d_nodeptr.alloc(h->P + 1));
  KB_CUDA(h, h->d_Moff.alloc(h->P + 1));
  KB_CUDA(h, cudaMemcpyAsync(h->d_nodeptr.p, h->nodeptr.data(), (h->P + 1) * sizeof(int64_t),
                             cudaMemcpyHostToDevice, s));
  KB_CUDA(h, cudaMemcpyAsync(h->d_Moff.p, h->Moff.data(), (h->P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  KB_CUDA(h, h->d_flags.alloc(256 * KB_FLAG_STRIDE));
  KB_CUDA(h, h->d_sweep_err.alloc(1));
  if (getenv("KB_SWEEP_TIMING")) KB_CUDA(h, h->d_sweep_timing.alloc(256 * 8));
  KB_CUDA(h, cudaMemsetAsync(h->d_sweep_err.p, 0, sizeof(int), s));
  if (h->sweep_grid == 0) {
    int dev = h->device, sms = 0, coop = 0;
    KB_CUDA(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    KB_CUDA(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    h->sweep_grid = coop ? (sms < KB_SWEEP_THREADS ? sms : KB_SWEEP_THREADS) : -1;
    if (h->sweep_grid > 0 && getenv("KB_SWEEP_GRID")) {
      int gsz = atoi(getenv("KB_SWEEP_GRID"));
      if (gsz >= 1 && gsz < h->sweep_grid) h->sweep_grid = gsz;
    }
  }
  return KB_OK;
}

// y <- T'^{-1} r in one cooperative launch.  y must have n+1 entries with y[n] == 0.
int kbi_sweep_persistent(kb_context* h, const double2* r, double2* y) {
  cudaStream_t s = h->stream;
  KbSweepParams q;
  q.M = h->d_M.p;
  q.Moff = h->d_Moff.p;
  q.nodeptr = h->d_nodeptr.p;
  q.P = (int)h->P;
  q.r = r;
  q.y = y;
  q.t = h->d_t.p;
  q.Lval = h->d_Lval.p;
  q.Lcol = h->d_Lcol.p;
  q.WL = h->WL > 0 ? h->WL : 1;
  q.Uval = h->d_Uval.p;
  q.Ucol = h->d_Ucol.p;
  q.WU = h->WU > 0 ? h->WU : 1;
  q.flags = h->d_flags.p;
  q.err = h->d_sweep_err.p;
  q.timing = h->d_sweep_timing.p;
  q.bmax = (int)h->bmax;
  const int G = h->sweep_grid;
  const int rpc = (int)((h->bmax + G - 1) / G);
  int slice_elems = (int)(((int64_t)rpc * h->bmax + 7) & ~(int64_t)7);
  const size_t tv = (size_t)((h->bmax + 7) & ~(int64_t)7) * sizeof(double2);
  const size_t tabs = (size_t)(h->P + 1) * (sizeof(int64_t) + sizeof(int)) + 16;
  size_t smem_staged = tv + 2 * (size_t)slice_elems * sizeof(double2) + tabs;
  const bool staged = smem_staged <= 200 * 1024;
  size_t smem = staged ? smem_staged : tv + tabs;
  if (smem > 220 * 1024) return kb_fail(h, KB_EINVAL, "chain too long for the persistent sweep tables");
  const void* fn = staged ? (const void*)kb_sweep_persistent<true> : (const void*)kb_sweep_persistent<false>;
  if (smem > 48 * 1024) KB_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KB_CUDA(h, cudaMemsetAsync(h->d_flags.p, 0, 256 * KB_FLAG_STRIDE * sizeof(unsigned), s));
  void* args[] = {(void*)&q, (void*)&slice_elems};
  KB_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(KB_SWEEP_THREADS), args, smem, s));
  h->launches++;
  return KB_OK;
}

// debug: cycle counters of the last persistent sweep (needs KB_SWEEP_TIMING=1 at factor time)
extern "C" int kb_dbg_sweep_timing(kb_handle h, long long* out, int max_ctas) {
  if (!h || !h->d_sweep_timing.p) return KB_EINVAL;
  cudaStreamSynchronize(h->stream);
  int g = h->sweep_grid < max_ctas ? h->sweep_grid : max_ctas;
  cudaMemcpy(out, h->d_sweep_timing.p, (size_t)g * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
  return g;
}

// y <- T'^{-1} r with the barrier-free two-sided dataflow kernel.  y has n+1 entries, y[n] == 0.
int kbi_sweep_dataflow(kb_context* h, const double2* r, double2* y) {
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  if (h->d_yf.count < (size_t)n + 1) {
    KB_CUDA(h, h->d_yf.alloc(n + 1));
    KB_CUDA(h, cudaMemsetAsync(h->d_yf.p + n, 0, sizeof(double2), s));
  }
  KB_CUDA(h, h->d_t2.alloc(n));
  kb_fill_sentinel<<<nblk(n, 256), 256, 0, s>>>(n, h->d_yf.p, y, h->d_t.p, h->d_t2.p);
  KbFlowParams q;
  q.M = h->d_M.p;
  q.Moff = h->d_Moff.p;
  q.nodeptr = h->d_nodeptr.p;
  q.P = (int)h->P;
  q.mid = (int)h->mid;
  q.r = r;
  q.yf = h->d_yf.p;
  q.x = y;
  q.t1 = h->d_t.p;
  q.t2 = h->d_t2.p;
  q.Lval = h->d_Lval.p;
  q.Lcol = h->d_Lcol.p;
  q.WL = h->WL > 0 ? h->WL : 1;
  q.Uval = h->d_Uval.p;
  q.Ucol = h->d_Ucol.p;
  q.WU = h->WU > 0 ? h->WU : 1;
  q.err = h->d_sweep_err.p;
  q.timing = h->d_sweep_timing.p;
  q.bmax = (int)h->bmax;
  q.pollw = getenv("KB_SWEEP_POLLW") ? atoi(getenv("KB_SWEEP_POLLW")) : 4;
  if (q.pollw < 1 || q.pollw > KB_SWEEP_THREADS / 32) q.pollw = 4;
  const int G = h->sweep_grid;
  const bool two = h->mid < h->P - 1 && G >= 2;
  if (!two && h->mid != h->P - 1) return kb_fail(h, KB_EINVAL, "two-sided factors need at least 2 CTAs");
  int G0 = two ? (G + 1) / 2 : G;
  const int gmin = two ? G - G0 : G;
  const int rpc = (int)((h->bmax + gmin - 1) / gmin);
  int slice_elems = (int)(((int64_t)rpc * h->bmax + 7) & ~(int64_t)7);
  const size_t tv = (size_t)((h->bmax + 7) & ~(int64_t)7) * sizeof(double2);
  const size_t tabs = (size_t)(h->P + 1) * (sizeof(int64_t) + sizeof(int)) + 16;
  size_t smem_staged = tv + 2 * (size_t)slice_elems * sizeof(double2) + tabs;
  const bool staged = smem_staged <= 200 * 1024 && !getenv("KB_SWEEP_NOSTAGE");
  size_t smem = staged ? smem_staged : tv + tabs;
  if (smem > 220 * 1024) return kb_fail(h, KB_EINVAL, "chain too long for the persistent sweep tables");
  const void* fn = staged ? (const void*)kb_sweep_dataflow<true> : (const void*)kb_sweep_dataflow<false>;
  if (smem > 48 * 1024) KB_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&q, (void*)&slice_elems, (void*)&G0};
  KB_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(KB_SWEEP_THREADS), args, smem, s));
  h->launches += 2;
  return KB_OK;
}
