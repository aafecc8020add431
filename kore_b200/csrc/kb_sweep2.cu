// Folded chain sweep: the two-sided forward + backward block substitution of one solve in ONE
// cooperative launch, split by ROWS, with a single one-way exchange of b values per chain step.
//
// Replaces the MUMPS solve phase inside every ST application of E.solve()
// (/root/reference/bin/solve.py:123) and K.solve (solve.py:227).  Same algebra as kb_sweep1.cu,
// rearranged so that the quantity that travels between the CTAs is the full input vector of the
// next step and nothing else:
//
//     forward    t_0 = r_0,            t_{p'} = r_{p'} - (C_{p',p} M_p) t_p        p' = node after p
//     middle     u_m = r_m - (L M_{m-1}) t_{m-1} - (U M_{m+1}) t_{m+1}
//     backward   u_{p''} = t_{p''} - (C_{p'',p} M_p) u_p                          p'' = node before p
//     solution   x_p = M_p u_p                                                     (no dependency)
//
// The products F = C M_p ("folded" couplings: FL_p = L_{p+1,p} M_p, FU_p = U_{p-1,p} M_p, dense,
// row-major) are formed once per factorisation by kb_fold_couplings -- a banded-times-dense
// product, 2 (2w+1) b^2 complex FMAs per node, bandwidth-bound.  In the sweep CTA c of a chain
// group owns a row slice (<= 10 rows) of every F.  Every element of F is used exactly once per
// solve, so the slice goes from L2 (bulk-prefetched three steps ahead) straight into registers
// -- two rows per warp, lane = column mod 32, issued before the step's exchange is awaited --
// and shared memory carries only the gathered vector.  (Measured alternatives: a TMA-staged
// copy of the slice with one row per warp reads slice and vector through the 128 B/clk
// shared-memory port, 2.2 K cycles per step; all rows per thread and a column split over the
// CTA needs 200 double shuffles per warp, 3.9 K.)  Lane 0 of a warp publishes its rows after one
// shuffle tree; every CTA of the group gathers the b entries.  There is no sparse coupling phase and no partial-sum reduction across CTAs.  The
// solution x_p = M_p u_p has no dependency between nodes: the backward steps leave u in global
// memory and kb_fold_solution forms all P products afterwards at full HBM bandwidth (lanes along
// the rows of M_p^T, fully coalesced).
//
// Exchange protocol: ONE hop.  An entry travels as a 32-byte element (re, tag, im, tag) written
// with one 256-bit store and polled with 256-bit loads by its consumers; tag = solve epoch +
// publication index, strictly increasing over the life of the handle, so a consumer can tell a
// fresh entry from whatever the ring slot held before without any flag, counter, fence or reset.
// Both 16-byte halves carry the tag: an entry is accepted only when both match, so the protocol
// needs 16-byte single-copy atomicity only (what kb_sweep.cu relies on as well).  Against the
// counter protocol of kb_sweep1.cu this removes the release fence behind 9.6 KB of partial sums,
// the counter round trip and the separate read of the data.
//
// Traffic: the chain reads one F per step (2 x 16 sum b^2 bytes per solve, as kb_sweep1.cu reads
// M_p twice) and the solution pass reads M_p once more: 3 x 16 sum b^2 in all, every byte of it
// streamed at HBM rate instead of waiting on an exchange.
#include <stdlib.h>

#include "kb_internal.cuh"

#define K2_CW 8                  // warps: K2_RW row warps + gather warps
#define K2_RW 6
#define K2_THREADS (K2_CW * 32)
#define K2_CPL 20                // columns of a row per lane: nodes up to 32 * K2_CPL = 640 wide
#define K2_CPL_WIDE 22           // ... and up to 704 wide (Kore's own rule at E = 1e-8 gives N = 676): second
                                 // instantiation, a few spilled registers, ONE slice stage (two do not fit)
// entries of the input vector polled per gather thread: 32 * CPL / 64 gather threads = CPL / 2
#define K2_RING 4                // publication ring slots (power of two)
#define K2_XR 10                 // most rows of a node a CTA may own: warps 0-3 take two, 4-5 one
#ifndef K2_AHEAD
#define K2_AHEAD 4               // L2 prefetch distance, in steps
#endif

struct alignas(32) K2Elem {
  double re, tag0, im, tag1;
};

// One step of a group's schedule (built on the host, kbi_fold_prepare).
struct K2Op {
  long long mat_off;  // offset of F in the folded buffer (elements), -1: no product (last step)
  int in_node;        // node whose vector (t or u) is the input; its length is the F column count
  int out_node;       // node whose rows this step produces (-1: none)
  int in_kind;        // 0: r of in_node; 1: own ring, publication `ia`; 2: xchg[0] + xchg[1]
  int ia;
  int base_kind;      // 0: none, 1: r, 2: saved t
  int save;           // 1: keep the output as the saved t of out_node, 2: as u of out_node
  int pub;            // publication index of the output (-1: none)
  int pubx;           // also publish into this group's cross-group buffer
  int xnode;          // the input vector is u of this node: keep this CTA's rows of it (-1: no)
  int pad;
};

struct K2Params {
  const double2* MT;
  const int64_t* Moff;
  const int64_t* nodeptr;
  int P;
  const double2* F;
  const K2Op* ops[2];
  int nops[2];
  const double2* r;
  double2* uvec;
  double2* tsave;
  K2Elem* ring[2];
  K2Elem* xchg[2];
  int RS;           // elements per CTA in a ring slot
  double tag0[2];   // tag of publication -1 of each group in this solve
  double pub_skew;  // 0; tests (KB_OPT_INJECT_FAULT = 3): the publishers' tags are off, nothing ever arrives
  int* err;
  unsigned long long wait_ns;  // time bound of every wait of the launch
  long long* timing;
  int bmax;
  int G0;
};

__device__ __forceinline__ void k2_publish(K2Elem* p, double2 v, double tag) {
  asm volatile("st.relaxed.gpu.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v.x), "d"(tag), "d"(v.y), "d"(tag)
               : "memory");
}

__device__ __forceinline__ double2 k2_poll(const K2Elem* p, double tag, int* err, unsigned long long wait_ns,
                                           volatile int* cta_failed) {
  double re, t0, im, t1;
  KbSpin sp;
  for (;;) {
    asm volatile("ld.relaxed.gpu.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(re), "=d"(t0), "=d"(im), "=d"(t1)
                 : "l"(p)
                 : "memory");
    if (t0 == tag && t1 == tag) break;
#if defined(K2_EXP) && (K2_EXP & 2)
    break;  // timing experiment: do not wait for the producers
#endif
    if (kb_spin_expired(sp, err, KB_WERR_XCHG, wait_ns, cta_failed)) break;
  }
  return zmake(re, im);
}

// owner CTA and local index of entry e of a b-row node split over `size` CTAs (inverse of
// kb_group_rows)
__device__ __forceinline__ void k2_owner(int b, int size, int e, int& c, int& li) {
  const int base = b / size, rem = b - base * size;
  const int cut = rem * (base + 1);
  if (e < cut) {
    c = e / (base + 1);
    li = e - c * (base + 1);
  } else {
    const int e2 = e - cut;
    c = rem + e2 / base;
    li = e2 - (c - rem) * base;
  }
}

// Named barriers.  2 = row warps only (their rows of the slice are in registers: the stage may be
// refilled).  The input vector is a two-slot producer / consumer buffer between the gather warps
// and the row warps with a FULL and an EMPTY barrier per slot (step parity): full (3, 5) = gather
// warps arrive, row warps wait; empty (6, 7) = row warps arrive when they have read the slot,
// gather warps wait before they overwrite it two steps later.  In a healthy launch the gather
// warps cannot run more than one step ahead anyway (an entry of step s+1 exists only after this
// CTA's row warps have published step s), but a launch whose waits have been abandoned (expired
// wait, watchdog) has no such data dependency left, and a barrier that receives the arrivals of
// two steps at once never completes again: the hand-shake must not depend on the data.
__device__ __forceinline__ void k2_bar_rows() { asm volatile("bar.sync 2, %0;" ::"n"(K2_RW * 32) : "memory"); }
__device__ __forceinline__ void k2_bar_full_arrive(int s) {
  asm volatile("bar.arrive %0, %1;" ::"r"((s & 1) ? 5 : 3), "n"(K2_THREADS) : "memory");
}
__device__ __forceinline__ void k2_bar_full_wait(int s) {
  asm volatile("bar.sync %0, %1;" ::"r"((s & 1) ? 5 : 3), "n"(K2_THREADS) : "memory");
}
__device__ __forceinline__ void k2_bar_empty_arrive(int s) {
  asm volatile("bar.arrive %0, %1;" ::"r"((s & 1) ? 7 : 6), "n"(K2_THREADS) : "memory");
}
__device__ __forceinline__ void k2_bar_empty_wait(int s) {
  asm volatile("bar.sync %0, %1;" ::"r"((s & 1) ? 7 : 6), "n"(K2_THREADS) : "memory");
}

// CPL = columns of a row per lane; NST = stages of the row slice in shared memory (2: the copy of
// step s+2 is issued when step s has its rows in registers; 1: the copy of step s+1, which then has
// the products, the reduction and the exchange of step s to arrive from L2 -- for slices too large
// for two stages)
template <int CPL, int NST>
__global__ void __launch_bounds__(K2_THREADS, 1) kb_sweep_fold(K2Params q, int slice_elems) {
  constexpr int K2_PPT = CPL / 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* stage0 = (double2*)smem_raw;                  // NST stages of this CTA's row slice of F
  const int bpad = (q.bmax + 7) & ~7;
  double2* vbuf = stage0 + NST * (size_t)slice_elems;    // 2 x bpad: input vector, by step parity
  int* s_nptr = (int*)(vbuf + 2 * (size_t)bpad);
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ int s_failed;  // some thread of this CTA has seen the launch fail
  volatile int* cta_failed = &s_failed;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int group = ((int)blockIdx.x < q.G0) ? 0 : 1;
  const int gsz[2] = {q.G0, (int)gridDim.x - q.G0};
  const int gsize = gsz[group];
  const int grank = group == 0 ? (int)blockIdx.x : (int)blockIdx.x - q.G0;
  const int S = q.nops[group];
  const K2Op* ops = q.ops[group];
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tc0 = clock64();
  // cycle counters only with -DKB_SWEEP_TICKS (make EXTRA=-DKB_SWEEP_TICKS): a disabled tick still
  // costs ~15 predicated instructions
#ifdef KB_SWEEP_TICKS
#define K2_TICK(k)              \
  do {                          \
    if (q.timing) {             \
      long long _t = clock64(); \
      tacc[k] += _t - tc0;      \
      tc0 = _t;                 \
    }                           \
  } while (0)
#else
#define K2_TICK(k) \
  do {             \
  } while (0)
#endif

  for (int i = tid; i <= q.P; i += K2_THREADS) s_nptr[i] = (int)q.nodeptr[i];
  if (tid == 0) {
    s_failed = 0;
    kb_mbar_init(&mbar[0], 1);
    kb_mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (wid >= K2_RW) {
    // ===== gather warps: poll the producers of every step's input vector =====
    // They hold no rows, so all of a thread's entries (<= K2_PPT) are in flight together: one
    // L2 round trip per attempt.  They run ahead of the row warps by construction: an entry of
    // step s+1 exists only after every row warp of the group is done with step s.
    const int pt = tid - K2_RW * 32;
    constexpr int NP = (K2_CW - K2_RW) * 32;
    int c_bi = -1, c_off[K2_PPT];
    K2Op op, nop;
    nop = ops[0];
    for (int s = 0; s < S; ++s) {
      op = nop;
      if (s + 1 < S) nop = ops[s + 1];
      const int oi = s_nptr[op.in_node], bi = s_nptr[op.in_node + 1] - oi;
      double2* v = vbuf + (size_t)(s & 1) * bpad;
      if (s >= 2) k2_bar_empty_wait(s);  // the row warps are done with the vector of step s - 2
      if (op.in_kind == 0) {
        for (int e = pt; e < bi; e += NP) v[e] = q.r[oi + e];
      } else if (op.in_kind == 1) {
        const K2Elem* slot = q.ring[group] + (size_t)(op.ia & (K2_RING - 1)) * gsize * q.RS;
        const double tag = q.tag0[group] + (double)(op.ia + 1);
        if (c_bi != bi) {  // owner of each of this thread's entries: divisions, once per node size
          c_bi = bi;
#pragma unroll
          for (int k = 0; k < K2_PPT; ++k) {
            const int e = pt + k * NP;
            int c = 0, li = 0;
            if (e < bi) k2_owner(bi, gsize, e, c, li);
            c_off[k] = c * q.RS + li;
          }
        }
        unsigned pend = 0u;
#pragma unroll
        for (int k = 0; k < K2_PPT; ++k)
          if (pt + k * NP < bi) pend |= 1u << k;
        KbSpin sp;
        while (pend) {
          double re[K2_PPT], t0[K2_PPT], im[K2_PPT], t1[K2_PPT];
#pragma unroll
          for (int k = 0; k < K2_PPT; ++k)
            if (pend & (1u << k))
              asm volatile("ld.relaxed.gpu.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                           : "=d"(re[k]), "=d"(t0[k]), "=d"(im[k]), "=d"(t1[k])
                           : "l"(slot + c_off[k])
                           : "memory");
#pragma unroll
          for (int k = 0; k < K2_PPT; ++k)
            if ((pend & (1u << k)) && t0[k] == tag && t1[k] == tag) {
              v[pt + k * NP] = zmake(re[k], im[k]);
              pend &= ~(1u << k);
            }
#if defined(K2_EXP) && (K2_EXP & 2)
          break;  // timing experiment: do not wait for the producers
#endif
          if (pend && kb_spin_expired(sp, q.err, KB_WERR_GATHER, q.wait_ns, cta_failed)) break;
        }
      } else {
        // t of the middle node = group 0's part (r - F t) + group 1's part (-F t), both in the
        // cross-group buffers under the first tag of the solve
        const double tagx0 = q.tag0[0] + 1.0, tagx1 = q.tag0[1] + 1.0;
        for (int e = pt; e < bi; e += NP) {
          int c, li;
          k2_owner(bi, gsz[0], e, c, li);
          const double2 a = k2_poll(q.xchg[0] + (size_t)c * q.RS + li, tagx0, q.err, q.wait_ns, cta_failed);
          k2_owner(bi, gsz[1], e, c, li);
          const double2 b2 = k2_poll(q.xchg[1] + (size_t)c * q.RS + li, tagx1, q.err, q.wait_ns, cta_failed);
          v[e] = zadd(a, b2);
        }
      }
      if (op.xnode >= 0) {
        // u of the middle node (assembled from both chains by every CTA): keep own rows
        asm volatile("bar.sync 4, %0;" ::"n"(NP) : "memory");
        int x0, x1;
        kb_group_rows(bi, gsize, grank, x0, x1);
        for (int e = x0 + pt; e < x1; e += NP) q.uvec[oi + e] = v[e];
      }
      __threadfence_block();
      k2_bar_full_arrive(s);
    }
    return;
  }

  // ===== row warps =====
  // rows 2w, 2w+1 of the slice belong to warp w < 4, rows 8 and 9 to warps 4 and 5: every
  // scheduler gets at most three rows
  auto slice_of = [&](int sn, const double2*& src, unsigned& bytes) {
    bytes = 0;
    src = nullptr;
    if (sn >= S) return;
    const long long mo = __ldg(&ops[sn].mat_off);
    if (mo < 0) return;
    const int inn = __ldg(&ops[sn].in_node), outn = __ldg(&ops[sn].out_node);
    const int bi = s_nptr[inn + 1] - s_nptr[inn], bo = s_nptr[outn + 1] - s_nptr[outn];
    int a0, a1;
    kb_group_rows(bo, gsize, grank, a0, a1);
    bytes = (unsigned)((size_t)(a1 - a0) * bi * sizeof(double2));
    src = q.F + mo + (size_t)a0 * bi;
  };
  auto issue_copy = [&](int sn) {
    const double2* src;
    unsigned bytes;
    slice_of(sn, src, bytes);
    if (bytes) {
      const int st = NST == 2 ? (sn & 1) : 0;
      uint64_t* mb = &mbar[st];
      kb_mbar_expect_tx(mb, bytes);
      kb_bulk_g2s(stage0 + (size_t)st * slice_elems, src, bytes, mb);
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
    if (NST == 2) issue_copy(1);
  }
  if (tid == 1)
    for (int sn = 2; sn < K2_AHEAD; ++sn) prefetch(sn);

  unsigned uses0 = 0u, uses1 = 0u;  // fills of each stage consumed so far (mbarrier parity)
  int c_bo = -1, c_a0 = 0, c_nr = 0;  // this CTA's rows of a bo-row node (division cached)
  const int row0 = wid < 4 ? 2 * wid : wid + 4;
  const int nrow = wid < 4 ? 2 : 1;

  K2Op op, nop;
  nop = ops[0];
  for (int s = 0; s < S; ++s) {
    op = nop;
    if (s + 1 < S) nop = ops[s + 1];
    const int oi = s_nptr[op.in_node], bi = s_nptr[op.in_node + 1] - oi;
    const double2* v = vbuf + (size_t)(s & 1) * bpad;

    // ---- 0. everything that does not depend on the exchange, in the shadow of its flight:
    //         this warp's rows of F from the staged slice into registers (after which the stage
    //         is free for the slice of step s+2), the base values
    int oo = 0, a0 = 0, nr = 0;
    double2 f[2][CPL];
    double2 base[2] = {zmake(0.0, 0.0), zmake(0.0, 0.0)};
    if (op.mat_off >= 0) {
      oo = s_nptr[op.out_node];
      const int bo = s_nptr[op.out_node + 1] - oo;
      if (bo != c_bo) {
        int a1;
        kb_group_rows(bo, gsize, grank, c_a0, a1);
        c_nr = a1 - c_a0;
        c_bo = bo;
      }
      a0 = c_a0;
      nr = c_nr;
      if (nr > 0) {
        if (NST == 2 && (s & 1)) {
          kb_mbar_wait(&mbar[1], uses1 & 1u, q.err, q.wait_ns);
          uses1++;
        } else {
          kb_mbar_wait(&mbar[0], uses0 & 1u, q.err, q.wait_ns);
          uses0++;
        }
        const double2* Fs = stage0 + (size_t)(NST == 2 ? (s & 1) : 0) * slice_elems + (size_t)row0 * bi + lane;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const bool rok = rr < nrow && row0 + rr < nr;
#pragma unroll
          for (int k = 0; k < CPL; ++k)
            f[rr][k] = (rok && lane + 32 * k < bi) ? Fs[(size_t)rr * bi + 32 * k] : zmake(0.0, 0.0);
          if (rok && lane == 0) {
            if (op.base_kind == 1) base[rr] = q.r[oo + a0 + row0 + rr];
            if (op.base_kind == 2) base[rr] = __ldcg(&q.tsave[oo + a0 + row0 + rr]);
          }
        }
      }
    }
    k2_bar_rows();  // every row warp holds its rows: the stage of step s may be refilled
    if (tid == 0) issue_copy(s + NST);
    if (tid == 1) prefetch(s + K2_AHEAD);
    K2_TICK(0);

    // ---- 1. the input vector (gather warps)
    k2_bar_full_wait(s);
    K2_TICK(1);

    // ---- 2. this warp's rows of  base - F v
    if (op.mat_off >= 0 && row0 < nr) {
      double2 acc[2][2];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) acc[rr][0] = acc[rr][1] = zmake(0.0, 0.0);
#if defined(K2_EXP) && (K2_EXP & 1)
      // timing experiment: no products
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        acc[0][k & 1] = zadd(acc[0][k & 1], f[0][k]);
        acc[1][k & 1] = zadd(acc[1][k & 1], f[1][k]);
      }
      acc[0][0] = zadd(acc[0][0], v[lane]);
#else
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int j = lane + 32 * k;
        const double2 vj = j < bi ? v[j] : zmake(0.0, 0.0);
        zfma(acc[0][k & 1], f[0][k], vj);
        zfma(acc[1][k & 1], f[1][k], vj);
      }
#endif
      K2_TICK(2);
      if (s + 2 < S) k2_bar_empty_arrive(s);  // this warp has read the vector of step s
      // packed shuffle tree: the halves of the warp swap rows first (lanes < 16 keep row 0,
      // lanes >= 16 row 1), then four stages inside each half: 10 double shuffles instead of 20
      double2 a0v = zadd(acc[0][0], acc[0][1]), a1v = zadd(acc[1][0], acc[1][1]);
      const bool hi = lane >= 16;
      double2 keep = hi ? a1v : a0v, give = hi ? a0v : a1v;
      keep.x += __shfl_xor_sync(0xffffffffu, give.x, 16);
      keep.y += __shfl_xor_sync(0xffffffffu, give.y, 16);
#pragma unroll
      for (int sft = 8; sft > 0; sft >>= 1) {
        keep.x += __shfl_xor_sync(0xffffffffu, keep.x, sft);
        keep.y += __shfl_xor_sync(0xffffffffu, keep.y, sft);
      }
      const int rr = hi ? 1 : 0;
      const int li = row0 + rr;
      const double bx = __shfl_sync(0xffffffffu, base[1].x, 0), by = __shfl_sync(0xffffffffu, base[1].y, 0);
      if ((lane & 15) == 0 && rr < nrow && li < nr) {
        const int gi = oo + a0 + li;
        const double2 bs = hi ? zmake(bx, by) : base[0];
        const double2 out = zsub(bs, keep);
        if (op.pub >= 0) {
          const double tag = q.tag0[group] + (double)(op.pub + 1) + q.pub_skew;
          k2_publish(q.ring[group] + ((size_t)(op.pub & (K2_RING - 1)) * gsize + grank) * q.RS + li, out, tag);
        }
        if (op.pubx) k2_publish(q.xchg[group] + (size_t)grank * q.RS + li, out, q.tag0[group] + 1.0);
        if (op.save == 1) __stcg(&q.tsave[gi], out);
        if (op.save == 2) q.uvec[gi] = out;
      }
      K2_TICK(3);
    } else if (s + 2 < S) {
      k2_bar_empty_arrive(s);  // no rows of this step: nothing to read
    }
  }
  if (q.timing && tid == 0)
    for (int k = 0; k < 8; ++k) q.timing[blockIdx.x * 8 + k] = tacc[k];
#undef K2_TICK
}

// Folded couplings of every node:  FL_p = L_{p+1,p} M_p  (rows of node p+1)  and
// FU_p = U_{p-1,p} M_p  (rows of node p-1), row-major, from the ELL copies of the couplings and
// the transposed factors.  One CTA = 32 output rows (lane = row, so that for a fixed column of
// F and coupling slot the warp reads consecutive entries of one row of M_p^T) x all columns in
// tiles of 32, transposed through shared memory for the stores.
__global__ void __launch_bounds__(256) kb_fold_couplings(int plo, int P, const int64_t* __restrict__ nodeptr,
                                                         const int64_t* __restrict__ Moff,
                                                         const double2* __restrict__ MT, const double2* __restrict__ Lval,
                                                         const int* __restrict__ Lcol, int WL,
                                                         const double2* __restrict__ Uval, const int* __restrict__ Ucol,
                                                         int WU, const int64_t* __restrict__ FLoff,
                                                         const int64_t* __restrict__ FUoff, double2* __restrict__ F) {
  extern __shared__ __align__(16) unsigned char fold_smem[];
  const int p = plo + blockIdx.y, kind = blockIdx.z;  // kind 0: FL_p, 1: FU_p
  const int pr = kind == 0 ? p + 1 : p - 1;           // node of the output rows
  if (pr < 0 || pr >= P) return;
  if ((kind == 0 ? FLoff[p] : FUoff[p]) < 0) return;  // not wanted (l-sharded: no separator on that side)
  const int o = (int)nodeptr[p], b = (int)(nodeptr[p + 1] - nodeptr[p]);
  const int orow = (int)nodeptr[pr], brow = (int)(nodeptr[pr + 1] - nodeptr[pr]);
  const int i0 = blockIdx.x * 32;
  if (i0 >= brow) return;
  const int W = kind == 0 ? WL : WU;
  const double2* val = kind == 0 ? Lval : Uval;
  const int* col = kind == 0 ? Lcol : Ucol;
  double2* out = F + (kind == 0 ? FLoff[p] : FUoff[p]);
  const double2* Mp = MT + Moff[p];
  double2* cv = (double2*)fold_smem;            // [W][32] coupling values of the 32 rows
  int* cc = (int*)(cv + (size_t)W * 32);        // [W][32] local columns (-1: padding)
  double2* tile = (double2*)(cc + (size_t)W * 32);  // [32 rows][33]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < W * 32; e += 256) {
    const int k = e >> 5, i = e & 31;
    double2 v = zmake(0.0, 0.0);
    int c = -1;
    if (i0 + i < brow) {
      const size_t at = (size_t)(orow + i0 + i) * W + k;
      const int gc = col[at] - o;
      if (gc >= 0 && gc < b) {
        c = gc;
        v = val[at];
      }
    }
    cv[e] = v;
    cc[e] = c;
  }
  __syncthreads();
  for (int j0 = 0; j0 < b; j0 += 32) {
    // warp w forms columns j0 + w, j0 + w + 8, ... of the tile for the 32 rows (lane = row)
    for (int jj = wid; jj < 32; jj += 8) {
      const int j = j0 + jj;
      double2 acc = zmake(0.0, 0.0);
      if (j < b) {
        const double2* mrow = Mp + (size_t)j * b;  // row j of M^T = column j of M
        for (int k = 0; k < W; ++k) {
          const int c = cc[k * 32 + lane];
          if (c >= 0) zfma(acc, cv[k * 32 + lane], __ldg(mrow + c));
        }
      }
      tile[lane * 33 + jj] = acc;
    }
    __syncthreads();
    for (int ii = wid; ii < 32; ii += 8) {
      const int i = i0 + ii, j = j0 + lane;
      if (i < brow && j < b) out[(size_t)i * b + j] = tile[ii * 33 + lane];
    }
    __syncthreads();
  }
}

// x_p = M_p u_p for every node, from the transposed factors: x_i = sum_j M^T[j][i] u_j.  One CTA
// = 32 consecutive rows i of one node (lane = row: a warp reads 512 contiguous bytes of a row
// of M_p^T per j), the 8 warps split j, partial sums meet in shared memory.
__global__ void __launch_bounds__(256) kb_fold_solution(int plo, const int64_t* __restrict__ nodeptr,
                                                        const int64_t* __restrict__ Moff,
                                                        const double2* __restrict__ MT, const double2* __restrict__ u,
                                                        double2* __restrict__ x) {
  extern __shared__ __align__(16) unsigned char xs_smem[];
  double2* us = (double2*)xs_smem;  // b
  __shared__ double2 part[8][32];
  const int p = plo + blockIdx.y;
  const int o = (int)nodeptr[p], b = (int)(nodeptr[p + 1] - nodeptr[p]);
  const int i0 = blockIdx.x * 32;
  if (i0 >= b) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < b; j += 256) us[j] = u[o + j];
  __syncthreads();
  const int i = i0 + lane;
  const double2* col = MT + Moff[p] + i;
  double2 acc0 = zmake(0.0, 0.0), acc1 = zmake(0.0, 0.0);
  if (i < b) {
    int j = wid;
    for (; j + 56 < b; j += 64) {
      double2 m[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = __ldcs(col + (size_t)(j + 8 * k) * b);
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        zfma(acc0, m[k], us[j + 8 * k]);
        zfma(acc1, m[k + 1], us[j + 8 * k + 8]);
      }
    }
    for (; j < b; j += 8) zfma(acc0, __ldcs(col + (size_t)j * b), us[j]);
  }
  part[wid][lane] = zadd(acc0, acc1);
  __syncthreads();
  if (wid == 0 && i < b) {
    double2 sum = part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) sum = zadd(sum, part[w][lane]);
    x[o + i] = sum;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// CTAs of the folded sweep.  Every CTA of a chain group must own at least one row of EVERY node
// (a CTA that publishes nothing is waited for by nobody, so nothing would bound how far it may
// fall behind the ring): the group size is capped by the smallest node.
int kbi_fold_grid(const kb_context* h, bool two_sided) {
  int64_t g = h->sweep_grid;
  const int64_t cap = two_sided ? 2 * h->bmin : h->bmin;
  if (g > cap) g = cap;
  return (int)g;
}

bool kbi_fold_supported(const kb_context* h, int G, bool two_sided, int* slice_elems_out, size_t* smem_out) {
  if (!h->M_transposed || G < 2 || h->P < 1) return false;
  const int gmin = two_sided ? G / 2 : G;
  if (gmin < 1) return false;
  const int64_t rpc = (h->bmax + gmin - 1) / gmin;
  if (rpc > K2_XR || h->bmax > 32 * K2_CPL_WIDE) return false;
  const int64_t slice_elems = (rpc * h->bmax + 7) & ~(int64_t)7;
  const size_t bpad = (size_t)((h->bmax + 7) & ~(int64_t)7);
  // two slice stages when they fit, one otherwise
  const size_t fixed = 2 * bpad * sizeof(double2) + (size_t)(h->P + 1) * sizeof(int) + 16;
  size_t smem = 2 * (size_t)slice_elems * sizeof(double2) + fixed;
  if (smem > 220 * 1024) smem = (size_t)slice_elems * sizeof(double2) + fixed;
  if (smem > 220 * 1024) return false;
  if (slice_elems_out) *slice_elems_out = (int)slice_elems;
  if (smem_out) *smem_out = smem;
  return true;
}

// Called at the end of kb_factor (after the ELL copies exist): folded couplings and the step
// schedules of the two groups.  Not being able to allocate the folded buffer is not an error:
// the solve then runs through kb_sweep1.cu.
int kbi_fold_prepare(kb_context* h) {
  h->fold_ready = false;
  if (h->opt_sweep != 1) return KB_OK;
  // the chain factored on this GPU: nodes [lo, hi) ([0, P) unless the pencil is l-sharded)
  const int64_t P = h->P, mid = h->mid, lo = h->ch_lo, hi = h->ch_hi;
  const bool two = mid < hi - 1;
  const int G = kbi_fold_grid(h, two);
  if (!kbi_fold_supported(h, G, two, nullptr, nullptr)) return KB_OK;
  cudaStream_t s = h->stream;
  // ---- folded buffer layout (couplings to nodes outside the chain are not folded: -1)
  std::vector<int64_t> FLoff(P, -1), FUoff(P, -1);
  int64_t tot = 0;
  auto bsz = [&](int64_t p) { return h->nodeptr[p + 1] - h->nodeptr[p]; };
  for (int64_t p = lo; p < hi; ++p) {
    if (p + 1 < hi) {
      FLoff[p] = tot;
      tot += bsz(p + 1) * bsz(p);
    }
    if (p > lo) {
      FUoff[p] = tot;
      tot += bsz(p - 1) * bsz(p);
    }
  }
  if (h->d_fold.alloc((size_t)(tot > 0 ? tot : 1)) != cudaSuccess) {
    cudaGetLastError();
    return KB_OK;
  }
  KB_CUDA(h, h->d_foldoff.alloc(2 * P));
  KB_CUDA(h, cudaMemcpyAsync(h->d_foldoff.p, FLoff.data(), P * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  KB_CUDA(h, cudaMemcpyAsync(h->d_foldoff.p + P, FUoff.data(), P * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  h->FLoff = FLoff;
  h->FUoff = FUoff;
  const int WL = h->WL > 0 ? h->WL : 1, WU = h->WU > 0 ? h->WU : 1;
  const int Wm = WL > WU ? WL : WU;
  const size_t fsm = (size_t)Wm * 32 * (sizeof(double2) + sizeof(int)) + 32 * 33 * sizeof(double2);
  if (fsm > 200 * 1024) return KB_OK;
  if (fsm > 48 * 1024)
    KB_CUDA(h, cudaFuncSetAttribute((const void*)kb_fold_couplings, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)fsm));
  dim3 grid((unsigned)((h->bmax + 31) / 32), (unsigned)(hi - lo), 2);
  kb_fold_couplings<<<grid, 256, fsm, s>>>((int)lo, (int)P, h->d_nodeptr.p, h->d_Moff.p, h->d_M.p, h->d_Lval.p,
                                           h->d_Lcol.p, WL, h->d_Uval.p, h->d_Ucol.p, WU, h->d_foldoff.p,
                                           h->d_foldoff.p + P, h->d_fold.p);
  h->launches++;
  KB_LAUNCH_CHECK(h);

  // ---- step schedules
  std::vector<K2Op> ops[2];
  auto mk = [](long long mat, int in_node, int out_node, int in_kind, int ia, int base, int save, int pub, int pubx,
               int xnode) {
    K2Op o;
    o.mat_off = mat;
    o.in_node = in_node;
    o.out_node = out_node;
    o.in_kind = in_kind;
    o.ia = ia;
    o.base_kind = base;
    o.save = save;
    o.pub = pub;
    o.pubx = pubx;
    o.xnode = xnode;
    o.pad = 0;
    return o;
  };
  {
    // group 0: nodes lo .. mid downwards, then back up
    int pub = 0;
    for (int64_t p = lo; p < mid; ++p) {
      const bool last = two && p == mid - 1;
      ops[0].push_back(mk(FLoff[p], (int)p, (int)p + 1, p == lo ? 0 : 1, pub - 1, 1, 1, pub, last ? 1 : 0, -1));
      ++pub;
    }
    for (int64_t p = mid; p >= lo + 1; --p) {
      int kind = 1;
      if (p == mid) kind = two ? 2 : (mid == lo ? 0 : 1);
      ops[0].push_back(
          mk(FUoff[p], (int)p, (int)p - 1, kind, pub - 1, p - 1 == lo ? 1 : 2, 2, pub, 0, p == mid ? (int)p : -1));
      ++pub;
    }
    if (mid == lo) ops[0].push_back(mk(-1, (int)lo, -1, 0, -1, 0, 0, -1, 0, (int)lo));  // single node: u = r
    h->fold_npub[0] = pub;
  }
  if (two) {
    // group 1: nodes hi-1 .. mid+1 upwards, the middle node's input, then back down
    int pub = 0;
    for (int64_t p = hi - 1; p > mid; --p) {
      const bool tomid = p - 1 == mid;
      ops[1].push_back(mk(FUoff[p], (int)p, (int)p - 1, p == hi - 1 ? 0 : 1, pub - 1, tomid ? 0 : 1, tomid ? 0 : 1, pub,
                          tomid ? 1 : 0, -1));
      ++pub;
    }
    ops[1].push_back(mk(FLoff[mid], (int)mid, (int)mid + 1, 2, -1, mid + 1 == hi - 1 ? 1 : 2, 2, pub, 0, -1));
    ++pub;
    for (int64_t p = mid + 1; p < hi - 1; ++p) {
      ops[1].push_back(mk(FLoff[p], (int)p, (int)p + 1, 1, pub - 1, p + 1 == hi - 1 ? 1 : 2, 2, pub, 0, -1));
      ++pub;
    }
    h->fold_npub[1] = pub;
  } else {
    h->fold_npub[1] = 0;
  }
  h->fold_nops[0] = (int)ops[0].size();
  h->fold_nops[1] = (int)ops[1].size();
  const size_t nall = ops[0].size() + ops[1].size();
  KB_CUDA(h, h->d_foldops.alloc(nall * sizeof(K2Op)));
  // pageable source: the copy is staged by the runtime before the call returns
  KB_CUDA(h, cudaMemcpyAsync(h->d_foldops.p, ops[0].data(), ops[0].size() * sizeof(K2Op), cudaMemcpyHostToDevice, s));
  if (!ops[1].empty())
    KB_CUDA(h, cudaMemcpyAsync(h->d_foldops.p + ops[0].size() * sizeof(K2Op), ops[1].data(),
                               ops[1].size() * sizeof(K2Op), cudaMemcpyHostToDevice, s));
  h->fold_ready = true;
  return KB_OK;
}

// y <- T'^{-1} r with the folded factors, on the chain [ch_lo, ch_hi) of this GPU (entries of r
// and y outside it are not touched).  y has n+1 entries, y[n] == 0.  ends_only: form the solution
// on the first and the last node of the chain only (first pass of the l-sharded solve: the
// separator equations need nothing else).
int kbi_sweep_fold(kb_context* h, const double2* r, double2* y, int ends_only) {
  cudaStream_t s = h->stream;
  const int n = (int)h->n;
  const bool two = h->mid < h->ch_hi - 1;
  const int G = kbi_fold_grid(h, two);
  int slice_elems = 0;
  size_t smem = 0;
  if (!h->fold_ready || !kbi_fold_supported(h, G, two, &slice_elems, &smem))
    return kb_fail(h, KB_EINVAL, "chain does not fit the folded sweep kernel");
  const int G0 = two ? (G + 1) / 2 : G;
  const int gmin = two ? G / 2 : G;
  const int RS = (int)(((h->bmax + gmin - 1) / gmin + 3) & ~(int64_t)3);
  const size_t ring_elems = (size_t)K2_RING * G * RS;  // both groups
  const size_t xchg_elems = (size_t)G * RS;
  const size_t need = (ring_elems + xchg_elems) * sizeof(K2Elem);
  if (h->d_foldring.count < need) {
    KB_CUDA(h, h->d_foldring.alloc(need));
    KB_CUDA(h, cudaMemsetAsync(h->d_foldring.p, 0, need, s));  // tag 0 never matches: tags start at 1
    h->fold_epoch[0] = h->fold_epoch[1] = 0;
  }
  if (h->d_yf.count < (size_t)n + 1) KB_CUDA(h, h->d_yf.alloc(n + 1));
  K2Params q;
  q.MT = h->d_M.p;
  q.Moff = h->d_Moff.p;
  q.nodeptr = h->d_nodeptr.p;
  q.P = (int)h->P;
  q.F = h->d_fold.p;
  q.ops[0] = (const K2Op*)h->d_foldops.p;
  q.ops[1] = (const K2Op*)h->d_foldops.p + h->fold_nops[0];
  q.nops[0] = h->fold_nops[0];
  q.nops[1] = h->fold_nops[1];
  q.r = r;
  KB_CUDA(h, h->d_uvec.alloc((size_t)n + 1));
  q.uvec = h->d_uvec.p;
  q.tsave = h->d_yf.p;
  K2Elem* base = (K2Elem*)h->d_foldring.p;
  q.ring[0] = base;
  q.ring[1] = base + (size_t)K2_RING * G0 * RS;
  q.xchg[0] = base + ring_elems;
  q.xchg[1] = base + ring_elems + (size_t)G0 * RS;
  q.RS = RS;
  // publication i of a group carries tag0 + i + 1; the cross-group entries carry tag0 + 1
  q.tag0[0] = (double)h->fold_epoch[0];
  q.tag0[1] = (double)h->fold_epoch[1];
  h->fold_epoch[0] += (unsigned long long)h->fold_npub[0] + 1ull;
  h->fold_epoch[1] += (unsigned long long)h->fold_npub[1] + 1ull;
  q.err = h->d_sweep_err.p;
  q.wait_ns = h->wait_ns;
  q.pub_skew = 0.0;
  if (h->inject_fault == 3) {
    q.pub_skew = 0.5;
    h->inject_fault = 0;
  }
  q.timing = h->d_sweep_timing.p;
  q.bmax = (int)h->bmax;
  q.G0 = G0;
  const bool one_stage = smem < 2 * (size_t)slice_elems * sizeof(double2);
  const bool wide = h->bmax > 32 * K2_CPL;
  const void* fn = wide ? (one_stage ? (const void*)kb_sweep_fold<K2_CPL_WIDE, 1> : (const void*)kb_sweep_fold<K2_CPL_WIDE, 2>)
                        : (one_stage ? (const void*)kb_sweep_fold<K2_CPL, 1> : (const void*)kb_sweep_fold<K2_CPL, 2>);
  if (smem > 48 * 1024) KB_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&q, (void*)&slice_elems};
  KB_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(K2_THREADS), args, smem, s));
  const size_t xsm = (size_t)h->bmax * sizeof(double2);
  if (!ends_only) {
    dim3 xgrid((unsigned)((h->bmax + 31) / 32), (unsigned)(h->ch_hi - h->ch_lo));
    kb_fold_solution<<<xgrid, 256, xsm, s>>>((int)h->ch_lo, h->d_nodeptr.p, h->d_Moff.p, h->d_M.p, h->d_uvec.p, y);
  } else {
    dim3 xgrid((unsigned)((h->bmax + 31) / 32), 1);
    kb_fold_solution<<<xgrid, 256, xsm, s>>>((int)h->ch_lo, h->d_nodeptr.p, h->d_Moff.p, h->d_M.p, h->d_uvec.p, y);
    if (h->ch_hi - 1 > h->ch_lo) {
      kb_fold_solution<<<xgrid, 256, xsm, s>>>((int)h->ch_hi - 1, h->d_nodeptr.p, h->d_Moff.p, h->d_M.p, h->d_uvec.p,
                                               y);
      h->launches++;
    }
  }
  h->launches += 2;
  KB_LAUNCH_CHECK(h);
  return KB_OK;
}
