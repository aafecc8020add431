// Internal declarations shared by the translation units of libkoreb200.so.
// Not part of the C ABI (see include/kore_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <complex>
#include <string>
#include <vector>

#include "../../include/kore_b200.h"

typedef std::complex<double> zcomplex;

// ---------------------------------------------------------------------------
// complex128 device arithmetic on double2 (x = re, y = im)
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ double2 zmake(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ double2 zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 zsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double2 zneg(double2 a) { return make_double2(-a.x, -a.y); }
__host__ __device__ __forceinline__ double2 zconj(double2 a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ double2 zmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ double2 zscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
// acc += a*b  (4 real FMAs)
__device__ __forceinline__ void zfma(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a*b
__device__ __forceinline__ void zfms(double2& acc, double2 a, double2 b) {
  acc.x = fma(-a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void zfmac(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
__host__ __device__ __forceinline__ double2 kb_as_complex(double v) { return make_double2(v, 0.0); }
__host__ __device__ __forceinline__ double2 kb_as_complex(double2 v) { return v; }
__host__ __device__ __forceinline__ double zabs2(double2 a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ double2 zinv(double2 p) {
  // scaled reciprocal (avoids overflow of |p|^2)
  double s = fmax(fabs(p.x), fabs(p.y));
  double xr = p.x / s, xi = p.y / s;
  double d = (xr * xr + xi * xi) * s;
  return make_double2(xr / d, -xi / d);
}

// reciprocal with a single division (operands are equilibrated: no scaling needed)
__host__ __device__ __forceinline__ double2 zinv1(double2 p) {
  double d = 1.0 / (p.x * p.x + p.y * p.y);
  return make_double2(p.x * d, -p.y * d);
}

// reciprocal through the hardware double-precision reciprocal (correctly rounded 1/d)
__device__ __forceinline__ double2 zinv_fast(double2 p) {
  double d = __drcp_rn(p.x * p.x + p.y * p.y);
  return make_double2(p.x * d, -p.y * d);
}

// ---------------------------------------------------------------------------
// spin-wait bound of every device-side wait (a peer that never arrives must not hang the GPU)
// ---------------------------------------------------------------------------
#define KB_SPIN_LIMIT (1 << 22)

// Every device-side wait of the persistent kernels is bounded in TIME, not in polls, and gives up
// at once when any other wait of the same launch has already failed (or the host watchdog has
// raised the flag): a protocol failure ends the kernel within milliseconds of the first time-out
// instead of paying one time-out per wait.  The flag word receives a code that says which wait
// failed and where:  code | (blockIdx.x << 8).
#define KF_ERR_WORD 190  // word of kb_context::d_kfsync that holds the factorisation kernel's flag
#define KB_WAIT_NS_DEFAULT 4000000000ull  // 4 s: longer than any legitimate wait of any kernel here
enum {
  KB_WERR_GROUP_BARRIER = 1,  // kb_chainfac: end-of-node barrier of a chain group
  KB_WERR_COUNTER = 2,        // kb_chainfac: step / chain counter
  KB_WERR_STREAM = 3,         // kb_chainfac: tagged element of the column stream
  KB_WERR_BACKPRESSURE = 4,   // kb_chainfac: consumers of the ring slot about to be rewritten
  KB_WERR_GATHER = 5,         // folded sweep: entry of the next input vector
  KB_WERR_XCHG = 6,           // folded sweep: cross-group entry of the middle node
  KB_WERR_MBAR = 7,           // bulk (TMA) copy completion
  KB_WERR_SWEEP = 8,          // row-split / one-hop sweeps (kb_sweep.cu, kb_sweep1.cu)
  KB_WERR_WATCHDOG = 9,       // raised by the host: kernel exceeded its deadline
  KB_WERR_MAILBOX = 10        // l-sharded path: message of a peer rank (peer-memory mailbox)
};

__device__ __forceinline__ unsigned long long kb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct KbSpin {
  unsigned n = 0;
  unsigned long long t0 = 0;
};
// true: stop waiting (the launch has failed, here or elsewhere).  The fast path of a wait is
// untouched: nothing happens before the first missed poll, the first missed poll reads a flag in
// SHARED memory (`cta_failed`, may be null: set once any thread of the CTA has seen the failure,
// so that a failed launch drains at one missed poll per wait), and only every 32nd missed poll
// reads the launch's flag in global memory and the clock.
__device__ __forceinline__ bool kb_spin_expired(KbSpin& s, int* err, int code, unsigned long long limit_ns,
                                                volatile int* cta_failed = nullptr) {
  const unsigned n = ++s.n;
  if (n == 1u) return cta_failed != nullptr && *cta_failed != 0;
  if ((n & 31u) != 0u) return false;
  if (*(volatile int*)err != 0) {
    if (cta_failed) *cta_failed = 1;
    return true;
  }
  const unsigned long long now = kb_globaltimer();
  if (s.t0 == 0) {
    s.t0 = now;
    return false;
  }
  if (now - s.t0 > limit_ns) {
    atomicCAS(err, 0, code | ((int)blockIdx.x << 8));
    if (cta_failed) *cta_failed = 1;
    return true;
  }
  return false;
}
__device__ __forceinline__ bool kb_launch_failed(const int* err) { return *(volatile const int*)err != 0; }

// ---------------------------------------------------------------------------
// Peer-memory mailboxes of the l-sharded path (kb_shard.cu): every rank owns a buffer that all
// other ranks of the node map through CUDA IPC and write with plain NVLink stores.  A collective
// of a few hundred bytes (Gram-Schmidt coefficients, a norm, separator contributions, a halo node)
// is then: store the payload into every peer's slot, fence, store the epoch into every peer's flag,
// poll the own flags, read the own slots -- inside the kernel that produced the payload, with no
// collective launch and no proxy thread.  Layout of one buffer:
//   data  [2 parities][KB_MB_MAXRANKS senders][KB_MB_SLOT bytes]
//   flags [2 parities][KB_MB_MAXRANKS senders] x one 128-byte line (unsigned epoch of the last message)
// Consecutive collectives alternate the parity; every collective raises a flag on EVERY peer, so a
// rank can never be more than one collective ahead of any other and two parities suffice.
// ---------------------------------------------------------------------------
#define KB_MB_MAXRANKS 16
#define KB_MB_SLOT 32768
#define KB_MB_BYTES ((size_t)2 * KB_MB_MAXRANKS * KB_MB_SLOT + (size_t)2 * KB_MB_MAXRANKS * 128)
struct KbMailbox {
  int rank, nranks;
  unsigned char* base[KB_MB_MAXRANKS];  // base[q]: rank q's buffer as mapped here (base[rank]: the local one)
};
__device__ __forceinline__ double2* kb_mb_data(const KbMailbox& m, int owner, int par, int sender) {
  return (double2*)(m.base[owner] + (size_t)(par * KB_MB_MAXRANKS + sender) * KB_MB_SLOT);
}
__device__ __forceinline__ unsigned* kb_mb_flag(const KbMailbox& m, int owner, int par, int sender) {
  return (unsigned*)(m.base[owner] + (size_t)2 * KB_MB_MAXRANKS * KB_MB_SLOT) + (par * KB_MB_MAXRANKS + sender) * 32;
}
// After the CTA has stored its payloads: make them visible system-wide, then raise this rank's
// flag on every peer (all threads call; one thread fences and stores).
__device__ __forceinline__ void kb_mb_signal_all(const KbMailbox& m, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const int par = (int)(epoch & 1u);
    for (int q = 0; q < m.nranks; ++q)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(kb_mb_flag(m, q, par, m.rank)), "r"(epoch) : "memory");
  }
}
// Wait until every rank's message of this epoch has landed in the own buffer (all threads call).
__device__ __forceinline__ void kb_mb_wait_all(const KbMailbox& m, unsigned epoch, int* err, unsigned long long wait_ns) {
  if ((int)threadIdx.x < m.nranks) {
    const unsigned* f = kb_mb_flag(m, m.rank, (int)(epoch & 1u), (int)threadIdx.x);
    KbSpin sp;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if (kb_spin_expired(sp, err, KB_WERR_MAILBOX, wait_ns)) break;
    }
  }
  __syncthreads();
}

// ---- mbarrier + 1-D bulk (TMA) copy helpers
__device__ __forceinline__ unsigned kb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kb_mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kb_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kb_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kb_smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void kb_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   kb_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(kb_smem_addr(bar))
               : "memory");
}
// Wait for the completion of a bulk copy.  Never abandoned because something ELSE has failed: a
// copy in flight lands on its own, and a CTA must not refill the stage or exit before it has (a
// launch that is draining after an expired wait still consumes every copy it has issued).  Only
// a copy that itself does not land within the bound raises the launch's flag.
__device__ __forceinline__ void kb_mbar_wait(uint64_t* bar, unsigned parity, int* err, unsigned long long limit_ns) {
  unsigned done = 0;
  unsigned tries = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(kb_smem_addr(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if ((++tries & 255u) == 0u) {
      const unsigned long long now = kb_globaltimer();
      if (t0 == 0) t0 = now;
      if (now - t0 > limit_ns) {
        atomicCAS(err, 0, KB_WERR_MBAR | ((int)blockIdx.x << 8));
        break;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
struct kb_context;
int kb_fail(kb_context* h, int code, const char* fmt, ...);

#define KB_CUDA(h, expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return kb_fail((h), KB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                  \
  } while (0)

#define KB_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != KB_OK) return _rc; \
  } while (0)

#define KB_LAUNCH_CHECK(h)                                                                     \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return kb_fail((h), KB_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                      \
  } while (0)

// ---------------------------------------------------------------------------
// device buffer that frees itself
// ---------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t count = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    count = 0;
  }
  cudaError_t alloc(size_t n) {
    if (n <= count && p) return cudaSuccess;
    release();
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e == cudaSuccess) count = n;
    return e;
  }
  size_t bytes() const { return count * sizeof(T); }
};

struct HostCSR {
  int64_t n = 0;
  std::vector<int64_t> indptr;
  std::vector<int64_t> indices;
  std::vector<zcomplex> values;
  bool present = false;
};

// The caller's CSR as uploaded by kb_set_pencil (device), input of the device-side layout build.
struct KbRawCSR {
  int64_t n = 0, nnz = 0;
  int index_bytes = 4;
  bool is_complex = false, present = false;
  DevBuf<int64_t> indptr;
  DevBuf<unsigned char> indices, values;
};

// One contiguous run of chain nodes factored by block-Thomas on this GPU.
struct kb_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;

  // options
  int opt_equil = 1;
  int opt_refine = 1;
  int opt_refine_eigs = -1;  // -1: automatic (kb_eigs probes the accuracy of the unrefined solve)
  int opt_purify = 1;
  int64_t opt_seed = 1;
  int opt_panel = 0;
  int opt_factor = 1;  // 1: persistent strip kernel (kb_chainfac.cu) when the nodes fit, 0: per-step kernels
  // Time bound of every device-side wait of the persistent kernels (KB_OPT_WAIT_MS).  A launch
  // whose wait expires raises its error flag and drains; the host then switches the handle to the
  // kernels without device-side waits (safe mode), refactors and repeats the call.
  unsigned long long wait_ns = KB_WAIT_NS_DEFAULT;
  bool safe_mode = false;
  int inject_fault = 0;  // KB_OPT_INJECT_FAULT (tests): 1 = the next chain sweep, 2 = the next factorisation reports a time-out

  // pencil as given (host, original ordering)
  int64_t n = 0;
  HostCSR A, B;
  KbRawCSR rawA, rawB;  // device copies (layout built on the device unless KB_HOST_LAYOUT is set)
  bool b_is_complex = false;

  // chain
  bool chain_set = false;
  int64_t P = 0;
  std::vector<int64_t> nodeptr;  // P+1
  std::vector<int64_t> perm;     // chain position -> original index
  int64_t bmax = 0, bmin = 0;

  // device: permutation
  DevBuf<int> d_perm;  // n

  // device: chain CSR on the union pattern of A and B (rows sorted by chain column)
  int64_t nnz = 0;
  DevBuf<int64_t> d_rowptr;  // n+1
  DevBuf<int> d_col;         // nnz (chain column)
  DevBuf<double2> d_Aval;    // nnz
  DevBuf<int> d_bmap;        // nnzB: position of every B entry in the union pattern
  DevBuf<double2> d_Tval;    // nnz (equilibrated A - sigma B)
  DevBuf<int64_t> d_dstart;  // n: first entry of row in its own node
  DevBuf<int64_t> d_ustart;  // n: first entry of row in the next node
  // U blocks (row node p, column node p+1) by column
  int64_t nnzU = 0;
  DevBuf<int64_t> d_ucptr;  // n+1
  DevBuf<int> d_urow;       // nnzU chain row
  DevBuf<int64_t> d_upos;   // nnzU position in d_Tval
  // L blocks (row node p, column node p-1) by column
  DevBuf<int64_t> d_lcptr;
  DevBuf<int> d_lrow;
  DevBuf<int64_t> d_lpos;
  // B alone in chain order for the Arnoldi SpMV
  int64_t nnzB = 0;
  DevBuf<int64_t> d_browptr;
  DevBuf<int> d_bcol;
  DevBuf<double> d_bval_r;    // real B (assemble.py writes float64)
  DevBuf<double2> d_bval_c;   // complex B

  // equilibration (powers of two), chain order
  DevBuf<double> d_rscale, d_cscale;
  DevBuf<unsigned long long> d_maxbits;  // n scratch for column maxima

  // factors
  bool factored = false;
  zcomplex sigma = 0;
  std::vector<int64_t> Moff;  // P+1 offsets (in complex elements) of M_p in d_M
  DevBuf<double2> d_M;
  // factor workspaces
  DevBuf<double2> d_S0, d_S1, d_W, d_Gp, d_PT;
  // second elimination chain (two-sided factorisation)
  cudaStream_t stream2 = nullptr;
  DevBuf<double2> d_S0b, d_S1b, d_Wb, d_Gpb, d_PTb;
  DevBuf<int> d_origb, d_srcrowb;
  DevBuf<unsigned> d_ready;
  // persistent strip factorisation (kb_chainfac.cu)
  DevBuf<double2> d_kfG;
  DevBuf<int> d_kfpiv;
  DevBuf<unsigned> d_kfsync;
  DevBuf<double> d_kfstream;  // column stream to the next strip owner (tagged 32-byte elements)
  bool M_transposed = false;  // d_M holds M_p^T (strip kernel + one-hop sweep, kb_sweep1.cu)
  DevBuf<double2> d_ring;     // partial-product ring + exchange buffers of the one-hop sweep
  DevBuf<unsigned> d_k1flags; // publication flags of the one-hop sweep (one 256-byte line per CTA)
  unsigned k1_epoch = 0, k1_epoch1 = 0;
  DevBuf<int4> d_rng;         // coupling ranges per (CTA, step) of the one-hop sweep
  bool rng_valid = false;
  // folded sweep (kb_sweep2.cu): FL_p = L M_p, FU_p = U M_p, step schedules, exchange ring
  int opt_fold = 1;
  bool fold_ready = false;
  bool fold_required = false;  // the factors are transposed for the folded sweep alone (no one-hop fall-back)
  int sweep_grid_hint = 0;
  DevBuf<double2> d_fold, d_uvec;
  DevBuf<int64_t> d_foldoff;
  DevBuf<unsigned char> d_foldops, d_foldring;
  int fold_npub[2] = {0, 0}, fold_nops[2] = {0, 0};
  unsigned long long fold_epoch[2] = {0, 0};
  int64_t mid = 0;  // middle node of the two-sided elimination (ch_hi-1: one-sided)
  int64_t ch_lo = 0, ch_hi = 0;  // the chain factored on this GPU: [0, P), or the rank's interior when l-sharded
  std::vector<int64_t> FLoff, FUoff;  // offsets of FL_p / FU_p in d_fold (-1: not folded)
  // ELL copies of the couplings + node tables for the persistent sweep kernel
  int WL = 0, WU = 0;
  DevBuf<double2> d_Lval, d_Uval;
  DevBuf<int> d_Lcol, d_Ucol;
  DevBuf<int64_t> d_nodeptr, d_Moff;
  DevBuf<unsigned> d_flags;
  DevBuf<int> d_sweep_err;
  DevBuf<long long> d_sweep_timing;
  int sweep_grid = 0;
  int opt_sweep = 1;
  DevBuf<int> d_orig, d_srcrow, d_piv, d_info;

  // solve workspaces (chain order, scaled space)
  DevBuf<double2> d_r, d_y, d_res, d_x0, d_in, d_out, d_t, d_t2, d_yf;
  DevBuf<double2> d_io_a, d_io_b;  // staging of the host-pointer entry points
  int64_t ws_n = -1;               // n the zero sentinels of d_y / d_x0 were written for
  DevBuf<double> d_partial;

  // Krylov workspaces
  DevBuf<double2> d_V;  // n x (ncv+1), column-major
  DevBuf<double2> d_w, d_w2, d_h, d_hpart, d_Q, d_xo;
  DevBuf<double> d_normpart, d_beta;
  void* pinned_h = nullptr;
  void* pinned_beta = nullptr;
  int pinned_ncv = 0;          // the pinned mirrors hold (pinned_ncv + 2)^2 / pinned_ncv + 2 entries
  int* pinned_err = nullptr;   // host mirror of the sweep error flag, read back with every restart
  cudaStream_t side_stream = nullptr;  // watchdog: raises the error flags while a kernel runs
  double watchdog_ms = 0.0;    // 0: 3 x the device-side wait bound + 5 s

  // captured sweep graphs, keyed by the (rhs, solution) buffers
  struct SweepGraph {
    const double2* r;
    double2* y;
    void* exec;
  };
  std::vector<SweepGraph> sweep_graphs;

  // CUDA-event timing of the chain sweeps inside kb_eigs
  bool time_sweeps = false;
  std::vector<cudaEvent_t> sweep_events;

  // sharding (l-sharded path, kb_shard.cu)
  int rank = 0, nranks = 1;
  void* nccl_comm = nullptr;
  std::vector<int64_t> seg_lo, seg_hi;        // node range of every rank
  int64_t int_lo = 0, int_hi = 0;             // this rank's interior nodes [int_lo, int_hi)
  int64_t top_sep = -1, bot_sep = -1;         // separator node above / below (-1: none)
  std::vector<int64_t> Voff;                  // offsets of V_p / G_p (interior nodes) in d_Vsp / d_Gsp
  DevBuf<double2> d_Vsp, d_Gsp;               // spikes: V_p (b_p x b_t), G_p (b_t x b_p)
  DevBuf<double2> d_F, d_H, d_Acc;            // factor-time workspaces
  DevBuf<double2> d_contrib, d_contrib_all;   // reduced-system blocks: mine / all ranks
  DevBuf<double2> d_Mr, d_Csub, d_Csup;       // reduced system: inverses and couplings (G-1 nodes)
  DevBuf<double2> d_sepvec, d_sepvec_all;     // per-solve separator contributions
  DevBuf<double2> d_redz;                     // reduced-solve scratch
  DevBuf<int64_t> d_redtab;                   // {0, b_j} per separator + a zero: one-node chains for the strip kernel
  // peer-memory mailboxes (CUDA IPC): own buffer, peers' mappings, collective counter
  DevBuf<unsigned char> d_mbox;
  void* mb_peer[KB_MB_MAXRANKS] = {nullptr};
  bool mb_ready = false;
  unsigned mb_epoch = 0;
  DevBuf<unsigned> d_redctr;                  // grid-barrier counter of the fused reduced solve
  unsigned red_epoch = 0;
  int red_grid = 0;
  bool keep_sharded = false;                  // inside kb_eigs: solves leave their result on the rank's segment only
  bool shard_fast = false;                    // interior factored by the strip kernel, solved by the folded sweep
  DevBuf<double2> d_spk;                      // fast path: identity-column chains and corner blocks of T_I^{-1}

  kb_stats stats;
  int64_t launches = 0;

  kb_context() { memset(&stats, 0, sizeof(stats)); }
};

static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// A pair of timing events that cannot leak on an early return.
struct KbEventPair {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KbEventPair() {}
  KbEventPair(const KbEventPair&) = delete;
  KbEventPair& operator=(const KbEventPair&) = delete;
  ~KbEventPair() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
  cudaError_t create() {
    cudaError_t e = cudaEventCreate(&e0);
    if (e != cudaSuccess) return e;
    return cudaEventCreate(&e1);
  }
  float ms() const {
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    return t;
  }
};

// ---- kb_factor.cu
int kbi_factor(kb_context* h, zcomplex sigma);
int kbi_build_T(kb_context* h, zcomplex sigma);
int kbi_factor_workspace(kb_context* h);
int kbi_panel_width(const kb_context* h, int n);
// ---- kb_layout.cu
int kbi_upload_raw(kb_context* h, KbRawCSR& M, int64_t n, int index_bytes, const void* indptr, const void* indices,
                   const void* values, bool is_complex, const char* name);
int kbi_layout_device(kb_context* h);
// ---- kb_chainfac.cu
bool kbi_chainfac_supported(const kb_context* h);
// a dense b x b block inverted by the strip kernel as a one-node chain (row-major in, row-major out)
struct KbDenseNode {
  const double2* S;
  double2* out;
  const int64_t* nodeptr;  // device: {0, b}
  const int64_t* moff;     // device: {0}
};
int kbi_chainfac_run(kb_context* h, bool two_sided, bool transposed, int64_t plo = 0, int64_t phi = -1,
                     const KbDenseNode* dense = nullptr);
// ---- kb_solve.cu
//  chain solve in scaled/permuted space: d_y <- T'^{-1} d_r (d_r preserved)
int kbi_chain_solve(kb_context* h, const double2* r_dev, double2* x_dev, int refine);
//  full operator on chain-ordered device vectors: out = C T'^{-1} R B in
int kbi_apply_op_chain(kb_context* h, const double2* in_chain, double2* out_chain, int refine);
int kbi_spmv_B_chain(kb_context* h, const double2* x, double2* y, bool scale_rows);
int kbi_spmv_A_chain(kb_context* h, const double2* x, double2* y);
int kbi_to_chain(kb_context* h, const double2* x_orig_dev, double2* x_chain_dev);
int kbi_from_chain(kb_context* h, const double2* x_chain_dev, double2* x_orig_dev);
int kbi_solve_workspace(kb_context* h);
int kbi_check_sweep_error(kb_context* h);
// cudaStreamSynchronize(h->stream) with a deadline: a kernel that outlives it has the error flags
// of the persistent kernels raised from a side stream (every device-side wait then gives up) and
// the call fails with KB_ECUDA once the stream has drained
int kbi_sync(kb_context* h);
int kbi_enter_safe_mode(kb_context* h, const char* what, int wait_code);
// internal return code: a persistent kernel reported a protocol time-out and the handle has been
// switched to safe mode and refactored; the caller repeats its operation
#define KB_EPROTOCOL_RETRY 1000
void kbi_drop_graphs(kb_context* h);
// ---- kb_sweep.cu
int kbi_sweep_prepare(kb_context* h);
int kbi_sweep_persistent(kb_context* h, const double2* r, double2* y);
int kbi_sweep_dataflow(kb_context* h, const double2* r, double2* y);
// ---- kb_sweep1.cu
bool kbi_onehop_supported(const kb_context* h, int G, bool two_sided, int* slice_elems_out, size_t* smem_out);
int kbi_sweep_onehop(kb_context* h, const double2* r, double2* y);
// ---- kb_sweep2.cu
int kbi_fold_grid(const kb_context* h, bool two_sided);
bool kbi_fold_supported(const kb_context* h, int G, bool two_sided, int* slice_elems_out, size_t* smem_out);
int kbi_fold_prepare(kb_context* h);
int kbi_sweep_fold(kb_context* h, const double2* r, double2* y, int ends_only = 0);
// ---- kb_shard.cu
int kbi_factor_sharded(kb_context* h, zcomplex sigma);
int kbi_chain_solve_sharded(kb_context* h, const double2* r_dev, double2* x_dev, int refine);
int kbi_sharded_sweeps(kb_context* h, const double2* r, double2* y);
void kbi_nccl_destroy(kb_context* h);
void kbi_shard_rows(const kb_context* h, int64_t* row_lo, int64_t* row_hi);
int kbi_shard_allreduce(kb_context* h, double* buf, size_t count);
int kbi_shard_halo(kb_context* h, double2* x);
int kbi_shard_gather_segments(kb_context* h, double2* x);
// fused local reduction + sum over the ranks (peer-memory mailboxes when mapped, NCCL otherwise)
int kbi_shard_reduce_cols(kb_context* h, int ncols, int nchunks, const double2* hpart, double2* hdev, double2* hsum,
                          double2* scratch, int accumulate);
int kbi_shard_norm(kb_context* h, int nparts, const double* normpart, double* beta_dev);
__global__ void kb_norm2_partial(int n, const double2* __restrict__ v, double* __restrict__ out);
