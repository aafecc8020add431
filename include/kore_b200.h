/*
 * kore_b200.h -- C ABI of libkoreb200.so, the B200-native shift-and-invert
 * eigen / linear solver that replaces the SLEPc EPS / ST / PETSc KSP + MUMPS
 * block of Kore's bin/solve.py.
 *
 * The reference has no FFI of its own for this path: it reaches PETSc/SLEPc
 * through petsc4py/slepc4py object calls.  Each entry point below names the
 * reference call(s) it replaces (file:line under /root/reference).
 *
 * Conventions
 *   - plain C, no torch / C++ types; complex128 is `double[2]` (re, im),
 *     passed as `const double*` of length 2*count (numpy complex128 layout);
 *   - every pointer is CALLER-OWNED HOST memory unless the name ends in
 *     `_dev`; the library copies during the call; outputs go to caller
 *     buffers;
 *   - every function returns 0 on success, a KB_E* code otherwise; the
 *     message is available from kb_last_error(h) (petsc4py would raise on a
 *     non-zero PetscErrorCode);
 *   - a handle owns one GPU, one CUDA stream and all its device memory;
 *     calls on one handle are not re-entrant, different handles may be driven
 *     from different host threads;
 *   - there is no CPU fallback: kb_create fails with KB_ENODEVICE when no
 *     sm_100 device is visible.
 */
#ifndef KORE_B200_H
#define KORE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kb_context* kb_handle;

enum {
  KB_OK = 0,
  KB_EINVAL = 1,      /* bad argument / call order                                   */
  KB_ENODEVICE = 2,   /* no usable GPU                                                */
  KB_ECUDA = 3,       /* CUDA runtime error                                           */
  KB_ENOMEM = 4,      /* device or host allocation failed                             */
  KB_ESINGULAR = 5,   /* zero / non-finite pivot (MUMPS would report INFOG(1) = -10)  */
  KB_ESTRUCTURE = 6,  /* (perm,nodeptr) does not make the pencil block tridiagonal    */
  KB_ENCCL = 7        /* NCCL error on the l-sharded path                             */
};

/* SLEPc.EPS.Which as selected at bin/solve.py:99-117 (par.which_eigenpairs). */
enum {
  KB_WHICH_LM = 0, KB_WHICH_SM = 1, KB_WHICH_LR = 2, KB_WHICH_SR = 3, KB_WHICH_LI = 4,
  KB_WHICH_SI = 5, KB_WHICH_TM = 6, KB_WHICH_TR = 7, KB_WHICH_TI = 8
};

/* Integer options for kb_set_option. */
enum {
  KB_OPT_EQUILIBRATE = 1, /* 0/1: power-of-two row+column scaling of A - sigma B (default 1)   */
  KB_OPT_REFINE = 2,      /* max iterative-refinement steps per solve (default 1)             */
  KB_OPT_PURIFY = 3,      /* 0/1: x <- OP x on extracted eigenvectors (SLEPc EPSSetPurify)    */
  KB_OPT_SEED = 4,        /* seed of the random Arnoldi start vector when v0 == NULL          */
  KB_OPT_PANEL = 5,       /* Gauss-Jordan panel width override (0 = automatic)                */
  KB_OPT_REFINE_EIGS = 6, /* refinement steps per operator application inside kb_eigs; -1 (default):
                             automatic -- the start vector's solve is refined once and, when that
                             correction exceeds 1e-10 of the solution (kb_stats.refine_resid), every
                             application of the run gets one step (ill-conditioned pencils: the
                             reference-assembled E = 1e-8 one; the Ritz residuals otherwise stall
                             above 1e-10), else none                                             */
  KB_OPT_SWEEP = 7,       /* chain-sweep kernels: 1 persistent two-sided (default), 2 persistent
                             one-sided with grid barriers, 0 one kernel pair per node (no device-side
                             waits).  Takes effect at the next kb_factor (the factors are laid out
                             for the sweep that will read them).                                 */
  KB_OPT_FACTOR = 8,      /* 1 persistent strip factorisation (default), 0 per-step kernels.
                             Next kb_factor.                                                     */
  KB_OPT_FOLD = 9,        /* 0/1: folded couplings for the persistent sweep (default 1; costs
                             2 x the factor memory).  Next kb_factor.                            */
  KB_OPT_WAIT_MS = 10,    /* time bound, in ms, of every device-side wait of the persistent
                             kernels (default 4000).  When a wait expires the launch drains, the
                             handle falls back to the kernels of KB_OPT_SWEEP = 0 / KB_OPT_FACTOR
                             = 0, refactors and repeats the call (kb_stats.protocol_fallbacks).   */
  KB_OPT_INJECT_FAULT = 11 /* tests: 1 = the next chain sweep, 2 = the next factorisation behaves
                             as if a device-side wait had expired; 3 = the next folded sweep
                             publishes unreadable tags, so that its waits really expire           */
};

/* PETSc.Sys / slepc4py.init (solve.py:15-16, 31-34): create a solver context
 * bound to CUDA device `device`. */
int kb_create(kb_handle* h, int device);

/* MA.destroy / MB.destroy / K.destroy (solve.py:201-202, 236-239). */
int kb_destroy(kb_handle h);

const char* kb_last_error(kb_handle h);

int kb_set_option(kb_handle h, int option, int64_t value);

/* Mat create/setValuesCSR/assembly for A and B (solve.py:43-59, 69-85).
 * n x n CSR, int32 or int64 indices selected by index_bytes (4 or 8).
 * A values are complex128.  B may be NULL (forced problem, solve.py:209-227);
 * b_is_complex selects float64 (0, what assemble.py writes) or complex128 (1). */
int kb_set_pencil(kb_handle h, int64_t n, int index_bytes,
                  const void* a_indptr, const void* a_indices, const double* a_values,
                  const void* b_indptr, const void* b_indices, const void* b_values,
                  int b_is_complex);

/* Replaces the fill-reducing ordering / analysis phase of MUMPS (PCSetUp(LU),
 * inside E.solve() solve.py:123 and K.solve solve.py:227): the caller supplies
 * the l-major chain ordering.  perm[k] = original index of chain position k;
 * node p owns chain rows nodeptr[p] .. nodeptr[p+1]-1; nnodes = P.
 * Fails with KB_ESTRUCTURE if a nonzero couples nodes further than 1 apart. */
int kb_set_chain(kb_handle h, const int64_t* perm, const int64_t* nodeptr, int64_t nnodes);

/* l-sharding across the ranks of one node: this rank owns the
 * contiguous node range of rank `rank` out of `nranks`; nccl_unique_id is the
 * 128-byte ncclUniqueId created by rank 0 (kb_nccl_unique_id) and distributed
 * by the host (torch.distributed).  Replaces PETSc's MPI row ownership
 * (solve.py:50, 76) and MUMPS' distributed fronts.  Call after kb_set_chain,
 * before kb_factor. */
int kb_nccl_unique_id(void* id128);
int kb_set_sharding(kb_handle h, int rank, int nranks, const void* nccl_unique_id);

/* STSetUp + KSPSetUp + PCSetUp(LU) numeric factorisation (inside E.solve(),
 * solve.py:123; K.solve, solve.py:227): T = A - sigma B in chain layout,
 * block LU with explicit inverses of the Schur blocks. sigma = {re, im}. */
int kb_factor(kb_handle h, const double* sigma);

/* KSP preonly + PC lu application (K.solve(bvec,x), solve.py:227; ST apply
 * inside EPSSolve): x = T^{-1} rhs for nrhs right-hand sides, column-major
 * n x nrhs complex128, original (Kore) ordering. */
int kb_solve(kb_handle h, const double* rhs, double* x, int nrhs);

/* y = (A - sigma B)^{-1} B x : one application of the Krylov operator
 * (MatMult + KSPSolve inside EPSSolve).  Exposed for tests and baselines. */
int kb_apply_op(kb_handle h, const double* x, double* y);

/* y = A x (which=0) or y = B x (which=1), original ordering (MatMult). */
int kb_matvec(kb_handle h, int which, const double* x, double* y);

/* E.setDimensions / setTolerances / setWhichEigenpairs / setTarget / solve /
 * getConverged / getEigenpair (solve.py:91-149).  Krylov-Schur on
 * (A - sigma B)^{-1} B with the shift of the last kb_factor; eigenvalues are
 * back-transformed lambda = sigma + 1/theta and sorted by `which` relative to
 * `target`.  ncv = 0 selects SLEPc's default max(2 nev, nev + 15).
 * v0 may be NULL (seeded random).  On return *nconv pairs (possibly > nev, as
 * EPSGetConverged, but <= max_pairs) are stored: evals[2*i], evecs column i
 * (n complex128, unit 2-norm, original ordering), resid[i] =
 * ||A x - lambda B x|| / (|lambda| ||B x||).  Non-convergence is not an error:
 * *nconv < nev is returned (solve.py:140-199). */
int kb_eigs(kb_handle h, int nev, int ncv, double tol, int maxit, int which,
            const double* target, int true_residual, const double* v0,
            int max_pairs, double* evals, double* evecs, int* nconv, int* its,
            double* resid);

/* Statistics of the last kb_factor / kb_solve / kb_eigs, for timing.dat-style
 * reporting and the roofline: all times are CUDA-event milliseconds on the
 * handle's stream. */
typedef struct {
  double factor_ms;       /* last kb_factor                                   */
  double solve_ms;        /* last chain solve (fwd+bwd, all refinement steps) */
  double eigs_ms;         /* last kb_eigs total                               */
  double eigs_solve_ms;   /* time of chain solves inside last kb_eigs         */
  int64_t op_applies;     /* operator applications in last kb_eigs            */
  int64_t solve_calls;    /* chain solves (incl. refinement) in last kb_eigs  */
  int64_t kernel_launches;/* kernels launched by this handle since creation   */
  int64_t factor_bytes;   /* device bytes held by the factors                 */
  double factor_flops;    /* real flops executed by the last kb_factor        */
  double solve_bytes;     /* algorithmic bytes of one chain solve             */
  double refine_resid;    /* kb_eigs, automatic refinement: |first correction| / |solution| of the
                             start vector's solve (the forward error of an unrefined solve)      */
  int64_t protocol_fallbacks; /* times a persistent kernel timed out and the handle fell back
                                 to the kernels without device-side waits (0 in a healthy run) */
  int64_t wait_error;     /* code | CTA << 8 of the last expired device-side wait (0: none)  */
  int64_t shard_path;     /* last kb_factor: 0 one GPU, 1 l-sharded on the per-node kernels, 2 l-sharded
                             on the strip factorisation + folded sweep (the fast path)           */
} kb_stats;
int kb_get_stats(kb_handle h, kb_stats* out);

/* Device-resident entry points for the benchmark's HBM-resident leg: vectors
 * are device pointers (n complex128) in ORIGINAL ordering. */
int kb_solve_dev(kb_handle h, const double* rhs_dev, double* x_dev, int nrhs);
int kb_stream(kb_handle h, void** cuda_stream);

/* np.savetxt(f, X) with its defaults, as the field and eigenvalue writers use it
 * (solve.py:275-311: '%.18e', one space, '\n'): element (i, j) = data[i*row_stride +
 * j*col_stride] (strides in doubles, so the real or imaginary part of a column-major
 * complex block is written without a copy).  Host code, `nthreads` formatting threads
 * (0 = all); the bytes are np.savetxt's.  append != 0 opens the file in append mode
 * (timing.dat, solve.py:309-311). */
int kb_savetxt(const char* path, const double* data, int64_t rows, int64_t cols, int64_t row_stride,
               int64_t col_stride, int append, int nthreads);

/* Device-side assembly: replaces the host loops of bin/assemble.py:432-1171
 * (B: :432-590, A: :600-1171; block formulas bin/operators.py:22-195, 386-405, 699-775; boundary
 * rows assemble.py:1188-1345); which set-ups have programs is kore_b200/assembly.py's business
 * (hydrodynamic, thermal, anelastic and degree-1 magnetic ones: DESIGN.md section 4b).  An assembly
 * program describes every N1 x N1 block of a matrix whose rows and columns are `nblockrows`
 * blocks of N1 radial coefficients in Kore's own (section-major) ordering:
 *   - radial operators as bands: ops[(k*N1 + i)*(2H+1) + d] = R_k[i][i + d - H] (0 where absent; 2H+1 <= 1023);
 *   - block row r has the blocks blk_ptr[r] .. blk_ptr[r+1]-1, block b in block column blk_col[b]
 *     (ascending inside a block row); its first br_chop[r] rows are boundary rows: dense rows
 *     bc[(br_bc[r] + q)*N1 + j] of the diagonal block, nothing elsewhere;
 *   - block b is the sum of the groups blk_grp[b] .. blk_grp[b+1]-1; group g adds
 *       grp_sign[g] * s_{k-1} * (... s_0 * ((c_0 x_0 + c_1 x_1) + c_2 x_2 ...))
 *     to the real (grp_part[g] = 0) or imaginary (1) part, with the scalar factors
 *     s_j = grp_sc[4*g + j], j < grp_nsc[g] <= 4, and the terms t = grp_term[g] .. grp_term[g+1]-1:
 *     c = term_coef[t], x = the entry of operator term_op[t].  Every product and sum is one IEEE
 *     double operation in exactly this order (no fused multiply-add): the sequence scipy.sparse
 *     performs for the reference's expressions, so the result equals the reference's to the bit;
 *   - when use_final is set every entry is finally multiplied by final_scale (1 / ||B||_F,
 *     assemble.py:583-585, 1163-1164).
 * Entries that evaluate to exactly zero are dropped (utils.py:164).  kb_assemble makes the result
 * the handle's pencil, as kb_set_pencil does with host CSR arrays: A (complex128) and / or B
 * (float64 when B->is_complex == 0) stay on the device in CSR with int32 indices; call
 * kb_set_chain next.  Either program may be NULL: the pencil then consists of the other matrix
 * alone (forced problems have no B; B alone is assembled first to get its norm).
 * kb_get_assembled copies the CSR of A (which = 0) or B (1) to the host: *nnz always; indptr
 * (n + 1 int64), indices (nnz int32) and values (nnz complex128 or float64) when not NULL. */
typedef struct {
  int32_t N1, nblockrows, H, is_complex;
  int32_t nop, nbc, nblk, ngrp, nterm, use_final;
  double final_scale;
  const double* ops;
  const double* bc;
  const int32_t* br_chop;
  const int32_t* br_bc;
  const int32_t* blk_ptr;
  const int32_t* blk_col;
  const int32_t* blk_grp;
  const int32_t* grp_part;
  const int32_t* grp_sign;
  const int32_t* grp_nsc;
  const double* grp_sc;
  const int32_t* grp_term;
  const double* term_coef;
  const int32_t* term_op;
} kb_asm_program;
int kb_assemble(kb_handle h, const kb_asm_program* A, const kb_asm_program* B);
int kb_get_assembled(kb_handle h, int which, int64_t* nnz, int64_t* indptr, int32_t* indices, double* values);

/* Post-processing diagnostics: the per-degree volume integrals
 * bin/utils4pp.py:794-858 (`diagnose`: flow_worker :426-488, thermal_worker :536-561) computes with a
 * multiprocessing pool for bin/spin_doctor.py:119-147, for hydrodynamic and Boussinesq thermal
 * solutions.  x: nsol solution vectors (complex128, Kore ordering [u | v | h], sizmat each, one after
 * the other: the columns of the eigenvector block kb_eigs returns).  nodes: 3 x N doubles -- the
 * Chebyshev-Gauss nodes mapped into the solution's Chebyshev domain, the radii r_k in [Ra, Rb], and the
 * weights (pi / N) sqrt(1 - x_k^2) (Rb - Ra) / 2 (utils4pp.py:54-64, 806-826).  Outputs, per solution:
 * flow[nll][6] for the nll = lmax - m + 1 degrees in ascending order = {kinetic energy, kinetic
 * dissipation, internal dissipation, 0 (Lorentz), buoyancy power, 0 (compositional)} -- the columns of
 * the reference's `udgn`; thermal[nb][3] for the poloidal degrees = {thermal energy, dissipation,
 * advection} (`tdgn`; may be NULL when thermal == 0).  heating: 0 differential, 1 internal
 * (utils4pp.py:404-411).  The host side (kore_b200/diagnostics.py) gets the rest out of this one entry
 * point: the compositional field of double-diffusive runs is a second call on [u | v | c] (the reference
 * uses the thermal worker for it, utils4pp.py:846-850), a background gradient of the run's own
 * ('two zone' / 'user defined' heating) a second call with the weights scaled by fr / r^2 under
 * heating = 1, the magnetic energy and diffusion the flow pass on [f | g] with the symmetry of b. */
typedef struct {
  int32_t N, N1, nb, m, lmax, symm, thermal, heating;
  double ricb, rcmb;
} kb_diag_params;
int kb_diagnose(kb_handle h, const kb_diag_params* p, const double* nodes, const double* x, int nsol,
                double* flow, double* thermal);

/* Debug / test hooks (not part of the drop-in surface).  kb_dbg_schur: the host-side complex
 * Schur form + ordering of the projected problem (what SLEPc's DS does with LAPACK), m x m
 * column-major complex128 in, T and Q out; which < 0: no ordering.  kb_dbg_*_timing: in-kernel
 * cycle counters of the last persistent factorisation / sweep (library built with
 * -DKB_FACTOR_TIMING / -DKB_SWEEP_TICKS, KB_SWEEP_TIMING=1 in the environment). */
int kb_dbg_schur(int m, const double* H, int which, const double* sigma, const double* tau,
                 double* T, double* Q);
int kb_dbg_factor_timing(kb_handle h, long long* out, int max_ctas);
int kb_dbg_sweep_timing(kb_handle h, long long* out, int max_ctas);
/* C = alpha op(A) B + beta C on host arrays through the batched complex128 product kernel of the
 * l-sharded path (row-major; transA: A is k x m); `batch` identical problems per launch, `reps`
 * timed launches, mean ms per launch in *ms (tools/dev_zgemm.py). */
/* The elimination rank `rank` of `nranks` would do on this pencil, on this one GPU without a communicator,
 * and the four blocks (4 x bmax x bmax complex128) it would contribute to the reduced separator system;
 * path 2: strip factorisation + folded couplings + DMMA product chains, 1: per-node kernels. */
int kb_dbg_shard_segment(kb_handle h, int rank, int nranks, const double* sigma, int path, double* blocks);
int kb_dbg_zgemm(kb_handle h, int m, int n, int k, int transA, const double* A, const double* B, double* C,
                 double alpha, double beta, int batch, int reps, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* KORE_B200_H */
