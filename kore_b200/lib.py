"""ctypes binding of libkoreb200.so (the C ABI declared in include/kore_b200.h).

This is the thin host layer the north-star asks for: Python calls a C-ABI
library; numpy owns the host buffers; there is NO CPU fallback -- if the
shared library is missing or no B200 is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("KB_LIB_PATH", os.path.join(_HERE, "libkoreb200.so"))

KB_OK, KB_EINVAL, KB_ENODEVICE, KB_ECUDA, KB_ENOMEM, KB_ESINGULAR, KB_ESTRUCTURE, KB_ENCCL = range(8)
ERRNAMES = ["KB_OK", "KB_EINVAL", "KB_ENODEVICE", "KB_ECUDA", "KB_ENOMEM", "KB_ESINGULAR",
            "KB_ESTRUCTURE", "KB_ENCCL"]

# SLEPc.EPS.Which as used at bin/solve.py:99-117
WHICH = {"LM": 0, "SM": 1, "LR": 2, "SR": 3, "LI": 4, "SI": 5, "TM": 6, "TR": 7, "TI": 8}

(OPT_EQUILIBRATE, OPT_REFINE, OPT_PURIFY, OPT_SEED, OPT_PANEL, OPT_REFINE_EIGS, OPT_SWEEP, OPT_FACTOR,
 OPT_FOLD, OPT_WAIT_MS, OPT_INJECT_FAULT) = range(1, 12)

# every symbol include/kore_b200.h declares
EXPORTS = [
    "kb_create", "kb_destroy", "kb_last_error", "kb_set_option", "kb_set_pencil", "kb_set_chain",
    "kb_nccl_unique_id", "kb_set_sharding", "kb_factor", "kb_solve", "kb_apply_op", "kb_matvec",
    "kb_eigs", "kb_get_stats", "kb_solve_dev", "kb_stream", "kb_savetxt", "kb_assemble", "kb_get_assembled", "kb_diagnose",
    "kb_dbg_schur", "kb_dbg_factor_timing", "kb_dbg_sweep_timing", "kb_dbg_zgemm", "kb_dbg_shard_segment",
]


class KoreB200Error(RuntimeError):
    def __init__(self, code, msg):
        self.code = code
        name = ERRNAMES[code] if 0 <= code < len(ERRNAMES) else str(code)
        super().__init__("%s: %s" % (name, msg))


class KbStats(C.Structure):
    _fields_ = [
        ("factor_ms", C.c_double), ("solve_ms", C.c_double), ("eigs_ms", C.c_double),
        ("eigs_solve_ms", C.c_double), ("op_applies", C.c_int64), ("solve_calls", C.c_int64),
        ("kernel_launches", C.c_int64), ("factor_bytes", C.c_int64), ("factor_flops", C.c_double),
        ("solve_bytes", C.c_double), ("refine_resid", C.c_double),
        ("protocol_fallbacks", C.c_int64), ("wait_error", C.c_int64), ("shard_path", C.c_int64),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class KbAsmProgram(C.Structure):
    """kb_asm_program of include/kore_b200.h."""
    _fields_ = [
        ("N1", C.c_int32), ("nblockrows", C.c_int32), ("H", C.c_int32), ("is_complex", C.c_int32),
        ("nop", C.c_int32), ("nbc", C.c_int32), ("nblk", C.c_int32), ("ngrp", C.c_int32),
        ("nterm", C.c_int32), ("use_final", C.c_int32), ("final_scale", C.c_double),
        ("ops", C.c_void_p), ("bc", C.c_void_p), ("br_chop", C.c_void_p), ("br_bc", C.c_void_p),
        ("blk_ptr", C.c_void_p), ("blk_col", C.c_void_p), ("blk_grp", C.c_void_p),
        ("grp_part", C.c_void_p), ("grp_sign", C.c_void_p), ("grp_nsc", C.c_void_p),
        ("grp_sc", C.c_void_p), ("grp_term", C.c_void_p), ("term_coef", C.c_void_p),
        ("term_op", C.c_void_p),
    ]


class KbDiagParams(C.Structure):
    """kb_diag_params of include/kore_b200.h."""
    _fields_ = [("N", C.c_int32), ("N1", C.c_int32), ("nb", C.c_int32), ("m", C.c_int32), ("lmax", C.c_int32),
                ("symm", C.c_int32), ("thermal", C.c_int32), ("heating", C.c_int32),
                ("ricb", C.c_double), ("rcmb", C.c_double)]


def _asm_struct(prog):
    """(KbAsmProgram, arrays kept alive) of a kore_b200.assembly.AsmProgram."""
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)    # noqa: E731
    keep = dict(ops=f64(prog.ops), bc=f64(prog.bc), br_chop=i32(prog.br_chop), br_bc=i32(prog.br_bc),
                blk_ptr=i32(prog.blk_ptr), blk_col=i32(prog.blk_col), blk_grp=i32(prog.blk_grp),
                grp_part=i32(prog.grp_part), grp_sign=i32(prog.grp_sign), grp_nsc=i32(prog.grp_nsc),
                grp_sc=f64(prog.grp_sc), grp_term=i32(prog.grp_term), term_coef=f64(prog.term_coef),
                term_op=i32(prog.term_op))
    st = KbAsmProgram(N1=prog.N1, nblockrows=prog.nblockrows, H=prog.H, is_complex=int(prog.is_complex),
                      nop=keep["ops"].shape[0], nbc=keep["bc"].shape[0], nblk=len(keep["blk_col"]),
                      ngrp=len(keep["grp_part"]), nterm=len(keep["term_op"]), use_final=int(prog.use_final),
                      final_scale=float(prog.final_scale))
    for k, a in keep.items():
        setattr(st, k, a.ctypes.data)
    return st, keep


def build(verbose=False):
    """Compile libkoreb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", _ROOT, "kore_b200/libkoreb200.so"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("building libkoreb200.so failed")
    return LIB_PATH


_lib = None


def load():
    """dlopen the library and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KoreB200Error(KB_ENODEVICE,
                            "libkoreb200.so is not built (run `make` or __graft_entry__.build()); "
                            "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i64, dp = C.c_void_p, C.c_int64, C.POINTER(C.c_double)
    lib.kb_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.kb_destroy.argtypes = [vp]
    lib.kb_last_error.argtypes = [vp]
    lib.kb_last_error.restype = C.c_char_p
    lib.kb_set_option.argtypes = [vp, C.c_int, i64]
    lib.kb_set_pencil.argtypes = [vp, i64, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
    lib.kb_set_chain.argtypes = [vp, vp, vp, i64]
    lib.kb_nccl_unique_id.argtypes = [vp]
    lib.kb_set_sharding.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.kb_factor.argtypes = [vp, vp]
    lib.kb_solve.argtypes = [vp, vp, vp, C.c_int]
    lib.kb_apply_op.argtypes = [vp, vp, vp]
    lib.kb_matvec.argtypes = [vp, C.c_int, vp, vp]
    lib.kb_eigs.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, vp, C.c_int, vp,
                            C.c_int, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), vp]
    lib.kb_get_stats.argtypes = [vp, C.POINTER(KbStats)]
    lib.kb_solve_dev.argtypes = [vp, vp, vp, C.c_int]
    lib.kb_stream.argtypes = [vp, C.POINTER(vp)]
    lib.kb_savetxt.argtypes = [C.c_char_p, vp, i64, i64, i64, i64, C.c_int, C.c_int]
    lib.kb_dbg_schur.argtypes = [C.c_int, vp, C.c_int, vp, vp, vp, vp]
    lib.kb_assemble.argtypes = [vp, C.POINTER(KbAsmProgram), C.POINTER(KbAsmProgram)]
    lib.kb_get_assembled.argtypes = [vp, C.c_int, C.POINTER(i64), vp, vp, vp]
    lib.kb_diagnose.argtypes = [vp, C.POINTER(KbDiagParams), vp, vp, C.c_int, vp, vp]
    for name in EXPORTS:
        if name != "kb_last_error":
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def savetxt(path, X, part="real", append=False, nthreads=0):
    """``np.savetxt(path, X.real | X.imag | X)`` with np.savetxt's default format, written by the
    library's threaded formatter (kb_savetxt) straight from X's memory (any strides that are
    multiples of 8 bytes; complex128 or float64; 1-D arrays are written as one column, like
    np.savetxt does)."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    if X.ndim != 2:
        raise ValueError("savetxt expects a 1-D or 2-D array")
    if np.iscomplexobj(X):
        X = np.asarray(X, dtype=np.complex128)
        base = X.real if part == "real" else X.imag  # strided float64 views, no copy
    else:
        base = np.asarray(X, dtype=np.float64)
    if any(st % 8 for st in base.strides) or any(st < 0 for st in base.strides):
        base = np.ascontiguousarray(base)
    rows, cols = base.shape
    rs, cs = (base.strides[0] // 8, base.strides[1] // 8) if base.size else (0, 0)
    rc = load().kb_savetxt(os.fsencode(path), C.c_void_p(base.ctypes.data if base.size else 0), rows, cols, rs, cs,
                           int(bool(append)), int(nthreads))
    if rc != KB_OK:
        raise KoreB200Error(rc, "kb_savetxt could not write %r" % (path,))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Solver:
    """One GPU, one pencil (A, B), one chain layout; factor / solve / eigs.

    Mirrors the sequence of calls bin/solve.py makes on PETSc/SLEPc objects:
    Mat assembly -> (analysis) -> numeric factorisation -> solve / EPSSolve."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.kb_create(C.byref(h), int(device))
        if rc != KB_OK:
            raise KoreB200Error(rc, self.lib.kb_last_error(None).decode())
        self.h = h
        self.n = 0
        self._keep = []

    def _check(self, rc):
        if rc != KB_OK:
            raise KoreB200Error(rc, self.lib.kb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.kb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, opt, value):
        self._check(self.lib.kb_set_option(self.h, int(opt), int(value)))

    def set_pencil(self, A, B=None):
        """A, B: scipy.sparse CSR (A complex128; B float64 or complex128 or None)."""
        A = A.tocsr()
        n = A.shape[0]
        idx_dtype = np.int64 if (A.indices.dtype == np.int64 or A.indptr.dtype == np.int64) else np.int32
        ap = np.ascontiguousarray(A.indptr, dtype=idx_dtype)
        ai = np.ascontiguousarray(A.indices, dtype=idx_dtype)
        av = np.ascontiguousarray(A.data, dtype=np.complex128)
        bp = bi = bv = None
        bcomplex = 0
        if B is not None:
            B = B.tocsr()
            if not B.has_canonical_format:
                # the library gives every B entry its own slot of A - sigma B (ut.load_csr /
                # ss.csr_matrix do not merge duplicate entries; kb_set_chain rejects them)
                B = B.copy()
                B.sum_duplicates()
            bp = np.ascontiguousarray(B.indptr, dtype=idx_dtype)
            bi = np.ascontiguousarray(B.indices, dtype=idx_dtype)
            if np.iscomplexobj(B.data):
                bv = np.ascontiguousarray(B.data, dtype=np.complex128)
                bcomplex = 1
            else:
                bv = np.ascontiguousarray(B.data, dtype=np.float64)
        self._check(self.lib.kb_set_pencil(self.h, n, np.dtype(idx_dtype).itemsize, _ptr(ap), _ptr(ai),
                                           _ptr(av), _ptr(bp), _ptr(bi), _ptr(bv), bcomplex))
        self.n = n
        self._b_complex = bool(bcomplex)

    def assemble(self, progA, progB=None):
        """Assemble the pencil on the GPU from assembly programs (kore_b200.assembly) instead of
        ingesting host CSR arrays; either may be None.  Follow with `set_chain`."""
        sa = sb = None
        if progA is not None:
            sa, keep_a = _asm_struct(progA)
        if progB is not None:
            sb, keep_b = _asm_struct(progB)
        self._check(self.lib.kb_assemble(self.h, C.byref(sa) if sa is not None else None,
                                         C.byref(sb) if sb is not None else None))
        self.n = (progA if progA is not None else progB).n
        self._b_complex = bool(progB.is_complex) if progB is not None else False

    def get_assembled(self, which="A"):
        """(indptr int64, indices int32, values) of the A or B held on the device in raw CSR form
        (assembled there by `assemble`, or uploaded by `set_pencil`)."""
        w = 0 if which == "A" else 1
        nnz = C.c_int64(0)
        self._check(self.lib.kb_get_assembled(self.h, w, C.byref(nnz), None, None, None))
        indptr = np.empty(self.n + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int32)
        # values: complex128 for A; float64 or complex128 for B -- sized for the larger, trimmed below
        raw = np.empty(2 * nnz.value, dtype=np.float64)
        self._check(self.lib.kb_get_assembled(self.h, w, C.byref(nnz), _ptr(indptr), _ptr(indices), _ptr(raw)))
        cplx = w == 0 or getattr(self, "_b_complex", False)
        values = raw.view(np.complex128) if cplx else raw[:nnz.value].copy()
        return indptr, indices, values

    def diagnose(self, params, nodes, X):
        """kb_diagnose: per-degree integrals of the solutions in the columns of X (sizmat x nsol).
        `params`: KbDiagParams; `nodes`: (3, N) float64.  Returns (flow[nsol, nll, 6], thermal[nsol, nb, 3])."""
        X = np.asarray(X, dtype=np.complex128)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        X = np.asfortranarray(X)
        nsol = X.shape[1]
        nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        flow = np.zeros((nsol, 2 * params.nb, 6))
        thermal = np.zeros((nsol, params.nb, 3))
        self._check(self.lib.kb_diagnose(self.h, C.byref(params), _ptr(nodes), _ptr(X), nsol, _ptr(flow), _ptr(thermal)))
        return flow, thermal

    def set_chain(self, perm, nodeptr):
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        nodeptr = np.ascontiguousarray(nodeptr, dtype=np.int64)
        self._check(self.lib.kb_set_chain(self.h, _ptr(perm), _ptr(nodeptr), len(nodeptr) - 1))
        self._nodeptr = nodeptr

    def set_sharding(self, rank, nranks, unique_id):
        self._check(self.lib.kb_set_sharding(self.h, int(rank), int(nranks), _ptr(unique_id)))

    def factor(self, sigma):
        s = np.array([complex(sigma)], dtype=np.complex128)
        self._check(self.lib.kb_factor(self.h, _ptr(s)))

    def solve(self, rhs):
        rhs = np.asarray(rhs, dtype=np.complex128)
        one = rhs.ndim == 1
        R = np.asfortranarray(rhs.reshape(self.n, -1))
        X = np.empty_like(R, order="F")
        self._check(self.lib.kb_solve(self.h, _ptr(R), _ptr(X), R.shape[1]))
        return X[:, 0].copy() if one else X

    def apply_op(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        y = np.empty_like(x)
        self._check(self.lib.kb_apply_op(self.h, _ptr(x), _ptr(y)))
        return y

    def matvec(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        y = np.empty_like(x)
        self._check(self.lib.kb_matvec(self.h, 0 if which == "A" else 1, _ptr(x), _ptr(y)))
        return y

    def eigs(self, nev, which="TM", target=None, ncv=0, tol=1e-15, maxit=50, true_residual=False,
             v0=None, max_pairs=None, want_vectors=True):
        """Returns (eigenvalues[nconv], eigenvectors[n, nconv], info)."""
        if ncv <= 0:
            ncv = max(2 * nev, nev + 15)
        max_pairs = max_pairs or ncv
        evals = np.zeros(max_pairs, dtype=np.complex128)
        evecs = None
        if want_vectors:
            # the receive buffer is kept on the Solver: a fresh 16 n ncv-byte allocation per call
            # costs more in page faults than the device-to-host copies themselves
            ev = getattr(self, "_evecs", None)
            if ev is None or ev.shape != (self.n, max_pairs):
                ev = np.empty((self.n, max_pairs), dtype=np.complex128, order="F")
                self._evecs = ev
            evecs = ev
        resid = np.zeros(max_pairs, dtype=np.float64)
        nconv = C.c_int(0)
        its = C.c_int(0)
        tgt = None if target is None else np.array([complex(target)], dtype=np.complex128)
        if v0 is not None:
            v0 = np.ascontiguousarray(v0, dtype=np.complex128)
        self._check(self.lib.kb_eigs(self.h, int(nev), int(ncv), float(tol), int(maxit), WHICH[which],
                                     _ptr(tgt), int(bool(true_residual)), _ptr(v0), int(max_pairs),
                                     _ptr(evals), _ptr(evecs), C.byref(nconv), C.byref(its), _ptr(resid)))
        k = nconv.value
        info = dict(nconv=k, its=its.value, ncv=ncv, resid=resid[:k].copy())
        info.update(self.stats())
        return evals[:k].copy(), (np.array(evecs[:, :k], order="F") if want_vectors else None), info

    def dbg_shard_segment(self, rank, nranks, sigma, path="fast"):
        """Test hook: the four blocks rank `rank` of `nranks` would contribute to the reduced
        (separator) system of the l-sharded factorisation, computed on this one GPU without a
        communicator (kb_dbg_shard_segment).  Returns an array (4, bmax * bmax) complex128."""
        bmax = int(np.diff(self._nodeptr).max())
        out = np.zeros((4, bmax * bmax), dtype=np.complex128)
        sg = np.array([complex(sigma)], dtype=np.complex128)
        f = self.lib.kb_dbg_shard_segment
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        self._check(f(self.h, int(rank), int(nranks), _ptr(sg), 2 if path == "fast" else 1, _ptr(out)))
        return out

    def stats(self):
        st = KbStats()
        self._check(self.lib.kb_get_stats(self.h, C.byref(st)))
        return st.asdict()

    def solve_dev(self, rhs_ptr, x_ptr, nrhs=1):
        """Device-pointer variant (torch tensors' data_ptr())."""
        self._check(self.lib.kb_solve_dev(self.h, C.c_void_p(rhs_ptr), C.c_void_p(x_ptr), int(nrhs)))

    def stream(self):
        s = C.c_void_p()
        self._check(self.lib.kb_stream(self.h, C.byref(s)))
        return s.value
