"""CPU oracle for Kore's shift-and-invert eigen / forced solve path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kore_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker (or as the timed CPU baseline), never as the product path.

What it restates
----------------
The reference's hot path is the block ``/root/reference/bin/solve.py:91-123``
(``SLEPc.EPS`` GNHEP + ``ST`` sinvert + ``KSP preonly``/``PC lu`` via MUMPS)
and ``solve.py:209-233`` (forced problem: ``KSP`` preonly + LU).  The
arithmetic itself lives in un-vendored third-party libraries that are ABSENT
from ``/root/reference`` and from this image: PETSc 3.12.5 / SLEPc 3.12.2 /
petsc4py+slepc4py 3.12.0 / MUMPS or SuperLU_DIST 5.4.0 (README.md:31-34,
65-70, 92, 104-105; CI uses distro 3.19 builds, .github/workflows/main.yml:
37-46).  Their published algorithm is restated with SciPy's bundled serial
SuperLU (``splu``) and ARPACK (``eigs``) on the EXPLICIT shift-invert operator

    x  ->  (A - sigma B)^{-1} (B x),        lambda = sigma + 1/theta

(``scipy.sparse.linalg.eigs(A, M=B, sigma=...)`` is NOT used: it returns wrong
pairs on this non-Hermitian pencil with singular B.)

Parity pinning
--------------
Pinned against every golden the reference's own tests hold for this path
(see tests/test_oracle_golden.py):
  * tests/spinover/reference.eig:1          (eigenvalue, rtol 1e-8)
  * tests/dormy2004/reference.dormy04:1     (Ra_c -> Re(lambda)=0, omega_c 5 digits)
  * tests/jones2000/reference.jones:1       (same, full sphere)
on matrices produced by the UNMODIFIED reference assembler
(tools/make_case.py).  Eigenvectors, magnetic and forced runs have no golden
in the reference; for those this oracle is "parity vs SciPy-SuperLU;
SLEPc+MUMPS parity unpinned (dependency absent)".
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as ss
import scipy.sparse.linalg as ssl

# solve.py:99-117 -- which_eigenpairs strings (SLEPc.EPS.Which)
WHICH = ("LM", "SM", "LR", "SR", "LI", "SI", "TM", "TR", "TI")


def load_csr(filename):
    """bin/utils.py:1253-1256."""
    z = np.load(filename)
    return ss.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))


def default_ncv(nev: int) -> int:
    """SLEPc default when only nev is set (solve.py:95-96): max(2 nev, nev+15)."""
    return max(2 * nev, nev + 15)


def which_key(lam: np.ndarray, which: str, tau: complex) -> np.ndarray:
    """Sort key (ascending = most wanted first) on BACK-TRANSFORMED eigenvalues,
    as SLEPc compares them for each EPS.Which (solve.py:99-117; SLEPc manual
    'Selection of eigenvalues')."""
    lam = np.asarray(lam)
    if which == "LM":
        return -np.abs(lam)
    if which == "SM":
        return np.abs(lam)
    if which == "LR":
        return -lam.real
    if which == "SR":
        return lam.real
    if which == "LI":
        return -lam.imag
    if which == "SI":
        return lam.imag
    if which == "TM":
        return np.abs(lam - tau)
    if which == "TR":
        return np.abs((lam - tau).real)
    if which == "TI":
        return np.abs((lam - tau).imag)
    raise ValueError("unknown which_eigenpairs %r" % which)


class ShiftInvert:
    """ST sinvert + KSP preonly + PC lu  (inside E.solve(), solve.py:123)."""

    def __init__(self, A, B, sigma):
        self.A = A.tocsr()
        self.B = B.tocsr() if B is not None else None
        self.sigma = complex(sigma)
        T = self.A.astype(np.complex128)
        if self.B is not None and self.sigma != 0:
            T = T - self.sigma * self.B
        self.T = T.tocsc()
        self.lu = ssl.splu(self.T)
        self.napply = 0

    def solve(self, rhs):
        return self.lu.solve(np.asarray(rhs, dtype=np.complex128))

    def apply(self, x):
        self.napply += 1
        return self.lu.solve(self.B @ x)


class BlockShiftInvert:
    """The same operator as :class:`ShiftInvert` with the LU done the way a multifrontal code
    (MUMPS, what the reference's scripts select: tests/dormy2004/find_Rac.py:18,
    tools/subramp.sh:50) does it on this structure: dense LAPACK (``zgetrf`` / ``zgetrs``,
    threaded BLAS-3) on the fronts of the l-chain instead of SciPy's serial scalar-supernode
    SuperLU.  In chain order T = A - sigma B is block tridiagonal; the elimination is

        S_0 = D_0,   S_p = D_p - L_p (S_{p-1}^{-1} U_{p-1}),      LU(S_p) kept,
        forward  y_p = r_p - L_p x~_{p-1},  x~_p = S_p^{-1} y_p,
        backward x_p = x~_p - S_p^{-1} (U_p x_{p+1}),

    pivoting inside the fronts only (as accurate as SuperLU on every reference
    matrix).  It is the CPU baseline bench.py times on the full benchmark size with all host
    threads (``cpu_baseline.kind = "port"``) and is checked against :class:`ShiftInvert` in
    tests/test_oracle_golden.py.  ``chain`` = (perm, nodeptr) as kb_set_chain takes them."""

    def __init__(self, A, B, sigma, chain, max_nodes=None):
        import scipy.linalg as sl
        self._sl = sl
        perm, nodeptr = chain
        self.perm = np.asarray(perm)
        self.nodeptr = np.asarray(nodeptr)
        self.A = A.tocsr()
        self.B = B.tocsr() if B is not None else None
        self.sigma = complex(sigma)
        T = self.A.astype(np.complex128)
        if self.B is not None and self.sigma != 0:
            T = T - self.sigma * self.B
        Tc = T.tocsr()[self.perm][:, self.perm].tocsr()
        P = len(self.nodeptr) - 1
        self.P = P if max_nodes is None else min(P, int(max_nodes))
        self.lu, self.Lc, self.Uc = [], [], []
        W = None
        for p in range(self.P):
            o0, o1 = self.nodeptr[p], self.nodeptr[p + 1]
            rows = Tc[o0:o1]
            S = rows[:, o0:o1].toarray()
            Lp = rows[:, self.nodeptr[p - 1]:o0].tocsr() if p > 0 else None
            Up = rows[:, o1:self.nodeptr[p + 2]].tocsr() if p + 1 < P else None
            if p > 0:
                S -= Lp @ W
            lu = sl.lu_factor(S, overwrite_a=True, check_finite=False)
            if Up is not None and p + 1 < self.P:
                W = sl.lu_solve(lu, Up.toarray(), overwrite_b=True, check_finite=False)
            self.lu.append(lu)
            self.Lc.append(Lp)
            self.Uc.append(Up)
        self.napply = 0

    def solve(self, rhs):
        sl, nodeptr = self._sl, self.nodeptr
        r = np.asarray(rhs, dtype=np.complex128)[self.perm]
        x = np.empty_like(r)
        prev = None
        for p in range(self.P):
            o0, o1 = nodeptr[p], nodeptr[p + 1]
            y = r[o0:o1] if p == 0 else r[o0:o1] - self.Lc[p] @ prev
            prev = sl.lu_solve(self.lu[p], y, check_finite=False)
            x[o0:o1] = prev
        for p in range(self.P - 2, -1, -1):
            o0, o1 = nodeptr[p], nodeptr[p + 1]
            x[o0:o1] -= sl.lu_solve(self.lu[p], self.Uc[p] @ x[o1:nodeptr[p + 2]], check_finite=False)
        out = np.empty_like(x)
        out[self.perm] = x
        return out

    def apply(self, x):
        self.napply += 1
        return self.solve(self.B @ x)


def residuals(A, B, lam, X):
    """BASELINE.json criterion: ||A x - lam B x|| / (|lam| ||B x||)."""
    out = np.empty(len(lam))
    for i, l in enumerate(lam):
        x = X[:, i]
        bx = B @ x
        out[i] = np.linalg.norm(A @ x - l * bx) / (abs(l) * np.linalg.norm(bx))
    return out


def eigs(A, B, tau, nev, which="TM", ncv=None, tol=0.0, maxiter=None, v0=None,
         nsearch=None, op=None):
    """Restatement of solve.py:91-149.

    For the target-magnitude family ARPACK's 'LM' on theta is the same
    ordering as SLEPc's TARGET_MAGNITUDE on lambda.  For the other ``which``
    values the reference sorts back-transformed eigenvalues inside the
    Krylov-Schur restart; ARPACK cannot, so the oracle computes ``nsearch``
    (default 40) pairs nearest the shift and then selects ``nev`` by
    ``which_key`` -- SURVEY.md App. E probe: test_dormy's TR pair is the
    largest-real-part member of the 40 nearest.

    Returns (lam[nev], X[n,nev] unit 2-norm, info dict).
    """
    op = op or ShiftInvert(A, B, tau)
    n = A.shape[0]
    if which == "TM":
        k = nev
    else:
        k = max(nev, nsearch or 40)
    if ncv is None:
        ncv = default_ncv(k)
    if v0 is None:
        rng = np.random.default_rng(1)
        v0 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    OP = ssl.LinearOperator((n, n), matvec=op.apply, dtype=np.complex128)
    theta, V = ssl.eigs(OP, k=k, which="LM", ncv=ncv, tol=tol, maxiter=maxiter, v0=v0)
    lam = op.sigma + 1.0 / theta
    order = np.argsort(which_key(lam, which, op.sigma), kind="stable")[:nev]
    lam = lam[order]
    V = V[:, order]
    V = V / np.linalg.norm(V, axis=0)
    info = dict(napply=op.napply, ncv=ncv, nsearch=k)
    return lam, V, info


def forced_solve(A, b):
    """solve.py:209-233: KSP preonly + LU on A x = b, b = B_forced.npz column."""
    lu = ssl.splu(A.tocsc().astype(np.complex128))
    b = np.asarray(b.todense()).ravel() if ss.issparse(b) else np.asarray(b).ravel()
    return lu.solve(b.astype(np.complex128))


def split_fields(vec, n, hydro, magnetic, thermal, compositional):
    """solve.py:163-190 / 243-261: row slices of the solution by field."""
    vec = np.asarray(vec)
    if vec.ndim == 1:
        vec = vec.reshape(-1, 1)
    out = {}
    if hydro == 1:
        out["flow"] = vec[: 2 * n, :]
    if magnetic == 1:
        o = 2 * n * hydro
        out["magnetic"] = vec[o: o + 2 * n, :]
    if thermal == 1:
        o = 2 * n * hydro + magnetic * 2 * n
        out["temperature"] = vec[o: o + n, :]
    if compositional == 1:
        o = 2 * n * hydro + magnetic * 2 * n + thermal * n
        out["composition"] = vec[o: o + n, :]
    return out
