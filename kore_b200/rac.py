"""Critical-Rayleigh-number search (BASELINE.json config 2).

The reference finds the onset of convection with an outer script
(/root/reference/tests/dormy2004/find_Rac.py:36-78, 93-113): for every trial
Rayleigh number it rewrites ``Ra_gap`` in parameters.py with sed, re-runs
``assemble.py`` and ``solve.py`` (one sparse LU + one Krylov-Schur each) in fresh
MPI jobs, reads ``eigenvalues0.dat`` back and feeds the largest growth rate to a
bracket search followed by Brent's method in ``log10(Ra)``.

Here the same search runs inside one process on one handle.  Only the buoyancy
term of the momentum equation depends on the Rayleigh number and it does so
linearly (operators.py:405, ``par.Beyonce * out``), so ``A(Ra) = A(Ra_0) +
(Ra - Ra_0) A_1`` with ``A_1 = dA/dRa`` on a subset of A's pattern: the pencil is
assembled once (twice, to difference out ``A_1``: `affine_from_two`), every
evaluation updates the values on the fixed pattern, re-ingests, factors and runs
Krylov-Schur on the GPU.  Evaluations are cached by Ra like the reference's
``ra_cache``.

The search itself (`bracket_brentq`) follows find_Rac.py:57-78: start at x1, walk
in steps of ``dx`` towards the sign change (down if the first growth rate is
positive), then ``scipy.optimize.brentq`` with ``xtol = rtol = tol``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import brentq

from . import lib as _lib


def bracket_brentq(f, x1, x2=None, dx=0.01, tol=1e-6, maxiter=200, args=(), max_walk=1000):
    """Root of ``f`` near ``x1``: bracket by walking, then Brent (find_Rac.py:57-78).

    ``max_walk`` bounds the bracket walk (the reference loops forever when there is no sign
    change); exceeding it raises RuntimeError."""
    y1 = f(x1, *args)
    dx = abs(dx)
    if x2 is not None:
        y2 = f(x2, *args)
        if y2 * y1 > 0:  # not a bracket: restart the walk from the end closest to the root
            if abs(y2) < abs(y1):
                x1, y1 = x2, y2
            x2 = None
    if x2 is None:
        x2 = x1
        if y1 > 0:
            dx = -dx
        for _ in range(max_walk):
            x2 += dx
            y2 = f(x2, *args)
            if y2 * y1 < 0:
                break
            x1, y1 = x2, y2
        else:
            raise RuntimeError("bracket_brentq: no sign change within %d steps of %g" % (max_walk, dx))
    return brentq(f, x1, x2, maxiter=maxiter, xtol=tol, rtol=tol, args=args)


def affine_from_two(A_lo, Ra_lo, A_hi, Ra_hi):
    """``A_1 = (A_hi - A_lo) / (Ra_hi - Ra_lo)`` as a CSR on A's own pattern positions.

    Both matrices come from the same assembler run at two Rayleigh numbers and therefore share
    indptr / indices exactly; returns (positions into A.data, values) of the nonzero slope."""
    A_lo, A_hi = A_lo.tocsr(), A_hi.tocsr()
    if not (np.array_equal(A_lo.indptr, A_hi.indptr) and np.array_equal(A_lo.indices, A_hi.indices)):
        raise ValueError("the two assemblies do not share a sparsity pattern")
    d = (A_hi.data - A_lo.data) / (Ra_hi - Ra_lo)
    pos = np.flatnonzero(d)
    return pos.astype(np.int64), d[pos].astype(np.complex128)


def slope_positions(A, A1):
    """Positions in ``A.data`` (CSR, sorted or not) of the entries of the sparse slope ``A1``
    (pattern(A1) must be a subset of pattern(A)) and the slope values in that order."""
    A = A.tocsr()
    C = A1.tocoo()
    n = A.shape[1]
    rows = np.repeat(np.arange(A.shape[0], dtype=np.int64), np.diff(A.indptr))
    keyA = rows * n + A.indices
    order = np.argsort(keyA, kind="stable")
    keyC = C.row.astype(np.int64) * n + C.col
    at = np.searchsorted(keyA[order], keyC)
    if np.any(at >= len(order)) or np.any(keyA[order][np.minimum(at, len(order) - 1)] != keyC):
        raise ValueError("pattern(A1) is not contained in pattern(A)")
    return order[at].astype(np.int64), C.data.astype(np.complex128)


class RayleighPencil:
    """``A(Ra) = A_ref + (Ra - Ra_ref) A_1`` on the fixed pattern of ``A_ref``; B constant."""

    def __init__(self, A_ref, Ra_ref, slope_pos, slope_val, B):
        self.A = sp.csr_matrix(A_ref, dtype=np.complex128, copy=True)
        self.base = self.A.data.copy()
        self.Ra_ref = float(Ra_ref)
        self.pos = np.asarray(slope_pos, dtype=np.int64)
        self.val = np.asarray(slope_val, dtype=np.complex128)
        self.B = B

    def at(self, Ra):
        """The CSR of A(Ra) (values updated in place on the shared pattern)."""
        self.A.data[:] = self.base
        self.A.data[self.pos] += (float(Ra) - self.Ra_ref) * self.val
        return self.A

    def install(self, solver, Ra):
        """Make (A(Ra), B) the pencil of `solver`: host update + re-ingest of the CSR."""
        solver.set_pencil(self.at(Ra), self.B)


def buoyancy_factor(Ra_gap, Ek, ricb, Prandtl=1):
    """``par.Beyonce`` of a parameters.py that sets the Rayleigh number through ``Ra_gap``
    (tests/dormy2004/params.dormy04:177-186: Ra = Ra_gap / (1 - ricb)^3, BV2 = -Ra Ek^2 / Prandtl,
    Beyonce = BV2), operation for operation, so that a trial of the search assembles the matrix the
    reference's sed + assemble.py round trip would."""
    Ra = Ra_gap / (1 - ricb) ** 3
    return -Ra * Ek ** 2 / Prandtl


class AssembledPencil:
    """The pencil at any Rayleigh number, assembled ON THE GPU (kore_b200.assembly, kb_assemble):
    a trial of the search costs a new buoyancy factor in the assembly program and one pass of the
    assembly kernels; no matrix is built, updated or copied on the host (find_Rac.py:36-55 re-runs
    assemble.py -- tens of seconds -- and writes / re-reads A.npz for every trial).

    ``pp``: `assembly.PhysicsParams` of the run; ``operators``: the radial operators
    (`assembly.load_operators`), or None to generate them from ``pp`` (kore_b200/radial.py: no
    submatrices.py run either); ``factor_of``: Ra -> the buoyancy factor (`buoyancy_factor`).
    B does not depend on Ra: its norm is computed at the first trial and reused."""

    def __init__(self, pp, operators, factor_of, bnorm=None):
        from . import assembly as _asm
        self._asm = _asm
        if operators is None:
            from . import radial as _radial
            operators = _radial.radial_operators(pp, radprofs=_radial.run_profiles(pp))
        self.pp, self.ops, self.factor_of = pp, operators, factor_of
        self.progB = _asm.build_program_B(pp, operators)
        self.bnorm = bnorm
        if bnorm is not None:
            self.progB = self.progB.with_final_scale(1. / bnorm)

    def install(self, solver, Ra):
        a = self._asm
        if self.bnorm is None:
            solver.assemble(None, self.progB)
            self.bnorm = a.frobenius_norm(solver.get_assembled("B")[2])
            self.progB = self.progB.with_final_scale(1. / self.bnorm)
        q = a.PhysicsParams.from_dict({**self.pp.__dict__, "Beyonce": self.factor_of(Ra)})  # unknown keys dropped
        solver.assemble(a.build_program_A(q, self.ops).with_final_scale(1. / self.bnorm), self.progB)


class GrowthRate:
    """Largest growth rate max Re(lambda) of the ``nev`` pairs selected by ``which`` around
    ``tau`` at a given log10(Ra_gap) -- `get_sigma` of find_Rac.py:36-55 on one GPU handle.

    ``true_residual`` is find_Rac.py:18's ``-eps_true_residual``; with the reference's tol = 1e-15
    a true residual cannot pass in double precision, so it is off unless asked for (pair it with
    a tol >= 1e-13)."""

    def __init__(self, pencil, perm, nodeptr, tau, nev, which="TR", tol=1e-15, maxit=100,
                 true_residual=False, device=0):
        self.p = pencil
        self.perm, self.nodeptr = perm, nodeptr
        self.tau, self.nev, self.which = complex(tau), int(nev), which
        self.tol, self.maxit, self.true_residual = tol, maxit, true_residual
        self.solver = _lib.Solver(device)
        self.cache = {}
        self.history = []  # (Ra, lambda of largest real part, factor_ms, eigs_ms)

    def close(self):
        self.solver.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def eigenvalues(self, Ra):
        s = self.solver
        self.p.install(s, Ra)
        s.set_chain(self.perm, self.nodeptr)
        s.factor(self.tau)
        lam, _, info = s.eigs(self.nev, which=self.which, target=self.tau, tol=self.tol,
                              maxit=self.maxit, true_residual=self.true_residual, want_vectors=False)
        if info["nconv"] == 0:
            raise RuntimeError("no converged eigenpair at Ra = %r" % Ra)
        return lam, info

    def __call__(self, log10_Ra):
        Ra = 10.0 ** log10_Ra
        if Ra in self.cache:
            return self.cache[Ra].real
        lam, info = self.eigenvalues(Ra)
        best = lam[np.argmax(lam.real)]
        self.cache[Ra] = best
        self.history.append((Ra, best, info["factor_ms"], info["eigs_ms"]))
        return best.real


def find_rac(growth, Ra_min, dx=0.01, tol=1e-6, maxiter=200):
    """Critical Rayleigh number and drift frequency (find_Rac.py:100-113).

    ``growth`` maps log10(Ra) to the largest growth rate and keeps ``cache`` (a `GrowthRate`, or
    any callable with a ``cache`` dict of Ra -> complex eigenvalue).  Returns
    ``(Ra_c, omega_c, sigma_c)`` where the eigenvalue at Ra_c is ``sigma_c + i omega_c``."""
    x = bracket_brentq(growth, np.log10(Ra_min), dx=dx, tol=tol, maxiter=maxiter)
    growth(x)  # `runKoreRes`: one more evaluation at the root (cached when Brent ended on it)
    Ra_c = 10.0 ** x
    lam = growth.cache[Ra_c]
    return Ra_c, lam.imag, lam.real


def write_critical_params(path, Ek, ricb, Ra_c, m, omega_c):
    """Append the ``critical_params.dat`` row of find_Rac.py:115-119 (same formats)."""
    X = np.array([Ek, ricb, Ra_c, int(m), omega_c])
    with open(path, "a") as f:
        np.savetxt(f, X.reshape(1, X.shape[0]), fmt=["%.3e", "%.2f", "%.5e", "%d", "%.5e"])
