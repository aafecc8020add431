"""EPS / KSP facades over libkoreb200 with the slepc4py / petsc4py call surface that
Kore's bin/solve.py uses.

A Kore maintainer can keep the body of solve.py and swap the two imports:
every method below names the slepc4py/petsc4py call it stands in for
(/root/reference/bin/solve.py line numbers).  Only what solve.py calls is
provided; this is not a PETSc re-implementation.  All numerics run on the GPU
through the C ABI (kore_b200/lib.py); nothing here computes on the CPU.
"""
from __future__ import annotations

import numpy as np

from . import chain as _chain
from . import lib as _lib
from .lib import savetxt  # noqa: F401  (np.savetxt-compatible threaded writer, kb_savetxt)


class Options:
    """PETSc.Options() stand-in: a parsed view of the `$opts` string on argv
    (solve.py:16, 34).  PETSc options are `-name [value]`."""

    def __init__(self, argv=None):
        self.opts = {}
        argv = list(argv or [])
        i = 0
        while i < len(argv):
            a = argv[i]
            if a.startswith("-") and not _is_number(a):
                key = a.lstrip("-")
                if i + 1 < len(argv) and (not argv[i + 1].startswith("-") or _is_number(argv[i + 1])):
                    self.opts[key] = argv[i + 1]
                    i += 2
                else:
                    self.opts[key] = True
                    i += 1
            else:
                i += 1

    def getInt(self, name, default=None):
        v = self.opts.get(name, default)
        return default if v is True else (int(v) if v is not None else None)

    def getReal(self, name, default=None):
        v = self.opts.get(name, default)
        return default if v is True else (float(v) if v is not None else None)

    def getScalar(self, name, default=None):
        v = self.opts.get(name, None)
        if v is None or v is True:
            return default
        return complex(str(v).replace("i", "j").replace(" ", ""))

    def getString(self, name, default=None):
        v = self.opts.get(name, default)
        return default if v is True else v

    def hasName(self, name):
        return name in self.opts


def _is_number(s):
    try:
        complex(s.replace("i", "j"))
        return True
    except ValueError:
        return False


class Which:
    """SLEPc.EPS.Which (solve.py:99-117)."""
    LARGEST_MAGNITUDE = "LM"
    SMALLEST_MAGNITUDE = "SM"
    LARGEST_REAL = "LR"
    SMALLEST_REAL = "SR"
    LARGEST_IMAGINARY = "LI"
    SMALLEST_IMAGINARY = "SI"
    TARGET_MAGNITUDE = "TM"
    TARGET_REAL = "TR"
    TARGET_IMAGINARY = "TI"


class ProblemType:
    GNHEP = "gnhep"


class ChainLayout:
    """What replaces the sparse solver's ordering phase: Kore's own l-major chain."""

    def __init__(self, perm, nodeptr):
        self.perm = np.ascontiguousarray(perm, dtype=np.int64)
        self.nodeptr = np.ascontiguousarray(nodeptr, dtype=np.int64)

    @classmethod
    def from_kore(cls, A, N1, m, lmax, symm, symmB0=-1, hydro=1, magnetic=0, thermal=0, compositional=0):
        perm, nodeptr = _chain.chain_from_params(N1, m, lmax, symm, symmB0, hydro, magnetic, thermal,
                                                 compositional)
        if perm.size != A.shape[0] or not _chain.check_block_tridiagonal(A.indptr, A.indices, perm, nodeptr):
            # e.g. quadrupolar B0 (l +- 2 couplings): derive the chain from the pattern instead
            perm, nodeptr = _chain.chain_from_pattern(A.indptr, A.indices, A.shape[0], N1)
        return cls(perm, nodeptr)

    @classmethod
    def from_params(cls, N1, m, lmax, symm, symmB0=-1, hydro=1, magnetic=0, thermal=0, compositional=0):
        """The l-major chain from the parameter set alone (no matrix on the host: device-side assembly)."""
        return cls(*_chain.chain_from_params(N1, m, lmax, symm, symmB0, hydro, magnetic, thermal, compositional))


class EPS:
    """SLEPc.EPS stand-in (solve.py:91-149): GNHEP, Krylov-Schur, shift-and-invert."""

    Which = Which
    ProblemType = ProblemType

    def __init__(self, device=0):
        self._solver = _lib.Solver(device)
        self._A = self._B = None
        self._layout = None
        self.nev, self.ncv = 1, 0
        self.tol, self.max_it = 1e-8, 100
        self.which = "LM"
        self.target = 0.0
        self.true_residual = False
        self.print_errors = False
        self._result = None

    def create(self, comm=None):        # E.create(SLEPc.COMM_WORLD)            solve.py:92
        return self

    def setOperators(self, A, B):       # E.setOperators(MA, MB)                solve.py:93
        self._A, self._B = A.tocsr(), B.tocsr()
        self._asm = None

    def setAssembly(self, params, operators):
        """(new) instead of setOperators: the pencil is assembled on the GPU from the radial
        operators (kore_b200.assembly; replaces the assemble.py run + A.npz / B.npz)."""
        self._asm = (params, operators)
        self._A = self._B = None

    def setChainLayout(self, layout):   # (new) replaces MUMPS' analysis phase
        self._layout = layout

    def setProblemType(self, t):        # E.setProblemType(GNHEP)               solve.py:94
        if t != ProblemType.GNHEP:
            raise ValueError("only GNHEP is supported (what Kore uses)")

    def setDimensions(self, nev, ncv=None):   # E.setDimensions(par.nev)        solve.py:96
        self.nev = int(nev)
        self.ncv = int(ncv) if ncv else 0

    def setTolerances(self, tol, max_it):     # E.setTolerances(par.tol, par.maxit)  solve.py:97
        self.tol, self.max_it = float(tol), int(max_it)

    def setWhichEigenpairs(self, which):      # solve.py:99-117
        self.which = which

    def setTarget(self, tau):                 # E.setTarget(par.tau)            solve.py:119
        self.target = complex(tau)

    def setFromOptions(self, options=None):   # E.setFromOptions()              solve.py:120
        o = options
        if o is None:
            return
        st = o.getString("st_type", None)
        if st not in (None, "sinvert"):
            raise ValueError("only -st_type sinvert is supported (every Kore invocation uses it)")
        if o.hasName("eps_nev"):
            self.nev = o.getInt("eps_nev")
        if o.hasName("eps_ncv"):
            self.ncv = o.getInt("eps_ncv")
        if o.hasName("eps_tol"):
            self.tol = o.getReal("eps_tol")
        if o.hasName("eps_max_it"):
            self.max_it = o.getInt("eps_max_it")
        if o.hasName("eps_target"):
            self.target = o.getScalar("eps_target")
        self.true_residual = o.hasName("eps_true_residual")
        self.print_errors = o.hasName("eps_error_relative")
        # accepted and ignored: solver-package selection and its knobs (the factorisation is
        # ours), -eps_balance (equilibration is always on), -nbl
        return

    def solve(self):                           # E.solve()                      solve.py:123
        if self._layout is None:
            raise RuntimeError("setChainLayout must be called before solve()")
        s = self._solver
        if getattr(self, "_asm", None) is not None:
            from . import assembly as _assembly
            _assembly.assemble(s, *self._asm)
        else:
            s.set_pencil(self._A, self._B)
        s.set_chain(self._layout.perm, self._layout.nodeptr)
        s.factor(self.target)
        lam, X, info = s.eigs(self.nev, which=self.which, target=self.target, ncv=self.ncv, tol=self.tol,
                              maxit=self.max_it, true_residual=self.true_residual)
        self._result = (lam, X, info)
        if self.print_errors:
            print(" %-28s %s" % ("k", "||Ax-kBx||/||kBx||"))
            for l, r in zip(lam, info["resid"]):
                print(" %+.9f%+.9fi   %.5g" % (l.real, l.imag, r))

    def getIterationNumber(self):   # solve.py:131
        return self._result[2]["its"]

    def getType(self):              # solve.py:132
        return "krylovschur"

    def getTolerances(self):        # solve.py:133
        return self.tol, self.max_it

    def getDimensions(self):        # solve.py:134
        ncv = self._result[2]["ncv"] if self._result else (self.ncv or max(2 * self.nev, self.nev + 15))
        return self.nev, ncv, ncv

    def getConverged(self):         # solve.py:135
        return self._result[2]["nconv"]

    def getTarget(self):            # solve.py:138
        return self.target

    def getEigenpair(self, i, vr=None, vi=None):   # solve.py:149
        lam, X, _ = self._result
        if vr is not None:
            vr[...] = X[:, i]
        return lam[i]

    def getStats(self):
        return self._result[2] if self._result else self._solver.stats()

    def destroy(self):
        self._solver.close()


class KSP:
    """PETSc.KSP stand-in for the forced problem (solve.py:220-227): preonly + LU."""

    def __init__(self, device=0):
        self._solver = _lib.Solver(device)
        self._A = None
        self._layout = None

    def create(self, comm=None):
        return self

    def setOperators(self, A):             # K.setOperators(MA)                  solve.py:222
        self._A = A.tocsr()
        self._asm = None

    def setAssembly(self, params, operators):
        """(new) instead of setOperators: A is assembled on the GPU (kore_b200.assembly)."""
        self._asm = (params, operators)
        self._A = None

    def setChainLayout(self, layout):
        self._layout = layout

    def setTolerances(self, rtol=None, max_it=None):   # solve.py:223 (direct solve: unused)
        return

    def setFromOptions(self, options=None):            # solve.py:224
        return

    def solve(self, b, x):                 # K.solve(bvec, x)                    solve.py:227
        s = self._solver
        if getattr(self, "_asm", None) is not None:
            from . import assembly as _assembly
            _assembly.assemble(s, *self._asm)
        else:
            s.set_pencil(self._A, None)
        s.set_chain(self._layout.perm, self._layout.nodeptr)
        s.factor(0.0)
        x[...] = s.solve(np.asarray(b, dtype=np.complex128))

    def getStats(self):
        return self._solver.stats()

    def destroy(self):
        self._solver.close()
