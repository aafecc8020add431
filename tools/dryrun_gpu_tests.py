#!/usr/bin/env python3
"""Dry run, on the CPU, of the Python side of GPU tests that have not run on a GPU yet (written after the
round's GPU budget was spent): `kore_b200.lib.Solver` is replaced by a stand-in that evaluates assembly
programs with the NumPy model of the assembly kernel (tests/assembly_model.py) and solves with the CPU oracle
(oracle/kore_oracle.py: SuperLU + ARPACK), and the test functions are called as pytest would call them.
It proves nothing about the kernels -- those are covered by the GPU runs of the other test files -- but it
does catch what a cross-compile cannot: wrong attribute names, shapes, file names, tolerances that the
reference data cannot meet.  Test infrastructure; never imported by the package.
Usage: tools/dryrun_gpu_tests.py [-m test_module] [test name substring ...]   (default module: the last-sorted
tests/test_zz_assembly_extensions.py; other modules' GPU tests may need calls the stand-in does not offer)"""
import inspect
import os
import pathlib
import sys
import tempfile

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import assembly_model as am  # noqa: E402
import kore_oracle as ko  # noqa: E402
from kore_b200 import lib as real_lib  # noqa: E402


class StandInSolver:
    def __init__(self, device=0):
        self.M, self.n, self.tau = {}, 0, 0.0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    def set_option(self, opt, value):
        pass

    def set_pencil(self, A, B=None):
        self.M = {"A": sp.csr_matrix(A, dtype=complex)}
        if B is not None:
            self.M["B"] = sp.csr_matrix(B)
        self.n = A.shape[0]

    def assemble(self, progA, progB=None):
        self.M = {}
        if progA is not None:
            self.M["A"] = am.evaluate(progA)
        if progB is not None:
            self.M["B"] = am.evaluate(progB)
        self.n = (progA if progA is not None else progB).n

    def get_assembled(self, which="A"):
        M = self.M[which]
        return M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data

    def set_chain(self, perm, nodeptr):
        assert len(perm) == self.n and nodeptr[-1] == self.n and sorted(perm) == list(range(self.n))

    def factor(self, sigma):
        self.tau = complex(sigma)
        T = self.M["A"] - self.tau * self.M["B"] if "B" in self.M else self.M["A"]
        self.lu = spl.splu(sp.csc_matrix(T))

    def solve(self, rhs):
        return self.lu.solve(np.asarray(rhs, dtype=complex))

    def eigs(self, nev, which="TM", target=None, ncv=0, tol=1e-15, maxit=50, true_residual=False, v0=None,
             max_pairs=None, want_vectors=True):
        lam, X, info = ko.eigs(self.M["A"], self.M["B"], self.tau, nev, which)
        res = ko.residuals(self.M["A"], self.M["B"], lam, X)
        X = X / np.linalg.norm(X, axis=0)
        return lam, (X if want_vectors else None), dict(nconv=len(lam), its=1, ncv=ncv or max(2 * nev, nev + 15),
                                                        resid=res, factor_ms=0.0, eigs_ms=0.0)

    def diagnose(self, p, nodes, X):
        # the NumPy model of kb_diagnose (tests/diag_model.py), whole radial domain as the tests use it
        import diag_model as dm
        X = np.asarray(X, dtype=complex).reshape(-1, 1) if np.ndim(X) == 1 else np.asarray(X, dtype=complex)
        meta = dict(N=p.N, N1=p.N1, n=p.N1 * p.nb, ricb=p.ricb, m=p.m, lmax=p.lmax, symm=p.symm, thermal=p.thermal)
        out = [dm.diagnose(X[:, k], meta, "differential" if p.heating == 0 else "internal") for k in range(X.shape[1])]
        return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])

    def stats(self):
        return {}


class StandInLib:
    Solver = StandInSolver
    LIB_PATH = real_lib.LIB_PATH


class MonkeyPatch:
    def chdir(self, d):
        os.chdir(str(d))


def main():
    only = sys.argv[1:]
    module = "test_zz_assembly_extensions"
    if only[:1] == ["-m"]:
        module, only = only[1], only[2:]
    real_lib.Solver = StandInSolver          # kore_b200.eps / rac / sweep construct their solvers through the module
    real_lib.savetxt = lambda path, X, part="real", append=False, nthreads=0: np.savetxt(
        path, (np.asarray(X).real if part == "real" else np.asarray(X).imag) if np.iscomplexobj(X) else X)
    import kore_b200.eps as eps
    eps.savetxt = real_lib.savetxt
    import importlib
    StandInLib.KoreB200Error = real_lib.KoreB200Error
    T = importlib.import_module(module)
    ran = failed = 0
    cwd = os.getcwd()
    for name, fn in sorted(vars(T).items()):
        if not (name.startswith("test_") and callable(fn)):
            continue
        marks = [m.name for m in getattr(fn, "pytestmark", [])]
        if "gpu" not in marks or (only and not any(o in name for o in only)):
            continue
        params = [()]
        for m in getattr(fn, "pytestmark", []):
            if m.name == "parametrize":
                params = [(v,) for v in m.args[1]]
        for p in params:
            kw = {}
            sig = inspect.signature(fn).parameters
            if "lib" in sig:
                kw["lib"] = StandInLib
            if "tmp_path" in sig:
                kw["tmp_path"] = pathlib.Path(tempfile.mkdtemp(prefix="dryrun_"))
            if "monkeypatch" in sig:
                kw["monkeypatch"] = MonkeyPatch()
            if p:
                kw[[k for k in sig if k not in kw][0]] = p[0]
            ran += 1
            try:
                fn(**kw)
                print("ok      %s%s" % (name, list(p) if p else ""), flush=True)
            except Exception as e:  # noqa: BLE001
                failed += 1
                import traceback
                print("FAILED  %s%s: %s: %s" % (name, list(p) if p else "", type(e).__name__, e), flush=True)
                traceback.print_exc(limit=4)
            finally:
                os.chdir(cwd)
                sys.modules.pop("parameters", None)
    print("%d dry runs, %d failed" % (ran, failed))
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
