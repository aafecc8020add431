#!/usr/bin/env python3
"""Dev: time the chain sweep variants on the synthetic pencil (folded vs one-hop) and check them
against each other.  usage: dev_fold_timing.py P b"""
# cycle counters need the timing build: `make timing`, then KB_LIB_PATH=kore_b200/libkoreb200_timing.so
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from kore_b200 import lib, synthetic
P, b = int(sys.argv[1]), int(sys.argv[2])
A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
rhs = B @ synthetic.start_vector(A.shape[0], 3)
T = (A - 1j * B).tocsr()
xs = {}
for fold in (0, 1):
    s = lib.Solver(0)
    s.set_option(lib.OPT_REFINE, 0)
    s.set_option(lib.OPT_FOLD, fold)
    s.set_pencil(A, B); s.set_chain(perm, nodeptr)
    s.factor(1j)
    f_ms = s.stats()["factor_ms"]
    for i in range(3):
        x = s.solve(rhs)
    xs[fold] = x
    res = np.linalg.norm(T @ x - rhs) / np.linalg.norm(rhs)
    for rep in range(2):
        t0 = time.perf_counter()
        lam, X, info = s.eigs(10, "TM", 1j, ncv=25, tol=1e-12, maxit=100, v0=synthetic.start_vector(A.shape[0]), want_vectors=False)
        print("   eigs call %d: wall %.1f ms, eigs_ms %.1f, sweeps %.1f ms" % (rep, (time.perf_counter() - t0) * 1e3, info["eigs_ms"], info["eigs_solve_ms"]))
    print("fold=%d factor_ms %.2f  residual %.2e  eigs_ms %.1f  sweeps %d  ms/sweep %.3f  nconv %d" % (
        fold, f_ms, res, info["eigs_ms"], info["solve_calls"], info["eigs_solve_ms"] / max(1, info["solve_calls"]), info["nconv"]))
    if os.environ.get("KB_SWEEP_TIMING") and fold == 1:
        L = lib.load()
        out = np.zeros(256 * 8, dtype=np.int64)
        L.kb_dbg_sweep_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        g = L.kb_dbg_sweep_timing(s.h, out.ctypes.data, 256)
        t = out[: g * 8].reshape(g, 8)
        names = ["stage->regs", "wait vector", "fma", "reduce+publish", "-", "-", "-", "-"]
        for k in range(5):
            print("  %-15s mean %8.0f  max %8.0f  (cycles per chain step)" % (names[k], t[:, k].mean() / P, t[:, k].max() / P))
    s.close()
print("fold vs onehop rel diff %.2e" % (np.linalg.norm(xs[0] - xs[1]) / np.linalg.norm(xs[0])))
