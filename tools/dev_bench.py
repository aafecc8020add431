#!/usr/bin/env python3
"""Developer timing probe (GPU box): synthetic chain P x b -> factor / solve / eigs stats."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from kore_b200 import lib, synthetic

def main():
    sizes = [tuple(map(int, a.split("x"))) for a in sys.argv[1:]] or [(64, 68)]
    for P, b in sizes:
        t = time.time(); A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b); tg = time.time() - t
        n = A.shape[0]
        s = lib.Solver(0)
        t = time.time(); s.set_pencil(A, B); t1 = time.time() - t
        t = time.time(); s.set_chain(perm, nodeptr); t2 = time.time() - t
        for rep in range(2):
            t = time.time(); s.factor(1j); tf = time.time() - t
        st = s.stats()
        rhs = B @ synthetic.start_vector(n, 3)
        for rep in range(2):
            t = time.time(); x = s.solve(rhs); ts = time.time() - t
        st2 = s.stats()
        T = (A - 1j * B).tocsr()
        res = np.linalg.norm(T @ x - rhs) / np.linalg.norm(rhs)
        t = time.time(); lam, X, info = s.eigs(10, "TM", 1j, ncv=25, tol=1e-12, maxit=100, v0=synthetic.start_vector(n)); te = time.time() - t
        print("P=%d b=%d n=%d nnz=%d | gen %.1fs set_pencil %.2fs set_chain %.2fs | factor %.1f ms (wall %.3f) %.2f TF/s | solve(refine) %.2f ms/rhs wall %.3f resid %.1e | eigs %.1f ms wall %.2f nconv %d its %d applies %d maxres %.1e launches %d"
              % (P, b, n, A.nnz, tg, t1, t2, st["factor_ms"], tf, st["factor_flops"] / st["factor_ms"] / 1e9, st2["solve_ms"], ts, res,
                 info["eigs_ms"], te, info["nconv"], info["its"], info["op_applies"], info["resid"].max(), info["kernel_launches"]), flush=True)
        s.close()

if __name__ == "__main__":
    main()
