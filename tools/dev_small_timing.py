#!/usr/bin/env python3
"""Dev: wall vs device time of kb_factor / kb_solve on the small forced-libration fixture (n = 1600)."""
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case
from kore_b200 import lib
ce, cf = load_case("forced_small_eig"), load_case("forced_small")
f = np.asarray(cf.bf.todense()).ravel().astype(np.complex128)
for refine in (1, 0):
    s = lib.Solver(0)
    s.set_option(lib.OPT_REFINE, refine)
    s.set_pencil(ce.A, ce.B); s.set_chain(ce.perm, ce.nodeptr)
    s.factor(0.3j); s.solve(f)
    tf = ts = df = ds = 0.0
    N = 100
    for k in range(N):
        t0 = time.perf_counter(); s.factor(1j * (0.1 + 0.01 * k)); t1 = time.perf_counter()
        x = s.solve(f); t2 = time.perf_counter()
        st = s.stats()
        tf += t1 - t0; ts += t2 - t1; df += st["factor_ms"]; ds += st["solve_ms"]
    print("refine %d: factor wall %.3f ms device %.3f ms | solve wall %.3f ms device %.3f ms | launches/point %.0f"
          % (refine, tf / N * 1e3, df / N, ts / N * 1e3, ds / N, 0), flush=True)
    s.close()
