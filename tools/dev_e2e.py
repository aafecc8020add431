#!/usr/bin/env python3
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from kore_b200 import lib, synthetic
P, b = 600, 600
A, B, perm, nodeptr = synthetic.synthetic_pencil(P, b)
v0 = synthetic.start_vector(A.shape[0], 1)
s = lib.Solver(0)
for rep in range(3):
    t = [time.perf_counter()]
    s.set_pencil(A, B); t.append(time.perf_counter())
    s.set_chain(perm, nodeptr); t.append(time.perf_counter())
    s.factor(1j); t.append(time.perf_counter())
    lam, X, info = s.eigs(10, "TM", 1j, ncv=25, tol=1e-12, maxit=100, v0=v0, want_vectors=True); t.append(time.perf_counter())
    d = np.diff(t)
    print("rep %d: set_pencil %.3f set_chain %.3f factor %.3f (dev %.3f) eigs %.3f (dev %.3f) total %.3f" % (rep, d[0], d[1], d[2], info["factor_ms"]/1e3, d[3], info["eigs_ms"]/1e3, sum(d)), flush=True)
